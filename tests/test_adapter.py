"""The C++ drop-in adapter (rovaca_b200/adapter/call_pairhmm_b200.cpp): same signature and types as the reference's
call_pairhmm, compiled against the reference's headers (oracle/Makefile), bound through the reference's global function
pointer.  CPU part: it builds, exports, and fails loudly without a GPU.  GPU part: byte-identical to the reference's own
call_pairhmm (oracle/_ref, the unmodified reference compiled in place) on the same ReadRecord/Haplotype objects."""
import ctypes

import numpy as np
import pytest

from oracle import bindings as ob
from rovaca_b200 import workloads as wl


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:       # noqa: BLE001
        return False


@pytest.fixture(scope="module")
def adapter():
    a = ob.adapter()
    if a is None:
        pytest.skip("oracle/_ref/libphmm_adapter_test.so not built (needs /root/reference)")
    return a


def test_adapter_builds_and_exports(adapter):
    for name in ob.Adapter.EXPORTS:
        assert hasattr(adapter.lib, name), name


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_adapter_fails_loudly_without_gpu(adapter):
    with pytest.raises(RuntimeError) as ei:
        adapter.init()
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(RuntimeError):
        adapter.region([dict(bases=b"ACGT", qual=np.full(4, 30, np.uint8), mapq=60)], [b"ACGTACGT"])


@pytest.mark.gpu
def test_adapter_equals_reference_call_pairhmm(adapter, ref):
    """rovaca::call_pairhmm re-bound to the B200 engine returns the same bytes and leaves the same reads in `rs` as the
    reference's own call_old_pairhmm / call_new_pairhmm002 on identical ReadRecord / Haplotype inputs."""
    adapter.init()
    rng = np.random.default_rng(31)
    specs = [dict(n_reads=40, n_haps=6, hap_lo=300, hap_hi=300, read_lo=150, read_hi=150, junk=0.1),
             dict(n_reads=33, n_haps=9, hap_lo=250, hap_hi=450, read_lo=100, read_hi=151, junk=0.2),
             dict(n_reads=7, n_haps=2, hap_lo=41, hap_hi=120, read_lo=20, read_hi=41, junk=0.0),
             dict(n_reads=20, n_haps=5, hap_lo=600, hap_hi=700, read_lo=200, read_hi=250, junk=0.1)]
    for spec in specs:
        for model in (ob.PCR_CONSERVATIVE, ob.PCR_NONE, ob.PCR_AGGRESSIVE):
            reads, haps = wl._one_region(rng, **spec)
            reads.sort(key=lambda r: len(r["bases"]))        # call_new_pairhmm002 asserts ascending lengths (:602)
            for path in (0, -1):
                m0, k0 = ref.region(reads, haps, 18, model, path=path)
                m1, k1 = adapter.region(reads, haps, 18, model)
                assert np.array_equal(k0, k1), (spec, model, path)
                assert np.array_equal(m0.view(np.uint64), m1.view(np.uint64)), (spec, model, path)
    # empty inputs give the reference's empty matrix
    m, k = adapter.region([], [b"ACGT"])
    assert len(k) == 0
    adapter.shutdown()


@pytest.mark.gpu
def test_adapter_window_and_devices(adapter, ref):
    """The asynchronous use of the boundary (call_pairhmm_b200_window: submit a window of regions, collect them in order) on an
    engine that drives several devices (init_pairhmm_ptr_b200_devices; two queues on the one GPU of a single-GPU box): the
    same bytes and the same surviving reads as the reference's own call_pairhmm, region by region."""
    import torch
    n = torch.cuda.device_count()
    adapter.shutdown()
    adapter.init_devices(list(range(n)) if n >= 2 else [0, 0])
    b = wl.stream("cfg5", 96, seed=17, post_trim_fraction=0.5)
    lik, kept, nk = adapter.window(b)
    lik0, kept0, nk0 = ref.regions_flat(b, 8)
    assert np.array_equal(nk, nk0) and np.array_equal(kept, kept0)
    assert np.array_equal(lik.view(np.uint64), lik0.view(np.uint64))
    reads, haps = b.region(3)                                     # the synchronous form still works on the same engine
    m0, k0 = ref.region(reads, haps, 18, ob.PCR_CONSERVATIVE)
    m1, k1 = adapter.region(reads, haps, 18, ob.PCR_CONSERVATIVE)
    assert np.array_equal(k0, k1) and np.array_equal(m0.view(np.uint64), m1.view(np.uint64))
    adapter.shutdown()
