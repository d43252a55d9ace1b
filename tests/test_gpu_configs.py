"""GPU: BASELINE.json configs[2..4] at their stated shapes, the region stream at >= 100 k regions, one engine over several
devices, and the out-of-envelope shapes — all through the C ABI, byte for byte against the CPU checker.

The checker is the reference's own call_pairhmm compiled in place (oracle/_ref) where it is present (it is on the GPU box),
else the scalar port on a smaller sample."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from rovaca_b200 import PairHmmEngine, streams
from rovaca_b200 import workloads as wl
from rovaca_b200.pairhmm import EPILOGUE_DEVICE, EPILOGUE_HOST, FP64_ALWAYS, FP64_WHEN_OBSERVABLE, FlatResult

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("PHMM_B200_POISON", "1")          # a pair no kernel wrote must be loud (NaN), not stale


def _check(batch, result, n_sample):
    from oracle import check
    rec = check.sample_parity(batch, result, n_sample=n_sample)
    assert rec["ok"], rec
    return rec


def _have_ref():
    from oracle import bindings as ob
    return ob.ref() is not None


@pytest.mark.parametrize("fp64_mode", [FP64_WHEN_OBSERVABLE, FP64_ALWAYS])
def test_cfg3_at_its_stated_shape(fp64_mode):
    """configs[2]: 500 reads x 32 haplotypes per region, R 100-151, H 250-450, 5 % junk reads that send pairs down the FP64
    branch — with the recompute skipped (default: its values cannot reach the matrix) and forced."""
    n = 6 if _have_ref() else 1
    b = wl.stream("cfg3", n, seed=3)
    assert np.all(np.diff(b.reg_read_off) == 500) and np.all(np.diff(b.reg_hap_off) == 32)
    eng = PairHmmEngine(fp64_mode=fp64_mode)
    try:
        res = eng.run_flat(b)
        assert res.n_fp64 > 0.03 * b.pairs()                      # the junk reads underflow against every haplotype
        _check(b, res, n)
        again = eng.run_flat(b)
        assert np.array_equal(res.likelihoods.view(np.uint64), again.likelihoods.view(np.uint64))
    finally:
        eng.close()


def test_cfg4_long_and_mixed(engine):
    """configs[3]: 250 bp x 1000 bp (16 haplotypes x 64 reads), alone and alternating with 150 x 300 regions."""
    n = 8 if _have_ref() else 2
    long_ = wl.stream("cfg4", n, seed=4)
    res = engine.run_flat(long_)
    _check(long_, res, n)
    both = wl.concat([wl.stream("cfg4", n // 2, seed=4), wl.stream("cfg2", n // 2, seed=4)])
    mix = wl.take(both, np.arange(n).reshape(2, n // 2).T.reshape(-1))
    res = engine.run_flat(mix)
    _check(mix, res, n)
    # the same regions give the same bytes whatever they are batched with
    lo, llo = res.lik_off, long_.lik_off()
    solo = engine.run_flat(long_)
    for k in range(n // 2):                                       # mix region 2k is long_ region k
        a = res.likelihoods[int(lo[2 * k]):int(lo[2 * k + 1])]
        c = solo.likelihoods[int(llo[k]):int(llo[k + 1])]
        assert np.array_equal(a.view(np.uint64), c.view(np.uint64))


@pytest.mark.parametrize("post_trim", [0.0, 0.5, 1.0])
def test_cfg5_stream_of_100k_regions(post_trim):
    """configs[4] reduced to matrix identity: a stream of 102 400 regions in pieces through phmm_run_flat; a sample of the
    regions byte-compared with the reference's CPU path, every region's digest identical on a second run and on an engine with
    the epilogue on the host."""
    spec = streams.StreamSpec("cfg5", 102400, 5, post_trim)
    eng = PairHmmEngine()
    try:
        kept = []
        first = streams.run_stream(eng, spec, keep_results=kept, piece_regions=51200)
        assert len(first.digests) == spec.n_regions and first.pieces == 2
        for ids, batch, res in kept:
            _check(batch, res, 64 if _have_ref() else 4)
        del kept
        second = streams.run_stream(eng, spec, warm=False, piece_regions=25600)          # other piece size, other batching
        assert np.array_equal(first.region_ids, second.region_ids) and np.array_equal(first.digests, second.digests)
    finally:
        eng.close()
    if post_trim == 0.5:
        host = PairHmmEngine(epilogue=EPILOGUE_HOST)
        try:
            sub = streams.StreamSpec("cfg5", 20480, 5, post_trim)
            third = streams.run_stream(host, sub, warm=False)
            assert np.array_equal(third.digests, first.digests[:20480])
        finally:
            host.close()


def _ordinals():
    import torch
    n = torch.cuda.device_count()
    return list(range(n)) if n >= 2 else [0, 0]           # one GPU: two independent queues on it exercise the same front


def test_one_engine_over_several_devices_flat_stream():
    """phmm_create_multi: one process, one engine, every GPU of the box (two queues on the one GPU of a single-GPU box).  The
    stream's digests equal a single-device engine's; every device did a share of the work."""
    spec = streams.StreamSpec("cfg5", 40960, 7, 0.5)
    one = PairHmmEngine()
    many = PairHmmEngine(devices=_ordinals())
    try:
        assert many.n_devices == len(_ordinals()) and many.device_ordinals() == _ordinals()
        a = streams.run_stream(one, spec, piece_regions=20480)
        b = streams.run_stream(many, spec, piece_regions=20480, warm=False)
        assert np.array_equal(a.digests, b.digests)
        per_dev = [many.device_stats(k) for k in range(many.n_devices)]
        total = many.stats()
        assert sum(d["cells"] for d in per_dev) == total["cells"] == b.cells
        share = np.array([d["cells"] for d in per_dev], float) / total["cells"]
        assert share.min() > 0.5 / many.n_devices, share                                   # least-outstanding-work dispatch
        assert all(d["batches"] > 0 for d in per_dev)
    finally:
        one.close()
        many.close()


def test_one_engine_over_several_devices_synchronous_callers(orc):
    """The reference's calling pattern (one region per call from many pool workers, haplotypecaller_engine.cpp:250) against a
    multi-device engine: every caller gets the bytes a single-device engine gives."""
    b = wl.stream("cfg5", 192, seed=11, post_trim_fraction=0.5)
    one = PairHmmEngine(epilogue=EPILOGUE_DEVICE)
    many = PairHmmEngine(devices=_ordinals(), epilogue=EPILOGUE_DEVICE)
    try:
        want = one.run_flat(b)
        got = [None] * b.n_regions

        def caller(t):
            for g in range(t, b.n_regions, 16):
                reads, haps = b.region(g)
                got[g] = many.call_pairhmm(haps, reads)
        th = [threading.Thread(target=caller, args=(t,)) for t in range(16)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for g in range(b.n_regions):
            m, kept = got[g]
            assert np.array_equal(m.view(np.uint64), want.matrix(g).view(np.uint64)) and np.array_equal(kept, want.kept(g))
        assert all(many.device_stats(k)["regions"] > 0 for k in range(many.n_devices))
        reads, haps = b.region(5)
        m0, k0, *_ = orc.region(reads, haps)
        assert np.array_equal(got[5][0].view(np.uint64), m0.view(np.uint64)) and np.array_equal(got[5][1], k0)
    finally:
        one.close()
        many.close()


def test_a_callers_own_region_runs_on_its_thread_and_leaves_its_device_alone(monkeypatch):
    """A synchronous caller whose batch is its own single region runs it on the calling thread (caller slots, results
    written to mapped pinned memory): same bytes as with every batch handed to a worker (PHMM_B200_CALLER_RUNS=0), the calling
    thread's current CUDA device — switched to the engine's for the call — is what it was afterwards, whichever device served."""
    import torch
    b = wl.stream("cfg5", 24, seed=13, post_trim_fraction=0.5)
    regions = [b.region(g) for g in range(b.n_regions)]
    ords = _ordinals()
    before = torch.cuda.current_device()
    eng = PairHmmEngine(devices=list(reversed(ords)))                # the engine's first device is not the caller's when there are two
    try:
        got = [eng.call_pairhmm(haps, reads) for reads, haps in regions]       # (no raw sums asked: the short device path)
        assert torch.cuda.current_device() == before
        assert eng.stats()["batches"] == len(regions)
    finally:
        eng.close()
    monkeypatch.setenv("PHMM_B200_CALLER_RUNS", "0")
    eng = PairHmmEngine(devices=ords)
    try:
        for (reads, haps), (m1, k1) in zip(regions, got):
            m0, k0 = eng.call_pairhmm(haps, reads)
            assert np.array_equal(m0.view(np.uint64), m1.view(np.uint64)) and np.array_equal(k0, k1)
    finally:
        eng.close()


def test_sharded_stream_under_torchrun():
    """bench.py's `sharded_stream` record with one rank per GPU: needs two GPUs."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: the sharded stream over ranks needs two (the multi-device engine tests above cover the dispatch)")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "1", "--warmup", "3", "--regions", "1024",
           "--no-workloads", "--no-cpu-baseline", "--sustain", "0", "--stream-regions", "32768"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    rec = line["sharded_stream"]
    assert rec["n_gpus"] == n and rec["identical_to_single_gpu"] is True and rec["regions_differing_from_single_gpu"] == 0


def test_haplotypes_outside_the_envelope_between_normal_ones(engine, orc):
    """an empty or over-long (> 22 000 columns) haplotype BETWEEN two normal ones of one chunk used to lose a pair
    (stale value).  [ok, over-long, ok], [ok, empty, ok], odd and even counts, first / last position — bit for bit."""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def hap(n):
        return acgt[rng.integers(0, 4, n)]
    reads, base_haps = wl._one_region(rng, 9, 2, 300, 300, 60, 140, junk=0.1)
    orders = [
        [base_haps[0], hap(23000), base_haps[1]],
        [base_haps[0], np.zeros(0, np.uint8), base_haps[1]],
        [hap(23000), base_haps[0], np.zeros(0, np.uint8), base_haps[1], hap(280)],
        [base_haps[0], base_haps[1], hap(310), hap(22500), hap(290), np.zeros(0, np.uint8)],
    ]
    for haps in orders:
        if any(len(h) == 0 for h in haps):
            # the reference divides by the haplotype length (2^120 / 0 = inf): only shapes it can run are compared with it
            m, kept, raw32, *_ = engine.call_pairhmm(haps, reads, raw=True)
            assert not np.isnan(raw32[:, [j for j, h in enumerate(haps) if len(h)]]).any()
            live = [h for h in haps if len(h)]
            m2, kept2, raw2, *_ = engine.call_pairhmm(live, reads, raw=True)
            cols = [j for j, h in enumerate(haps) if len(h)]
            assert np.array_equal(raw32[:, cols].view(np.uint32), raw2.view(np.uint32))
            continue
        m0, k0, o32, *_ = orc.region(reads, haps)
        m, kept, raw32, *_ = engine.call_pairhmm(haps, reads, raw=True)
        assert np.array_equal(raw32.view(np.uint32), o32.view(np.uint32))
        assert np.array_equal(m.view(np.uint64), m0.view(np.uint64)) and np.array_equal(kept, k0)


def test_several_very_long_haplotypes_fit_shared_memory(engine, orc):
    """three 15 kb haplotypes used to land head to tail in one stream (240 KB of shared memory): launch failure for
    the whole batch.  Now the chunk takes fewer of them."""
    rng = np.random.default_rng(78)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    haps = [acgt[rng.integers(0, 4, n)] for n in (15000, 14800, 15100)]
    reads = []
    for k in range(5):
        o = int(rng.integers(0, 14000))
        reads.append(dict(bases=haps[k % 3][o:o + 100 + k].copy(), qual=np.full(100 + k, 30, np.uint8), mapq=60))
    m0, k0, o32, *_ = orc.region(reads, haps)
    m, kept, raw32, *_ = engine.call_pairhmm(haps, reads, raw=True)
    assert np.array_equal(raw32.view(np.uint32), o32.view(np.uint32))
    assert np.array_equal(m.view(np.uint64), m0.view(np.uint64)) and np.array_equal(kept, k0)


def test_unknown_pcr_model_is_a_no_op(engine, orc):
    """apply_pcr_error_model's switch has a default: break (pairhmm_internal.cpp:121-159): any value outside 1..3 leaves the
    gap qualities alone, like PHMM_PCR_NONE — it must not index the cache table."""
    rng = np.random.default_rng(79)
    reads, haps = wl._one_region(rng, 12, 3, 200, 200, 80, 120)
    want = engine.call_pairhmm(haps, reads, 18, 0, raw=True)
    for model in (4, 7, 15, -1, 99):
        got = engine.call_pairhmm(haps, reads, 18, model, raw=True)
        assert np.array_equal(got[2].view(np.uint32), want[2].view(np.uint32))
    m0, k0, o32, *_ = orc.region(reads, haps, 18, 0)
    assert np.array_equal(want[2].view(np.uint32), o32.view(np.uint32))


def test_gpu_ms_is_a_union_not_a_sum(engine):
    """phmm_stats.gpu_ms over concurrent batches must not exceed the wall time they took."""
    import time
    b = wl.stream("cfg2", 2048, seed=2)
    engine.run_flat(b)
    s0 = engine.stats()
    t0 = time.perf_counter()
    for _ in range(3):
        engine.run_flat(b)
    wall = (time.perf_counter() - t0) * 1e3
    s1 = engine.stats()
    gpu = s1["gpu_ms"] - s0["gpu_ms"]
    assert 0 < gpu <= wall * 1.02, (gpu, wall)
    assert s1["batches"] - s0["batches"] >= 6


def test_device_marginalisation_of_resident_matrices(ref):
    """§8f item 4: AlleleLikelihoods::marginalize (genotype/allele_likelihoods.hpp:189-216, 702-732) for many sites on the
    matrices a resident batch left on the device, against tests/golden/alleles.npz — the UNMODIFIED reference's
    AlleleLikelihoods on the reference's own call_pairhmm matrices.  The regions are regenerated with the generating script's
    seed (tests/golden/make_golden.py::make_alleles), so the device's matrices must equal the golden ones first."""
    import _golden
    gold = _golden.alleles()
    rng = np.random.default_rng(20261019)
    regs, sites = [], []
    for k, gcase in enumerate(gold):
        reads, haps = wl._one_region(rng, int(rng.integers(1, 60)), int(rng.integers(1, 9)), 100, 150, 30, 90, junk=0.15)
        haps = list(haps)
        for _ in range(int(rng.integers(0, 4))):
            haps.insert(int(rng.integers(0, len(haps) + 1)), haps[int(rng.integers(0, len(haps)))].copy())
        rng.integers(-3, 3, len(haps))                       # the priorities the script drew
        n_new = int(rng.integers(1, 5))
        mp = rng.integers(-1, n_new, len(haps)).astype(np.int32)
        assert np.array_equal(mp, gcase["old_to_new"]) and gcase["matrix"].shape == (len(haps), len(reads))
        regs.append((reads, haps))
        sites.append((k, mp, n_new + 1))
    batch = wl.RegionBatch.from_regions(regs)
    for epilogue in (EPILOGUE_DEVICE, 0):                    # forced, and as a default engine places it (resident batches: device)
        eng = PairHmmEngine(epilogue=epilogue)
        try:
            res = eng.upload(batch)
            res.run()
            got = res.marginalize(sites)
            full = res.download()
            for k, gcase in enumerate(gold):
                n = gcase["n_evidence"]
                assert int(full.n_kept[k]) == n
                assert np.array_equal(full.matrix(k).view(np.uint64), gcase["matrix"].view(np.uint64))
                assert np.array_equal(got[k][:, :n].view(np.uint64), gcase["marginal"].view(np.uint64)), k
                assert np.isnan(got[k][:, n:]).all()
            # several sites on one region, in any order; the host form gives the same
            from rovaca_b200 import marginalize
            k = max(range(len(gold)), key=lambda i: gold[i]["matrix"].shape[0])
            nh = gold[k]["matrix"].shape[0]
            maps = [np.arange(nh, dtype=np.int32) % 3, np.full(nh, -1, np.int32), np.zeros(nh, np.int32)]
            out = res.marginalize([(k, maps[2], 1), (0, sites[0][1], sites[0][2]), (k, maps[0], 3), (k, maps[1], 2)])
            n = gold[k]["n_evidence"]
            for o, (mp, n_new) in zip([out[2], out[3], out[0]], [(maps[0], 3), (maps[1], 2), (maps[2], 1)]):
                want = marginalize(full.matrix(k), n, mp, n_new)
                assert np.array_equal(o[:, :n].view(np.uint64), want.view(np.uint64))
            assert np.array_equal(out[1].view(np.uint64), got[0].view(np.uint64))
            with pytest.raises(Exception):
                res.marginalize([(k, np.full(nh, 5, np.int32), 2)])
            res.free()
        finally:
            eng.close()
    host = PairHmmEngine(epilogue=EPILOGUE_HOST)
    try:
        res = host.upload(batch)
        res.run()
        with pytest.raises(Exception) as ei:
            res.marginalize(sites[:1])
        assert "phmm_marginalize" in str(ei.value)
        res.free()
    finally:
        host.close()


def test_fast_fp64_kernel_equals_the_general_one(orc, monkeypatch):
    """The FP64 recompute has a fast kernel (one strip per pair, no flush test per operation, a range check on M instead) and a
    general one (any shape, flushes after every operation as the reference's FTZ does).  Same bits from both on junk reads of
    1 to 5 rows per lane, on reads beyond the fast kernel's 159 bases, and on pairs the fast kernel hands back because its
    range check trips (forced here by raising the check's floor from 2^-830 to 2^+900: with the default gap-continuation
    quality a read of <= 159 bases cannot get anywhere near the real floor); all within 1e-12 of the scalar FP64 oracle."""
    rng = np.random.default_rng(91)
    regs = [wl._one_region(rng, 40, 6, 300, 420, 100, 151, junk=0.6), wl._one_region(rng, 30, 5, 80, 120, 10, 100, junk=0.7),
            wl._one_region(rng, 12, 3, 500, 700, 140, 250, junk=0.8)]                     # some reads over 159 bases
    acgt = np.frombuffer(b"ACGT", np.uint8)
    haps = [np.full(400, ord("C"), np.uint8), np.full(350, ord("G"), np.uint8), acgt[rng.integers(0, 4, 380)]]
    reads = [dict(bases=np.full(n, ord("A"), np.uint8), qual=np.full(n, 40, np.uint8), mapq=60) for n in (150, 151, 120, 159, 90)]
    regs.append((reads, haps))                                                            # every base a mismatch
    batch = wl.RegionBatch.from_regions(regs)
    eng = PairHmmEngine(fp64_mode=FP64_ALWAYS)
    try:
        fast = eng.run_flat(batch, raw=True)
        monkeypatch.setenv("PHMM_B200_FP64_GENERAL", "1")
        general = eng.run_flat(batch, raw=True)
        monkeypatch.delenv("PHMM_B200_FP64_GENERAL")
        monkeypatch.setenv("PHMM_B200_FP64_TRIP_HI", hex((1023 + 900) << 20))
        tripped = eng.run_flat(batch, raw=True)
        monkeypatch.delenv("PHMM_B200_FP64_TRIP_HI")
    finally:
        eng.close()
    took = fast.raw_f32 < np.float32(1e-28)
    assert took.sum() == fast.n_fp64 == general.n_fp64 == tripped.n_fp64 > 200
    for other in (general, tripped):
        assert np.array_equal(fast.raw_f64[took].view(np.uint64), other.raw_f64[took].view(np.uint64))
        assert np.array_equal(fast.likelihoods.view(np.uint64), other.likelihoods.view(np.uint64))
    assert np.isnan(fast.raw_f64[~took]).all()
    assert (fast.raw_f64[took] < 2.0 ** 900).sum() > 150, "most pairs must fall under the raised floor"
    for g in range(batch.n_regions):
        rd, hp = batch.region(g)
        _, _, o32, o64, n64 = orc.region(rd, hp)
        t = o32 < np.float32(1e-28)
        a, b = fast.raw64(g)[t], o64[t]
        nz = b != 0
        assert np.all(np.abs(a[nz] - b[nz]) <= 1e-12 * np.abs(b[nz])) and np.all(a[~nz] == 0)
