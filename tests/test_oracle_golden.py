"""CPU: the scalar restatement (oracle/phmm_oracle.c) against the golden vectors produced by the compiled
reference (tests/golden/make_golden.py).  This is what pins the oracle where /root/reference is absent."""
import numpy as np

import _golden
from oracle import bindings as ob


def test_pairs_bitwise(orc):
    gp = _golden.pairs()
    assert len(gp) >= 100
    for p in gp:
        args = (p["hap"], p["rs"], p["q"], p["ins"], p["dele"], p["gcp"])
        a = orc.pair_f32(*args, ob.FMA_AVX512).view(np.uint32)
        c = orc.pair_f32(*args, ob.FMA_NONE).view(np.uint32)
        assert a == p["ref_avx512"], "contracted variant differs from core/avx512_s.cpp"
        assert a == p["ref_1xn"], "contracted variant differs from rovaca/avx_512_float.cc"
        assert c == p["ref_avx2"], "un-contracted variant differs from core/avx2_s.cpp"


def test_regions_byte_identical(orc):
    batch, meta, mats, kept = _golden.regions()
    for g in range(batch.n_regions):
        reads, haps = batch.region(g)
        m, k, raw32, raw64, n64 = orc.region(reads, haps, int(meta[g][0]), int(meta[g][1]))
        assert np.array_equal(k, kept[g])
        assert np.array_equal(m.reshape(-1).view(np.uint64), mats[g]), f"region {g}: matrix bits differ from call_pairhmm"


def test_pcr_model(orc):
    for rec in _golden.pcr():
        seq = rec["seq"]
        n = len(seq)
        for m, model in enumerate((ob.PCR_HOSTILE, ob.PCR_AGGRESSIVE, ob.PCR_CONSERVATIVE)):
            got = orc.apply_pcr_error_model(seq, np.full(n, 45, np.uint8), model)
            assert np.array_equal(got, rec["capped"][m])
        got = np.array([orc.find_tandem_repeat_units(seq, o) for o in range(n)], np.int32)
        assert np.array_equal(got, rec["repeats"])


def test_underflow_policy(orc):
    # pairhmm_internal.cpp:578-584: float subtract then widen; FP64 branch below 1e-28f
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.log10f.restype, libm.log10f.argtypes = ctypes.c_float, [ctypes.c_float]
    f = np.float32(3.5e26)
    assert orc.log10_likelihood(f) == float(np.float32(libm.log10f(float(f))) - np.float32(36.1236000061))
    d = 1e250
    assert orc.log10_likelihood(np.float32(1e-30), d) == np.log10(d) - 307.050595577260822
    assert orc.log10_likelihood(np.float32(0.0), 0.0) == -np.inf


def test_edge_cases(orc):
    hap = b"ACGTACGTAC"
    one = dict(bases=b"ACGT", qual=bytes([30] * 4), mapq=60)
    m, k, *_ = orc.region([], [hap])
    assert m.size == 0 and len(k) == 0
    m, k, *_ = orc.region([one], [])
    assert m.size == 0 and len(k) == 0
    # N matches anything, other letters match nothing (one-hot rule of the default path)
    q = bytes([30] * 4)
    g = bytes([45] * 4)
    c = bytes([10] * 4)
    base = orc.pair_f32(b"ACGT", b"ACGT", q, g, g, c)
    assert orc.pair_f32(b"ACGT", b"ANGT", q, g, g, c) >= base
    assert orc.pair_f32(b"ANGT", b"ACGT", q, g, g, c) >= base
    assert orc.pair_f32(b"ANGT", b"ACGT", q, g, g, c) > orc.pair_f32(b"AXGT", b"ACGT", q, g, g, c)
    assert orc.pair_f32(b"ACGT", b"AXGT", q, g, g, c) < base
    # qualities are masked & 127
    q2 = bytes([30 | 128] * 4)
    assert orc.pair_f32(b"ACGT", b"ACGT", q2, g, g, c) == base


def test_best_alleles_and_marginalisation(orc):
    """SURVEY.md §8f item 4 against golden vectors from the reference's own AlleleLikelihoods class (make_golden.py --alleles):
    the oracle's restatements AND the product's host implementations (phmm_best_haplotypes through the epilogue is covered on
    the GPU; phmm_marginalize here)."""
    from rovaca_b200 import pairhmm as ph
    cases = _golden.alleles()
    assert len(cases) >= 40
    ties = 0
    for c in cases:
        m, n = c["matrix"], c["n_evidence"]
        best, second, bl, sl, conf = orc.best_alleles(m, n, c["priorities"])
        assert np.array_equal(best, c["best"]) and np.array_equal(second, c["second"])
        assert np.array_equal(bl.view(np.uint64), c["best_lik"]) and np.array_equal(sl.view(np.uint64), c["second_lik"])
        assert np.array_equal(conf.view(np.uint64), c["confidence"])
        ties += int((best != orc.best_alleles(m, n, None)[0]).sum())
        n_new = c["marginal"].shape[0]
        assert np.array_equal(orc.marginalize(m, n, c["old_to_new"], n_new).view(np.uint64), c["marginal"])
        assert np.array_equal(np.ascontiguousarray(ph.marginalize(m, n, c["old_to_new"], n_new)).view(np.uint64), c["marginal"])
    assert ties > 0
