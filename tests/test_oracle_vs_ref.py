"""CPU, build container only: fresh random inputs through the compiled reference (oracle/_ref) and the scalar
restatement.  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

from oracle import bindings as ob
from rovaca_b200 import workloads as wl

B = np.frombuffer(b"ACGT", np.uint8)


def _pair(rng, R, H, mut, generic, orc):
    hap = B[rng.integers(0, 4, H)]
    rs = hap[:R].copy() if H >= R else B[rng.integers(0, 4, R)]
    m = rng.random(R) < mut
    rs[m] = B[rng.integers(0, 4, int(m.sum()))]
    q = rng.integers(2, 41, R).astype(np.uint8)
    if generic:
        ins, dele, gcp = (rng.integers(lo, hi, R).astype(np.uint8) for lo, hi in ((10, 46), (10, 46), (5, 30)))
    else:
        ins = orc.apply_pcr_error_model(rs, np.full(R, 45, np.uint8), ob.PCR_CONSERVATIVE)
        dele, gcp = ins, np.full(R, 10, np.uint8)
    return hap, rs, q, ins, dele, gcp


def test_tables(orc, ref):
    ph, mm = ref.tables_f32()
    assert np.array_equal(ph.view(np.uint32), orc.ph2pr_f32().view(np.uint32))
    assert np.array_equal(mm.view(np.uint32), orc.mm_tri_f32().view(np.uint32))


def test_pairs_random(orc, ref):
    rng = np.random.default_rng(7)
    for (R, H) in [(150, 300), (9, 33), (120, 260), (250, 600)]:
        for it in range(10):
            p = _pair(rng, R, H, [0.0, 0.02, 0.3][it % 3], it % 2 == 0, orc)
            assert orc.pair_f32(*p, ob.FMA_NONE).view(np.uint32) == ref.pair_f32(0, *p).view(np.uint32)
            if ref.avx512:
                assert orc.pair_f32(*p, ob.FMA_AVX512).view(np.uint32) == ref.pair_f32(1, *p).view(np.uint32)


def test_reference_fp64_is_defective(orc, ref):
    """the as-built FP64 kernels return 0 for multi-stripe reads; the scalar FP64 is the oracle."""
    rng = np.random.default_rng(8)
    p = _pair(rng, 150, 300, 0.01, False, orc)
    d = orc.pair_f64(*p)
    f = orc.pair_f32(*p)
    assert abs((np.log10(d) - 307.050595577260822) - orc.log10_likelihood(f)) < 1e-5
    assert ref.pair_f64(0, *p) == 0.0


@pytest.mark.parametrize("path", [0, 1])
def test_regions_random(orc, ref, path):
    if path == 1 and not ref.avx512:
        pytest.skip("call_new_pairhmm002 needs AVX-512")
    rng = np.random.default_rng(9 + path)
    for sp in (dict(n_reads=40, n_haps=5, hap_lo=200, hap_hi=320, read_lo=90, read_hi=151, junk=0.1),
               dict(n_reads=12, n_haps=3, hap_lo=41, hap_hi=80, read_lo=10, read_hi=41, junk=0.1)):
        reads, haps = wl._one_region(rng, **sp)
        m0, k0, *_ = orc.region(reads, haps, 18, ob.PCR_CONSERVATIVE)
        m1, k1 = ref.region(reads, haps, 18, ob.PCR_CONSERVATIVE, path)
        assert np.array_equal(k0, k1)
        assert np.array_equal(m0.view(np.uint64), m1.view(np.uint64))


def test_best_alleles_and_marginalisation_vs_the_references_own_class(orc, ref):
    """the oracle's restatements of search_best_allele (genotype/allele_likelihoods.hpp:589-657, as
    best_alleles_breaking_ties :364-383 calls it) and marginal_likelihoods_direct (:702-732) against the reference's
    AlleleLikelihoods template itself, instantiated in place by oracle/al_shim.cpp — on the matrices call_pairhmm returns
    for regions with duplicated haplotypes (ties), under several priority vectors, with alleles that map to nothing."""
    rng = np.random.default_rng(83)
    broke = 0
    for case in range(30):
        reads, haps = wl._one_region(rng, int(rng.integers(1, 60)), int(rng.integers(1, 9)), 100, 150, 30, 90, junk=0.15)
        haps = list(haps)
        for _ in range(int(rng.integers(0, 4))):
            haps.insert(int(rng.integers(0, len(haps) + 1)), haps[int(rng.integers(0, len(haps)))].copy())
        m, kept, *_ = orc.region(reads, haps)
        n = len(kept)
        pri = rng.integers(-3, 3, len(haps)).astype(np.float64)
        # random priorities; the reference's default (reference allele first, :349); all equal = no tie-break at all
        for p in (pri, np.array([1.0] + [0.0] * (len(haps) - 1)), np.zeros(len(haps))):
            want = ref.best_alleles(m, n, p, reference_index=0)
            got = orc.best_alleles(m, n, p)
            for w, g in zip(want, got):
                assert np.array_equal(np.asarray(w).view(np.uint8), np.asarray(g).view(np.uint8)), case
        plain = orc.best_alleles(m, n, None)
        for w, g in zip(got, plain):
            assert np.array_equal(np.asarray(w).view(np.uint8), np.asarray(g).view(np.uint8)), case
        broke += int((orc.best_alleles(m, n, pri)[0] != orc.best_alleles(m, n, None)[0]).sum())
        n_new = int(rng.integers(1, 5))
        mp = rng.integers(-1, n_new, len(haps)).astype(np.int32)
        assert np.array_equal(ref.marginalize(m, n, mp, n_new + 1).view(np.uint64), orc.marginalize(m, n, mp, n_new + 1).view(np.uint64))
    assert broke > 0
