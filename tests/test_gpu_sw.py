"""GPU: the batched Smith-Waterman (phmm_sw_align) against the golden vectors of the reference's sw_avx.c and the scalar
oracle: score, alignment offset and CIGAR identical, for both of the reference's parameter sets, every kernel width."""
import importlib.util
import os

import numpy as np
import pytest

import _golden
from oracle import bindings as ob
from rovaca_b200 import pairhmm as ph

pytestmark = pytest.mark.gpu
B = np.frombuffer(b"ACGT", np.uint8)


def _same(got, want):
    return np.array_equal(got[0], want[0]) and got[1] == want[1] and got[2] == want[2]


def test_sw_golden(engine):
    cases = _golden.sw()
    for name, prm in (("realign", ph.SW_REALIGN), ("assemble", ph.SW_ASSEMBLE)):
        got = engine.sw_align([(c["hap"], c["read"]) for c in cases], prm, cigar_cap=512)     # unrelated sequences give long CIGARs
        for k, (c, g) in enumerate(zip(cases, got)):
            assert g[3] == 0 and _same(g, c[name]), (k, name, ob.cigar_string(g[0]), ob.cigar_string(c[name][0]), g[1:], c[name][1:])


def test_sw_random_vs_oracle_all_widths(engine, orc):
    """Reads on both sides of every columns-per-lane width of the fill kernel (64, 160, 256, 512, 1024, 2047 bases), haplotypes
    up to 2047 rows, indels, overhangs, low-complexity ties; and the lengths at which the reference gives up."""
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rng = np.random.default_rng(9)
    pairs = mg.sw_cases(rng, 300)
    for R in (1, 2, 63, 64, 65, 159, 160, 161, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047):
        H = int(min(2047, R + rng.integers(0, 300)))
        hap = B[rng.integers(0, 4, H)]
        o = int(rng.integers(0, max(H - R, 0) + 1))
        rd = hap[o:o + R].copy() if H >= R else B[rng.integers(0, 4, R)]
        m = rng.random(len(rd)) < 0.02
        rd[m] = B[rng.integers(0, 4, int(m.sum()))]
        if len(rd) > 20:
            p = int(rng.integers(5, len(rd) - 5))
            rd = np.concatenate([rd[:p], rd[p + 3:], B[rng.integers(0, 4, 3)]])
        pairs.append((hap, rd))
    for prm in (ph.SW_REALIGN, ph.SW_ASSEMBLE):
        got = engine.sw_align(pairs, prm, cigar_cap=1024)
        for k, ((hap, rd), g) in enumerate(zip(pairs, got)):
            want = orc.sw_align(hap, rd, prm)
            assert g[3] == 0 and _same(g, want), (k, len(hap), len(rd), ob.cigar_string(g[0]), ob.cigar_string(want[0]), g[1:3], want[1:])
    long_hap = B[rng.integers(0, 4, 2048)]
    got = engine.sw_align([(long_hap, long_hap[:100].copy()), (long_hap[:100].copy(), long_hap), (long_hap[:50].copy(), long_hap[:50].copy())])
    assert got[0][3] == 1 and got[0][1] == 1 and len(got[0][0]) == 0 and got[1][3] == 1       # sw_avx.c:494-499
    assert got[2][3] == 0 and ob.cigar_string(got[2][0]) == "50M"
    with pytest.raises(ph.PhmmError):
        engine.sw_align([(b"", b"ACGT")])
    tiny = engine.sw_align([(b"ACGTACGTAAAAACGT", b"ACGTACGTACGT")], cigar_cap=1)
    assert tiny[0][3] == 2                                                                       # caller buffer too small


def test_sw_realigns_reads_of_a_region(engine, orc, ref):
    """The use the reference makes of it: every read of a region against its best haplotype (assembly_based_caller_utils.cpp:
    146-170), through the reference's own sw_avx.c and through the device."""
    from rovaca_b200 import workloads as wl
    rng = np.random.default_rng(10)
    reads, haps = wl._one_region(rng, 80, 6, 280, 340, 90, 151, junk=0.05)
    # the chain the reference runs: call_pairhmm -> best_alleles_breaking_ties (reference haplotype first, then the fewest
    # CIGAR elements: assembly_based_caller_utils.cpp:150-155) -> one Smith-Waterman per read against its best haplotype
    pri = [1.0] + [float(1 - int(rng.integers(1, 6))) for _ in haps[1:]]
    m, kept, best, second = engine.call_pairhmm(haps, reads, priorities=pri)
    b0, s0, *_ = orc.best_alleles(m, len(kept), pri)
    assert np.array_equal(best, b0) and np.array_equal(second, s0)
    pairs = [(haps[int(best[c])], reads[int(kept[c])]["bases"]) for c in range(len(kept))]
    got = engine.sw_align(pairs)
    for (hap, rd), g in zip(pairs, got):
        want = ref.sw_align(hap, rd, False)
        assert _same(g, want), (ob.cigar_string(g[0]), ob.cigar_string(want[0]))


def test_sw_many_alignments_cross_the_sub_batches(engine, orc):
    """phmm_sw_align runs big calls as sub-batches of 32 768 alignments on two streams: 80 000 small alignments (five hundred
    distinct ones, shuffled, with a few the reference gives up on in between) must come back in the caller's order."""
    rng = np.random.default_rng(91)
    distinct = []
    for _ in range(500):
        H = int(rng.integers(30, 70))
        hap = B[rng.integers(0, 4, H)]
        R = int(rng.integers(10, H + 1))
        o = int(rng.integers(0, H - R + 1))
        rd = hap[o:o + R].copy()
        m = rng.random(R) < 0.05
        rd[m] = B[rng.integers(0, 4, int(m.sum()))]
        distinct.append((hap, rd))
    want = [orc.sw_align(h, r, ph.SW_REALIGN) for h, r in distinct]
    long_hap = B[rng.integers(0, 4, 2100)]
    pick = rng.integers(0, 500, 80000)
    pairs = [distinct[int(k)] for k in pick]
    gave_up = (5, 32767, 32768, 65536, 79999)
    for k in gave_up:
        pairs[k] = (long_hap, long_hap[:40].copy())
    got = engine.sw_align(pairs, ph.SW_REALIGN, cigar_cap=32)
    for k, g in enumerate(got):
        if k in gave_up:
            assert g[3] == 1 and g[1] == 1 and len(g[0]) == 0
        else:
            assert g[3] == 0 and _same(g, want[int(pick[k])]), k


def test_sw_narrow_and_wide_kernels_agree(engine, orc, monkeypatch):
    """Alignments whose scores fit 16 bits take the two-alignments-per-warp kernel (packed 16-bit halves), the rest the
    32-bit one: identical CIGAR, offset and score from both, and from the oracle — pairs of unequal shapes in one warp (so
    that one alignment's rows run on after the other's end), odd counts, reads near the 16-bit kernel's bound, the
    assembler's parameter set (never narrow)."""
    rng = np.random.default_rng(123)
    acgt = np.frombuffer(b"ACGT", np.uint8)

    def seq(n):
        return acgt[rng.integers(0, 4, n)]
    pairs = []
    for _ in range(101):                                    # odd count: the last warp holds one alignment
        h = seq(int(rng.integers(30, 420)))
        n = int(rng.integers(8, 159))
        if rng.random() < 0.7 and len(h) > n:               # a read of the haplotype with substitutions and an indel
            o = int(rng.integers(0, len(h) - n + 1))
            r = h[o:o + n].copy()
            m = rng.random(n) < 0.05
            r[m] = acgt[rng.integers(0, 4, int(m.sum()))]
            if n > 20 and rng.random() < 0.5:
                p = int(rng.integers(5, n - 5))
                r = np.concatenate([r[:p], seq(int(rng.integers(1, 6))), r[p:]]) if rng.random() < 0.5 else np.concatenate([r[:p], r[p + 3:]])
        else:
            r = seq(n)
        pairs.append((h, r))
    pairs += [(seq(900), seq(790)), (seq(790), seq(900)), (seq(1200), seq(810)), (seq(60), seq(1000))]   # around the bound (798)
    for prm in (ph.SW_REALIGN, ph.SW_ASSEMBLE):
        got = engine.sw_align(pairs, prm, cigar_cap=128)
        monkeypatch.setenv("PHMM_B200_SW_WIDE", "1")
        wide = engine.sw_align(pairs, prm, cigar_cap=128)
        monkeypatch.delenv("PHMM_B200_SW_WIDE")
        for k, ((h, r), g, w) in enumerate(zip(pairs, got, wide)):
            assert np.array_equal(g[0], w[0]) and g[1:] == w[1:], (k, len(h), len(r), prm)
        for k in range(0, len(pairs), 7):
            want = orc.sw_align(pairs[k][0], pairs[k][1], prm)
            assert np.array_equal(got[k][0], want[0]) and got[k][1:3] == want[1:3], k
