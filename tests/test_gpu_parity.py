"""GPU (B200): the CUDA path, called through the C ABI, against the CPU oracle and the golden vectors.

Bars (integer/bit work unless said otherwise):
  raw FP32 forward sums ........ bit-identical to the reference's AVX-512 (or AVX2) kernels
  FP64 recompute ............... <= 1e-12 relative vs the scalar FP64 oracle (tolerance stated here; the reference's own
                                 FP64 kernels are defective and unobservable downstream)
  PCR model / quality capping .. covered through the region results (any slip changes the matrices)
  post-processed matrix ........ byte-identical incl. NaN tail, and the same set of surviving reads
"""
import threading

import numpy as np
import pytest

import _golden
from oracle import bindings as ob
from rovaca_b200 import workloads as wl
from rovaca_b200 import PairHmmEngine, PCR_NONE, PCR_CONSERVATIVE, PCR_HOSTILE, PCR_AGGRESSIVE

pytestmark = pytest.mark.gpu
B = np.frombuffer(b"ACGT", np.uint8)
F64_RTOL = 1e-12


def _bits32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _bits64(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _check_region_vs_oracle(eng, orc, reads, haps, thr=18, model=PCR_CONSERVATIVE, fma=ob.FMA_AVX512):
    m1, k1, r32, r64, n64 = eng.call_pairhmm(haps, reads, thr, model, raw=True)
    m0, k0, o32, o64, o_n64 = orc.region(reads, haps, thr, model, fma)
    assert np.array_equal(_bits32(r32), _bits32(o32)), "raw FP32 sums differ"
    assert n64 == o_n64
    took = o32 < np.float32(1e-28)
    if took.any():
        a, b = r64[took], o64[took]
        nz = b != 0
        assert np.all(np.abs(a[nz] - b[nz]) <= F64_RTOL * np.abs(b[nz])) and np.all(a[~nz] == 0)
        assert np.all(np.isnan(r64[~took]))
    assert np.array_equal(k1, k0), "different reads survive the filter"
    assert np.array_equal(_bits64(m1), _bits64(m0)), "likelihood matrix bytes differ"
    return n64


def test_golden_pairs_raw_bits(engine, engine_nofma):
    """Every golden pair as a 1x1 region with explicit ins/del/gcp and no read preparation (thr 0, mapq 255, PCR NONE)."""
    gp = _golden.pairs()
    regs = [([dict(bases=p["rs"], qual=p["q"], ins=p["ins"], gcp=p["gcp"], mapq=255, **{"del": p["dele"]})], [p["hap"]]) for p in gp]
    batch = wl.RegionBatch.from_regions(regs)
    res = engine.run_flat(batch, 0, PCR_NONE, raw=True)
    got = np.array([res.raw32(g)[0, 0] for g in range(batch.n_regions)], np.float32)
    assert np.array_equal(_bits32(got), np.array([p["ref_avx512"] for p in gp], np.uint32))
    res2 = engine_nofma.run_flat(batch, 0, PCR_NONE, raw=True)
    got2 = np.array([res2.raw32(g)[0, 0] for g in range(batch.n_regions)], np.float32)
    assert np.array_equal(_bits32(got2), np.array([p["ref_avx2"] for p in gp], np.uint32))


def test_golden_regions_byte_identical(engine):
    batch, meta, mats, kept = _golden.regions()
    for g in range(batch.n_regions):
        reads, haps = batch.region(g)
        m, k = engine.call_pairhmm(haps, reads, int(meta[g][0]), int(meta[g][1]))
        assert np.array_equal(k, kept[g]), f"region {g}"
        assert np.array_equal(_bits64(m).reshape(-1), mats[g]), f"region {g}"
    # and all of them in one flat call per (threshold, model)
    for model in sorted(set(int(x) for x in meta[:, 1])):
        idx = [g for g in range(batch.n_regions) if int(meta[g][1]) == model]
        sub = batch.subset(idx)
        res = engine.run_flat(sub, 18, model)
        for j, g in enumerate(idx):
            assert np.array_equal(res.kept(j), kept[g])
            assert np.array_equal(_bits64(res.matrix(j)).reshape(-1), mats[g])


@pytest.mark.parametrize("spec", [
    dict(n_reads=40, n_haps=6, hap_lo=300, hap_hi=300, read_lo=150, read_hi=150, junk=0.05),
    dict(n_reads=37, n_haps=9, hap_lo=250, hap_hi=450, read_lo=100, read_hi=151, junk=0.1),       # odd haplotype count
    dict(n_reads=25, n_haps=3, hap_lo=41, hap_hi=120, read_lo=1, read_hi=41, junk=0.1),           # post-trim short shapes
    dict(n_reads=30, n_haps=16, hap_lo=1000, hap_hi=1000, read_lo=250, read_hi=250, junk=0.1),    # cfg4 shape
    dict(n_reads=12, n_haps=1, hap_lo=200, hap_hi=200, read_lo=60, read_hi=151, junk=0.0),        # single haplotype
    dict(n_reads=20, n_haps=4, hap_lo=500, hap_hi=700, read_lo=300, read_hi=420, junk=0.1),       # whole-warp bins
    dict(n_reads=6, n_haps=3, hap_lo=800, hap_hi=900, read_lo=600, read_hi=700, junk=0.2),        # above the envelope: per-pair kernel
])
@pytest.mark.parametrize("model", [PCR_CONSERVATIVE, PCR_NONE, PCR_HOSTILE, PCR_AGGRESSIVE])
def test_random_regions_vs_oracle(engine, orc, spec, model):
    rng = np.random.default_rng(hash((spec["n_reads"], spec["n_haps"], model)) % (2**32))
    reads, haps = wl._one_region(rng, **spec)
    _check_region_vs_oracle(engine, orc, reads, haps, 18, model)


def test_uncontracted_variant(engine_nofma, orc):
    rng = np.random.default_rng(5)
    reads, haps = wl._one_region(rng, 40, 7, 200, 330, 90, 151, junk=0.05)
    _check_region_vs_oracle(engine_nofma, orc, reads, haps, 18, PCR_CONSERVATIVE, fma=ob.FMA_NONE)


def test_explicit_gap_qualities(engine, orc):
    """Callers may hand in per-base ins/del/gcp (the generic path of the kernel, not just the reference's constants)."""
    rng = np.random.default_rng(6)
    reads, haps = wl._one_region(rng, 24, 5, 260, 300, 120, 151, junk=0.05)
    for r in reads:
        n = len(r["bases"])
        r["ins"] = rng.integers(10, 60, n).astype(np.uint8)
        r["del"] = rng.integers(10, 46, n).astype(np.uint8)
        r["gcp"] = rng.integers(3, 40, n).astype(np.uint8)
    _check_region_vs_oracle(engine, orc, reads, haps, 18, PCR_CONSERVATIVE)
    _check_region_vs_oracle(engine, orc, reads, haps, 10, PCR_NONE)


def test_edge_cases(engine, orc):
    q = lambda n, v=30: np.full(n, v, np.uint8)
    # empty inputs -> empty matrix (pairhmm_internal.cpp:508-510)
    m, k = engine.call_pairhmm([b"ACGT"], [])
    assert m.size == 0 and len(k) == 0
    m, k = engine.call_pairhmm([], [dict(bases=b"ACGT", qual=q(4), mapq=60)])
    assert m.size == 0 and len(k) == 0
    # lengths on the bin boundaries of the wavefront kernel and beyond it
    rng = np.random.default_rng(11)
    hap = B[rng.integers(0, 4, 420)]
    reads = []
    for n in (1, 2, 7, 8, 15, 16, 31, 32, 63, 64, 151, 152, 159, 160, 191, 192, 255, 256, 303, 304, 319, 320, 383, 384, 400):
        o = int(rng.integers(0, 420 - n + 1))
        reads.append(dict(bases=hap[o:o + n].copy(), qual=q(n, 35), mapq=60))
    _check_region_vs_oracle(engine, orc, reads, [hap, hap[:100].copy(), hap[5:6].copy()])
    # N on either side, bytes outside ACGTN, lower case (match nothing: one-hot rule), quals with the top bit set
    reads = [dict(bases=b"ACGTNNACGTACGTTTGACA", qual=q(20), mapq=60), dict(bases=b"acgtACGTXXACGTACGTAC", qual=q(20, 200), mapq=255),
             dict(bases=b"ACGTACGTACGTACGTACGT", qual=q(20, 1), mapq=0)]
    haps = [b"TTACGTNNACGTACGTTTGACATT", b"NNNNNNNNNNNNNNNNNNNNNNNN", b"TTACGTACACGTACGTTTGACATTGG"]
    _check_region_vs_oracle(engine, orc, reads, haps)
    _check_region_vs_oracle(engine, orc, reads, haps, thr=0, model=PCR_NONE)
    # everything filtered: junk reads only -> n_kept == 0, matrix all NaN
    reads = [dict(bases=B[rng.integers(0, 4, 120)], qual=q(120, 38), mapq=60) for _ in range(5)]
    haps = [B[rng.integers(0, 4, 300)] for _ in range(3)]
    n64 = _check_region_vs_oracle(engine, orc, reads, haps)
    m, k = engine.call_pairhmm(haps, reads)
    assert len(k) == 0 and np.all(np.isnan(m)) and n64 >= 0
    # homopolymers / short tandem repeats drive the PCR model to its caps
    reads = [dict(bases=b"A" * 150, qual=q(150), mapq=60), dict(bases=b"AC" * 75, qual=q(150), mapq=60),
             dict(bases=b"ACGTTGCA" * 18 + b"ACGTTG", qual=q(150), mapq=60), dict(bases=b"G", qual=q(1), mapq=60)]
    haps = [b"A" * 300, b"AC" * 150, (b"ACGTTGCA" * 40)[:300]]
    for model in (PCR_CONSERVATIVE, PCR_HOSTILE, PCR_AGGRESSIVE):
        _check_region_vs_oracle(engine, orc, reads, haps, 18, model)


def _bin_boundary_reads(rng, hap, caps, explicit_gaps):
    reads = []
    for c in caps:
        for n in (c, c + 1):
            if n > len(hap):
                continue
            o = int(rng.integers(0, len(hap) - n + 1))
            rd = hap[o:o + n].copy()
            rd[rng.integers(0, n, max(1, n // 40))] = B[rng.integers(0, 4, max(1, n // 40))]
            r = dict(bases=rd, qual=rng.integers(8, 41, n).astype(np.uint8), mapq=int(rng.integers(20, 61)))
            if explicit_gaps:
                r["ins"] = rng.integers(20, 50, n).astype(np.uint8)
                r["del"] = rng.integers(20, 50, n).astype(np.uint8)
                r["gcp"] = rng.integers(5, 30, n).astype(np.uint8)
            reads.append(r)
    return reads


def test_every_length_bin_and_partial_groups(engine, engine_nofma, orc):
    """One read on each side of every read-length bin of the wavefront kernel, for each of its three variants (reference
    gap qualities: 8/16/32 lanes x up to 19 rows; generic per-base gap qualities; un-contracted arithmetic), and read
    counts that leave sub-warp groups and whole warps of a CTA without a read."""
    rng = np.random.default_rng(17)
    hap = B[rng.integers(0, 4, 640)]
    haps = [hap, hap[3:590].copy(), hap[::-1].copy()]
    caps_ref = [8 * r - 1 for r in range(1, 20)] + [16 * r - 1 for r in (10, 12, 14, 16, 18, 19)] + [32 * r - 1 for r in (10, 12, 14, 16, 18, 19)]
    caps_gen = [16 * r - 1 for r in range(1, 13)] + [32 * r - 1 for r in range(7, 13)]
    _check_region_vs_oracle(engine, orc, _bin_boundary_reads(rng, hap, caps_ref, False), haps)
    _check_region_vs_oracle(engine, orc, _bin_boundary_reads(rng, hap, caps_gen, True), haps)
    _check_region_vs_oracle(engine_nofma, orc, _bin_boundary_reads(rng, hap, caps_gen, False), haps, fma=ob.FMA_NONE)
    for count in (1, 2, 3, 5, 7, 9, 15, 17, 33):
        sub = [dict(bases=B[rng.integers(0, 4, 150)], qual=rng.integers(8, 41, 150).astype(np.uint8), mapq=60) for _ in range(count)]
        for r in sub[: max(1, count // 2)]:           # half of them real reads of the haplotype, the rest junk
            o = int(rng.integers(0, 250))
            r["bases"] = hap[o:o + 150].copy()
        _check_region_vs_oracle(engine, orc, sub, haps)
        for r in sub:
            r["gcp"] = np.full(150, 12, np.uint8)      # explicit gcp: the generic variant, 16 lanes per read
        _check_region_vs_oracle(engine, orc, sub, haps)


def test_pcr_model_tandem_repeats(engine, orc):
    """The PCR-indel model's tandem-repeat scan (pairhmm_internal.cpp:65-159) on reads built from repeats of every unit
    length 1..9 and count 1..25 between random flanks, at lengths on both sides of the prep kernel's bit-parallel limit
    (1024 bases); any slip in a repeat length changes an insertion quality and with it the raw sums."""
    rng = np.random.default_rng(23)
    reads = []
    for unit_len in range(1, 10):
        for _ in range(3):
            unit = B[rng.integers(0, 4, unit_len)]
            parts = []
            while sum(len(x) for x in parts) < 140:
                parts.append(B[rng.integers(0, 4, int(rng.integers(0, 7)))])
                parts.append(np.tile(unit, int(rng.integers(1, 26))))
            rs = np.concatenate(parts)[: int(rng.integers(100, 152))]
            reads.append(dict(bases=rs, qual=rng.integers(10, 41, len(rs)).astype(np.uint8), mapq=60))
    hap = np.concatenate([r["bases"] for r in reads[:3]])[:400]
    for model in (PCR_CONSERVATIVE, PCR_HOSTILE, PCR_AGGRESSIVE):
        _check_region_vs_oracle(engine, orc, reads, [hap, hap[::-1].copy()], 18, model)
    long_reads = []
    for n in (1023, 1024, 1025, 1100):
        unit = B[rng.integers(0, 4, 3)]
        rs = np.concatenate([B[rng.integers(0, 4, 200)], np.tile(unit, 30), B[rng.integers(0, 4, 900)]])[:n]
        rs[-40:] = rs[-41]                               # homopolymer running into the end of the read
        long_reads.append(dict(bases=rs, qual=rng.integers(10, 41, n).astype(np.uint8), mapq=60))
    _check_region_vs_oracle(engine, orc, long_reads, [long_reads[0]["bases"][100:700].copy()], 18, PCR_CONSERVATIVE)


@pytest.mark.parametrize("lanes", [16, 32])
def test_latency_lane_families(orc, lanes):
    """The reference-gap kernels with 16 and 32 lanes per read (what a small batch on an idle device gets): every bin boundary,
    odd read counts, a mixed region, and the default engine on a one-region batch (which picks them by itself)."""
    eng = PairHmmEngine(lanes_per_read=lanes)
    try:
        rng = np.random.default_rng(100 + lanes)
        hap = B[rng.integers(0, 4, 640)]
        haps = [hap, hap[3:590].copy(), hap[::-1].copy()]
        caps = [lanes * r - 1 for r in list(range(1, 10)) + [10, 12, 14, 16, 18, 19]]
        _check_region_vs_oracle(eng, orc, _bin_boundary_reads(rng, hap, caps, False), haps)
        for count in (1, 2, 3, 5):
            sub = [dict(bases=hap[o:o + 150].copy(), qual=rng.integers(8, 41, 150).astype(np.uint8), mapq=60) for o in rng.integers(0, 400, count)]
            _check_region_vs_oracle(eng, orc, sub, haps)
        reads, hs = wl._one_region(rng, 37, 9, 250, 450, 30, 151, junk=0.1)
        _check_region_vs_oracle(eng, orc, reads, hs)
    finally:
        eng.close()


def test_paired_reads_on_the_odd_haplotype(orc):
    """Throughput layout (8 lanes pinned, as a big batch gets): a region with an odd number of haplotypes has its shortest one
    swept by pairs of reads — one read in each half of every lane, row parameters and prior tables per half.  Same bits as the
    oracle for 1 / 3 / 5 / 9 haplotypes, full and odd read counts (a pair of one), reads on either side of every lane-count
    boundary of the 12-row layout, short post-trim shapes, junk reads that underflow in either half (FP64 list), and a region
    whose longest read does not fit 32 lanes at 12 rows (which keeps the two-stream chunks)."""
    eng = PairHmmEngine(lanes_per_read=8)
    try:
        rng = np.random.default_rng(77)
        for n_haps in (1, 3, 5, 9):
            reads, haps = wl._one_region(rng, 64, n_haps, 280, 330, 150, 150)
            _check_region_vs_oracle(eng, orc, reads, haps)
            _check_region_vs_oracle(eng, orc, reads[:63], haps)
            reads, haps = wl._one_region(rng, 41, n_haps, 250, 450, 30, 151, junk=0.15)
            _check_region_vs_oracle(eng, orc, reads, haps)
            reads, haps = wl._one_region(rng, 37, n_haps, 41, 120, 10, 100)                  # post-trim shapes
            _check_region_vs_oracle(eng, orc, reads, haps, 18, PCR_HOSTILE)
        hap = B[rng.integers(0, 4, 500)]
        haps = [hap, hap[5:480].copy(), hap[::-1][:300].copy()]
        edge = [11, 11, 12, 12, 23, 23, 24, 24, 35, 36, 36, 371, 371, 372, 372, 383, 383, 1, 1, 2]
        reads = [dict(bases=hap[o:o + n].copy(), qual=rng.integers(5, 41, n).astype(np.uint8), mapq=int(rng.integers(20, 61)))
                 for n, o in zip(edge, rng.integers(0, 100, len(edge)))]
        _check_region_vs_oracle(eng, orc, reads, haps)
        reads.append(dict(bases=hap[:400].copy(), qual=np.full(400, 30, np.uint8), mapq=60))   # 34 lanes at 12 rows: no pairs here
        _check_region_vs_oracle(eng, orc, reads, haps)
        # a flat batch of many such regions equals the per-region calls (CTAs of several warps, several regions per launch)
        regs = [wl._one_region(rng, int(rng.integers(1, 60)), int(rng.integers(1, 12)), 60, 400, 10, 151, junk=0.05) for _ in range(48)]
        flat = eng.run_flat(wl.RegionBatch.from_regions(regs))
        for g, (rd, hp) in enumerate(regs):
            m0, k0, *_ = orc.region(rd, hp)
            assert np.array_equal(k0, flat.kept(g)) and np.array_equal(_bits64(m0), _bits64(flat.matrix(g))), g
    finally:
        eng.close()


def test_random_awkward_batches_vs_oracle(engine, orc):
    """Seeded random flat batches of awkward shapes — odd and even haplotype counts, haplotypes of 1 to 2399 columns, reads of 1 to 400
    bases sorted or of one length, junk reads, N bases — through the default engine and one pinned to the throughput layout (paired
    reads, fixed lanes, mixed rows per lane): every region byte-identical to the oracle."""
    rng = np.random.default_rng(4242)
    pinned = PairHmmEngine(lanes_per_read=8)
    try:
        for trial in range(6):
            regs = []
            for _ in range(int(rng.integers(6, 14))):
                n_reads, n_haps = int(rng.integers(1, 50)), int(rng.integers(1, 10))
                H = int(rng.choice([1, 7, 40, 80, 120, 300, 450, 1000, 2399], p=[.03, .05, .1, .2, .1, .3, .1, .08, .04]))
                ref = B[rng.integers(0, 4, H)]
                haps = []
                for j in range(n_haps):
                    h = ref.copy()
                    if H > 4:
                        h[rng.integers(0, H, 2)] = B[rng.integers(0, 4, 2)]
                        if rng.random() < 0.3:
                            h = h[: H - int(rng.integers(1, min(8, H - 1)))]
                        if rng.random() < 0.05:
                            h[int(rng.integers(0, len(h)))] = ord("N")
                    haps.append(h)
                lens = [int(rng.choice([100, 150, 151, 250]))] * n_reads if rng.random() < 0.3 else \
                    sorted(int(rng.choice([1, 2, 11, 12, 23, 24, 35, 36, 60, 100, 150, 151, 250, 383, 384, 400])) for _ in range(n_reads))
                reads = []
                for n in lens:
                    if rng.random() < 0.1 or n > H:
                        rs = B[rng.integers(0, 4, n)]
                    else:
                        o = int(rng.integers(0, H - n + 1))
                        rs = ref[o:o + n].copy()
                    reads.append(dict(bases=rs, qual=rng.integers(2, 41, n).astype(np.uint8), mapq=int(rng.integers(0, 61))))
                regs.append((reads, haps))
            batch = wl.RegionBatch.from_regions(regs)
            for eng in (engine, pinned):
                flat = eng.run_flat(batch)
                for g, (rd, hp) in enumerate(regs):
                    m0, k0, *_ = orc.region(rd, hp)
                    assert np.array_equal(k0, flat.kept(g)), (trial, g)
                    assert np.array_equal(_bits64(m0), _bits64(flat.matrix(g))), (trial, g)
    finally:
        pinned.close()


def test_fp64_recompute_exercised(engine, orc):
    """cfg3-style: junk reads underflow FP32 (< 1e-28f) and are recomputed in FP64 on the device."""
    rng = np.random.default_rng(12)
    reads, haps = wl._one_region(rng, 60, 8, 300, 420, 120, 151, junk=0.4)
    n64 = _check_region_vs_oracle(engine, orc, reads, haps)
    assert n64 > 50


def test_matrix_does_not_depend_on_the_fp64_recompute(engine, orc):
    """By default the FP64 recompute runs only when its values can be seen (raw_f64 asked for): an underflowing pair is
    below -64, so its read is dropped or its entry capped at best - 4.5 (consequence of a11-a14).  The
    matrices and the surviving reads must be the ones the oracle gets WITH the recompute, for reads on both sides of the
    51-bp filter limit, all-junk regions, flat batches and device-resident batches; PHMM_FP64_ALWAYS gives the same bytes."""
    from rovaca_b200.pairhmm import FP64_ALWAYS
    always = PairHmmEngine(fp64_mode=FP64_ALWAYS)
    try:
        rng = np.random.default_rng(31)
        specs = [dict(n_reads=60, n_haps=8, hap_lo=300, hap_hi=420, read_lo=120, read_hi=151, junk=0.4),
                 dict(n_reads=40, n_haps=5, hap_lo=200, hap_hi=260, read_lo=20, read_hi=60, junk=0.5),     # limit -4 below 51 bp
                 dict(n_reads=9, n_haps=3, hap_lo=300, hap_hi=300, read_lo=150, read_hi=150, junk=1.0),    # nothing survives
                 dict(n_reads=30, n_haps=4, hap_lo=500, hap_hi=700, read_lo=300, read_hi=420, junk=0.3)]
        seen = 0
        for spec in specs:
            reads, haps = wl._one_region(rng, **spec)
            m0, k0, o32, _, o_n64 = orc.region(reads, haps)
            seen += o_n64
            for eng in (engine, always):
                m, k = eng.call_pairhmm(haps, reads)
                assert np.array_equal(k, k0)
                assert np.array_equal(_bits64(m), _bits64(m0))
        assert seen > 200
        batch = wl.cfg3(2, seed=33, n_reads=120, n_haps=12)
        before = engine.stats()["fp64_pairs"]
        lazy, full, raw = engine.run_flat(batch), always.run_flat(batch), engine.run_flat(batch, raw=True)
        assert lazy.n_fp64 == full.n_fp64 == raw.n_fp64 > 50                        # counted either way
        assert engine.stats()["fp64_pairs"] - before == 2 * raw.n_fp64
        assert np.array_equal(_bits64(lazy.likelihoods), _bits64(full.likelihoods))
        assert np.array_equal(_bits64(lazy.likelihoods), _bits64(raw.likelihoods))
        assert np.array_equal(lazy.n_kept, full.n_kept) and np.array_equal(lazy.kept_idx, full.kept_idx)
        resident = engine.upload(batch)
        resident.run()
        d0 = resident.download()                  # no raw sums asked for: nothing recomputed
        d1 = resident.download(raw=True)          # asked for after the run: recomputed now
        resident.free()
        assert np.array_equal(_bits64(d0.likelihoods), _bits64(raw.likelihoods))
        assert np.array_equal(_bits64(d1.likelihoods), _bits64(raw.likelihoods))
        assert np.array_equal(_bits64(d1.raw_f64), _bits64(raw.raw_f64))
        took = raw.raw_f32 < np.float32(1e-28)
        assert took.sum() == raw.n_fp64 and not np.isnan(raw.raw_f64[took]).any() and np.isnan(raw.raw_f64[~took]).all()
    finally:
        always.close()


def test_device_log10_is_the_hosts_libm(engine):
    """The device epilogue's log10f (glibc's algorithm restated, phmm_kernels.cu log10_of_sum) against the host's libm: every
    float of [0.5, 2) — all the formula ever hands to logf — and four million floats spread over everything above 1e-28."""
    from rovaca_b200 import pairhmm as ph
    _, on_device = engine.device_log10(np.ones(1, np.float32))
    assert on_device, "this host's libm is not glibc's algorithm: the engine must have kept its epilogue on the host"
    for lo in range(0x3F000000, 0x40000000, 1 << 22):
        x = np.arange(lo, lo + (1 << 22), dtype=np.uint32).view(np.float32)
        got, _ = engine.device_log10(x)
        assert np.array_equal(got.view(np.uint64), ph.host_log10(x, 0).view(np.uint64)), hex(lo)
    rng = np.random.default_rng(5)
    x = rng.integers(0x0D000000, 0x7F800000, 1 << 22, dtype=np.uint32).view(np.float32)
    got, _ = engine.device_log10(x)
    want = ph.host_log10(x, 0)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert np.isneginf(got[x < np.float32(1e-28)]).all()


def test_device_epilogue_equals_host_epilogue_and_oracle(engine, engine_host_epilogue, orc):
    """Matrices, surviving reads, FP64-branch counts and best / second-best haplotypes from the device
    epilogue, from the host epilogue and from the oracle (orc.region + its restatement of search_best_allele), on regions with
    duplicate haplotypes (ties for the priorities to break), junk reads, reads under 51 bp and a single-haplotype region."""
    rng = np.random.default_rng(47)
    regs, pris = [], []
    for g in range(24):
        reads, haps = wl._one_region(rng, int(rng.integers(1, 140)), int(rng.integers(2, 9)), 150, 260, 30, 151, junk=0.15)
        haps = list(haps)
        for _ in range(int(rng.integers(0, 4))):
            haps.insert(int(rng.integers(0, len(haps) + 1)), haps[int(rng.integers(0, len(haps)))].copy())
        if g == 3:
            haps = haps[:1]
        regs.append((reads, haps))
        pris.append(rng.integers(-3, 3, len(haps)).astype(np.float64))
    regs.insert(5, ([], [b"ACGT"]))                            # an empty region in the middle
    pris.insert(5, np.zeros(1))
    batch = wl.RegionBatch.from_regions(regs)
    batch.hap_priority = np.concatenate(pris)
    dev = engine.run_flat(batch, best=True, raw=True)
    host = engine_host_epilogue.run_flat(batch, best=True, raw=True)
    assert np.array_equal(dev.n_kept, host.n_kept) and dev.n_fp64 == host.n_fp64
    assert np.array_equal(_bits64(dev.likelihoods), _bits64(host.likelihoods))
    assert np.array_equal(_bits32(dev.raw_f32), _bits32(host.raw_f32))
    assert np.array_equal(_bits64(dev.raw_f64), _bits64(host.raw_f64))
    ties = 0
    for g, (reads, haps) in enumerate(regs):
        assert np.array_equal(dev.kept(g), host.kept(g))
        for a, b in zip(dev.best(g), host.best(g)):
            assert np.array_equal(a, b), g
        if not reads:
            continue
        m0, k0, *_ = orc.region(reads, haps)
        assert np.array_equal(dev.kept(g), k0) and np.array_equal(_bits64(dev.matrix(g)), _bits64(m0))
        b0, s0, *_ = orc.best_alleles(m0, len(k0), pris[g])
        ties += int((b0 != orc.best_alleles(m0, len(k0), None)[0]).sum())
        assert np.array_equal(dev.best(g)[0], b0) and np.array_equal(dev.best(g)[1], s0), g
        # the per-region call, with and without priorities
        m, k, b, s2 = engine.call_pairhmm(haps, reads, priorities=pris[g])
        assert np.array_equal(k, k0) and np.array_equal(_bits64(m), _bits64(m0)) and np.array_equal(b, b0) and np.array_equal(s2, s0)
    assert ties > 0
    # without best haplotypes asked for, and a resident batch that learns about them only at download time
    plain = engine.run_flat(batch)
    assert np.array_equal(_bits64(plain.likelihoods), _bits64(host.likelihoods)) and np.array_equal(plain.kept_idx, host.kept_idx)
    resident = engine.upload(batch)
    resident.run()
    down = resident.download(best=True)
    resident.free()
    assert np.array_equal(_bits64(down.likelihoods), _bits64(host.likelihoods))
    assert np.array_equal(down.best_hap, host.best_hap) and np.array_equal(down.second_hap, host.second_hap)


def test_default_engine_picks_the_epilogue_by_batch_size(engine):
    """PHMM_EPILOGUE_AUTO: the device epilogue for batches of 32 768 pairs or more, the host's for smaller ones; same bytes."""
    auto = PairHmmEngine()
    try:
        small, big = wl.cfg2(8, seed=61), wl.cfg2(1024, seed=62)          # 4 096 pairs; 524 288 pairs in sub-batches of >= 65 000
        n0 = auto.stats()["device_epilogue_batches"]
        a = auto.run_flat(small)
        n1 = auto.stats()["device_epilogue_batches"]
        b = auto.run_flat(big)
        n2 = auto.stats()["device_epilogue_batches"]
        assert n1 == n0 and n2 > n1
        for got, batch in ((a, small), (b, big)):
            want = engine.run_flat(batch)
            assert np.array_equal(_bits64(got.likelihoods), _bits64(want.likelihoods))
            assert np.array_equal(got.kept_idx, want.kept_idx) and np.array_equal(got.n_kept, want.n_kept)
    finally:
        auto.close()


def test_forcing_libm_keeps_the_epilogue_off_the_device(monkeypatch):
    """PHMM_B200_LIBM_LOG10F pins the host's own log10f: nothing restated may run, so PHMM_EPILOGUE_DEVICE must refuse to
    start and a default engine must keep every batch's epilogue on the host."""
    from rovaca_b200 import PhmmError
    from rovaca_b200.pairhmm import EPILOGUE_DEVICE
    monkeypatch.setenv("PHMM_B200_LIBM_LOG10F", "1")
    with pytest.raises(PhmmError, match="log10f"):
        PairHmmEngine(epilogue=EPILOGUE_DEVICE)
    eng = PairHmmEngine()
    try:
        eng.run_flat(wl.cfg2(256, seed=63))
        assert eng.stats()["device_epilogue_batches"] == 0
    finally:
        eng.close()


def test_epilogue_with_many_reads_and_haplotypes(engine, engine_host_epilogue, orc):
    """More reads than the epilogue kernel's CTA has threads (its block scan runs over several tiles, with dropped reads in
    every tile) and more haplotypes than a warp has lanes; priorities that flip many of the ties."""
    rng = np.random.default_rng(71)
    reads, haps = wl._one_region(rng, 300, 40, 90, 110, 30, 70, junk=0.3)
    haps = list(haps) + [h.copy() for h in haps[:8]]
    pri = rng.integers(-2, 3, len(haps)).astype(np.float64)
    m0, k0, *_ = orc.region(reads, haps)
    b0, s0, *_ = orc.best_alleles(m0, len(k0), pri)
    assert 50 < len(k0) < 290
    for eng in (engine, engine_host_epilogue):
        m, k, b, s2 = eng.call_pairhmm(haps, reads, priorities=pri)
        assert np.array_equal(k, k0) and np.array_equal(_bits64(m), _bits64(m0))
        assert np.array_equal(b, b0) and np.array_equal(s2, s0)


def test_long_haplotypes_multi_chunk(engine, orc):
    """More haplotype columns than one shared-memory chunk holds: the region is swept in several chunks."""
    rng = np.random.default_rng(13)
    reads, haps = wl._one_region(rng, 6, 40, 900, 1000, 100, 151, junk=0.0)
    _check_region_vs_oracle(engine, orc, reads, haps)


def test_flat_equals_per_region_and_resident(engine, orc):
    batch = wl.cfg5(40, seed=21)
    flat = engine.run_flat(batch, raw=True)
    resident = engine.upload(batch)
    resident.run()
    down = resident.download(raw=True)
    resident.free()
    assert np.array_equal(_bits32(flat.raw_f32), _bits32(down.raw_f32))
    assert np.array_equal(flat.n_kept, down.n_kept)
    for g in range(batch.n_regions):
        reads, haps = batch.region(g)
        m, k = engine.call_pairhmm(haps, reads)
        assert np.array_equal(k, flat.kept(g)) and np.array_equal(k, down.kept(g))
        assert np.array_equal(_bits64(m), _bits64(flat.matrix(g)))
        assert np.array_equal(_bits64(m), _bits64(down.matrix(g)))
    for g in (0, 7, 19, 33):
        reads, haps = batch.region(g)
        m0, k0, *_ = orc.region(reads, haps)
        assert np.array_equal(k0, flat.kept(g)) and np.array_equal(_bits64(m0), _bits64(flat.matrix(g)))


def test_flat_batches_with_empty_regions(engine, orc):
    """Empty regions (no reads or no haplotypes) inside a flat batch, a batch of nothing but those, and an empty batch."""
    rng = np.random.default_rng(29)
    r1, h1 = wl._one_region(rng, 9, 3, 120, 160, 40, 100)
    r2, h2 = wl._one_region(rng, 5, 2, 300, 300, 150, 150)
    regs = [([], h1), (r1, h1), (r2, []), ([], []), (r2, h2)]
    batch = wl.RegionBatch.from_regions(regs)
    res = engine.run_flat(batch, raw=True)
    for g, (reads, haps) in enumerate(regs):
        if reads and haps:
            m0, k0, *_ = orc.region(reads, haps)
            assert np.array_equal(k0, res.kept(g)) and np.array_equal(_bits64(m0), _bits64(res.matrix(g)))
        else:
            assert res.n_kept[g] == 0
    only_empty = engine.run_flat(wl.RegionBatch.from_regions([([], h1), (r1, [])]))
    assert list(only_empty.n_kept[:2]) == [0, 0]
    nothing = engine.run_flat(wl.RegionBatch.from_regions([]))
    assert nothing.n_fp64 == 0


def test_pinned_flat_batch_is_copied_directly(engine):
    """PHMM_FLAT_PINNED: blobs in page-locked memory are copied to the device without the staging gather; same bytes out."""
    batch = wl.cfg5(30, seed=23)
    want = engine.run_flat(batch, raw=True)
    s0 = engine.stats()
    got = engine.run_flat(wl.cfg5(30, seed=23).pin(), raw=True)
    s1 = engine.stats()
    assert np.array_equal(_bits32(want.raw_f32), _bits32(got.raw_f32)) and np.array_equal(_bits64(want.likelihoods), _bits64(got.likelihoods))
    assert np.array_equal(want.n_kept, got.n_kept) and s1["h2d_bytes"] > s0["h2d_bytes"]


def test_concurrent_callers(engine, orc):
    """The reference calls call_pairhmm from up to --nthreads pool workers; submit/wait must be re-entrant."""
    batch = wl.cfg5(24, seed=22)
    want = {}
    for g in range(batch.n_regions):
        reads, haps = batch.region(g)
        want[g] = orc.region(reads, haps)[:2]
    errs = []

    def work(ids):
        try:
            for g in ids:
                reads, haps = batch.region(g)
                m, k = engine.call_pairhmm(haps, reads)
                assert np.array_equal(k, want[g][1]) and np.array_equal(_bits64(m), _bits64(want[g][0]))
        except Exception as e:       # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(list(range(t, batch.n_regions, 8)),)) for t in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[0]


def test_many_synchronous_callers_match_the_flat_path(engine):
    """48 threads calling call_pairhmm at once (the reference's pool workers; regions of different callers share GPU batches
    under the batch-while-busy policy): every caller gets exactly what the flat path computes for its region."""
    batch = wl.cfg5(48 * 8, seed=31)
    flat = engine.run_flat(batch)
    errs = []

    def work(ids):
        try:
            for g in ids:
                reads, haps = batch.region(g)
                m, k = engine.call_pairhmm(haps, reads)
                assert np.array_equal(k, flat.kept(g)), g
                assert np.array_equal(_bits64(m), _bits64(flat.matrix(g))), g
        except Exception as e:       # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(list(range(t, batch.n_regions, 48)),)) for t in range(48)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[0]
    s = engine.stats()
    assert s["batches"] > 0


def test_full_size_cfg2_properties(engine, orc):
    """BASELINE.json configs[1] at bench size: a sample of regions against the oracle, plus size-independent properties:
    determinism across runs, independence from batching, and column permutation when the reads are permuted."""
    batch = wl.cfg2(1024, seed=2)
    r1 = engine.run_flat(batch, raw=True)
    r2 = engine.run_flat(batch, raw=True)
    assert np.array_equal(_bits32(r1.raw_f32), _bits32(r2.raw_f32)) and np.array_equal(_bits64(r1.likelihoods), _bits64(r2.likelihoods))
    assert np.all(np.isfinite(r1.raw_f32)) and np.all(r1.raw_f32 >= 0)
    rng = np.random.default_rng(3)
    for g in rng.choice(batch.n_regions, 12, replace=False):
        reads, haps = batch.region(int(g))
        m0, k0, o32, *_ = orc.region(reads, haps)
        assert np.array_equal(_bits32(r1.raw32(int(g))), _bits32(o32))
        assert np.array_equal(k0, r1.kept(int(g))) and np.array_equal(_bits64(m0), _bits64(r1.matrix(int(g))))
    g = 5
    reads, haps = batch.region(g)
    perm = rng.permutation(len(reads))
    m, k, raw, _, _ = engine.call_pairhmm(haps[::-1], [reads[i] for i in perm], raw=True)
    assert np.array_equal(_bits32(raw), _bits32(r1.raw32(g)[perm][:, ::-1]))


def test_stats_and_kernel_info(engine):
    s = engine.stats()
    assert s["cells"] > 0 and s["kernel_launches"] > 0 and s["h2d_bytes"] > 0
    ki = engine.kernel_info(150)
    assert ki["lanes_per_read"] == 8 and ki["rows_per_lane"] == 19
    assert 0 < ki["regs"] <= 255 and ki["spill_bytes"] == 0, "the 150-bp forward kernel must not spill"
    assert engine.kernel_info(151)["rows_per_lane"] == 19
    for n in (10, 44, 100, 152, 250, 400, 607):            # every read gets the lanes it needs inside one warp
        k = engine.kernel_info(n)
        assert k["lanes_per_read"] * k["rows_per_lane"] >= n + 1 and k["lanes_per_read"] <= 32 and k["spill_bytes"] == 0, (n, k)
    assert engine.kernel_info(44)["rows_per_lane"] == 12      # the most efficient instantiation where the lanes pack well
    assert engine.kernel_info(250)["lanes_per_read"] == 16
