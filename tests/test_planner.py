"""CPU: the engine's planner (rows per lane per region, reads packed into warps lane by lane, haplotypes chunked into balanced
column streams, CTA sizing) audited without a GPU through phmm_debug_plan: every (read, haplotype) pair is computed exactly
once, every read has the lanes it needs inside one warp, and the lanes are used as well as DESIGN.md says."""
import numpy as np
import pytest

from rovaca_b200 import workloads as wl
from rovaca_b200 import pairhmm as ph

B = np.frombuffer(b"ACGT", np.uint8)


def _sound(a):
    assert a["covered_twice"] == 0 and a["missed"] == 0, a
    assert a["wavefront_pairs"] + a["per_pair_kernel_pairs"] == a["pairs"], a
    assert a["reads_in_too_small_a_bin"] == 0 and a["shape_violations"] == 0, a


@pytest.mark.parametrize("cta_warps", [0, 1, 2, 4])
def test_cfg2_plan_is_tight(cta_warps):
    a = ph.plan_audit(wl.cfg2(48, seed=3), cta_warps=cta_warps, busy_batches=8)
    _sound(a)
    assert a["pairs"] == 48 * 64 * 8 and a["per_pair_kernel_pairs"] == 0
    assert a["slot_fill"] == 1.0                      # 64 reads: whole warps of four reads on 8 lanes each
    assert a["lane_utilisation"] > 0.96, a            # 150 rows in 152, 7 fill steps per 1200 columns


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_every_lane_policy_covers_every_pair(lanes):
    """lanes_per_read pins the latency policy: the longest read of a region spreads over about that many lanes (fewer rows
    per lane, a shorter wavefront step, more warps); 8 = the throughput search."""
    b = wl.cfg5(40, seed=12, post_trim_fraction=0.5)
    a = ph.plan_audit(b, lanes_per_read=lanes, busy_batches=4)
    _sound(a)
    wide = ph.plan_audit(b, lanes_per_read=32, busy_batches=4)
    assert a["warps"] <= wide["warps"] and a["lane_steps"] >= wide["lane_steps"] * (0.99 if lanes == 32 else 0.5)


def test_an_idle_device_gets_the_wide_layouts():
    """One region on an idle device: reads spread over 32 lanes (5 rows per lane for 150 bp) — a quarter of the rows per lane of
    the throughput layout, so a quarter of the time per wavefront step; a big batch keeps 8 lanes x 19 rows."""
    one = ph.plan_audit(wl.cfg2(1, seed=3), busy_batches=1)
    many = ph.plan_audit(wl.cfg2(256, seed=3), busy_batches=1)
    _sound(one)
    _sound(many)
    assert one["read_slots"] == 32 * one["warps"] and one["warps"] == 64 * one["chunks"]       # one read per warp (31 lanes x 5 rows)
    assert many["warps"] == 16 * many["chunks"]                                                  # four reads per warp


def test_small_batches_are_cut_finer():
    """A batch with fewer than sixteen waves of warps gets shorter chunks (its last warps would otherwise finish alone); a batch
    that fills the device many times over keeps one full-length chunk per cfg2 region."""
    few, many = ph.plan_audit(wl.cfg2(4, seed=3), busy_batches=1), ph.plan_audit(wl.cfg2(2048, seed=3), busy_batches=1)
    _sound(few)
    _sound(many)
    assert few["chunks"] == 4 * 4 and many["chunks"] == 2048          # 8 haplotypes: four chunks of one pair each
    lone = ph.plan_audit(wl.cfg2(1, seed=3), busy_batches=1)          # a lone call on an idle device: a haplotype per chunk,
    _sound(lone)                                                      # half the sweep (and half the lanes idle, which nobody needs)
    assert lone["chunks"] == 8 and lone["warps"] == 8 * 64
    mid = ph.plan_audit(wl.cfg2(256, seed=3), busy_batches=1)
    _sound(mid)
    assert 256 < mid["chunks"] <= 4 * 256


@pytest.mark.parametrize("make", [lambda: wl.cfg3(2, seed=3, n_reads=120, n_haps=11), lambda: wl.cfg4(6, seed=4, mix=True),
                                  lambda: wl.cfg5(60, seed=5, post_trim_fraction=0.5), lambda: wl.cfg5(40, seed=6, post_trim_fraction=1.0)])
@pytest.mark.parametrize("fma", [ph.FMA_AVX512, ph.FMA_NONE])
def test_mixed_workloads_cover_every_pair_once(make, fma):
    b = make()
    for cta_warps in (0, 1, 4):
        a = ph.plan_audit(b, fma_mode=fma, cta_warps=cta_warps)
        _sound(a)
        assert a["pairs"] == b.pairs()


def test_reads_of_any_length_mix_share_warps():
    """A region whose reads scatter over many lengths: every read takes the lanes it needs and the warps are packed lane by
    lane — one launch, no warp per length class."""
    rng = np.random.default_rng(7)
    hap = B[rng.integers(0, 4, 300)]
    reads = [dict(bases=B[rng.integers(0, 4, n)], qual=np.full(n, 30, np.uint8), mapq=60) for n in range(60, 152, 3)]   # 31 reads, 31 lengths
    b = wl.RegionBatch.from_regions([(reads, [hap, hap[::-1].copy()])] * 20)
    a = ph.plan_audit(b, cta_warps=1, busy_batches=64)   # a busy device: one chunk per region
    _sound(a)
    assert a["slot_fill"] > 0.9, a                    # at 12 rows per lane: 6 to 13 lanes per read, 291 lanes in 10 warps
    assert a["chunks"] == 20 and a["warps"] <= 20 * 10 and a["launches"] == 1


def test_out_of_envelope_shapes_go_to_the_pair_kernel():
    rng = np.random.default_rng(8)
    long_hap = B[rng.integers(0, 4, 23000)]            # beyond one shared-memory stream
    reads = [dict(bases=B[rng.integers(0, 4, n)], qual=np.full(n, 30, np.uint8), mapq=60) for n in (1, 7, 8, 151, 152, 607, 608, 900)]
    regs = [(reads, [long_hap, long_hap[:300].copy(), long_hap[:1].copy()])]
    a = ph.plan_audit(wl.RegionBatch.from_regions(regs))
    _sound(a)
    assert a["per_pair_kernel_pairs"] == 8 + 2 * 2     # every read x the long haplotype, the two long reads x the other two
    for r in reads:                                    # explicit per-base gap qualities: at most 12 rows per lane (<= 383 bp)
        r["gcp"] = np.full(len(r["bases"]), 11, np.uint8)
    a = ph.plan_audit(wl.RegionBatch.from_regions(regs))
    _sound(a)
    assert a["per_pair_kernel_pairs"] == 8 + 3 * 2
    empty = wl.RegionBatch.from_regions([([], [long_hap[:50].copy()]), (reads[:2], [])])
    a = ph.plan_audit(empty)
    assert a["pairs"] == 0 and a["warps"] == 0


_STREAMS = {}


def _stream(name):
    if name not in _STREAMS:
        _STREAMS[name] = {"cfg2": lambda: wl.cfg2(4096, seed=2), "cfg5": lambda: wl.cfg5(2048, seed=5, post_trim_fraction=0.0),
                          "cfg5t": lambda: wl.cfg5(300, seed=6, post_trim_fraction=1.0), "one": lambda: wl.cfg2(1, seed=3),
                          "cfg3": lambda: wl.cfg3(5, seed=3)}[name]()
    return _STREAMS[name]


@pytest.mark.parametrize("name", ["cfg2", "cfg5", "cfg5t", "one", "cfg3"])
def test_flat_partition_covers_the_stream_in_order(name):
    """phmm_run_flat's sub-batches: contiguous, in order, every region exactly once; sizes ramp up from a quarter of the full
    size (so the device starts after one small plan; round 2: from an eighth, several of each size) and down again at the end; no sub-batch above batch_cells unless a
    single region is."""
    b = _stream(name)
    for workers in (2, 4, 6):
        for batch_cells in (0, int(4e9)):
            bounds, work = ph.flat_partition(b, workers, batch_cells)
            assert bounds[0, 0] == 0 and bounds[-1, 1] == b.n_regions
            assert np.array_equal(bounds[1:, 0], bounds[:-1, 1]) and np.all(bounds[:, 1] > bounds[:, 0])
            cap = batch_cells or int(32e9)
            one_region_slack = 1.2 * cap / 4 if b.n_regions > 1 else float("inf")
            assert np.all((work <= cap + one_region_slack) | (bounds[:, 1] - bounds[:, 0] == 1))
            if len(work) >= 12:
                assert work[0] <= work[5] * 0.6 <= work[8] * 0.6 and work[-1] <= work.max() * 0.6          # ramp up, ramp down


def test_planner_quality_does_not_regress():
    """How well a plan uses the lanes it occupies (useful cells / lane-steps, lanes that hold a read) on the shapes the workload
    table watches: a guard for later planner changes — the numbers of the round-2 planner, rounded down.  (Lane utilisation
    is per padded row; the planner trades some of it for more rows per lane, i.e. fewer shuffles and loads per cell.)"""
    for name, util, fill in (("cfg2", 0.97, 0.999), ("cfg5", 0.86, 0.92), ("cfg5t", 0.77, 0.89), ("cfg3", 0.80, 0.85)):
        a = ph.plan_audit(_stream(name), busy_batches=6)
        _sound(a)
        assert a["lane_utilisation"] >= util and a["slot_fill"] >= fill, (name, a["lane_utilisation"], a["slot_fill"])
        assert a["launches"] <= 4, (name, a["launches"])    # a few well-separated rows-per-lane counts (+ the paired reads), not one per length


def test_an_odd_haplotype_is_swept_by_pairs_of_reads():
    """A region with an odd number of haplotypes: the shortest one leaves the two-stream chunks (which then balance) and is
    swept by the reads two at a time, one read in each half of every lane — a launch of its own, every pair still covered
    exactly once, fewer lane-steps than the same region pays for one haplotype more (which makes its count even)."""
    rng = np.random.default_rng(21)

    def region(n_haps, lens):
        hap = B[rng.integers(0, 4, 320)]
        haps = [hap[: 300 + 2 * j].copy() for j in range(n_haps)]
        return ([dict(bases=B[rng.integers(0, 4, n)], qual=np.full(n, 30, np.uint8), mapq=60) for n in lens], haps)

    mixed = [int(x) for x in rng.integers(30, 152, 41)]
    for n_haps in (1, 3, 5, 9):
        for lens in ([150] * 64, [150] * 63, mixed, [100, 101] * 20 + [17]):
            odd = ph.plan_audit(wl.RegionBatch.from_regions([region(n_haps, lens)] * 40), lanes_per_read=8, busy_batches=4)
            even = ph.plan_audit(wl.RegionBatch.from_regions([region(n_haps + 1, lens)] * 40), lanes_per_read=8, busy_batches=4)
            _sound(odd)
            _sound(even)
            assert odd["pairs"] == 40 * n_haps * len(lens)
            assert odd["launches"] == even["launches"] + (1 if n_haps > 1 else 0), (n_haps, odd["launches"], even["launches"])
            assert odd["lane_steps"] < even["lane_steps"], (n_haps, odd["lane_steps"], even["lane_steps"])
    # reads that do not pair up (every one its own lane count): the odd haplotype stays in the two-stream chunks
    lone = ph.plan_audit(wl.RegionBatch.from_regions([region(3, [20, 40, 60, 80, 100, 120, 140])] * 40), lanes_per_read=8, busy_batches=4)
    _sound(lone)
    assert lone["launches"] == 1


def test_random_batches_are_covered_exactly_once():
    """Seeded random batches of awkward shapes — odd and even haplotype counts, haplotypes of length 0, 1 and beyond one stream, reads of
    length 0 up to beyond the wavefront envelope, regions without reads or haplotypes, explicit gap qualities now and then — through
    every layout policy and both arithmetic flavours: every pair computed exactly once, every read on enough lanes, every CTA in shape."""
    rng = np.random.default_rng(2024)
    pool = B[rng.integers(0, 4, 24000)]
    for trial in range(300):
        regs = []
        for _ in range(int(rng.integers(1, 12))):
            n_reads, n_haps = int(rng.integers(0, 70)), int(rng.integers(0, 12))
            hap_lens = [int(rng.choice([0, 1, 7, 40, 80, 300, 301, 450, 1000, 2399, 23000], p=[.02, .02, .04, .1, .2, .3, .1, .1, .06, .04, .02]))
                        for _ in range(n_haps)]
            if rng.random() < 0.3:
                read_lens = [int(rng.choice([100, 150, 151, 250]))] * n_reads
            else:
                read_lens = sorted(int(rng.choice([0, 1, 11, 12, 23, 24, 60, 100, 150, 151, 250, 383, 384, 607, 608, 900])) for _ in range(n_reads))
            explicit = rng.random() < 0.1
            reads = []
            for n in read_lens:
                r = dict(bases=pool[:n].copy(), qual=np.full(n, 30, np.uint8), mapq=60)
                if explicit:
                    r["gcp"] = np.full(n, 10, np.uint8)
                reads.append(r)
            regs.append((reads, [pool[:h].copy() for h in hap_lens]))
        b = wl.RegionBatch.from_regions(regs)
        lanes = int(rng.choice([0, 8, 16, 32]))
        a = ph.plan_audit(b, fma_mode=ph.FMA_AVX512 if trial % 5 else ph.FMA_NONE, cta_warps=int(rng.choice([0, 1, 2, 4])),
                          busy_batches=int(rng.choice([1, 4, 32])), lanes_per_read=lanes)
        _sound(a)
        assert a["pairs"] == b.pairs(), trial
