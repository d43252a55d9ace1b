"""CPU: workload generator, flat-batch bookkeeping and the multi-GPU shard plan (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from rovaca_b200 import workloads as wl
from rovaca_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg2_shape_and_cells():
    b = wl.cfg2(8, seed=1)
    assert b.n_regions == 8 and b.n_reads == 8 * 64 and b.n_haps == 8 * 8
    assert b.cells() == 8 * 64 * 8 * 150 * 300
    assert b.pairs() == 8 * 64 * 8
    assert set(np.unique(b.read_bases[: b.n_reads * 150])) <= set(b"ACGT")
    b2 = wl.cfg2(8, seed=1)
    assert np.array_equal(b.read_bases, b2.read_bases) and np.array_equal(b.read_qual, b2.read_qual)


def test_from_regions_roundtrip():
    b = wl.cfg5(6, seed=3)
    regs = [b.region(g) for g in range(b.n_regions)]
    b2 = wl.RegionBatch.from_regions(regs)
    for name in ("reg_read_off", "reg_hap_off", "read_len", "read_mapq", "hap_len"):
        assert np.array_equal(getattr(b, name), getattr(b2, name))
    for g in range(b.n_regions):
        r1, h1 = b.region(g)
        r2, h2 = b2.region(g)
        assert all(np.array_equal(x["bases"], y["bases"]) and np.array_equal(x["qual"], y["qual"]) for x, y in zip(r1, r2))
        assert all(np.array_equal(x, y) for x, y in zip(h1, h2))
    lik = b.lik_off()
    assert lik[-1] == sum(len(r) * len(h) for r, h in regs)


def test_cfg3_reads_sorted_by_length():
    b = wl.cfg3(1, seed=2, n_reads=50, n_haps=4)
    assert np.all(np.diff(b.read_len.astype(np.int64)) >= 0)


def test_shard_plan_covers_everything_once():
    b = wl.cfg5(37, seed=5)
    for world in (1, 2, 3, 8):
        seen = np.zeros(b.n_regions, int)
        loads = []
        for rank in range(world):
            idx = sharding.shard_regions(b, rank, world)
            seen[idx] += 1
            loads.append(sum(sharding.region_cells(b)[idx]))
        assert np.all(seen == 1)
        if world > 1:
            assert max(loads) <= 1.35 * (sum(loads) / world) + max(sharding.region_cells(b))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = wl.cfg5(16, seed=11)
        idx = sharding.shard_regions(b, rank, world)
        local = b.subset(idx)
        # stand-in for the GPU work: a per-region checksum every rank can compute
        vals = np.array([float(local.read_len[int(local.reg_read_off[g]):int(local.reg_read_off[g + 1])].sum()) for g in range(local.n_regions)])
        gathered = sharding.gather_region_values(idx, vals, b.n_regions)
        t = sharding.max_over_ranks(float(rank + 1))
        total_cells = sharding.sum_over_ranks(float(local.cells()))
        q.put((rank, gathered, t, total_cells))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = wl.cfg5(16, seed=11)
    expect = np.array([float(b.read_len[int(b.reg_read_off[g]):int(b.reg_read_off[g + 1])].sum()) for g in range(b.n_regions)])
    for rank, gathered, t, total_cells in res:
        if rank == 0:
            assert np.array_equal(gathered, expect)
        assert t == 2.0
        assert total_cells == float(b.cells())


@pytest.mark.parametrize("mode", [0, 1, 2, -1])
def test_host_epilogue_on_the_oracles_raw_sums(orc, mode):
    """The engine's host side after the kernels (pairhmm_internal.cpp:578-584 conversion, :53-63 normalize, :161-182 filter,
    :184-193 transpose) fed with the ORACLE's raw sums must give the oracle's matrices byte for byte — in every log10f mode
    (libm, the formula around libm's logf, the AVX-512 restatement), with the FP64 sums and without them (underflowing pairs
    then stand at -inf and never reach the matrix), NaN tail and surviving reads included."""
    import ctypes
    from rovaca_b200 import pairhmm as ph
    if mode == 2 and ph.load_library().phmm_debug_log10f_check(1 << 16, 2, None) == 2**64 - 1:
        pytest.skip("no AVX-512 on this host")
    rng = np.random.default_rng(41)
    specs = [dict(n_reads=24, n_haps=5, hap_lo=120, hap_hi=160, read_lo=40, read_hi=76, junk=0.3),
             dict(n_reads=17, n_haps=9, hap_lo=90, hap_hi=120, read_lo=20, read_hi=60, junk=0.4),      # limit -4 below 51 bp
             dict(n_reads=6, n_haps=3, hap_lo=100, hap_hi=100, read_lo=70, read_hi=70, junk=1.0),      # nothing survives
             dict(n_reads=9, n_haps=1, hap_lo=80, hap_hi=80, read_lo=30, read_hi=60, junk=0.0),
             dict(n_reads=1, n_haps=17, hap_lo=60, hap_hi=90, read_lo=50, read_hi=50, junk=0.0)]
    regs = [wl._one_region(rng, **sp) for sp in specs]
    batch = wl.RegionBatch.from_regions(regs)
    want = [orc.region(reads, haps) for reads, haps in regs]
    raw32 = np.concatenate([w[2].reshape(-1) for w in want])
    raw64 = np.concatenate([w[3].reshape(-1) for w in want])
    assert sum(w[4] for w in want) > 30
    for f64 in (raw64, None):
        res, _ = ph.host_epilogue(batch, raw32, f64, mode=mode)
        assert res.n_fp64 == sum(w[4] for w in want)
        for g, (m0, k0, *_) in enumerate(want):
            assert np.array_equal(res.kept(g), k0), (g, f64 is None)
            assert np.array_equal(res.matrix(g).view(np.uint64), m0.view(np.uint64)), (g, f64 is None)


def _regions_with_ties(rng, n_regions=6):
    """Regions whose haplotype lists hold exact duplicates and near-duplicates, so that best / second-best haplotypes tie and
    the priorities decide (allele_likelihoods.hpp:624-645)."""
    regs, pris = [], []
    for g in range(n_regions):
        reads, haps = wl._one_region(rng, int(rng.integers(5, 40)), int(rng.integers(2, 7)), 120, 160, 40, 90, junk=0.15)
        haps = list(haps)
        for _ in range(int(rng.integers(1, 4))):
            haps.insert(int(rng.integers(0, len(haps) + 1)), haps[int(rng.integers(0, len(haps)))].copy())
        if g == 0:
            haps = haps[:1]                                   # a single haplotype: second best = best, likelihood -inf
        regs.append((reads, haps))
        pris.append(rng.integers(-3, 3, len(haps)).astype(np.float64))
    return regs, pris


@pytest.mark.parametrize("with_priorities", [True, False])
def test_host_best_haplotypes_match_search_best_allele(orc, with_priorities):
    """phmm_result.best_hap / second_hap from the host epilogue against the oracle's restatement of
    AlleleLikelihoods::search_best_allele (genotype/allele_likelihoods.hpp:589-657) on the oracle's own matrices."""
    from rovaca_b200 import pairhmm as ph
    rng = np.random.default_rng(43)
    regs, pris = _regions_with_ties(rng)
    batch = wl.RegionBatch.from_regions(regs)
    if with_priorities:
        batch.hap_priority = np.concatenate(pris)
    want = [orc.region(reads, haps) for reads, haps in regs]
    raw32 = np.concatenate([w[2].reshape(-1) for w in want])
    res, _ = ph.host_epilogue(batch, raw32, None, best=True)
    tie_broken = 0
    for g, (m0, k0, *_) in enumerate(want):
        assert np.array_equal(res.kept(g), k0)
        b0, s0, *_ = orc.best_alleles(m0, len(k0), pris[g] if with_priorities else None)
        b_plain, *_ = orc.best_alleles(m0, len(k0), None)
        tie_broken += int((b0 != b_plain).sum())
        b1, s1 = res.best(g)
        assert np.array_equal(b1, b0) and np.array_equal(s1, s0), g
    assert (tie_broken > 0) == with_priorities


def test_marginalize_matches_marginal_likelihoods_direct(orc):
    """phmm_marginalize against the oracle's restatement of AlleleLikelihoods::marginal_likelihoods_direct
    (genotype/allele_likelihoods.hpp:702-732) on the oracle's matrices: several haplotypes per allele, haplotypes that map
    to nothing, an allele nobody maps to (-inf row)."""
    from rovaca_b200 import pairhmm as ph
    rng = np.random.default_rng(53)
    for _ in range(8):
        reads, haps = wl._one_region(rng, int(rng.integers(1, 50)), int(rng.integers(1, 12)), 100, 140, 30, 80, junk=0.1)
        m, kept, *_ = orc.region(reads, haps)
        n_new = int(rng.integers(1, 5))
        mp = rng.integers(-1, n_new, len(haps)).astype(np.int32)
        got = ph.marginalize(m, len(kept), mp, n_new + 1)          # allele n_new has no haplotype
        want = orc.marginalize(m, len(kept), mp, n_new + 1)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
        assert np.isneginf(got[n_new]).all()
    with pytest.raises(ph.PhmmError):
        ph.marginalize(np.zeros((2, 3)), 3, [0, 5], 2)
