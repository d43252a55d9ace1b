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
