"""CPU: the stream generator (csrc/workgen.cpp), the per-region digests and the sharded-stream bookkeeping (gloo, world_size 2).

The GPU work is stood in for by the CPU checker (the reference compiled in place, or the scalar port): what is tested here is
that N ranks which shard ONE common stream, generate their own blocks and exchange nothing but digests reproduce the
single-process result region for region — the host logic of BASELINE.json configs[4].  The same code runs on the GPUs in
tests/test_gpu_configs.py and in bench.py's `sharded_stream` record."""
import os
import socket

import numpy as np
import pytest

from rovaca_b200 import sharding, streams
from rovaca_b200 import workloads as wl


def test_stream_slices_concatenate_to_the_stream():
    whole = wl.stream("cfg5", 3000, post_trim_fraction=0.5, seed=9)
    a = wl.stream("cfg5", 1000, first=500, post_trim_fraction=0.5, seed=9)
    c = wl.take(whole, np.arange(500, 1500))
    for f in ("read_len", "hap_len", "read_mapq", "reg_read_off", "reg_hap_off", "read_off", "hap_off"):
        assert np.array_equal(getattr(a, f), getattr(c, f)), f
    nb, hb = int(a.read_len.sum()), int(a.hap_len.sum())
    assert np.array_equal(a.read_bases[:nb], c.read_bases[:nb]) and np.array_equal(a.read_qual[:nb], c.read_qual[:nb])
    assert np.array_equal(a.hap_bases[:hb], c.hap_bases[:hb])
    both = wl.concat([wl.stream("cfg5", 500, post_trim_fraction=0.5, seed=9), a])
    d = wl.take(whole, np.arange(0, 1500))
    assert np.array_equal(both.read_bases[: int(both.read_len.sum())], d.read_bases[: int(d.read_len.sum())])
    assert np.array_equal(both.reg_read_off, d.reg_read_off)
    # one thread or many, the same bytes
    e = wl.stream("cfg5", 1000, first=500, post_trim_fraction=0.5, seed=9, threads=1)
    assert np.array_equal(a.read_qual, e.read_qual) and np.array_equal(a.hap_bases, e.hap_bases)
    assert np.array_equal(wl.stream_cells("cfg5", 1000, 500, 9, 0.5).astype(np.int64), sharding.region_cells(a))


def test_stream_shapes_follow_the_survey():
    """~8.8 haplotypes (2..24) of 201-499 bp (+ indels of up to 8), Poisson(41) reads of 60-151 bp sorted by
    length; post-trim regions H 41-120 (+-8) with R <= min(H, 100); only ACGT; qualities 2..40; mapq 60 or 20..40."""
    b = wl.stream("cfg5", 4000, seed=3)
    nh, nr = np.diff(b.reg_hap_off), np.diff(b.reg_read_off)
    assert 8.3 < nh.mean() < 9.3 and nh.min() >= 2 and nh.max() <= 24
    assert 40 < nr.mean() < 42 and nr.min() >= 1
    assert b.hap_len.min() >= 193 and b.hap_len.max() <= 507 and b.read_len.min() >= 60 and b.read_len.max() <= 151
    assert set(np.unique(b.read_bases[: int(b.read_len.sum())])) <= set(b"ACGT")
    assert set(np.unique(b.hap_bases[: int(b.hap_len.sum())])) <= set(b"ACGT")
    q = b.read_qual[: int(b.read_len.sum())]
    assert q.min() >= 2 and q.max() <= 40 and 0.78 < (q >= 30).mean() < 0.82
    assert set(np.unique(b.read_mapq)) <= set(range(20, 41)) | {60}
    for g in range(0, 4000, 97):                                    # the caller's sort, haplotypecaller_engine.cpp:245-248
        rl = b.read_len[int(b.reg_read_off[g]):int(b.reg_read_off[g + 1])]
        assert np.all(np.diff(rl.astype(np.int64)) >= 0)
    t = wl.stream("cfg5", 2000, seed=3, post_trim_fraction=1.0)
    assert t.hap_len.min() >= 33 and t.hap_len.max() <= 128 and t.read_len.min() >= 10 and t.read_len.max() <= 100
    for kind, reads, haps in (("cfg2", 64, 8), ("cfg3", 500, 32), ("cfg4", 64, 16)):
        x = wl.stream(kind, 3)
        assert np.all(np.diff(x.reg_read_off) == reads) and np.all(np.diff(x.reg_hap_off) == haps)
    assert wl.stream("cfg2", 5).cells() == 5 * 64 * 8 * 150 * 300


def test_stream_regions_are_valid_input(orc, ref):
    """Regions of the C generator through the reference's call_pairhmm and through the scalar port: byte-identical."""
    b = wl.stream("cfg5", 24, seed=21, post_trim_fraction=0.5)
    lik, kept, nk = ref.regions_flat(b, 4)
    lo = b.lik_off()
    for g in range(0, 24, 3):
        reads, haps = b.region(g)
        m, k, *_ = orc.region(reads, haps)
        got = lik[int(lo[g]):int(lo[g + 1])].reshape(len(haps), len(reads))
        assert np.array_equal(got.view(np.uint64), m.view(np.uint64))
        r0 = int(b.reg_read_off[g])
        assert np.array_equal(kept[r0:r0 + int(nk[g])], k)


def test_block_sharding_balances_and_covers():
    spec = streams.StreamSpec("cfg5", 20000, 5, 0.5)
    cells = spec.block_cells()
    assert len(cells) == spec.n_blocks == 40 and cells.sum() == int(wl.stream_cells("cfg5", 20000, 0, 5, 0.5).sum())
    for world in (2, 4, 8):
        owned = [sharding.shard_blocks(cells, r, world) for r in range(world)]
        assert sorted(np.concatenate(owned).tolist()) == list(range(spec.n_blocks))
        load = np.array([cells[o].sum() for o in owned], float)
        assert load.max() / load.mean() < 1.05
    pieces = list(streams._pieces(spec, [0, 1, 2, 5, 9, 39], 1024))
    assert pieces == [[(0, 1024)], [(1024, 512), (2560, 512)], [(4608, 512), (19968, 32)]]


class CpuStandInEngine:
    """``run_flat`` of the engine, computed by the CPU checker — so that the stream / shard / digest logic runs without a GPU."""

    def __init__(self):
        from oracle import bindings as ob
        self.ref, self.orc = ob.ref(), ob.oracle()

    def run_flat(self, batch, out=None, **_):
        if self.ref is not None:
            lik, kept, nk = self.ref.regions_flat(batch, 2)
            out.likelihoods[: len(lik)] = lik
            out.kept_idx[: len(kept)] = kept
            out.n_kept[: len(nk)] = nk
        else:
            lo = out.lik_off
            for g in range(batch.n_regions):
                reads, haps = batch.region(g)
                m, k, *_ = self.orc.region(reads, haps)
                out.likelihoods[int(lo[g]):int(lo[g + 1])] = m.reshape(-1)
                r0 = int(batch.reg_read_off[g])
                out.kept_idx[r0:r0 + len(k)] = k
                out.n_kept[g] = len(k)
        out.n_fp64 = 0
        return out


SPEC = dict(kind="cfg5", n_regions=1100, seed=13, post_trim_fraction=1.0)     # post-trim shapes: small regions, quick on a CPU


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec = streams.sharded_stream(CpuStandInEngine(), streams.StreamSpec(**SPEC), rank, world, piece_regions=600)
        q.put((rank, rec))
    finally:
        dist.destroy_process_group()


def test_sharded_stream_two_ranks_gloo():
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
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] is None
    rec = res[0]
    assert rec["n_gpus"] == 2 and rec["regions"] == SPEC["n_regions"]
    assert rec["identical_to_single_gpu"] is True and rec["regions_differing_from_single_gpu"] == 0
    assert rec["cells"] == int(wl.stream_cells(SPEC["kind"], SPEC["n_regions"], 0, SPEC["seed"], SPEC["post_trim_fraction"]).sum())
    # the same digest of digests as one process on its own
    solo = streams.sharded_stream(CpuStandInEngine(), streams.StreamSpec(**SPEC), 0, 1, piece_regions=600)
    assert solo["digest_of_digests"] == rec["digest_of_digests"]


def test_digests_see_every_byte():
    b = wl.stream("cfg5", 8, seed=2, post_trim_fraction=1.0)
    from rovaca_b200.pairhmm import FlatResult
    res = CpuStandInEngine().run_flat(b, out=FlatResult(b))
    d0 = wl.region_digests(res).copy()
    assert len(set(d0.tolist())) == 8
    res.likelihoods[int(res.lik_off[3])] = np.nextafter(res.likelihoods[int(res.lik_off[3])], 0)
    res.n_kept[5] -= 1
    d1 = wl.region_digests(res)
    assert [i for i in range(8) if d0[i] != d1[i]] == [3, 5]
