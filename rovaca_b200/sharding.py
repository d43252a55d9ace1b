"""Multi-GPU: active regions are independent, so the path shards with NO data-path collective.

One process per GPU (torchrun); every rank derives the same deterministic plan, runs its own regions through its own
engine, and only scalars (timings, counts) or — when a caller wants the whole result on one host — the per-region
outputs travel, over ``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests).  NVLink is deliberately unused.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .workloads import RegionBatch


def region_cells(batch: RegionBatch) -> np.ndarray:
    rl = np.concatenate([[0], np.cumsum(batch.read_len.astype(np.int64))])
    hl = np.concatenate([[0], np.cumsum(batch.hap_len.astype(np.int64))])
    r0, r1 = batch.reg_read_off[:-1].astype(np.int64), batch.reg_read_off[1:].astype(np.int64)
    h0, h1 = batch.reg_hap_off[:-1].astype(np.int64), batch.reg_hap_off[1:].astype(np.int64)
    return (rl[r1] - rl[r0]) * (hl[h1] - hl[h0])


def shard_regions(batch: RegionBatch, rank: int, world: int) -> np.ndarray:
    """Indices of the regions rank ``rank`` owns: longest-processing-time-first onto the least loaded rank, so that
    mixed shapes (cfg4/cfg5) balance by cells, not by count.  Deterministic; identical on every rank."""
    if world <= 1:
        return np.arange(batch.n_regions)
    cells = region_cells(batch)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.zeros(batch.n_regions, np.int64)
    for g in order:
        r = int(np.argmin(load))
        owner[g] = r
        load[r] += cells[g]
    return np.nonzero(owner == rank)[0]


def shard_blocks(block_cells: np.ndarray, rank: int, world: int) -> np.ndarray:
    """The same longest-processing-time-first assignment over BLOCKS of consecutive regions of a stream (``block_cells[b]`` =
    cells of block b): a rank generates and runs whole blocks, so a stream of a million regions is never held or shipped in
    one piece.  Returns the ascending block indices rank ``rank`` owns; deterministic, identical on every rank."""
    n = len(block_cells)
    if world <= 1:
        return np.arange(n)
    order = np.argsort(-np.asarray(block_cells, np.int64), kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.zeros(n, np.int64)
    for b in order:
        r = int(np.argmin(load))
        owner[b] = r
        load[r] += int(block_cells[b])
    return np.nonzero(owner == rank)[0]


def gather_digests(region_ids: np.ndarray, digests: np.ndarray, n_regions: int) -> Optional[np.ndarray]:
    """Per-region 64-bit digests of the regions this rank ran -> on rank 0 the digest of every region of the stream (None
    elsewhere).  Host-side gather of a few bytes per region: the only thing that travels between ranks."""
    d = _dist()
    if d is None:
        out = np.zeros(n_regions, np.uint64)
        out[region_ids] = digests
        return out
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, (np.asarray(region_ids, np.int64), np.asarray(digests, np.uint64)))
    if d.get_rank() != 0:
        return None
    out = np.zeros(n_regions, np.uint64)
    seen = np.zeros(n_regions, np.int64)
    for ids, dg in parts:
        out[ids] = dg
        seen[ids] += 1
    if not np.all(seen == 1):
        raise RuntimeError("sharded stream: a region was run by %s ranks" % ("no" if seen.min() == 0 else "several"))
    return out


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device():
    import torch
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def max_over_ranks(x: float) -> float:
    import torch
    d = _dist()
    if d is None:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    import torch
    d = _dist()
    if d is None:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=_device())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def gather_region_values(idx: np.ndarray, values: np.ndarray, n_regions: int) -> Optional[np.ndarray]:
    """Host-side gather of one float per owned region onto rank 0 (returns None elsewhere)."""
    import torch
    d = _dist()
    if d is None:
        out = np.zeros(n_regions)
        out[idx] = values
        return out
    full = torch.zeros(n_regions, dtype=torch.float64, device=_device())
    full[torch.as_tensor(idx, dtype=torch.long, device=full.device)] = torch.as_tensor(values, dtype=torch.float64, device=full.device)
    d.reduce(full, dst=0, op=d.ReduceOp.SUM)
    return full.cpu().numpy() if d.get_rank() == 0 else None
