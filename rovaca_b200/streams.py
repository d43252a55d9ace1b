"""Region streams through the engine: BASELINE.json configs[4] ("whole-genome-scale stream of active regions sharded across
2/4/8 B200") reduced, as SURVEY.md §7/§8d explains, to matrix identity on a synthetic stream — the full BAM -> GVCF pipeline
around the path cannot be built here.

A stream is never held in one piece: it is cut into blocks of consecutive regions (``BLOCK`` regions each), every rank or
device takes whole blocks (``sharding.shard_blocks``: longest-processing-time-first by cells), generates them itself
(``workloads.stream``: region g is a pure function of the stream's parameters and g), runs them through ``phmm_run_flat`` in
pieces of a few ten thousand regions, and keeps one 64-bit digest per region (matrix bytes with the NaN tail, kept indices,
kept count).  Only digests travel between ranks; equal digests = byte-identical results.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import sharding, workloads
from .pairhmm import FlatResult, PairHmmEngine

BLOCK = 512              # regions per block of a stream: the unit ranks / devices are dealt
PAIR_WEIGHT = 16384      # cells a pair is charged for the host's work on it (kPairWeight in the engine)
PIECE = 65536            # regions per phmm_run_flat call (one pinned input set, one result set: ~0.6 GB for cfg5)


@dataclass
class StreamSpec:
    kind: str = "cfg5"
    n_regions: int = 131072
    seed: int = 5
    post_trim_fraction: float = 0.5

    @property
    def n_blocks(self) -> int:
        return (self.n_regions + BLOCK - 1) // BLOCK

    def block_range(self, b: int):
        first = b * BLOCK
        return first, min(BLOCK, self.n_regions - first)

    def block_cells(self) -> np.ndarray:
        c = workloads.stream_cells(self.kind, self.n_regions, 0, self.seed, self.post_trim_fraction).astype(np.int64)
        pad = self.n_blocks * BLOCK - self.n_regions
        return np.concatenate([c, np.zeros(pad, np.int64)]).reshape(self.n_blocks, BLOCK).sum(axis=1)

    def block_work(self) -> np.ndarray:
        """What a block costs end to end, in cell units: its cells plus a per-pair charge for the host's share (planning and
        result hand-over cost about as much per pair as PAIR_WEIGHT cells cost the GPU — phmm_run_flat cuts its sub-batches by
        the same measure).  Small post-trim regions are host-heavy: balancing ranks by cells alone leaves them uneven."""
        c, p = workloads.stream_work(self.kind, self.n_regions, 0, self.seed, self.post_trim_fraction)
        w = c.astype(np.int64) + PAIR_WEIGHT * p.astype(np.int64)
        pad = self.n_blocks * BLOCK - self.n_regions
        return np.concatenate([w, np.zeros(pad, np.int64)]).reshape(self.n_blocks, BLOCK).sum(axis=1)

    def describe(self) -> str:
        return (f"{self.kind} stream of {self.n_regions} regions (seed {self.seed}, {self.post_trim_fraction:.0%} post-trim shapes), "
                f"blocks of {BLOCK}")


@dataclass
class StreamRun:
    region_ids: np.ndarray                 # global region indices this run covered, in the order run
    digests: np.ndarray                    # one per region
    cells: int = 0
    pairs: int = 0
    seconds: float = 0.0                   # inside phmm_run_flat only (host buffers in, host results out)
    n_fp64: int = 0
    pieces: int = 0
    piece_seconds: List[float] = field(default_factory=list)


def _pieces(spec: StreamSpec, blocks: Sequence[int], piece_regions: int):
    """Consecutive owned blocks merged into contiguous ranges, ranges grouped into pieces of at most piece_regions regions."""
    ranges = []
    for b in blocks:
        first, n = spec.block_range(int(b))
        if ranges and ranges[-1][0] + ranges[-1][1] == first:
            ranges[-1][1] += n
        else:
            ranges.append([first, n])
    piece, count = [], 0
    for first, n in ranges:
        while n > 0:
            take = min(n, piece_regions - count)
            piece.append((first, take))
            first, n, count = first + take, n - take, count + take
            if count == piece_regions:
                yield piece
                piece, count = [], 0
    if piece:
        yield piece


def run_stream(eng: PairHmmEngine, spec: StreamSpec, blocks: Optional[Sequence[int]] = None, piece_regions: int = PIECE,
               pinned: bool = True, warm: bool = True, keep_results: Optional[list] = None) -> StreamRun:
    """Run the given blocks of the stream (default: all of it) through ``eng.run_flat``; -> digests and timing.
    ``warm``: the first piece is run once untimed first (arena growth, clocks)."""
    if blocks is None:
        blocks = np.arange(spec.n_blocks)
    piece_regions *= max(1, int(getattr(eng, "n_devices", 1)))      # an engine over N devices takes N times the piece per call
    if pinned:                                  # page-locked inputs need the CUDA runtime (the CPU tests stand the engine in)
        import torch
        pinned = torch.cuda.is_available()
    ids, digs = [], []
    out = StreamRun(np.zeros(0, np.int64), np.zeros(0, np.uint64))
    for piece in _pieces(spec, blocks, piece_regions):
        parts = [workloads.stream(spec.kind, n, first, spec.seed, spec.post_trim_fraction, pinned=pinned and len(piece) == 1)
                 for first, n in piece]
        batch = parts[0] if len(parts) == 1 else workloads.concat(parts)
        if pinned and not batch.pinned:
            batch.pin()
        res = FlatResult(batch)
        if warm:
            eng.run_flat(batch, out=res)
            warm = False
        t0 = time.perf_counter()
        eng.run_flat(batch, out=res)
        dt = time.perf_counter() - t0
        out.seconds += dt
        out.piece_seconds.append(dt)
        out.cells += batch.cells()
        out.pairs += batch.pairs()
        out.n_fp64 += res.n_fp64
        out.pieces += 1
        ids.append(np.concatenate([np.arange(first, first + n, dtype=np.int64) for first, n in piece]))
        digs.append(workloads.region_digests(res))
        if keep_results is not None:
            keep_results.append((ids[-1], batch, res))
    if ids:
        out.region_ids, out.digests = np.concatenate(ids), np.concatenate(digs)
    return out


def sharded_stream(eng: PairHmmEngine, spec: StreamSpec, rank: int, world: int, piece_regions: int = PIECE) -> Optional[dict]:
    """One common stream sharded over ``world`` ranks (one engine each) with no data-path collective; rank 0 then runs the
    whole stream alone on its own GPU and compares every region's digest.  -> the record (rank 0) or None."""
    mine = sharding.shard_blocks(spec.block_work(), rank, world)
    sharding.barrier()
    run = run_stream(eng, spec, mine, piece_regions)
    secs = sharding.max_over_ranks(run.seconds)
    cells = sharding.sum_over_ranks(float(run.cells))
    all_dig = sharding.gather_digests(run.region_ids, run.digests, spec.n_regions)
    rank_secs = sharding.gather_region_values(np.array([rank]), np.array([run.seconds]), world)
    if rank != 0:
        sharding.barrier()                    # rank 0 is running the whole stream alone: leave it the host's cores
        return None
    rec = {"stream": spec.describe(), "regions": spec.n_regions, "cells": int(cells), "n_gpus": world,
           "e2e_gcups": round(cells / secs / 1e9, 2), "seconds_max_over_ranks": round(secs, 4),
           "seconds_per_rank": [round(float(x), 4) for x in rank_secs], "sharding": f"LPT by work (cells + {PAIR_WEIGHT} per pair) over blocks of {BLOCK} regions, no collective",
           "what_travels": "one 64-bit digest per region to rank 0 (host-side gather)"}
    if world > 1:
        solo = run_stream(eng, spec, None, piece_regions, warm=False)
        order = np.argsort(solo.region_ids)
        solo_dig = solo.digests[order]
        rec["single_gpu_e2e_gcups"] = round(solo.cells / solo.seconds / 1e9, 2)
        rec["regions_differing_from_single_gpu"] = int(np.count_nonzero(solo_dig != all_dig))
        rec["identical_to_single_gpu"] = bool(np.array_equal(solo_dig, all_dig))
    sharding.barrier()
    rec["digest_of_digests"] = f"{int(np.bitwise_xor.reduce(all_dig * np.uint64(0x9E3779B97F4A7C15))):016x}"
    return rec
