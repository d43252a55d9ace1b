"""Synthetic active-region workloads and the flat batch container the C ABI takes.

Everything is seeded numpy; no reference data is needed.  The shapes follow BASELINE.json's configs:
  cfg2  150 bp reads x 300 bp haplotypes, 8 haplotypes and 64 reads per region (the headline metric)
  cfg3  high depth: 500 reads x 32 haplotypes, mixed lengths, 5 % junk reads to force the FP64 recompute
  cfg4  long-read-like: 250 bp reads x 1000 bp haplotypes, optionally mixed with 150x300 regions
  cfg5  a stream of regions with realistic (post-assembly) haplotype counts and lengths
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class RegionBatch:
    """Flat arrays describing many regions (mirrors ``phmm_flat_batch`` in include/phmm_b200.h)."""
    reg_read_off: np.ndarray   # u64 [n_regions+1]
    reg_hap_off: np.ndarray    # u64 [n_regions+1]
    read_len: np.ndarray       # u32 [n_reads]
    read_mapq: np.ndarray      # u8  [n_reads]
    read_off: np.ndarray       # u64 [n_reads]  byte offset in every read blob
    read_bases: np.ndarray     # u8 blob
    read_qual: np.ndarray      # u8 blob
    hap_len: np.ndarray        # u32 [n_haps]
    hap_off: np.ndarray        # u64 [n_haps]
    hap_bases: np.ndarray      # u8 blob
    read_ins: Optional[np.ndarray] = None
    read_del: Optional[np.ndarray] = None
    read_gcp: Optional[np.ndarray] = None
    name: str = ""
    pinned: bool = False       # the byte blobs are views of page-locked memory (see pin())

    def pin(self) -> "RegionBatch":
        """Move the byte blobs into page-locked host memory (pinned torch tensors, viewed as numpy arrays) so that the
        engine can copy them to the device without staging them first (PHMM_FLAT_PINNED)."""
        import torch
        keep = []
        for name in ("read_bases", "read_qual", "read_ins", "read_del", "read_gcp", "hap_bases"):
            a = getattr(self, name)
            if a is None:
                continue
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            keep.append(t)
            setattr(self, name, t.numpy())
        self._pinned_tensors = keep      # the tensors own the memory
        self.pinned = True
        return self

    @property
    def n_regions(self) -> int:
        return len(self.reg_read_off) - 1

    @property
    def n_reads(self) -> int:
        return len(self.read_len)

    @property
    def n_haps(self) -> int:
        return len(self.hap_len)

    def lik_off(self) -> np.ndarray:
        nr = np.diff(self.reg_read_off).astype(np.uint64)
        nh = np.diff(self.reg_hap_off).astype(np.uint64)
        off = np.zeros(self.n_regions + 1, np.uint64)
        np.cumsum(nr * nh, out=off[1:])
        return off

    def cells(self) -> int:
        """sum over pairs of read_len * hap_len — the numerator of GCUPS."""
        tot = 0
        rl = self.read_len.astype(np.int64)
        hl = self.hap_len.astype(np.int64)
        rcs = np.concatenate([[0], np.cumsum(rl)])
        hcs = np.concatenate([[0], np.cumsum(hl)])
        r0, r1 = self.reg_read_off[:-1].astype(np.int64), self.reg_read_off[1:].astype(np.int64)
        h0, h1 = self.reg_hap_off[:-1].astype(np.int64), self.reg_hap_off[1:].astype(np.int64)
        tot = int(np.sum((rcs[r1] - rcs[r0]) * (hcs[h1] - hcs[h0])))
        return tot

    def pairs(self) -> int:
        return int(self.lik_off()[-1])

    def input_bytes(self) -> int:
        n = self.read_bases.nbytes + self.read_qual.nbytes + self.hap_bases.nbytes
        for a in (self.read_ins, self.read_del, self.read_gcp):
            if a is not None:
                n += a.nbytes
        return n

    def region(self, g: int):
        """(reads, haps) of region g as python objects (dict per read) — for the per-region API and the oracle."""
        r0, r1 = int(self.reg_read_off[g]), int(self.reg_read_off[g + 1])
        h0, h1 = int(self.reg_hap_off[g]), int(self.reg_hap_off[g + 1])
        reads = []
        for k in range(r0, r1):
            o, n = int(self.read_off[k]), int(self.read_len[k])
            d = dict(bases=self.read_bases[o:o + n], qual=self.read_qual[o:o + n], mapq=int(self.read_mapq[k]))
            if self.read_ins is not None:
                d["ins"] = self.read_ins[o:o + n]
            if self.read_del is not None:
                d["del"] = self.read_del[o:o + n]
            if self.read_gcp is not None:
                d["gcp"] = self.read_gcp[o:o + n]
            reads.append(d)
        haps = [self.hap_bases[int(self.hap_off[j]):int(self.hap_off[j]) + int(self.hap_len[j])] for j in range(h0, h1)]
        return reads, haps

    def subset(self, regions: Sequence[int]) -> "RegionBatch":
        return RegionBatch.from_regions([self.region(int(g)) for g in regions], name=self.name)

    @staticmethod
    def from_regions(regions: Sequence[tuple], name: str = "") -> "RegionBatch":
        """regions: sequence of (reads, haps); reads are dicts (bases, qual, mapq[, ins, del, gcp])."""
        rro, rho = [0], [0]
        rlen, rmapq, rb, rq, ri, rd, rg = [], [], [], [], [], [], []
        hlen, hb = [], []
        any_i = any("ins" in r and r["ins"] is not None for reads, _ in regions for r in reads)
        any_d = any("del" in r and r["del"] is not None for reads, _ in regions for r in reads)
        any_g = any("gcp" in r and r["gcp"] is not None for reads, _ in regions for r in reads)
        for reads, haps in regions:
            for r in reads:
                b = _as_u8(r["bases"])
                rlen.append(len(b))
                rmapq.append(int(r.get("mapq", 60)))
                rb.append(b)
                rq.append(_as_u8(r["qual"]))
                if any_i:
                    ri.append(_as_u8(r["ins"]) if r.get("ins") is not None else np.full(len(b), 45, np.uint8))
                if any_d:
                    if r.get("del") is None:
                        raise ValueError("a batch with explicit deletion qualities needs them on every read")
                    rd.append(_as_u8(r["del"]))
                if any_g:
                    rg.append(_as_u8(r["gcp"]) if r.get("gcp") is not None else np.full(len(b), 10, np.uint8))
            for h in haps:
                a = _as_u8(h)
                hlen.append(len(a))
                hb.append(a)
            rro.append(len(rlen))
            rho.append(len(hlen))
        rlen_a = np.array(rlen, np.uint32)
        hlen_a = np.array(hlen, np.uint32)
        return RegionBatch(
            reg_read_off=np.array(rro, np.uint64), reg_hap_off=np.array(rho, np.uint64), read_len=rlen_a,
            read_mapq=np.array(rmapq, np.uint8), read_off=_excl_cumsum(rlen_a), read_bases=_cat(rb), read_qual=_cat(rq),
            hap_len=hlen_a, hap_off=_excl_cumsum(hlen_a), hap_bases=_cat(hb),
            read_ins=_cat(ri) if any_i else None, read_del=_cat(rd) if any_d else None, read_gcp=_cat(rg) if any_g else None,
            name=name)


def _as_u8(x) -> np.ndarray:
    if isinstance(x, str):
        x = x.encode()
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _cat(parts: List[np.ndarray]) -> np.ndarray:
    if not parts:
        return np.zeros(16, np.uint8)
    return np.concatenate(parts + [np.zeros(16, np.uint8)])   # a little slack past the last record


def _excl_cumsum(lens: np.ndarray) -> np.ndarray:
    off = np.zeros(len(lens), np.uint64)
    if len(lens) > 1:
        np.cumsum(lens[:-1], dtype=np.uint64, out=off[1:])
    return off


def _quals(rng: np.random.Generator, shape) -> np.ndarray:
    """80 % Q30-40, 15 % Q20-29, 5 % Q2-19."""
    u = rng.random(shape)
    q = rng.integers(30, 41, shape)
    mid = rng.integers(20, 30, shape)
    low = rng.integers(2, 20, shape)
    q = np.where(u < 0.80, q, np.where(u < 0.95, mid, low))
    return q.astype(np.uint8)


def _mapq(rng: np.random.Generator, n: int) -> np.ndarray:
    m = np.full(n, 60, np.uint8)
    low = rng.random(n) < 0.05
    m[low] = rng.integers(20, 41, int(low.sum())).astype(np.uint8)
    return m


def uniform_regions(n_regions: int, n_reads: int, n_haps: int, read_len: int, hap_len: int, seed: int = 0,
                    sub_rate: float = 0.01, name: str = "") -> RegionBatch:
    """Regions of identical shape, fully vectorised: one reference haplotype of ``hap_len`` uniform ACGT plus
    n_haps-1 alternates carrying 1-3 SNPs; reads sampled from a random haplotype/offset with 1 % substitutions."""
    rng = np.random.default_rng(seed)
    G, H, R = n_regions, hap_len, read_len
    ref = rng.integers(0, 4, (G, 1, H), dtype=np.uint8)
    hap_idx = np.repeat(ref, n_haps, axis=1)                       # [G, n_haps, H] indices 0..3
    for j in range(1, n_haps):
        n_snp = rng.integers(1, 4, G)
        for s in range(3):
            use = n_snp > s
            pos = rng.integers(0, H, G)
            shift = rng.integers(1, 4, G).astype(np.uint8)
            g = np.nonzero(use)[0]
            hap_idx[g, j, pos[g]] = (hap_idx[g, j, pos[g]] + shift[g]) & 3
    which = rng.integers(0, n_haps, (G, n_reads))
    off = rng.integers(0, max(H - R, 0) + 1, (G, n_reads)) if H >= R else np.zeros((G, n_reads), np.int64)
    if H >= R:
        cols = off[:, :, None] + np.arange(R)[None, None, :]
        rd = hap_idx[np.arange(G)[:, None, None], which[:, :, None], cols]
    else:
        rd = rng.integers(0, 4, (G, n_reads, R), dtype=np.uint8)
    sub = rng.random((G, n_reads, R)) < sub_rate
    rd = np.where(sub, (rd + rng.integers(1, 4, (G, n_reads, R)).astype(np.uint8)) & 3, rd)
    read_bases = _ACGT[rd].reshape(-1)
    read_qual = _quals(rng, (G, n_reads, R)).reshape(-1)
    n_r, n_h = G * n_reads, G * n_haps
    rlen = np.full(n_r, R, np.uint32)
    hlen = np.full(n_h, H, np.uint32)
    pad = np.zeros(16, np.uint8)
    return RegionBatch(
        reg_read_off=(np.arange(G + 1, dtype=np.uint64) * np.uint64(n_reads)),
        reg_hap_off=(np.arange(G + 1, dtype=np.uint64) * np.uint64(n_haps)),
        read_len=rlen, read_mapq=_mapq(rng, n_r), read_off=_excl_cumsum(rlen),
        read_bases=np.concatenate([read_bases, pad]), read_qual=np.concatenate([read_qual, pad]),
        hap_len=hlen, hap_off=_excl_cumsum(hlen), hap_bases=np.concatenate([_ACGT[hap_idx].reshape(-1), pad]),
        name=name or f"uniform {R}x{H}, {n_haps} haps x {n_reads} reads x {G} regions")


def cfg2(n_regions: int = 4096, seed: int = 2) -> RegionBatch:
    """BASELINE.json configs[1]: 150 bp reads x 300 bp haplotypes, 8 haplotypes, 64 reads per region."""
    return uniform_regions(n_regions, 64, 8, 150, 300, seed, name=f"cfg2 150x300, 8 haps x 64 reads x {n_regions} regions")


def _one_region(rng: np.random.Generator, n_reads: int, n_haps: int, hap_lo: int, hap_hi: int, read_lo: int, read_hi: int,
                junk: float = 0.0, indel: float = 0.3, sort_reads: bool = True):
    H = int(rng.integers(hap_lo, hap_hi + 1))
    ref = _ACGT[rng.integers(0, 4, H)]
    haps = [ref]
    for _ in range(n_haps - 1):
        h = ref.copy()
        for _s in range(int(rng.integers(1, 4))):
            p = int(rng.integers(0, len(h)))
            h[p] = _ACGT[(int(np.searchsorted(_ACGT, h[p])) + int(rng.integers(1, 4))) % 4]
        if rng.random() < indel and len(h) > 20:
            p = int(rng.integers(1, len(h) - 1))
            if rng.random() < 0.5:
                h = np.concatenate([h[:p], _ACGT[rng.integers(0, 4, int(rng.integers(1, 9)))], h[p:]])
            else:
                h = np.concatenate([h[:p], h[p + int(rng.integers(1, 9)):]])
        haps.append(h)
    reads = []
    for _ in range(n_reads):
        R = int(rng.integers(read_lo, read_hi + 1))
        h = haps[int(rng.integers(0, len(haps)))]
        if rng.random() < junk or len(h) < R:
            rs = _ACGT[rng.integers(0, 4, R)]
        else:
            o = int(rng.integers(0, len(h) - R + 1))
            rs = h[o:o + R].copy()
            m = rng.random(R) < 0.01
            rs[m] = _ACGT[rng.integers(0, 4, int(m.sum()))]
        reads.append(dict(bases=rs, qual=_quals(rng, R), mapq=int(_mapq(rng, 1)[0])))
    if sort_reads:   # the caller's sort on AVX-512 hosts, haplotypecaller_engine.cpp:245-248
        reads.sort(key=lambda r: len(r["bases"]))
    return reads, haps


def cfg3(n_regions: int = 16, seed: int = 3, n_reads: int = 500, n_haps: int = 32) -> RegionBatch:
    """BASELINE.json configs[2]: high-depth regions, mixed lengths, 5 % junk reads (FP64 recompute exercised)."""
    rng = np.random.default_rng(seed)
    regs = [_one_region(rng, n_reads, n_haps, 250, 450, 100, 151, junk=0.05) for _ in range(n_regions)]
    return RegionBatch.from_regions(regs, name=f"cfg3 high-depth {n_reads} reads x {n_haps} haps x {n_regions} regions")


def cfg4(n_regions: int = 64, seed: int = 4, mix: bool = True) -> RegionBatch:
    """BASELINE.json configs[3]: 250 bp reads x 1000 bp haplotypes (16 haps x 64 reads), 50/50 with 150x300 regions."""
    rng = np.random.default_rng(seed)
    regs = []
    for g in range(n_regions):
        if mix and g % 2 == 1:
            regs.append(_one_region(rng, 64, 8, 300, 300, 150, 150, indel=0.0))
        else:
            regs.append(_one_region(rng, 64, 16, 1000, 1000, 250, 250, indel=0.0))
    return RegionBatch.from_regions(regs, name=f"cfg4 long 250x1000 (+150x300 mix={mix}) x {n_regions} regions")


def cfg5(n_regions: int = 256, seed: int = 5, post_trim_fraction: float = 0.5) -> RegionBatch:
    """BASELINE.json configs[4] stand-in: a stream of regions with ~8.8 haplotypes (2..24), haplotype length 201-499,
    Poisson(41) reads of 60-151 bp; ``post_trim_fraction`` of the regions are short post-trim shapes (H 41-120, R <= H)."""
    rng = np.random.default_rng(seed)
    regs = []
    for _ in range(n_regions):
        n_haps = int(np.clip(rng.poisson(7.8) + 1, 2, 24))
        n_reads = int(max(1, rng.poisson(41)))
        if rng.random() < post_trim_fraction:
            H = int(rng.integers(41, 121))
            regs.append(_one_region(rng, n_reads, n_haps, H, H, 10, min(H, 100), junk=0.02))
        else:
            regs.append(_one_region(rng, n_reads, n_haps, 201, 499, 60, 151, junk=0.02))
    return RegionBatch.from_regions(regs, name=f"cfg5 region stream x {n_regions}")


# ---- big streams: the C generator (csrc/workgen.cpp -> libphmm_workgen.so) -----------------------------------------------
import ctypes as _C
import os as _os


class _WgSpec(_C.Structure):
    _fields_ = [("seed", _C.c_uint64), ("hap_lambda", _C.c_double), ("hap_min", _C.c_uint32), ("hap_max", _C.c_uint32),
                ("read_lambda", _C.c_double), ("read_fixed", _C.c_uint32), ("hap_lo", _C.c_uint32), ("hap_hi", _C.c_uint32),
                ("read_lo", _C.c_uint32), ("read_hi", _C.c_uint32), ("post_trim", _C.c_double), ("pt_hap_lo", _C.c_uint32),
                ("pt_hap_hi", _C.c_uint32), ("pt_read_lo", _C.c_uint32), ("pt_read_hi", _C.c_uint32), ("junk", _C.c_double),
                ("indel", _C.c_double), ("sub_rate", _C.c_double)]


# name -> (hap_lambda, hap_min, hap_max, read_lambda, read_fixed, hap_lo, hap_hi, read_lo, read_hi, junk, indel): SURVEY.md §8d
STREAM_KINDS = {
    "cfg2": (0.0, 8, 8, 0.0, 64, 300, 300, 150, 150, 0.0, 0.0),
    "cfg3": (0.0, 32, 32, 0.0, 500, 250, 450, 100, 151, 0.05, 0.3),
    "cfg4": (0.0, 16, 16, 0.0, 64, 1000, 1000, 250, 250, 0.0, 0.0),
    "cfg5": (7.8, 2, 24, 41.0, 0, 201, 499, 60, 151, 0.02, 0.3),
}

_wg = None


def _workgen():
    global _wg
    if _wg is None:
        path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "libphmm_workgen.so")
        if not _os.path.exists(path):
            raise RuntimeError(f"{path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        L = _C.CDLL(path)
        vp, u32, u64 = _C.c_void_p, _C.c_uint32, _C.c_uint64
        L.wg_counts.argtypes = [_C.POINTER(_WgSpec), u64, u32, vp, vp, u32]
        L.wg_lengths.argtypes = [_C.POINTER(_WgSpec), u64, u32, vp, vp, vp, vp, u32]
        L.wg_fill.argtypes = [_C.POINTER(_WgSpec), u64, u32, vp, vp, vp, vp, vp, vp, vp, vp, u32]
        L.wg_digest_regions.argtypes = [u32, vp, vp, vp, vp, vp, vp, vp, u32]
        L.wg_region_cells.argtypes = [_C.POINTER(_WgSpec), u64, u32, vp, vp, u32]
        for f in (L.wg_counts, L.wg_lengths, L.wg_fill, L.wg_digest_regions, L.wg_region_cells):
            f.restype = None
        _wg = L
    return _wg


def _threads(threads: Optional[int]) -> int:
    if threads:
        return int(threads)
    try:
        return max(1, len(_os.sched_getaffinity(0)))
    except AttributeError:
        return _os.cpu_count() or 1


def _spec(kind: str, seed: int, post_trim_fraction: float) -> _WgSpec:
    k = STREAM_KINDS[kind]
    return _WgSpec(seed, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], post_trim_fraction, 41, 120, 10, 100, k[9], k[10], 0.01)


def stream_cells(kind: str, n_regions: int, first: int = 0, seed: int = 5, post_trim_fraction: float = 0.0,
                 threads: Optional[int] = None) -> np.ndarray:
    """Cells of every region of [first, first + n_regions) of the stream, without generating it (for sharding plans)."""
    return stream_work(kind, n_regions, first, seed, post_trim_fraction, threads)[0]


def stream_work(kind: str, n_regions: int, first: int = 0, seed: int = 5, post_trim_fraction: float = 0.0,
                threads: Optional[int] = None):
    """(cells, pairs) of every region of [first, first + n_regions) of the stream, without generating it."""
    sp = _spec(kind, seed, post_trim_fraction)
    cells, pairs = np.zeros(max(int(n_regions), 1), np.uint64), np.zeros(max(int(n_regions), 1), np.uint64)
    _workgen().wg_region_cells(_C.byref(sp), first, int(n_regions), cells.ctypes.data, pairs.ctypes.data, _threads(threads))
    return cells[: int(n_regions)], pairs[: int(n_regions)]


def stream(kind: str, n_regions: int, first: int = 0, seed: int = 5, post_trim_fraction: float = 0.0, pinned: bool = False,
           threads: Optional[int] = None) -> RegionBatch:
    """Regions [first, first + n_regions) of the synthetic stream ``kind`` (STREAM_KINDS).  Region g is a pure function of
    (kind, seed, post_trim_fraction, g): slices generated by different processes concatenate to the one stream, so N ranks
    can shard a common stream without shipping it.  ``pinned``: the byte blobs are allocated page-locked (torch) and filled
    in place."""
    sp = _spec(kind, seed, post_trim_fraction)
    L, n, th = _workgen(), int(n_regions), _threads(threads)
    n_reads, n_haps = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
    L.wg_counts(_C.byref(sp), first, n, n_reads.ctypes.data, n_haps.ctypes.data, th)
    rro, rho = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint64)
    np.cumsum(n_reads[:n], out=rro[1:])
    np.cumsum(n_haps[:n], out=rho[1:])
    read_len, hap_len = np.zeros(max(int(rro[-1]), 1), np.uint32), np.zeros(max(int(rho[-1]), 1), np.uint32)
    L.wg_lengths(_C.byref(sp), first, n, rro.ctypes.data, rho.ctypes.data, read_len.ctypes.data, hap_len.ctypes.data, th)
    read_len, hap_len = read_len[: int(rro[-1])], hap_len[: int(rho[-1])]
    read_off, hap_off = _excl_cumsum(read_len), _excl_cumsum(hap_len)
    rb_n, hb_n = int(read_len.sum(dtype=np.uint64)) + 16, int(hap_len.sum(dtype=np.uint64)) + 16
    keep = []

    def blob(nbytes):
        if pinned:
            import torch
            t = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
            keep.append(t)
            return t.numpy()
        return np.zeros(nbytes, np.uint8)

    read_bases, read_qual, hap_bases = blob(rb_n), blob(rb_n), blob(hb_n)
    read_mapq = np.zeros(max(len(read_len), 1), np.uint8)
    L.wg_fill(_C.byref(sp), first, n, rro.ctypes.data, rho.ctypes.data, read_off.ctypes.data, hap_off.ctypes.data,
              read_bases.ctypes.data, read_qual.ctypes.data, read_mapq.ctypes.data, hap_bases.ctypes.data, th)
    b = RegionBatch(reg_read_off=rro, reg_hap_off=rho, read_len=read_len, read_mapq=read_mapq[: len(read_len)], read_off=read_off,
                    read_bases=read_bases, read_qual=read_qual, hap_len=hap_len, hap_off=hap_off, hap_bases=hap_bases,
                    name=f"{kind} stream, regions {first}..{first + n} (seed {seed}, post-trim {post_trim_fraction:g})", pinned=pinned)
    b._pinned_tensors = keep
    return b


def region_digests(result, threads: Optional[int] = None) -> np.ndarray:
    """One 64-bit digest per region of a FlatResult (matrix bytes incl. the NaN tail, kept indices, kept count): what ranks
    and runs exchange to prove their results identical without shipping the matrices."""
    b = result.batch
    out = np.zeros(max(b.n_regions, 1), np.uint64)
    _workgen().wg_digest_regions(b.n_regions, b.reg_read_off.ctypes.data, b.reg_hap_off.ctypes.data, result.lik_off.ctypes.data,
                                 result.likelihoods.ctypes.data, result.kept_idx.ctypes.data, result.n_kept.ctypes.data,
                                 out.ctypes.data, _threads(threads))
    return out[: b.n_regions]


def concat(batches: Sequence[RegionBatch], name: str = "") -> RegionBatch:
    """Regions of several batches, in order, as one batch (default gap qualities only)."""
    rro, rho, ro, ho = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint64)], [], []
    r_at = h_at = rb_at = hb_at = 0
    for b in batches:
        rro.append(b.reg_read_off[1:] + np.uint64(r_at))
        rho.append(b.reg_hap_off[1:] + np.uint64(h_at))
        ro.append(b.read_off + np.uint64(rb_at))
        ho.append(b.hap_off + np.uint64(hb_at))
        r_at += b.n_reads
        h_at += b.n_haps
        rb_at += int(b.read_len.sum(dtype=np.uint64))
        hb_at += int(b.hap_len.sum(dtype=np.uint64))
    pad = np.zeros(16, np.uint8)
    cut = lambda a, lens: a[: int(lens.sum(dtype=np.uint64))]
    return RegionBatch(
        reg_read_off=np.concatenate(rro), reg_hap_off=np.concatenate(rho), read_len=np.concatenate([b.read_len for b in batches]),
        read_mapq=np.concatenate([b.read_mapq for b in batches]), read_off=np.concatenate(ro),
        read_bases=np.concatenate([cut(b.read_bases, b.read_len) for b in batches] + [pad]),
        read_qual=np.concatenate([cut(b.read_qual, b.read_len) for b in batches] + [pad]),
        hap_len=np.concatenate([b.hap_len for b in batches]), hap_off=np.concatenate(ho),
        hap_bases=np.concatenate([cut(b.hap_bases, b.hap_len) for b in batches] + [pad]), name=name)


def take(batch: RegionBatch, regions) -> RegionBatch:
    """The given regions of ``batch`` as a batch of their own, vectorised (``RegionBatch.subset`` goes through Python objects
    per read; this one is for shards of big streams).  Default gap qualities only."""
    regions = np.asarray(regions, np.int64)
    nr = (batch.reg_read_off[regions + 1] - batch.reg_read_off[regions]).astype(np.int64)
    nh = (batch.reg_hap_off[regions + 1] - batch.reg_hap_off[regions]).astype(np.int64)

    def expand(starts, counts):                       # indices starts[i] .. starts[i] + counts[i], concatenated
        total = int(counts.sum())
        if total == 0:
            return np.zeros(0, np.int64)
        ends = np.cumsum(counts)
        idx = np.arange(total, dtype=np.int64) - np.repeat(ends - counts, counts)
        return idx + np.repeat(starts, counts)

    ri = expand(batch.reg_read_off[regions].astype(np.int64), nr)
    hi = expand(batch.reg_hap_off[regions].astype(np.int64), nh)
    read_len, hap_len = batch.read_len[ri], batch.hap_len[hi]
    rbi = expand(batch.read_off[ri].astype(np.int64), read_len.astype(np.int64))
    hbi = expand(batch.hap_off[hi].astype(np.int64), hap_len.astype(np.int64))
    pad = np.zeros(16, np.uint8)
    rro, rho = np.zeros(len(regions) + 1, np.uint64), np.zeros(len(regions) + 1, np.uint64)
    np.cumsum(nr, out=rro[1:])
    np.cumsum(nh, out=rho[1:])
    return RegionBatch(reg_read_off=rro, reg_hap_off=rho, read_len=read_len, read_mapq=batch.read_mapq[ri], read_off=_excl_cumsum(read_len),
                       read_bases=np.concatenate([batch.read_bases[rbi], pad]), read_qual=np.concatenate([batch.read_qual[rbi], pad]),
                       hap_len=hap_len, hap_off=_excl_cumsum(hap_len), hap_bases=np.concatenate([batch.hap_bases[hbi], pad]),
                       name=batch.name + f" [{len(regions)} regions]")
