"""ctypes doors to the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Two libraries live behind this module:

* ``liboracle.so``  — our scalar C restatement (``oracle/phmm_oracle.c``), always buildable.
* ``_ref/libphmm_ref.so`` — the reference's *own* PairHMM sources compiled in place from
  ``/root/reference`` by ``oracle/Makefile`` (kernels, ``pairhmm_internal.cpp`` wrapper,
  ``ReadRecord``).  Git-ignored, travels to the GPU box as a built file; absent ⇒ ``ref()`` is None.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package ``rovaca_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

PCR_NONE, PCR_HOSTILE, PCR_AGGRESSIVE, PCR_CONSERVATIVE = 0, 1, 2, 3
FMA_NONE, FMA_AVX512 = 0, 1

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _u8(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray, str)):
        x = np.frombuffer(x.encode() if isinstance(x, str) else bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def build(ref: bool = True) -> None:
    """Compile the checkers (``make -C oracle``); ``_ref`` only if /root/reference is present."""
    target = "all" if ref else "oracle"
    subprocess.run(["make", "-s", "-C", _HERE, target], check=True)


class _OrcRead(C.Structure):
    _fields_ = [("bases", _u8p), ("qual", _u8p), ("ins", _u8p), ("del_", _u8p), ("gcp", _u8p),
                ("len", C.c_uint32), ("mapq", C.c_uint8)]


class _OrcHap(C.Structure):
    _fields_ = [("bases", _u8p), ("len", C.c_uint32)]


SW_REALIGN = (10, -15, -30, -5)        # sw_parameters_genotype, common/base/lib/sw_avx/sw_avx.c:14
SW_ASSEMBLE = (200, -150, -260, -11)   # sw_parameters_assemble, :13


class _SwParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32)]


def cigar_string(cigar) -> str:
    return "".join(f"{int(c) >> 4}{'MIDNSHP=X'[int(c) & 15]}" for c in cigar)


class Oracle:
    """Scalar restatement (phmm_oracle.c)."""

    def __init__(self, path: Optional[str] = None):
        path = path or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.orc_init.restype = None
        for name, rt in (("orc_ph2pr_f32", _f32p), ("orc_ph2pr_f64", _f64p), ("orc_mm_tri_f32", _f32p), ("orc_mm_tri_f64", _f64p)):
            getattr(L, name).restype = rt
        L.orc_pair_f32.restype = C.c_float
        L.orc_pair_f32.argtypes = [_u8p, C.c_uint32, _u8p, _u8p, _u8p, _u8p, _u8p, C.c_uint32, C.c_int]
        L.orc_pair_f64.restype = C.c_double
        L.orc_pair_f64.argtypes = [_u8p, C.c_uint32, _u8p, _u8p, _u8p, _u8p, _u8p, C.c_uint32]
        L.orc_log10_likelihood.restype = C.c_double
        L.orc_log10_likelihood.argtypes = [C.c_float, C.c_double]
        L.orc_cap_quals.restype = None
        L.orc_cap_quals.argtypes = [_u8p, C.c_uint32, C.c_uint8, C.c_int32]
        L.orc_find_tandem_repeat_units.restype = C.c_int32
        L.orc_find_tandem_repeat_units.argtypes = [_u8p, C.c_uint32, C.c_uint32]
        L.orc_apply_pcr_error_model.restype = None
        L.orc_apply_pcr_error_model.argtypes = [_u8p, _u8p, C.c_uint32, C.c_int]
        L.orc_region.restype = C.c_int64
        L.orc_region.argtypes = [C.POINTER(_OrcRead), C.c_uint32, C.POINTER(_OrcHap), C.c_uint32, C.c_int32, C.c_int, C.c_int,
                                 _f64p, _u32p, _u32p, _f32p, _f64p]
        L.orc_best_alleles.restype = None
        L.orc_best_alleles.argtypes = [_f64p, C.c_uint32, C.c_uint32, C.c_uint32, _f64p, _u32p, _u32p, _f64p, _f64p, _f64p]
        L.orc_marginalize.restype = None
        L.orc_marginalize.argtypes = [_f64p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, _f64p]
        L.orc_sw_align.restype = C.c_int32
        L.orc_sw_align.argtypes = [_u8p, C.c_int32, _u8p, C.c_int32, C.c_void_p, _u32p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_init()

    # tables -------------------------------------------------------------------------------
    def ph2pr_f32(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.orc_ph2pr_f32(), shape=(128,)).copy()

    def mm_tri_f32(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.orc_mm_tri_f32(), shape=(8256,)).copy()

    def ph2pr_f64(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.orc_ph2pr_f64(), shape=(128,)).copy()

    def mm_tri_f64(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.orc_mm_tri_f64(), shape=(8256,)).copy()

    # pairs --------------------------------------------------------------------------------
    def pair_f32(self, hap, rs, q, ins, dele, gcp, fma_mode: int = FMA_AVX512) -> np.float32:
        hap, rs, q, ins, dele, gcp = map(_u8, (hap, rs, q, ins, dele, gcp))
        return np.float32(self.lib.orc_pair_f32(_p(hap, _u8p), len(hap), _p(rs, _u8p), _p(q, _u8p), _p(ins, _u8p),
                                                _p(dele, _u8p), _p(gcp, _u8p), len(rs), fma_mode))

    def pair_f64(self, hap, rs, q, ins, dele, gcp) -> float:
        hap, rs, q, ins, dele, gcp = map(_u8, (hap, rs, q, ins, dele, gcp))
        return float(self.lib.orc_pair_f64(_p(hap, _u8p), len(hap), _p(rs, _u8p), _p(q, _u8p), _p(ins, _u8p),
                                           _p(dele, _u8p), _p(gcp, _u8p), len(rs)))

    def sw_align(self, seq1, seq2, params=SW_REALIGN):
        """Scalar restatement of the reference's Smith-Waterman (oracle/sw_oracle.c) -> (cigar, alignment_offset, score)."""
        a, b = _u8(seq1), _u8(seq2)
        cig = np.zeros(4096, np.uint32)
        off, score = C.c_int32(0), C.c_int32(0)
        prm = _SwParams(*params)
        n = self.lib.orc_sw_align(_p(a, _u8p), len(a), _p(b, _u8p), len(b), C.byref(prm), _p(cig, _u32p), len(cig), C.byref(off),
                                  C.byref(score))
        return (None if n < 0 else cig[:n].copy()), int(off.value), int(score.value)

    def log10_likelihood(self, raw_f32, raw_f64=float("nan")) -> float:
        return float(self.lib.orc_log10_likelihood(C.c_float(float(raw_f32)), C.c_double(raw_f64)))

    # read preparation ---------------------------------------------------------------------
    def cap_quals(self, qual, mapq: int, thr: int) -> np.ndarray:
        q = _u8(qual).copy()
        self.lib.orc_cap_quals(_p(q, _u8p), len(q), mapq, thr)
        return q

    def find_tandem_repeat_units(self, bases, offset: int) -> int:
        b = _u8(bases)
        return int(self.lib.orc_find_tandem_repeat_units(_p(b, _u8p), len(b), offset))

    def apply_pcr_error_model(self, bases, gap_qual, model: int) -> np.ndarray:
        b, g = _u8(bases), _u8(gap_qual).copy()
        self.lib.orc_apply_pcr_error_model(_p(b, _u8p), _p(g, _u8p), len(b), model)
        return g

    def prepare_read(self, bases, qual, mapq: int, thr: int, model: int, ins=None, dele=None, gcp=None):
        """(q, ins, del, gcp) exactly as the wrapper feeds them to the kernels."""
        b = _u8(bases)
        q = self.cap_quals(qual, mapq, thr)
        i = np.full(len(b), 45, np.uint8) if ins is None else _u8(ins).copy()
        if model != PCR_NONE:
            i = self.apply_pcr_error_model(b, i, model)
        d = i if dele is None else _u8(dele)
        c = np.full(len(b), 10, np.uint8) if gcp is None else _u8(gcp)
        return q, i, d, c

    # region -------------------------------------------------------------------------------
    def marginalize(self, matrix: np.ndarray, n_evidence: int, old_to_new, n_new: int) -> np.ndarray:
        """marginal_likelihoods_direct on the first n_evidence columns of matrix[n_old, stride] -> [n_new, n_evidence]."""
        m = np.ascontiguousarray(matrix, np.float64)
        mp = np.ascontiguousarray(old_to_new, np.int32)
        out = np.zeros((n_new, n_evidence), np.float64)
        self.lib.orc_marginalize(_p(m, _f64p), m.shape[0], m.shape[1], n_evidence, mp.ctypes.data_as(C.POINTER(C.c_int32)), n_new,
                                 _p(out, _f64p))
        return out

    def best_alleles(self, matrix: np.ndarray, n_evidence: int, priorities=None):
        """search_best_allele for the first n_evidence columns of matrix[n_alleles, stride] ->
        (best, second, best_lik, second_lik, confidence)."""
        m = np.ascontiguousarray(matrix, np.float64)
        na, stride = m.shape
        pri = None if priorities is None else np.ascontiguousarray(priorities, np.float64)
        best, second = np.zeros(max(n_evidence, 1), np.uint32), np.zeros(max(n_evidence, 1), np.uint32)
        bl, sl, conf = (np.zeros(max(n_evidence, 1), np.float64) for _ in range(3))
        self.lib.orc_best_alleles(_p(m, _f64p), na, stride, n_evidence, _p(pri, _f64p) if pri is not None else None,
                                  _p(best, _u32p), _p(second, _u32p), _p(bl, _f64p), _p(sl, _f64p), _p(conf, _f64p))
        return best[:n_evidence], second[:n_evidence], bl[:n_evidence], sl[:n_evidence], conf[:n_evidence]

    def region(self, reads: Sequence[dict], haps: Sequence, thr: int = 18, model: int = PCR_CONSERVATIVE,
               fma_mode: int = FMA_AVX512):
        """reads: dicts with bases, qual, mapq and optional ins/del/gcp.  Returns
        (matrix[n_haps, n_reads] with NaN tail, kept_idx, raw_f32[n_reads, n_haps], raw_f64, n_fp64)."""
        n_r, n_h = len(reads), len(haps)
        keep = []
        ra = (_OrcRead * max(n_r, 1))()
        for k, r in enumerate(reads):
            b, q = _u8(r["bases"]), _u8(r["qual"])
            keep += [b, q]
            ra[k].bases, ra[k].qual = _p(b, _u8p), _p(q, _u8p)
            for name, field in (("ins", "ins"), ("del", "del_"), ("gcp", "gcp")):
                v = r.get(name)
                if v is not None:
                    a = _u8(v)
                    keep.append(a)
                    setattr(ra[k], field, _p(a, _u8p))
            ra[k].len, ra[k].mapq = len(b), int(r.get("mapq", 60))
        ha = (_OrcHap * max(n_h, 1))()
        for j, h in enumerate(haps):
            a = _u8(h)
            keep.append(a)
            ha[j].bases, ha[j].len = _p(a, _u8p), len(a)
        out = np.full((n_h, n_r), np.nan, np.float64)
        kept = np.zeros(max(n_r, 1), np.uint32)
        n_kept = C.c_uint32(0)
        raw32 = np.zeros((n_r, n_h), np.float32)
        raw64 = np.full((n_r, n_h), np.nan, np.float64)
        n64 = self.lib.orc_region(ra, n_r, ha, n_h, thr, model, fma_mode, _p(out, _f64p), _p(kept, _u32p), C.byref(n_kept),
                                  _p(raw32, _f32p), _p(raw64, _f64p))
        return out, kept[: n_kept.value].copy(), raw32, raw64, int(n64)


class Ref:
    """The unmodified reference, compiled in place (oracle/_ref/libphmm_ref.so)."""

    def __init__(self, path: str):
        L = self.lib = C.CDLL(path)
        L.ref_has_avx512.restype = C.c_int
        pair = [C.c_int, _u8p, C.c_uint32, _u8p, _u8p, _u8p, _u8p, _u8p, C.c_uint32]
        L.ref_pair_f32.restype, L.ref_pair_f32.argtypes = C.c_float, pair
        L.ref_pair_f64.restype, L.ref_pair_f64.argtypes = C.c_double, pair
        pp = C.POINTER(_u8p)
        L.ref_1xn_f32.restype = C.c_int
        L.ref_1xn_f32.argtypes = [_u8p, C.c_uint32, C.c_uint32, pp, pp, pp, pp, pp, _u32p, _f32p]
        L.ref_tables_f32.restype, L.ref_tables_f32.argtypes = None, [_f32p, _f32p]
        L.ref_find_tandem_repeat_units.restype = C.c_int32
        L.ref_find_tandem_repeat_units.argtypes = [_u8p, C.c_uint32, C.c_uint32]
        L.ref_apply_pcr_error_model.restype = None
        L.ref_apply_pcr_error_model.argtypes = [_u8p, _u8p, C.c_uint32, C.c_int]
        L.ref_region.restype = C.c_int
        L.ref_region.argtypes = [C.c_int, C.c_uint32, _u8p, _u8p, _u64p, _u32p, _u8p, C.c_uint32, _u8p, _u64p, _u32p, C.c_int32,
                                 C.c_int, _f64p, _u32p, _u32p]
        L.ref_regions_flat.restype = C.c_int
        L.ref_regions_flat.argtypes = [C.c_int, C.c_int, C.c_uint32, _u64p, _u64p, _u8p, _u8p, _u64p, _u32p, _u8p, _u8p, _u64p,
                                       _u32p, C.c_int32, C.c_int, _u64p, _f64p, _u32p, _u32p]
        L.ref_bench_regions.restype = C.c_double
        L.ref_bench_regions.argtypes = [C.c_int, C.c_int, C.c_uint32, _u64p, _u64p, _u8p, _u8p, _u64p, _u32p, _u8p, _u8p, _u64p,
                                        _u32p, C.c_int32, C.c_int, _f64p]
        self.avx512 = bool(L.ref_has_avx512())
        L.ref_best_alleles.restype = C.c_int
        L.ref_best_alleles.argtypes = [_f64p, C.c_uint32, C.c_uint32, C.c_uint32, _f64p, C.c_int32, _u32p, _u32p, _f64p, _f64p, _f64p]
        L.ref_marginalize.restype = C.c_int
        L.ref_marginalize.argtypes = [_f64p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, _f64p]
        L.ref_sw_align.restype = C.c_int
        L.ref_sw_align.argtypes = [C.c_int, C.c_int, _u8p, C.c_int, _u8p, C.c_int, _u32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int)]

    def sw_bench(self, haps, reads, n_threads: int, assemble: bool = False):
        """n alignments through the reference's Smith-Waterman on n_threads host threads -> (seconds, checksum)."""
        hb = [_u8(h) for h in haps]
        rb = [_u8(r) for r in reads]
        hl, rl = np.array([len(x) for x in hb], np.uint32), np.array([len(x) for x in rb], np.uint32)
        ho, ro = np.zeros(len(hb), np.uint64), np.zeros(len(rb), np.uint64)
        ho[1:], ro[1:] = np.cumsum(hl[:-1], dtype=np.uint64), np.cumsum(rl[:-1], dtype=np.uint64)
        hcat, rcat = np.concatenate(hb), np.concatenate(rb)
        chk = C.c_longlong(0)
        self.lib.ref_sw_bench.restype = C.c_double
        self.lib.ref_sw_bench.argtypes = [C.c_int, C.c_int, C.c_uint32, _u8p, _u64p, _u32p, _u8p, _u64p, _u32p, C.POINTER(C.c_longlong)]
        secs = self.lib.ref_sw_bench(n_threads, int(assemble), len(hb), _p(hcat, _u8p), _p(ho, _u64p), _p(hl, _u32p), _p(rcat, _u8p),
                                     _p(ro, _u64p), _p(rl, _u32p), C.byref(chk))
        return float(secs), int(chk.value)

    def best_alleles(self, matrix: np.ndarray, n_evidence: int, priorities, reference_index: int = 0):
        """The reference's own AlleleLikelihoods::best_alleles_breaking_ties (al_shim.cpp) on matrix[n_alleles, stride] ->
        (best, second, best_lik, second_lik, confidence) for the first n_evidence columns."""
        m = np.ascontiguousarray(matrix, np.float64)
        na, stride = m.shape
        pri = np.ascontiguousarray(priorities, np.float64)
        best, second = np.zeros(max(n_evidence, 1), np.uint32), np.zeros(max(n_evidence, 1), np.uint32)
        bl, sl, conf = (np.zeros(max(n_evidence, 1), np.float64) for _ in range(3))
        rc = self.lib.ref_best_alleles(_p(m, _f64p), na, stride, n_evidence, _p(pri, _f64p),
                                       reference_index, _p(best, _u32p), _p(second, _u32p), _p(bl, _f64p), _p(sl, _f64p), _p(conf, _f64p))
        if rc != 0:
            raise RuntimeError("ref_best_alleles failed rc=%d" % rc)
        return best[:n_evidence], second[:n_evidence], bl[:n_evidence], sl[:n_evidence], conf[:n_evidence]

    def marginalize(self, matrix: np.ndarray, n_evidence: int, old_to_new, n_new: int) -> np.ndarray:
        """The reference's own AlleleLikelihoods::marginalize (al_shim.cpp) -> [n_new, n_evidence]."""
        m = np.ascontiguousarray(matrix, np.float64)
        mp = np.ascontiguousarray(old_to_new, np.int32)
        out = np.zeros((n_new, max(n_evidence, 1)), np.float64)[:, :n_evidence].copy()
        rc = self.lib.ref_marginalize(_p(m, _f64p), m.shape[0], m.shape[1], n_evidence, mp.ctypes.data_as(C.POINTER(C.c_int32)), n_new,
                                      _p(out, _f64p))
        if rc != 0:
            raise RuntimeError("ref_marginalize failed rc=%d" % rc)
        return out

    def sw_align(self, seq1, seq2, assemble: bool = False, isa: int = -1):
        """The reference's Smith-Waterman (sw_avx.c / sw_avx2.c): seq2 (read) against seq1 (haplotype).
        -> (cigar uint32[], alignment_offset, score), cigar None where the reference gives up (>= 2048 bases)."""
        a, b = _u8(seq1), _u8(seq2)
        cig = np.zeros(4096, np.uint32)
        off, score, bound = C.c_int(0), C.c_int(0), C.c_int(0)
        n = self.lib.ref_sw_align(isa, int(assemble), _p(a, _u8p), len(a), _p(b, _u8p), len(b), _p(cig, _u32p), len(cig), C.byref(off),
                                  C.byref(score), C.byref(bound))
        if n == -2:
            raise RuntimeError("ref_sw_align: bad arguments")
        return (None if n < 0 else cig[:n].copy()), int(off.value), int(score.value)

    def pair_f32(self, isa: int, hap, rs, q, ins, dele, gcp) -> np.float32:
        hap, rs, q, ins, dele, gcp = map(_u8, (hap, rs, q, ins, dele, gcp))
        return np.float32(self.lib.ref_pair_f32(isa, _p(hap, _u8p), len(hap), _p(rs, _u8p), _p(q, _u8p), _p(ins, _u8p),
                                                _p(dele, _u8p), _p(gcp, _u8p), len(rs)))

    def pair_f64(self, isa: int, hap, rs, q, ins, dele, gcp) -> float:
        hap, rs, q, ins, dele, gcp = map(_u8, (hap, rs, q, ins, dele, gcp))
        return float(self.lib.ref_pair_f64(isa, _p(hap, _u8p), len(hap), _p(rs, _u8p), _p(q, _u8p), _p(ins, _u8p),
                                           _p(dele, _u8p), _p(gcp, _u8p), len(rs)))

    def one_by_n_f32(self, hap, reads: Sequence[tuple]) -> np.ndarray:
        """reads: up to 16 tuples (rs, q, ins, del, gcp) -> raw float sums through the 16x1 kernel."""
        hap = _u8(hap)
        n = len(reads)
        cols = [[_u8(r[c]) for r in reads] for c in range(5)]
        arrs = [(_u8p * n)(*[_p(a, _u8p) for a in col]) for col in cols]
        lens = np.array([len(a) for a in cols[0]], np.uint32)
        out = np.zeros(16, np.float32)
        rc = self.lib.ref_1xn_f32(_p(hap, _u8p), len(hap), n, *arrs, _p(lens, _u32p), _p(out, _f32p))
        if rc != 0:
            raise RuntimeError("ref_1xn_f32 needs AVX-512 (rc=%d)" % rc)
        return out[:n].copy()

    def tables_f32(self):
        ph, mm = np.zeros(128, np.float32), np.zeros(8256, np.float32)
        self.lib.ref_tables_f32(_p(ph, _f32p), _p(mm, _f32p))
        return ph, mm

    def find_tandem_repeat_units(self, bases, offset: int) -> int:
        b = _u8(bases)
        return int(self.lib.ref_find_tandem_repeat_units(_p(b, _u8p), len(b), offset))

    def apply_pcr_error_model(self, bases, gap_qual, model: int) -> np.ndarray:
        b, g = _u8(bases), _u8(gap_qual).copy()
        self.lib.ref_apply_pcr_error_model(_p(b, _u8p), _p(g, _u8p), len(b), model)
        return g

    def region(self, reads: Sequence[dict], haps: Sequence, thr: int = 18, model: int = PCR_CONSERVATIVE, path: int = -1):
        """One region through the reference's call_pairhmm (path 0 = old, 1 = new002, -1 = as bound
        on this CPU).  Reads must be sorted by ascending length for path 1."""
        n_r, n_h = len(reads), len(haps)
        rb = [_u8(r["bases"]) for r in reads]
        rq = [_u8(r["qual"]) for r in reads]
        rlen = np.array([len(b) for b in rb], np.uint32)
        roff = np.zeros(n_r, np.uint64)
        if n_r:
            roff[1:] = np.cumsum(rlen[:-1], dtype=np.uint64)
        rbase = np.concatenate(rb) if n_r else np.zeros(1, np.uint8)
        rqual = np.concatenate(rq) if n_r else np.zeros(1, np.uint8)
        rmapq = np.array([int(r.get("mapq", 60)) for r in reads], np.uint8)
        hb = [_u8(h) for h in haps]
        hlen = np.array([len(b) for b in hb], np.uint32)
        hoff = np.zeros(n_h, np.uint64)
        if n_h:
            hoff[1:] = np.cumsum(hlen[:-1], dtype=np.uint64)
        hbase = np.concatenate(hb) if n_h else np.zeros(1, np.uint8)
        out = np.full((n_h, n_r), np.nan, np.float64)
        kept = np.zeros(max(n_r, 1), np.uint32)
        n_kept = C.c_uint32(0)
        rc = self.lib.ref_region(path, n_r, _p(rbase, _u8p), _p(rqual, _u8p), _p(roff, _u64p), _p(rlen, _u32p), _p(rmapq, _u8p),
                                 n_h, _p(hbase, _u8p), _p(hoff, _u64p), _p(hlen, _u32p), thr, model, _p(out, _f64p),
                                 _p(kept, _u32p), C.byref(n_kept))
        if rc != 0:
            raise RuntimeError("ref_region failed rc=%d" % rc)
        return out, kept[: n_kept.value].copy()

    def regions_flat(self, batch, n_threads: int, thr: int = 18, model: int = PCR_CONSERVATIVE, path: int = -1):
        """Every region of a RegionBatch through the reference's call_pairhmm on n_threads threads ->
        (likelihoods flat at batch.lik_off(), kept_idx flat at reg_read_off, n_kept per region)."""
        lik_off = batch.lik_off()
        lik = np.empty(max(int(lik_off[-1]), 1), np.float64)
        kept = np.zeros(max(batch.n_reads, 1), np.uint32)
        n_kept = np.zeros(max(batch.n_regions, 1), np.uint32)
        rc = self.lib.ref_regions_flat(path, n_threads, batch.n_regions, _p(batch.reg_read_off, _u64p), _p(batch.reg_hap_off, _u64p),
                                       _p(batch.read_bases, _u8p), _p(batch.read_qual, _u8p), _p(batch.read_off, _u64p),
                                       _p(batch.read_len, _u32p), _p(batch.read_mapq, _u8p), _p(batch.hap_bases, _u8p),
                                       _p(batch.hap_off, _u64p), _p(batch.hap_len, _u32p), thr, model, _p(lik_off, _u64p),
                                       _p(lik, _f64p), _p(kept, _u32p), _p(n_kept, _u32p))
        if rc != 0:
            raise RuntimeError("ref_regions_flat failed rc=%d" % rc)
        return lik, kept, n_kept

    def bench_regions(self, batch, n_threads: int, thr: int = 18, model: int = PCR_CONSERVATIVE, path: int = -1):
        """batch: a rovaca_b200.workloads.RegionBatch (flat arrays).  Returns (seconds, checksum)."""
        chk = C.c_double(0.0)
        secs = self.lib.ref_bench_regions(path, n_threads, batch.n_regions, _p(batch.reg_read_off, _u64p),
                                          _p(batch.reg_hap_off, _u64p), _p(batch.read_bases, _u8p), _p(batch.read_qual, _u8p),
                                          _p(batch.read_off, _u64p), _p(batch.read_len, _u32p), _p(batch.read_mapq, _u8p),
                                          _p(batch.hap_bases, _u8p), _p(batch.hap_off, _u64p), _p(batch.hap_len, _u32p),
                                          thr, model, C.byref(chk))
        if secs < 0:
            raise RuntimeError("ref_bench_regions failed (%f)" % secs)
        return float(secs), float(chk.value)


class Adapter:
    """rovaca_b200/adapter/call_pairhmm_b200.cpp (the drop-in with the reference's signature) driven through the reference's
    own ReadRecord / Haplotype / RovacaMemoryPool objects (oracle/_ref/libphmm_adapter_test.so).  Needs a B200 to run."""

    EXPORTS = ["adapter_init", "adapter_init_devices", "adapter_shutdown", "adapter_region", "adapter_window", "adapter_last_error"]

    def __init__(self, path: str):
        L = self.lib = C.CDLL(path)
        L.adapter_last_error.restype = C.c_char_p
        L.adapter_init.restype, L.adapter_init.argtypes = C.c_int, [C.c_int]
        L.adapter_init_devices.restype, L.adapter_init_devices.argtypes = C.c_int, [C.POINTER(C.c_int), C.c_int]
        L.adapter_window.restype = C.c_int
        L.adapter_window.argtypes = [C.c_uint32, _u64p, _u64p, _u8p, _u8p, _u64p, _u32p, _u8p, _u8p, _u64p, _u32p, C.c_int32, C.c_int,
                                     _u64p, _f64p, _u32p, _u32p]
        L.adapter_shutdown.restype = None
        L.adapter_region.restype = C.c_int
        L.adapter_region.argtypes = [C.c_uint32, _u8p, _u8p, _u64p, _u32p, _u8p, C.c_uint32, _u8p, _u64p, _u32p, C.c_int32, C.c_int,
                                     _f64p, _u32p, _u32p]

    def init_devices(self, devices: Sequence[int]):
        arr = (C.c_int * max(len(devices), 1))(*devices)
        rc = self.lib.adapter_init_devices(arr, len(devices))
        if rc != 0:
            raise RuntimeError("init_pairhmm_ptr_b200_devices failed (%d): %s" % (rc, self.lib.adapter_last_error().decode()))

    def window(self, batch, thr: int = 18, model: int = PCR_CONSERVATIVE):
        """Every region of a RegionBatch through call_pairhmm_b200_window (submit all, collect in order) ->
        (likelihoods flat at batch.lik_off(), kept_idx flat at reg_read_off, n_kept per region)."""
        lik_off = batch.lik_off()
        lik = np.empty(max(int(lik_off[-1]), 1), np.float64)
        kept = np.zeros(max(batch.n_reads, 1), np.uint32)
        n_kept = np.zeros(max(batch.n_regions, 1), np.uint32)
        rc = self.lib.adapter_window(batch.n_regions, _p(batch.reg_read_off, _u64p), _p(batch.reg_hap_off, _u64p),
                                     _p(batch.read_bases, _u8p), _p(batch.read_qual, _u8p), _p(batch.read_off, _u64p),
                                     _p(batch.read_len, _u32p), _p(batch.read_mapq, _u8p), _p(batch.hap_bases, _u8p),
                                     _p(batch.hap_off, _u64p), _p(batch.hap_len, _u32p), thr, model, _p(lik_off, _u64p),
                                     _p(lik, _f64p), _p(kept, _u32p), _p(n_kept, _u32p))
        if rc != 0:
            raise RuntimeError("adapter_window failed (%d): %s" % (rc, self.lib.adapter_last_error().decode()))
        return lik, kept, n_kept

    def init(self, device: int = -1):
        rc = self.lib.adapter_init(device)
        if rc != 0:
            raise RuntimeError("init_pairhmm_ptr_b200 failed (%d): %s" % (rc, self.lib.adapter_last_error().decode()))

    def shutdown(self):
        self.lib.adapter_shutdown()

    def region(self, reads: Sequence[dict], haps: Sequence, thr: int = 18, model: int = PCR_CONSERVATIVE):
        """One region through the global rovaca::call_pairhmm pointer after init_pairhmm_ptr_b200 re-bound it."""
        n_r, n_h = len(reads), len(haps)
        rb = [_u8(r["bases"]) for r in reads]
        rq = [_u8(r["qual"]) for r in reads]
        rlen = np.array([len(b) for b in rb], np.uint32)
        roff = np.zeros(n_r, np.uint64)
        if n_r:
            roff[1:] = np.cumsum(rlen[:-1], dtype=np.uint64)
        rbase = np.concatenate(rb) if n_r else np.zeros(1, np.uint8)
        rqual = np.concatenate(rq) if n_r else np.zeros(1, np.uint8)
        rmapq = np.array([int(r.get("mapq", 60)) for r in reads], np.uint8)
        hb = [_u8(h) for h in haps]
        hlen = np.array([len(b) for b in hb], np.uint32)
        hoff = np.zeros(n_h, np.uint64)
        if n_h:
            hoff[1:] = np.cumsum(hlen[:-1], dtype=np.uint64)
        hbase = np.concatenate(hb) if n_h else np.zeros(1, np.uint8)
        out = np.full((n_h, n_r), np.nan, np.float64)
        kept = np.zeros(max(n_r, 1), np.uint32)
        n_kept = C.c_uint32(0)
        rc = self.lib.adapter_region(n_r, _p(rbase, _u8p), _p(rqual, _u8p), _p(roff, _u64p), _p(rlen, _u32p), _p(rmapq, _u8p), n_h,
                                     _p(hbase, _u8p), _p(hoff, _u64p), _p(hlen, _u32p), thr, model, _p(out, _f64p), _p(kept, _u32p),
                                     C.byref(n_kept))
        if rc != 0:
            raise RuntimeError("adapter_region failed rc=%d: %s" % (rc, self.lib.adapter_last_error().decode()))
        return out, kept[: n_kept.value].copy()


def adapter() -> Optional[Adapter]:
    """The adapter test library, or None when oracle/_ref/libphmm_adapter_test.so is not there (needs /root/reference to build)."""
    path = os.path.join(_HERE, "_ref", "libphmm_adapter_test.so")
    if not os.path.exists(path):
        return None
    try:
        return Adapter(path)
    except OSError:
        return None


_oracle: Optional[Oracle] = None
_ref: Optional[Ref] = None
_ref_tried = False


def oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def ref() -> Optional[Ref]:
    """The compiled reference, or None when oracle/_ref/libphmm_ref.so is not there (or cannot load)."""
    global _ref, _ref_tried
    if not _ref_tried:
        _ref_tried = True
        path = os.path.join(_HERE, "_ref", "libphmm_ref.so")
        if os.path.exists(path):
            try:
                _ref = Ref(path)
            except OSError:
                _ref = None
    return _ref
