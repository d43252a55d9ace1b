"""Python door to ``libphmm_b200.so`` — mirrors the reference's PairHMM interface.

Reference seam: ``rovaca::call_pairhmm(hs, rs, min_quality_threshold, pcr_option, pool)``
(/root/reference/src/haplotypecaller/pairhmm/pairhmm_engine.h:14-15), bound once by
``init_pairhmm_ptr`` (pairhmm_engine.cpp:21-53).  Here: ``PairHmmEngine()`` is the binding,
``PairHmmEngine.call_pairhmm(haps, reads, min_quality_threshold, pcr_option)`` the call; it returns
the ``[hap][read]`` log10 likelihood matrix with the reference's NaN tail plus the indices of the reads
the call left in ``rs``.  All arithmetic happens in the CUDA library; nothing here computes
likelihoods and nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

from .workloads import RegionBatch, _as_u8

PCR_NONE, PCR_HOSTILE, PCR_AGGRESSIVE, PCR_CONSERVATIVE = 0, 1, 2, 3      # common/enum.h:8
FMA_AVX512, FMA_NONE = 1, 2
FLAT_PINNED = 1                                                          # phmm_flat_batch.flags
FP64_WHEN_OBSERVABLE, FP64_ALWAYS = 0, 1                                     # phmm_config.fp64_mode
EPILOGUE_AUTO, EPILOGUE_HOST, EPILOGUE_DEVICE = 0, 1, 2                      # phmm_config.epilogue

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


class PhmmError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("fma_mode", C.c_int32), ("n_workers", C.c_uint32), ("batch_cells", C.c_uint64),
                ("linger_us", C.c_uint32), ("cta_warps", C.c_uint32), ("pipeline_depth", C.c_uint32), ("lanes_per_read", C.c_uint32), ("fp64_mode", C.c_uint32), ("epilogue", C.c_uint32), ("rsv", C.c_uint32 * 4)]


class _Read(C.Structure):
    _fields_ = [("bases", _u8p), ("qual", _u8p), ("ins", _u8p), ("del_", _u8p), ("gcp", _u8p), ("len", C.c_uint32),
                ("mapq", C.c_uint8)]


class _Hap(C.Structure):
    _fields_ = [("bases", _u8p), ("len", C.c_uint32), ("priority", C.c_double)]


class _Region(C.Structure):
    _fields_ = [("reads", C.POINTER(_Read)), ("n_reads", C.c_uint32), ("haps", C.POINTER(_Hap)), ("n_haps", C.c_uint32),
                ("min_quality_threshold", C.c_int32), ("pcr_model", C.c_int32)]


class _Result(C.Structure):
    _fields_ = [("likelihoods", _f64p), ("kept_idx", _u32p), ("n_kept", C.c_uint32), ("raw_f32", _f32p), ("raw_f64", _f64p),
                ("n_fp64", C.c_uint32), ("best_hap", _u32p), ("second_hap", _u32p)]


class _FlatBatch(C.Structure):
    _fields_ = [("n_regions", C.c_uint32), ("reg_read_off", _u64p), ("reg_hap_off", _u64p), ("read_len", _u32p),
                ("read_mapq", _u8p), ("read_off", _u64p), ("read_bases", _u8p), ("read_qual", _u8p), ("read_ins", _u8p),
                ("read_del", _u8p), ("read_gcp", _u8p), ("hap_len", _u32p), ("hap_off", _u64p), ("hap_bases", _u8p),
                ("min_quality_threshold", C.c_int32), ("pcr_model", C.c_int32), ("flags", C.c_uint32), ("hap_priority", _f64p)]


class _FlatResult(C.Structure):
    _fields_ = [("likelihoods", _f64p), ("lik_off", _u64p), ("kept_idx", _u32p), ("n_kept", _u32p), ("raw_f32", _f32p),
                ("raw_f64", _f64p), ("n_fp64", C.c_uint64), ("best_hap", _u32p), ("second_hap", _u32p)]


class _SwParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32)]


class _SwPair(C.Structure):
    _fields_ = [("ref", _u8p), ("ref_len", C.c_uint32), ("read", _u8p), ("read_len", C.c_uint32)]


class _SwResult(C.Structure):
    _fields_ = [("cigar", _u32p), ("cigar_cap", C.c_uint32), ("n_cigar", C.c_uint32), ("alignment_offset", C.c_int32),
                ("score", C.c_int32), ("status", C.c_int32)]


SW_REALIGN = (10, -15, -30, -5)        # sw_parameters_genotype, common/base/lib/sw_avx/sw_avx.c:14
SW_ASSEMBLE = (200, -150, -260, -11)   # sw_parameters_assemble, :13


class _MargSite(C.Structure):
    _fields_ = [("region", C.c_uint32), ("n_new", C.c_uint32), ("old_to_new", C.POINTER(C.c_int32))]


class _Stats(C.Structure):
    _fields_ = [("regions", C.c_uint64), ("pairs", C.c_uint64), ("cells", C.c_uint64), ("fp64_pairs", C.c_uint64),
                ("slow_pairs", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("batches", C.c_uint64),
                ("gpu_ms", C.c_double), ("kernel_launches", C.c_uint64), ("host_plan_ms", C.c_double), ("host_epilogue_ms", C.c_double),
                ("device_epilogue_batches", C.c_uint64)]


EXPORTS = ["phmm_resident_marginalize", "phmm_create_multi", "phmm_device_count", "phmm_device_ordinal", "phmm_get_device_stats", "phmm_resident_upload_on", "phmm_marginalize", "phmm_sw_align", "phmm_sw_last_timing", "phmm_create", "phmm_destroy", "phmm_call_region", "phmm_submit", "phmm_wait", "phmm_flush", "phmm_run_flat",
           "phmm_resident_upload", "phmm_resident_run", "phmm_resident_download", "phmm_resident_free",
           "phmm_resident_launches", "phmm_get_stats", "phmm_kernel_info", "phmm_ubench", "phmm_last_error", "phmm_version"]


def library_path() -> str:
    return os.environ.get("PHMM_B200_LIB") or os.path.join(_HERE, "libphmm_b200.so")      # the override is for kernel experiments


def build_library(force: bool = False) -> str:
    """Compile the CUDA kernels (sm_100a) and the host engine in-tree.  nvcc cross-compiles without a GPU."""
    if force:
        subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc"), "clean"], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)
    return library_path()


_lib = None


def load_library():
    """dlopen libphmm_b200.so; raises PhmmError (never falls back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise PhmmError(f"{path} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                        "rovaca_b200 has no CPU fallback")
    L = C.CDLL(path)
    L.phmm_last_error.restype = C.c_char_p
    L.phmm_version.restype = C.c_char_p
    L.phmm_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
    L.phmm_create_multi.argtypes = [C.POINTER(_Config), C.POINTER(C.c_int32), C.c_uint32, C.POINTER(C.c_void_p)]
    L.phmm_device_count.argtypes = [C.c_void_p]
    L.phmm_device_count.restype = C.c_uint32
    L.phmm_device_ordinal.argtypes = [C.c_void_p, C.c_uint32]
    L.phmm_get_device_stats.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(_Stats)]
    L.phmm_resident_upload_on.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(_FlatBatch), C.POINTER(C.c_void_p)]
    L.phmm_destroy.argtypes = [C.c_void_p]
    L.phmm_destroy.restype = None
    L.phmm_call_region.argtypes = [C.c_void_p, C.POINTER(_Region), C.POINTER(_Result)]
    L.phmm_submit.argtypes = [C.c_void_p, C.POINTER(_Region), C.POINTER(_Result), _u64p]
    L.phmm_wait.argtypes = [C.c_void_p, C.c_uint64]
    L.phmm_flush.argtypes = [C.c_void_p]
    L.phmm_run_flat.argtypes = [C.c_void_p, C.POINTER(_FlatBatch), C.POINTER(_FlatResult)]
    L.phmm_resident_upload.argtypes = [C.c_void_p, C.POINTER(_FlatBatch), C.POINTER(C.c_void_p)]
    L.phmm_resident_run.argtypes = [C.c_void_p, C.c_void_p, _f32p, _f32p]
    L.phmm_resident_download.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_FlatResult)]
    L.phmm_resident_marginalize.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_MargSite), C.c_uint32, _u64p, _f64p]
    L.phmm_resident_free.argtypes = [C.c_void_p, C.c_void_p]
    L.phmm_resident_free.restype = None
    L.phmm_resident_launches.argtypes = [C.c_void_p]
    L.phmm_resident_launches.restype = C.c_uint32
    L.phmm_get_stats.argtypes = [C.c_void_p, C.POINTER(_Stats)]
    L.phmm_kernel_info.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.phmm_ubench.argtypes = [C.c_void_p, C.c_int, C.c_uint32, _f32p, _f64p]
    L.phmm_sw_align.argtypes = [C.c_void_p, C.POINTER(_SwPair), C.c_uint32, C.POINTER(_SwParams), C.POINTER(_SwResult)]
    L.phmm_sw_last_timing.argtypes = [C.c_void_p, _f32p, _u64p]
    L.phmm_debug_tables.argtypes = [_f32p, _f32p, _f64p, _f64p]
    L.phmm_debug_tables.restype = None
    L.phmm_debug_log10f_check.restype = C.c_uint64
    L.phmm_debug_log10f_check.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_int)]
    L.phmm_marginalize.argtypes = [_f64p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, _f64p, C.c_uint32]
    L.phmm_debug_partition.argtypes = [C.POINTER(_FlatBatch), C.c_uint32, C.c_uint64, _u32p, _u64p, C.c_uint32, _u32p]
    L.phmm_debug_host_log10.argtypes = [_f32p, C.c_uint32, C.c_int, _f64p]
    L.phmm_debug_device_log10.argtypes = [C.c_void_p, _f32p, C.c_uint32, _f64p, C.POINTER(C.c_int)]
    L.phmm_debug_logf_check.restype = C.c_uint64
    L.phmm_debug_logf_check.argtypes = [C.c_uint32]
    L.phmm_debug_plan.argtypes = [C.POINTER(_FlatBatch), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _u64p]
    L.phmm_debug_epilogue.argtypes = [C.POINTER(_FlatBatch), _f32p, _f64p, C.POINTER(_FlatResult), C.c_int, C.c_uint32, C.POINTER(C.c_double)]
    _lib = L
    return L


def host_tables():
    """The libm tables the engine uploads (core/context.h) — host-only, for table-level tests."""
    L = load_library()
    ph, mm = np.zeros(128, np.float32), np.zeros(8256, np.float32)
    phd, mmd = np.zeros(128, np.float64), np.zeros(8256, np.float64)
    L.phmm_debug_tables(ph.ctypes.data_as(_f32p), mm.ctypes.data_as(_f32p), phd.ctypes.data_as(_f64p), mmd.ctypes.data_as(_f64p))
    return ph, mm, phd, mmd


PLAN_AUDIT_FIELDS = ("pairs", "wavefront_pairs", "per_pair_kernel_pairs", "covered_twice", "missed", "warps", "read_slots",
                     "read_slots_filled", "reads_in_too_small_a_bin", "launches", "chunks", "lane_steps", "wavefront_cells",
                     "shape_violations")


def plan_audit(batch: RegionBatch, fma_mode: int = FMA_AVX512, cta_warps: int = 0, busy_batches: int = 1,
               min_quality_threshold: int = 18, pcr_option: int = PCR_CONSERVATIVE, lanes_per_read: int = 0) -> dict:
    """Host only (no GPU): run the engine's planner on ``batch`` and audit its output — every pair computed exactly once,
    every read in a bin that holds it, CTA shapes respected — plus the numbers that tell how well the lanes are used."""
    L = load_library()
    fb = _flat_c(batch, min_quality_threshold, pcr_option)
    out = np.zeros(16, np.uint64)
    rc = L.phmm_debug_plan(C.byref(fb), fma_mode, cta_warps, busy_batches, lanes_per_read, out.ctypes.data_as(_u64p))
    if rc != 0:
        raise PhmmError(f"phmm_debug_plan failed ({rc}): {L.phmm_last_error().decode()}")
    d = {k: int(v) for k, v in zip(PLAN_AUDIT_FIELDS, out)}
    d["slot_fill"] = d["read_slots_filled"] / max(d["read_slots"], 1)             # lanes that hold a read / lanes of the warps launched
    d["lane_utilisation"] = d["wavefront_cells"] / max(2 * d["lane_steps"], 1)      # two haplotype streams per lane
    return d


def flat_partition(batch: RegionBatch, n_workers: int = 0, batch_cells: int = 0):
    """Host only: how ``run_flat`` would cut ``batch`` into sub-batches -> (bounds[k] = (first, end), work[k])."""
    L = load_library()
    fb = _flat_c(batch, 18, PCR_CONSERVATIVE)
    cap = max(batch.n_regions, 1)
    bounds, work, n = np.zeros(2 * cap, np.uint32), np.zeros(cap, np.uint64), C.c_uint32(0)
    rc = L.phmm_debug_partition(C.byref(fb), n_workers, int(batch_cells), _ptr(bounds, _u32p), _ptr(work, _u64p), cap, C.byref(n))
    if rc != 0:
        raise PhmmError(f"phmm_debug_partition failed ({rc}): {L.phmm_last_error().decode()}")
    return bounds[: 2 * n.value].reshape(-1, 2), work[: n.value]


def marginalize(matrix: np.ndarray, n_evidence: int, old_to_new, n_new: int) -> np.ndarray:
    """``AlleleLikelihoods::marginalize`` on a matrix ``call_pairhmm`` returned: [n_new, n_evidence], each entry the largest
    likelihood over the old alleles mapped to it (-inf when none is)."""
    L = load_library()
    m = np.ascontiguousarray(matrix, np.float64)
    mp = np.ascontiguousarray(old_to_new, np.int32)
    out = np.zeros((n_new, max(n_evidence, 1)), np.float64)
    rc = L.phmm_marginalize(_ptr(m, _f64p), m.shape[0], m.shape[1] if m.ndim == 2 else 0, n_evidence,
                            mp.ctypes.data_as(C.POINTER(C.c_int32)), n_new, _ptr(out, _f64p), out.shape[1])
    if rc != 0:
        raise PhmmError(f"phmm_marginalize failed ({rc}): {L.phmm_last_error().decode()}")
    return out[:, :n_evidence]


def host_log10(x: np.ndarray, mode: int = 0) -> np.ndarray:
    """(double)(log10f(x) - 36.1236f) on the host in ``mode`` (0 = libm itself), -inf under 1e-28f."""
    L = load_library()
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size, np.float64)
    rc = L.phmm_debug_host_log10(_ptr(x, _f32p), x.size, mode, _ptr(out, _f64p))
    if rc != 0:
        raise PhmmError(f"phmm_debug_host_log10 failed ({rc})")
    return out


def host_epilogue(batch: RegionBatch, raw_f32: np.ndarray, raw_f64: Optional[np.ndarray] = None, mode: int = -1, reps: int = 1,
                  min_quality_threshold: int = 18, pcr_option: int = PCR_CONSERVATIVE, best: bool = False):
    """Host only (no GPU): the engine's epilogue — log10 conversion, normalisation, filter, transpose — on raw sums laid out
    like ``FlatResult.raw_f32``.  ``mode``: how log10f is computed (0 libm, 1 around libm's logf, 2 AVX-512, -1 as an engine
    would pick).  -> (FlatResult, mean milliseconds per run)."""
    L = load_library()
    fb = _flat_c(batch, min_quality_threshold, pcr_option)
    res = FlatResult(batch, best=best)
    f32 = np.ascontiguousarray(raw_f32, np.float32).reshape(-1)
    f64 = None if raw_f64 is None else np.ascontiguousarray(raw_f64, np.float64).reshape(-1)
    ms = C.c_double(0)
    rc = L.phmm_debug_epilogue(C.byref(fb), _ptr(f32, _f32p), _ptr(f64, _f64p), C.byref(res.c), mode, reps, C.byref(ms))
    if rc != 0:
        raise PhmmError(f"phmm_debug_epilogue failed ({rc}): {L.phmm_last_error().decode()}")
    res.n_fp64 = int(res.c.n_fp64)
    return res, float(ms.value)


def _ptr(a: Optional[np.ndarray], t):
    return a.ctypes.data_as(t) if a is not None else t()


class FlatResult:
    """Host buffers one flat call fills; ``matrix(g)`` gives region g's [n_haps][n_reads] view."""

    def __init__(self, batch: RegionBatch, raw: bool = False, best: bool = False):
        self.batch = batch
        self.lik_off = batch.lik_off()
        n = int(self.lik_off[-1])
        self.likelihoods = np.empty(max(n, 1), np.float64)
        self.kept_idx = np.zeros(max(batch.n_reads, 1), np.uint32)
        self.n_kept = np.zeros(max(batch.n_regions, 1), np.uint32)
        self.raw_f32 = np.zeros(max(n, 1), np.float32) if raw else None
        self.raw_f64 = np.zeros(max(n, 1), np.float64) if raw else None
        self.n_fp64 = 0
        self.best_hap = np.zeros(max(batch.n_reads, 1), np.uint32) if best else None
        self.second_hap = np.zeros(max(batch.n_reads, 1), np.uint32) if best else None
        self.c = _FlatResult(_ptr(self.likelihoods, _f64p), _ptr(self.lik_off, _u64p), _ptr(self.kept_idx, _u32p),
                             _ptr(self.n_kept, _u32p), _ptr(self.raw_f32, _f32p), _ptr(self.raw_f64, _f64p), 0,
                             _ptr(self.best_hap, _u32p), _ptr(self.second_hap, _u32p))

    def _dims(self, g: int) -> Tuple[int, int]:
        b = self.batch
        return int(b.reg_hap_off[g + 1] - b.reg_hap_off[g]), int(b.reg_read_off[g + 1] - b.reg_read_off[g])

    def matrix(self, g: int) -> np.ndarray:
        nh, nr = self._dims(g)
        o = int(self.lik_off[g])
        return self.likelihoods[o:o + nh * nr].reshape(nh, nr)

    def kept(self, g: int) -> np.ndarray:
        r0 = int(self.batch.reg_read_off[g])
        return self.kept_idx[r0:r0 + int(self.n_kept[g])]

    def best(self, g: int) -> Tuple[np.ndarray, np.ndarray]:
        """(best, second-best) haplotype index of every surviving read of region g."""
        r0 = int(self.batch.reg_read_off[g])
        n = int(self.n_kept[g])
        return self.best_hap[r0:r0 + n], self.second_hap[r0:r0 + n]

    def raw32(self, g: int) -> np.ndarray:
        nh, nr = self._dims(g)
        o = int(self.lik_off[g])
        return self.raw_f32[o:o + nh * nr].reshape(nr, nh)

    def raw64(self, g: int) -> np.ndarray:
        nh, nr = self._dims(g)
        o = int(self.lik_off[g])
        return self.raw_f64[o:o + nh * nr].reshape(nr, nh)


def _flat_c(batch: RegionBatch, thr: int, pcr: int) -> _FlatBatch:
    return _FlatBatch(batch.n_regions, _ptr(batch.reg_read_off, _u64p), _ptr(batch.reg_hap_off, _u64p), _ptr(batch.read_len, _u32p),
                      _ptr(batch.read_mapq, _u8p), _ptr(batch.read_off, _u64p), _ptr(batch.read_bases, _u8p),
                      _ptr(batch.read_qual, _u8p), _ptr(batch.read_ins, _u8p), _ptr(batch.read_del, _u8p), _ptr(batch.read_gcp, _u8p),
                      _ptr(batch.hap_len, _u32p), _ptr(batch.hap_off, _u64p), _ptr(batch.hap_bases, _u8p), thr, pcr,
                      FLAT_PINNED if getattr(batch, "pinned", False) else 0, _ptr(getattr(batch, "hap_priority", None), _f64p))


class Resident:
    """A batch living in HBM: ``run()`` executes the device path only (kernel-timing leg of the bench)."""

    def __init__(self, eng: "PairHmmEngine", batch: RegionBatch, thr: int, pcr: int, device_index: int = 0):
        self.eng, self.batch = eng, batch
        self.h = C.c_void_p()
        fb = _flat_c(batch, thr, pcr)
        eng._check(eng.L.phmm_resident_upload_on(eng.h, device_index, C.byref(fb), C.byref(self.h)))
        self.launches = int(eng.L.phmm_resident_launches(self.h))

    def run(self) -> Tuple[float, float]:
        """-> (all kernels ms, forward kernels ms) from CUDA events on the batch's stream."""
        a, b = C.c_float(0), C.c_float(0)
        self.eng._check(self.eng.L.phmm_resident_run(self.eng.h, self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def download(self, raw: bool = False, best: bool = False) -> FlatResult:
        res = FlatResult(self.batch, raw, best)
        self.eng._check(self.eng.L.phmm_resident_download(self.eng.h, self.h, C.byref(res.c)))
        res.n_fp64 = int(res.c.n_fp64)
        return res

    def marginalize(self, sites: Sequence[Tuple[int, Sequence[int], int]]):
        """``AlleleLikelihoods::marginalize`` for many variant sites on the matrices this batch left on the device
        (``phmm_resident_marginalize``).  sites: (region index, old_to_new per haplotype of that region, n_new).
        -> per site the [n_new, n_reads of the region] array: first n_kept columns valid, NaN tail like the matrix."""
        b = self.batch
        n = len(sites)
        arr = (_MargSite * max(n, 1))()
        keep, off, at = [], np.zeros(max(n, 1), np.uint64), 0
        for k, (g, mp, n_new) in enumerate(sites):
            m = np.ascontiguousarray(mp, np.int32)
            keep.append(m)
            arr[k] = _MargSite(int(g), int(n_new), m.ctypes.data_as(C.POINTER(C.c_int32)))
            off[k] = at
            at += int(n_new) * int(b.reg_read_off[g + 1] - b.reg_read_off[g])
        out = np.zeros(max(at, 1), np.float64)
        self.eng._check(self.eng.L.phmm_resident_marginalize(self.eng.h, self.h, arr, n, _ptr(off, _u64p), _ptr(out, _f64p)))
        res = []
        for k, (g, mp, n_new) in enumerate(sites):
            nr = int(b.reg_read_off[g + 1] - b.reg_read_off[g])
            res.append(out[int(off[k]):int(off[k]) + int(n_new) * nr].reshape(int(n_new), nr))
        return res

    def free(self):
        if self.h:
            self.eng.L.phmm_resident_free(self.eng.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PairHmmEngine:
    """``init_pairhmm_ptr`` + ``call_pairhmm`` of the reference, on one B200 or on several behind one engine."""

    def __init__(self, device: int = -1, fma_mode: int = FMA_AVX512, n_workers: int = 0, batch_cells: int = 0,
                 linger_us: int = 0, cta_warps: int = 0, lanes_per_read: int = 0, fp64_mode: int = 0, epilogue: int = 0,
                 devices: Optional[Sequence[int]] = None, pipeline_depth: int = 0):
        """``devices``: CUDA ordinals this one engine drives (``phmm_create_multi``; ``"all"`` or ``[]`` = every visible
        device, an ordinal may repeat); without it the engine owns the single ``device``."""
        self.L = load_library()
        cfg = _Config(device, fma_mode, n_workers, batch_cells, linger_us, cta_warps, pipeline_depth, lanes_per_read, fp64_mode, epilogue)
        self.h = C.c_void_p()
        if devices is None:
            rc = self.L.phmm_create(C.byref(cfg), C.byref(self.h))
        else:
            ords = [] if isinstance(devices, str) else [int(d) for d in devices]
            arr = (C.c_int32 * max(len(ords), 1))(*ords)
            rc = self.L.phmm_create_multi(C.byref(cfg), arr, len(ords), C.byref(self.h))
        if rc != 0:
            raise PhmmError(f"phmm_create failed ({rc}): {self.L.phmm_last_error().decode()}")

    @property
    def n_devices(self) -> int:
        return int(self.L.phmm_device_count(self.h))

    def device_ordinals(self):
        return [int(self.L.phmm_device_ordinal(self.h, k)) for k in range(self.n_devices)]

    def device_stats(self, index: int) -> dict:
        s = _Stats()
        self._check(self.L.phmm_get_device_stats(self.h, index, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _Stats._fields_}

    def _check(self, rc: int):
        if rc != 0:
            raise PhmmError(f"libphmm_b200 error {rc}: {self.L.phmm_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.phmm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the reference's call ---------------------------------------------------------------------
    def call_pairhmm(self, haps: Sequence, reads: Sequence[dict], min_quality_threshold: int = 18,
                     pcr_option: int = PCR_CONSERVATIVE, raw: bool = False, priorities: Optional[Sequence[float]] = None):
        """One region.  reads: dicts with bases, qual, mapq and optionally ins/del/gcp (bytes / uint8 arrays).
        Returns (matrix[n_haps, n_reads] with NaN tail, kept_idx) and, with raw=True, also (raw_f32, raw_f64, n_fp64); with
        ``priorities`` (one per haplotype, the tie-break of best_alleles_breaking_ties) also (best_hap, second_hap) of the
        surviving reads."""
        n_r, n_h = len(reads), len(haps)
        keep = []
        ra = (_Read * max(n_r, 1))()
        for k, r in enumerate(reads):
            b, q = _as_u8(r["bases"]), _as_u8(r["qual"])
            if len(b) != len(q):
                raise ValueError("bases and qual differ in length")
            keep += [b, q]
            ra[k].bases, ra[k].qual = b.ctypes.data_as(_u8p), q.ctypes.data_as(_u8p)
            for name, field in (("ins", "ins"), ("del", "del_"), ("gcp", "gcp")):
                v = r.get(name)
                if v is not None:
                    a = _as_u8(v)
                    keep.append(a)
                    setattr(ra[k], field, a.ctypes.data_as(_u8p))
            ra[k].len, ra[k].mapq = len(b), int(r.get("mapq", 60))
        ha = (_Hap * max(n_h, 1))()
        for j, h in enumerate(haps):
            a = _as_u8(h)
            keep.append(a)
            ha[j].bases, ha[j].len = a.ctypes.data_as(_u8p), len(a)
            ha[j].priority = float(priorities[j]) if priorities is not None else 0.0
        best = np.zeros(max(n_r, 1), np.uint32) if priorities is not None else None
        second = np.zeros(max(n_r, 1), np.uint32) if priorities is not None else None
        out = np.full((n_h, n_r), np.nan, np.float64)
        kept = np.zeros(max(n_r, 1), np.uint32)
        raw32 = np.zeros((n_r, n_h), np.float32) if raw else None
        raw64 = np.zeros((n_r, n_h), np.float64) if raw else None
        reg = _Region(ra, n_r, ha, n_h, min_quality_threshold, pcr_option)
        res = _Result(_ptr(out, _f64p), _ptr(kept, _u32p), 0, _ptr(raw32, _f32p), _ptr(raw64, _f64p), 0, _ptr(best, _u32p),
                      _ptr(second, _u32p))
        self._check(self.L.phmm_call_region(self.h, C.byref(reg), C.byref(res)))
        if n_r == 0 or n_h == 0:       # pairhmm_internal.cpp:508-510 returns an empty matrix
            out = np.zeros((0, 0), np.float64)
        ret = (out, kept[: res.n_kept].copy())
        if raw:
            ret += (raw32, raw64, int(res.n_fp64))
        if priorities is not None:
            ret += (best[: res.n_kept].copy(), second[: res.n_kept].copy())
        return ret

    # -- the step after the path: batched Smith-Waterman ------------------------------------------
    def sw_prepare(self, pairs: Sequence[Tuple], cigar_cap: int = 64):
        """Marshal (haplotype, read) pairs into the C arrays phmm_sw_align takes (Python-side cost, kept out of timings)."""
        n = len(pairs)
        keep, arr = [], (_SwPair * max(n, 1))()
        res = (_SwResult * max(n, 1))()
        cig = np.zeros((max(n, 1), cigar_cap), np.uint32)
        for k, (ref, read) in enumerate(pairs):
            a, b = _as_u8(ref), _as_u8(read)
            keep += [a, b]
            arr[k] = _SwPair(a.ctypes.data_as(_u8p), len(a), b.ctypes.data_as(_u8p), len(b))
            res[k].cigar = cig[k].ctypes.data_as(_u32p)
            res[k].cigar_cap = cigar_cap
        return dict(n=n, pairs=arr, results=res, cigars=cig, keep=keep)

    def sw_run(self, prepared: dict, params=SW_REALIGN):
        """One phmm_sw_align call on prepared arrays -> per pair (cigar, alignment_offset, score, status)."""
        prm = _SwParams(*params)
        self._check(self.L.phmm_sw_align(self.h, prepared["pairs"], prepared["n"], C.byref(prm), prepared["results"]))
        res, cig = prepared["results"], prepared["cigars"]
        return [(cig[k, : res[k].n_cigar].copy(), int(res[k].alignment_offset), int(res[k].score), int(res[k].status))
                for k in range(prepared["n"])]

    def sw_align(self, pairs: Sequence[Tuple], params=SW_REALIGN, cigar_cap: int = 64):
        """pairs: (haplotype, read) byte sequences.  One device pass; per pair -> (cigar uint32[], alignment_offset, score,
        status), exactly what the reference's sw->avx_function(sw) leaves in lib_sw_avx (alignment_utils.cpp:574-579)."""
        return self.sw_run(self.sw_prepare(pairs, cigar_cap), params)

    def sw_last_timing(self) -> Tuple[float, int]:
        ms, cells = C.c_float(0), C.c_uint64(0)
        self._check(self.L.phmm_sw_last_timing(self.h, C.byref(ms), C.byref(cells)))
        return float(ms.value), int(cells.value)

    # -- batch forms ------------------------------------------------------------------------------
    def run_flat(self, batch: RegionBatch, min_quality_threshold: int = 18, pcr_option: int = PCR_CONSERVATIVE,
                 raw: bool = False, out: Optional[FlatResult] = None, best: bool = False) -> FlatResult:
        """Host buffers in, host buffers out: packing, H2D, kernels, D2H and the epilogue, pipelined.  ``best``: also the best /
        second-best haplotype of every surviving read (ties broken by ``batch.hap_priority`` when the batch has one)."""
        res = out or FlatResult(batch, raw, best)
        fb = _flat_c(batch, min_quality_threshold, pcr_option)
        self._check(self.L.phmm_run_flat(self.h, C.byref(fb), C.byref(res.c)))
        res.n_fp64 = int(res.c.n_fp64)
        return res

    def upload(self, batch: RegionBatch, min_quality_threshold: int = 18, pcr_option: int = PCR_CONSERVATIVE,
               device_index: int = 0) -> Resident:
        return Resident(self, batch, min_quality_threshold, pcr_option, device_index)

    def stats(self) -> dict:
        s = _Stats()
        self._check(self.L.phmm_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _Stats._fields_}

    def device_log10(self, x: np.ndarray) -> Tuple[np.ndarray, bool]:
        """(double)(log10f(x) - 36.1236f) as the device epilogue computes it, and whether this engine's epilogue runs there."""
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(x.size, np.float64)
        use = C.c_int(0)
        self._check(self.L.phmm_debug_device_log10(self.h, _ptr(x, _f32p), x.size, _ptr(out, _f64p), C.byref(use)))
        return out, bool(use.value)

    def ubench(self, mode: int, iters: int = 4096) -> Tuple[float, float]:
        """-> (ms, FMA lane-ops per second) of the FP32 issue-rate micro-benchmark."""
        ms, rate = C.c_float(0), C.c_double(0)
        self._check(self.L.phmm_ubench(self.h, mode, iters, C.byref(ms), C.byref(rate)))
        return float(ms.value), float(rate.value)

    def kernel_info(self, read_len: int, fma_mode: int = FMA_AVX512) -> dict:
        """Static facts of the forward-kernel instantiation serving reads of ``read_len`` bases."""
        regs, loc, lanes, rpt = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self.L.phmm_kernel_info(read_len, fma_mode, C.byref(regs), C.byref(loc), C.byref(lanes), C.byref(rpt)))
        return {"regs": regs.value, "spill_bytes": loc.value, "lanes_per_read": lanes.value, "rows_per_lane": rpt.value}
