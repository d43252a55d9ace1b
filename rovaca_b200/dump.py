"""Dump / replay formats for the PairHMM path (SURVEY.md §8c "dump format to adopt", §8f item 2).

Three formats, so that batches captured wherever the full reference *can* be built can be replayed here:

* **pair text, 7-line records** — exactly what the reference's ``debug_tools_print_temp`` prints for one kernel call
  (``pairhmm_internal.cpp:493-503``): ``"<rslen> \\t <haplen>"``, haplotype, read, then base/insertion/deletion/gcp
  qualities as ASCII+33.  These are *prepared* inputs (already quality-capped and PCR-adjusted).
* **pair text, one line** — the GKL/GATK PairHMM test-vector line ``hap read qual ins del gcp [expected]`` (ASCII+33
  qualities, optional expected value: a log10 likelihood or ``0x........`` raw float bits).
* **region binary** (``PHMMREG1``) — what ``call_pairhmm`` receives for one region, *before* read preparation: written by
  the C++ adapter when ``ROVACA_B200_DUMP`` names a file (``rovaca_b200/adapter/call_pairhmm_b200.cpp``), little endian::

      "PHMMREG1" u32 n_reads  u32 n_haps  i32 min_quality_threshold  i32 pcr_model
      n_reads x { u32 len  u8 mapq  u8 bases[len]  u8 qual[len] }     n_haps x { u32 len  u8 bases[len] }

Nothing here computes likelihoods: replay goes through ``PairHmmEngine`` (the CUDA library).
"""
from __future__ import annotations

import struct
import sys
from typing import BinaryIO, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .workloads import RegionBatch

MAGIC = b"PHMMREG1"


def _q(s: str) -> np.ndarray:
    return (np.frombuffer(s.encode("latin-1"), np.uint8).astype(np.int16) - 33).astype(np.uint8)


def _qs(a) -> str:
    return (np.asarray(a, np.uint8).astype(np.int16) + 33).astype(np.uint8).tobytes().decode("latin-1")


def _b(s) -> np.ndarray:
    return np.frombuffer(s.encode("latin-1") if isinstance(s, str) else bytes(s), np.uint8)


# ---- prepared pairs ---------------------------------------------------------------------------------------------------

def parse_pairs(text: str) -> List[dict]:
    """Either text format (auto-detected per record) -> dicts with hap, rs, q, ins, dele, gcp (uint8 arrays) and, when the
    line carried one, ``expected`` (float log10 likelihood) or ``expected_bits`` (raw float32 bits)."""
    lines = [ln.rstrip("\r\n") for ln in text.splitlines()]
    out, i = [], 0
    while i < len(lines):
        ln = lines[i]
        if not ln.strip():
            i += 1
            continue
        head = ln.split()
        if len(head) == 2 and all(t.isdigit() for t in head):                 # "<rslen> \t <haplen>" record
            rslen, haplen = int(head[0]), int(head[1])
            hap, rs, q, ins, dele, gcp = lines[i + 1:i + 7]
            if len(hap) != haplen or any(len(x) != rslen for x in (rs, q, ins, dele, gcp)):
                raise ValueError(f"pair record at line {i + 1}: lengths disagree with the header")
            out.append(dict(hap=_b(hap), rs=_b(rs), q=_q(q), ins=_q(ins), dele=_q(dele), gcp=_q(gcp)))
            i += 7
            continue
        f = ln.split()
        if len(f) not in (6, 7):
            raise ValueError(f"line {i + 1}: expected 'hap read qual ins del gcp [expected]'")
        hap, rs, q, ins, dele, gcp = f[:6]
        if any(len(x) != len(rs) for x in (q, ins, dele, gcp)):
            raise ValueError(f"line {i + 1}: quality strings differ in length from the read")
        rec = dict(hap=_b(hap), rs=_b(rs), q=_q(q), ins=_q(ins), dele=_q(dele), gcp=_q(gcp))
        if len(f) == 7:
            if f[6].lower().startswith("0x"):
                rec["expected_bits"] = int(f[6], 16)
            else:
                rec["expected"] = float(f[6])
        out.append(rec)
        i += 1
    return out


def format_pairs(pairs: Iterable[dict], style: str = "line", expected_bits: Optional[Sequence[int]] = None) -> str:
    """style 'line' (GKL test-vector lines, optionally with raw float bits) or 'record' (the reference's 7-line dump)."""
    rows = []
    for k, p in enumerate(pairs):
        hap, rs = bytes(p["hap"]).decode("latin-1"), bytes(p["rs"]).decode("latin-1")
        q, ins, dele, gcp = (_qs(p[x]) for x in ("q", "ins", "dele", "gcp"))
        if style == "record":
            rows += [f"{len(rs)} \t {len(hap)}", hap, rs, q, ins, dele, gcp]
        else:
            tail = f" 0x{int(expected_bits[k]):08x}" if expected_bits is not None else ""
            rows.append(f"{hap} {rs} {q} {ins} {dele} {gcp}{tail}")
    return "\n".join(rows) + "\n"


def pairs_to_batch(pairs: Sequence[dict]) -> RegionBatch:
    """Every prepared pair as a 1x1 region with explicit gap qualities; replay with threshold 0, PCR model NONE (the values
    are already prepared) and mapq 255 so that nothing is touched again."""
    regs = [([dict(bases=p["rs"], qual=p["q"], ins=p["ins"], gcp=p["gcp"], mapq=255, **{"del": p["dele"]})], [p["hap"]]) for p in pairs]
    return RegionBatch.from_regions(regs, name=f"{len(pairs)} dumped pairs")


def replay_pairs(engine, pairs: Sequence[dict]) -> Tuple[np.ndarray, np.ndarray]:
    """-> (raw float32 forward sums, log10 likelihoods as the reference reports them) for prepared pairs."""
    from .pairhmm import PCR_NONE
    batch = pairs_to_batch(pairs)
    res = engine.run_flat(batch, 0, PCR_NONE, raw=True)
    raw = np.array([res.raw32(g)[0, 0] for g in range(batch.n_regions)], np.float32)
    raw64 = np.array([res.raw64(g)[0, 0] for g in range(batch.n_regions)], np.float64)
    return raw, np.array([_log10_likelihood(r, d) for r, d in zip(raw, raw64)], np.float64)


_libm = None


def _log10_likelihood(raw: np.float32, raw64: float) -> float:
    """pairhmm_internal.cpp:578-584 with the host's own libm (log10f / log10), as the reference and the engine do: numpy's
    float32 log10 is a different implementation and differs in the last bit."""
    global _libm
    if _libm is None:
        import ctypes
        import ctypes.util
        _libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _libm.log10f.restype, _libm.log10f.argtypes = ctypes.c_float, [ctypes.c_float]
        _libm.log10.restype, _libm.log10.argtypes = ctypes.c_double, [ctypes.c_double]
    if raw < np.float32(1e-28):
        return float(_libm.log10(float(raw64))) - 307.050595577260822 if raw64 > 0 else float("-inf")
    return float(np.float32(_libm.log10f(float(raw))) - np.float32(36.1236000061))


# ---- whole regions ----------------------------------------------------------------------------------------------------

def write_region(f: BinaryIO, reads: Sequence[dict], haps: Sequence, min_quality_threshold: int = 18, pcr_model: int = 3) -> None:
    f.write(MAGIC + struct.pack("<IIii", len(reads), len(haps), min_quality_threshold, pcr_model))
    for r in reads:
        b, q = np.asarray(r["bases"], np.uint8), np.asarray(r["qual"], np.uint8)
        f.write(struct.pack("<IB", len(b), int(r.get("mapq", 60))) + b.tobytes() + q.tobytes())
    for h in haps:
        h = _b(h) if isinstance(h, (str, bytes, bytearray)) else np.asarray(h, np.uint8)
        f.write(struct.pack("<I", len(h)) + h.tobytes())


def read_regions(f: BinaryIO) -> List[Tuple[List[dict], List[np.ndarray], int, int]]:
    """All regions of a dump file -> [(reads, haps, min_quality_threshold, pcr_model)]."""
    data, at, out = f.read(), 0, []
    while at < len(data):
        if data[at:at + 8] != MAGIC:
            raise ValueError(f"offset {at}: not a PHMMREG1 record")
        n_r, n_h, thr, pcr = struct.unpack_from("<IIii", data, at + 8)
        at += 24
        reads, haps = [], []
        for _ in range(n_r):
            n, mapq = struct.unpack_from("<IB", data, at)
            at += 5
            reads.append(dict(bases=np.frombuffer(data, np.uint8, n, at).copy(), qual=np.frombuffer(data, np.uint8, n, at + n).copy(), mapq=mapq))
            at += 2 * n
        for _ in range(n_h):
            (n,) = struct.unpack_from("<I", data, at)
            haps.append(np.frombuffer(data, np.uint8, n, at + 4).copy())
            at += 4 + n
        out.append((reads, haps, thr, pcr))
    return out


def main(argv: Sequence[str]) -> int:
    """python -m rovaca_b200.dump pairs <file>    replay a text pair dump, print raw bits + log10 (+ check if expected given)
       python -m rovaca_b200.dump regions <file>  replay a PHMMREG1 dump, print one matrix summary line per region"""
    if len(argv) != 3 or argv[1] not in ("pairs", "regions"):
        print(main.__doc__)
        return 2
    from .pairhmm import PairHmmEngine
    eng = PairHmmEngine()
    bad = 0
    if argv[1] == "pairs":
        pairs = parse_pairs(open(argv[2], encoding="latin-1").read())
        raw, lk = replay_pairs(eng, pairs)
        for p, r, v in zip(pairs, raw, lk):
            note = ""
            if "expected_bits" in p:
                ok = int(r.view(np.uint32)) == p["expected_bits"]
                note, bad = (" ok" if ok else " MISMATCH"), bad + (not ok)
            elif "expected" in p:
                ok = abs(v - p["expected"]) <= 1e-4 * max(1.0, abs(p["expected"]))      # the north-star tolerance for log10 values
                note, bad = (" ok" if ok else " MISMATCH"), bad + (not ok)
            print(f"0x{int(r.view(np.uint32)):08x} {v:.9f}{note}")
    else:
        for g, (reads, haps, thr, pcr) in enumerate(read_regions(open(argv[2], "rb"))):
            m, kept = eng.call_pairhmm(haps, reads, thr, pcr)
            print(f"region {g}: {len(reads)} reads x {len(haps)} haplotypes, {len(kept)} kept, checksum {np.nansum(m):.9f}")
    eng.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
