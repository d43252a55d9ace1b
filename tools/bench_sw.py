#!/usr/bin/env python
"""Throughput of the batched Smith-Waterman (phmm_sw_align) on the realignment workload of one cfg2 step: every read of 4096
regions (64 x 150-bp reads) against one of its region's 300-bp haplotypes.  Device time (fill + trace kernels), end to end
(host buffers in, CIGARs out), and the reference's own sw_avx.c on the host cores beside it.   python tools/bench_sw.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                      # noqa: E402
from rovaca_b200 import PairHmmEngine, workloads as wl   # noqa: E402
from oracle import bindings as ob                       # noqa: E402

n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = wl.cfg2(n_regions, seed=2)
rng = np.random.default_rng(1)
pairs = []
for g in range(b.n_regions):
    reads, haps = b.region(g)
    pick = rng.integers(0, len(haps), len(reads))
    pairs += [(haps[int(p)], r["bases"]) for p, r in zip(pick, reads)]
cells = sum(len(h) * len(r) for h, r in pairs)
eng = PairHmmEngine()
import ctypes as C
from rovaca_b200.pairhmm import _SwParams, SW_REALIGN
prep = eng.sw_prepare(pairs, cigar_cap=16)       # a 150-bp read against its own haplotype needs a handful of elements
prm = _SwParams(*SW_REALIGN)
for _ in range(2):
    eng._check(eng.L.phmm_sw_align(eng.h, prep["pairs"], prep["n"], C.byref(prm), prep["results"]))
t0 = time.perf_counter()
eng._check(eng.L.phmm_sw_align(eng.h, prep["pairs"], prep["n"], C.byref(prm), prep["results"]))      # the C call alone
e2e = time.perf_counter() - t0
out = eng.sw_run(prep)
ms, dev_cells = eng.sw_last_timing()
assert dev_cells == cells and all(o[3] == 0 for o in out)
line = {"alignments": len(pairs), "cells": cells, "device_ms": round(ms, 3), "device_gcups": round(cells / ms / 1e6, 1),
        "device_alignments_per_s": round(len(pairs) / (ms * 1e-3)), "e2e_ms": round(e2e * 1e3, 2), "e2e_gcups": round(cells / e2e / 1e9, 1),
        "e2e_is": "one phmm_sw_align call: host gather, H2D, fill + trace kernels, D2H of CIGARs, scatter into the caller's buffers; sub-batches of 32768 alignments pipelined on two streams",
        "device_is": "union of the sub-batches' kernel spans (includes the gaps the host leaves; one un-pipelined batch of this size ran 18.8 ms = 627 GCUPS)"}
ref = ob.ref()
if ref is not None:
    cores = len(os.sched_getaffinity(0))
    n_cpu = min(len(pairs), 4096 * cores)
    secs, _ = ref.sw_bench([p[0] for p in pairs[:n_cpu]], [p[1] for p in pairs[:n_cpu]], cores)
    cpu_cells = sum(len(h) * len(r) for h, r in pairs[:n_cpu])
    line["cpu_reference"] = {"gcups": round(cpu_cells / secs / 1e9, 2), "alignments_per_s": round(n_cpu / secs), "cores": cores,
                             "kind": "reference sw_avx.c (avx512)" if ref.avx512 else "reference sw_avx2.c", "sample": f"{n_cpu} alignments"}
    k = rng.integers(0, n_cpu, 64)
    for i in k:
        w = ref.sw_align(pairs[int(i)][0], pairs[int(i)][1])
        assert np.array_equal(w[0], out[int(i)][0]) and w[1:] == out[int(i)][1:3]
    line["spot_check_vs_reference"] = "64 alignments identical"
import json
print(json.dumps(line))
eng.close()
