#!/usr/bin/env python
"""ONE process, ONE engine, every GPU of the box (phmm_create_multi): a cfg5 region stream through phmm_run_flat, compared
region by region (digests) with a single-device engine's results; per-device shares of the work.
   python tools/multi_gpu_stream.py [regions] [post_trim] [devices]      devices: "all" (default) or e.g. 0,1,2,3
This is the reference's own shape — one multi-threaded process that binds the path once (haplotypecaller.cpp:49,97) — as
opposed to bench.py's one rank per GPU."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                       # noqa: E402
from rovaca_b200 import PairHmmEngine, streams          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
pt = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
devs = sys.argv[3] if len(sys.argv) > 3 else "all"
devices = "all" if devs == "all" else [int(x) for x in devs.split(",")]
spec = streams.StreamSpec("cfg5", n, 5, pt)
one = PairHmmEngine(device=0)
t0 = time.perf_counter()
solo = streams.run_stream(one, spec)
one.close()
many = PairHmmEngine(devices=devices)
multi = streams.run_stream(many, spec)
per_dev = [many.device_stats(k) for k in range(many.n_devices)]
ords = many.device_ordinals()
many.close()
rec = {"stream": spec.describe(), "cells": multi.cells, "process": "one process, one engine (phmm_create_multi)", "devices": ords,
       "single_device_e2e_gcups": round(solo.cells / solo.seconds / 1e9, 1), "multi_device_e2e_gcups": round(multi.cells / multi.seconds / 1e9, 1),
       "speedup": round(solo.seconds / multi.seconds, 2), "e2e_efficiency": round(solo.seconds / multi.seconds / len(ords), 3),
       "identical_to_single_device": bool(np.array_equal(solo.digests, multi.digests)),
       "regions_differing": int(np.count_nonzero(solo.digests != multi.digests)),
       "share_of_cells_per_device": [round(d["cells"] / max(sum(x["cells"] for x in per_dev), 1), 4) for d in per_dev],
       "batches_per_device": [int(d["batches"]) for d in per_dev], "host_cores": len(os.sched_getaffinity(0)),
       "wall_seconds_total": round(time.perf_counter() - t0, 1)}
print(json.dumps(rec), flush=True)
