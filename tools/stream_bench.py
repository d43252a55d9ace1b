#!/usr/bin/env python
"""End-to-end rate of a cfg5 region stream through phmm_run_flat with the host side's own accounting (planning, epilogue /
hand-over, device time), e.g. under `taskset -c 0-3` with PHMM_B200_WORKERS=4 to see what one rank of an 8-GPU box has.
   python tools/stream_bench.py [regions] [post_trim] [piece]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rovaca_b200 import PairHmmEngine, streams
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
pt = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
piece = int(sys.argv[3]) if len(sys.argv) > 3 else streams.PIECE
eng = PairHmmEngine()
spec = streams.StreamSpec("cfg5", n, 5, pt)
streams.run_stream(eng, streams.StreamSpec("cfg5", min(n, piece), 5, pt), warm=True, piece_regions=piece)     # same piece size: the arenas are at size
s0 = eng.stats()
t0 = time.perf_counter()
run = streams.run_stream(eng, spec, warm=False, piece_regions=piece)
wall = time.perf_counter() - t0
s1 = eng.stats()
d = {k: s1[k] - s0[k] for k in s1}
print(json.dumps({"regions": n, "post_trim": pt, "cores": len(os.sched_getaffinity(0)), "workers_env": os.environ.get("PHMM_B200_WORKERS"),
                  "e2e_gcups": round(run.cells / run.seconds / 1e9, 1), "run_flat_seconds": round(run.seconds, 4), "wall_incl_generation": round(wall, 3),
                  "device_busy_s": round(d["gpu_ms"] / 1e3, 4), "host_plan_cpu_s": round(d["host_plan_ms"] / 1e3, 4),
                  "host_epilogue_cpu_s": round(d["host_epilogue_ms"] / 1e3, 4), "batches": int(d["batches"]),
                  "plan_us_per_region": round(d["host_plan_ms"] * 1e3 / n, 2), "epilogue_us_per_region": round(d["host_epilogue_ms"] * 1e3 / n, 2)}), flush=True)
eng.close()
