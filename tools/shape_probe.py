#!/usr/bin/env python
"""One workload shape, looked at closely: device-resident step time, the engine's own phase timers for the end-to-end
call, and (under ncu --metrics gpu__time_duration.sum) the per-kernel launch list of exactly three resident steps.
   python tools/shape_probe.py cfg3|cfg5|cfg5t [n_regions] [--resident-only] [--epilogue host|device]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rovaca_b200 import PairHmmEngine, workloads as wl  # noqa: E402
from rovaca_b200.pairhmm import FlatResult             # noqa: E402

shape = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
make = {"cfg2": lambda: wl.cfg2(n or 4096, seed=2),
        "cfg3": lambda: wl.cfg3(n or 24, seed=3),
        "cfg5": lambda: wl.cfg5(n or 2048, seed=5, post_trim_fraction=0.0),
        "cfg5t": lambda: wl.cfg5(n or 4096, seed=5, post_trim_fraction=1.0)}[shape]
b = make()
cells = b.cells()
epi = {"host": 1, "device": 2}.get(sys.argv[sys.argv.index("--epilogue") + 1], 0) if "--epilogue" in sys.argv else 0
def _opt(name, default=0):
    return int(float(sys.argv[sys.argv.index(name) + 1])) if name in sys.argv else default


eng = PairHmmEngine(epilogue=epi, n_workers=_opt("--workers"), batch_cells=_opt("--batch-cells"))
res = eng.upload(b)
t = [res.run() for _ in range(3)]
print(f"{shape}: {cells / 1e9:.2f} Gcells, {b.pairs()} pairs; resident step {t[-1][0]:.3f} ms (forward {t[-1][1]:.3f} ms) "
      f"= {cells / t[-1][0] / 1e6:.0f} GCUPS")
res.free()
if "--resident-only" not in sys.argv:
    out = FlatResult(b)
    for _ in range(3):
        eng.run_flat(b, out=out)
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.run_flat(b, out=out)
    e2e = (time.perf_counter() - t0) / 5
    s1 = eng.stats()
    d = {k: s1[k] - s0[k] for k in s1 if isinstance(s1[k], (int, float))}
    nb = max(1, d.get("batches", 1))
    print(f"  e2e {e2e * 1e3:.3f} ms = {cells / e2e / 1e9:.0f} GCUPS; per call: batches {nb / 5:.1f}, "
          + ", ".join(f"{k} {d[k] / 5:.3f}" for k in d if k.endswith("_ms"))
          + f", fp64 pairs {d.get('fp64_pairs', 0) // 5}, slow pairs {d.get('slow_pairs', 0) // 5}")
eng.close()
