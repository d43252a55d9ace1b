#!/usr/bin/env python
"""Forward-kernel efficiency against the length of a warp's sweep: uniform regions of 64 reads x 151 bp, n_haps haplotypes
of H columns, device-resident, one B200; the planner's lane utilisation beside it, so that what is left is the cost of a
warp's prologue and epilogue.   python tools/chunk_curve.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rovaca_b200 import PairHmmEngine, workloads as wl  # noqa: E402
from rovaca_b200.pairhmm import plan_audit             # noqa: E402

eng = PairHmmEngine()
print(f"{'R':>4s} {'H':>5s} {'haps':>4s} {'Gcells':>8s} {'fwd ms':>8s} {'GCUPS':>8s} {'frac':>6s} {'util':>6s} {'frac/util':>9s}")
for R, H, nh in [(151, 300, 1), (151, 300, 2), (151, 300, 3), (151, 300, 4), (151, 300, 6), (151, 300, 8), (151, 300, 12), (151, 300, 16),
                 (151, 1000, 2), (151, 1000, 8), (100, 100, 2), (100, 100, 4), (100, 100, 8), (100, 100, 16), (60, 60, 8), (40, 41, 8)]:
    n = max(256, int(40e9 / (64 * nh * R * H)) // 256 * 256)
    n = min(n, 32768)
    b = wl.uniform_regions(n, 64, nh, R, H, 7)
    util = plan_audit(b)["lane_utilisation"]
    res = eng.upload(b)
    for _ in range(2):
        res.run()
    t = min(res.run()[1] for _ in range(3))
    res.free()
    cells = b.cells()
    frac = cells * 8 / (t * 1e-3) / (148 * 128 * 1.965e9)
    print(f"{R:4d} {H:5d} {nh:4d} {cells / 1e9:8.2f} {t:8.3f} {cells / t / 1e6:8.0f} {frac:6.3f} {util:6.3f} {frac / util:9.3f}", flush=True)
eng.close()
