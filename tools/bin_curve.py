#!/usr/bin/env python
"""Forward-kernel efficiency per length bin: uniform regions (64 reads x 8 haplotypes x 300 columns) at read lengths that
fill a bin exactly, device-resident, one B200.   python tools/bin_curve.py [lengths...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rovaca_b200 import PairHmmEngine, workloads as wl  # noqa: E402

lens = [int(x) for x in sys.argv[1:]] or [32, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 136, 144, 151, 152, 200, 250, 300]
eng = PairHmmEngine()
print(f"{'R':>4s} {'bin':>6s} {'regs':>5s} {'Gcells':>8s} {'fwd ms':>8s} {'GCUPS':>8s} {'frac':>6s}")
for R in lens:
    n = max(256, int(60e9 / (64 * 8 * R * 300)) // 256 * 256)
    b = wl.uniform_regions(n, 64, 8, R, 300, 7)
    res = eng.upload(b)
    for _ in range(2):
        res.run()
    t = min(res.run()[1] for _ in range(3))
    res.free()
    ki = eng.kernel_info(R)
    cells = b.cells()
    print(f"{R:4d} {ki['lanes_per_read']:3d}x{ki['rows_per_lane']:<2d} {ki['regs']:5d} {cells / 1e9:8.2f} {t:8.3f} {cells / t / 1e6:8.0f} "
          f"{cells * 8 / (t * 1e-3) / (148 * 128 * 1.965e9):6.3f}", flush=True)
eng.close()
