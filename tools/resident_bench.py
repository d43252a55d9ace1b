#!/usr/bin/env python
"""Device-resident timing of one workload (kernel experiments):
   python tools/resident_bench.py <kind> <regions> [post_trim] [reps]      kind: cfg2..cfg5, or uniform:<reads>x<haps>x<R>x<H>
Prints device ms, forward ms, forward roofline fraction, launches.  Under ncu: reps 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rovaca_b200 import PairHmmEngine, workloads as wl
kind, n = sys.argv[1], int(sys.argv[2])
pt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
if kind.startswith("uniform:"):
    nr, nh, R, H = (int(x) for x in kind.split(":")[1].split("x"))
    b = wl.uniform_regions(n, nr, nh, R, H, seed=5).pin()
else:
    b = wl.stream(kind, n, seed=5, post_trim_fraction=pt).pin()
eng = PairHmmEngine(fp64_mode=1 if os.environ.get('FP64_ALWAYS') else 0)
r = eng.upload(b)
for _ in range(2):
    r.run()
t = [r.run() for _ in range(reps)]
dev, fwd = min(x[0] for x in t), min(x[1] for x in t)
cells = b.cells()
env = " ".join(f"{k[10:]}={v}" for k, v in os.environ.items() if k.startswith("PHMM_B200_")) + (" FP64_ALWAYS" if os.environ.get("FP64_ALWAYS") else "")
print(f"{kind} {n} pt={pt} [{env}]: {cells/1e9:.2f} Gcells dev {dev:.3f} ms fwd {fwd:.3f} ms  {cells/dev/1e6:.0f} GCUPS  fwd_frac {cells*8/(fwd*1e-3)/(148*128*1.965e9):.4f}  launches {r.launches}", flush=True)
r.free(); eng.close()
