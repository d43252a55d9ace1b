#!/usr/bin/env python
"""Throughput of the engine on every BASELINE.json shape (cfg2..cfg5), one GPU: device-resident GCUPS, end-to-end GCUPS
through phmm_run_flat, share of FP64 / per-pair-kernel pairs.  bench.py measures the headline (cfg2) only; this table is
how the other shapes are watched (profiles/*_workloads.txt).   python tools/bench_workloads.py [--quick]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                     # noqa: E402
from rovaca_b200 import PairHmmEngine, workloads as wl  # noqa: E402
from rovaca_b200.pairhmm import FlatResult             # noqa: E402

quick = "--quick" in sys.argv
cta_warps = int(sys.argv[sys.argv.index("--cta-warps") + 1]) if "--cta-warps" in sys.argv else 0
n_workers = int(sys.argv[sys.argv.index("--workers") + 1]) if "--workers" in sys.argv else 0
fp64_always = "--fp64-always" in sys.argv           # recompute underflowing pairs in FP64 even when nobody reads the values
specs = [
    ("cfg2 150x300, 8 haps x 64 reads", lambda: wl.cfg2(1024 if quick else 4096, seed=2)),
    ("cfg3 high depth 500 reads x 32 haps, mixed, 5% junk", lambda: wl.cfg3(4 if quick else 24, seed=3)),
    ("cfg4 long 250x1000, 16 haps x 64 reads", lambda: wl.cfg4(8 if quick else 48, seed=4, mix=False)),
    ("cfg4 mix 250x1000 + 150x300", lambda: wl.cfg4(16 if quick else 96, seed=4, mix=True)),
    ("cfg5 stream, no post-trim shapes", lambda: wl.cfg5(256 if quick else 2048, seed=5, post_trim_fraction=0.0)),
    ("cfg5 stream, 50% post-trim shapes", lambda: wl.cfg5(256 if quick else 2048, seed=5, post_trim_fraction=0.5)),
    ("cfg5 stream, 100% post-trim shapes", lambda: wl.cfg5(512 if quick else 4096, seed=5, post_trim_fraction=1.0)),
    ("uniform 100x250, 6 haps x 40 reads", lambda: wl.uniform_regions(1024 if quick else 8192, 40, 6, 100, 250, 7)),
    ("uniform 151x400, 12 haps x 50 reads", lambda: wl.uniform_regions(512 if quick else 2048, 50, 12, 151, 400, 8)),
]
eng = PairHmmEngine(cta_warps=cta_warps, n_workers=n_workers, fp64_mode=1 if fp64_always else 0)
print(f'# cta_warps={cta_warps or "default"} n_workers={n_workers or "default"} fp64={"always" if fp64_always else "when observable"}')
print(f"{'workload':54s} {'Gcells':>8s} {'pairs':>9s} {'dev ms':>8s} {'fwd ms':>8s} {'dev GCUPS':>10s} {'fwd frac':>8s} {'e2e GCUPS':>10s} {'fp64':>7s} {'slow':>6s}")
for name, make in specs:
    b = make()
    cells = b.cells()
    res = eng.upload(b)
    for _ in range(2):
        res.run()
    t = [res.run() for _ in range(3)]
    dev = sum(x[0] for x in t) / 3
    fwd = sum(x[1] for x in t) / 3
    res.free()
    out = FlatResult(b)
    for _ in range(3):                                  # the first calls size the pinned / device arenas of every worker
        eng.run_flat(b, out=out)
    s0 = eng.stats()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.run_flat(b, out=out)
    e2e = (time.perf_counter() - t0) / 3
    s1 = eng.stats()
    frac = cells * 8 / (fwd * 1e-3) / (148 * 128 * 1.965e9)
    print(f"{name:54s} {cells / 1e9:8.2f} {b.pairs():9d} {dev:8.3f} {fwd:8.3f} {cells / dev / 1e6:10.1f} {frac:8.3f} {cells / e2e / 1e9:10.1f} "
          f"{(s1['fp64_pairs'] - s0['fp64_pairs']) // 3:7d} {(s1['slow_pairs'] - s0['slow_pairs']) // 3:6d}", flush=True)
eng.close()
