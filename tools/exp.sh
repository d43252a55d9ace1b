#!/bin/bash
# scratch: ncu of the forward launches of a post-trim-only cfg5 step
mkdir -p gpurun_out
python tools/resident_bench.py cfg5 16384 1.0 1 > gpurun_out/${1:-exp}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:phmm_fwd -s 4 -c 4 -f -o gpurun_out/${1:-exp}_fwd_trim100 python tools/resident_bench.py cfg5 16384 1.0 1 > gpurun_out/${1:-exp}_ncu.log 2>&1
echo "ncu exit $?"; cat gpurun_out/${1:-exp}_plain.log
python - <<PY
import sys; sys.path.insert(0, '.')
from rovaca_b200 import workloads as wl, pairhmm as ph
b = wl.stream("cfg5", 16384, seed=5, post_trim_fraction=1.0)
a = ph.plan_audit(b, busy_batches=1)
print({k: a[k] for k in ("warps", "lane_steps", "lane_utilisation", "slot_fill", "launches", "chunks", "cells")})
PY
