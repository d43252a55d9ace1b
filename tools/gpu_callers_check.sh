#!/bin/bash
# last check: smoke, the whole GPU suite, the callers table
TAG=${1:-callers_check}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
bash tools/gpu_callers.sh ${TAG}_callers > /dev/null 2>&1
PHMM_B200_CALLER_RUNS=0 tools/caller_bench --threads 1 --regions 4096 >> gpurun_out/${TAG}_callers.txt 2>&1
python - <<PY
import json
for l in open("gpurun_out/${TAG}_callers.txt"):
    d = json.loads(l); print(d["mode"], d["threads"], d["window"], d["gcups"], d["mean_call_latency_us"], d["regions_per_batch"], d["per_batch_us"])
PY
