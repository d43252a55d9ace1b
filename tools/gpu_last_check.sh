#!/bin/bash
# last check of the round on one GPU: smoke, the whole GPU suite, the default bench line
TAG=${1:-last_check}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
for w in d["workloads"]:
    print(w["workload"], w["regions"], w["gcells"], w["device_gcups"], w["fwd_frac"], w["e2e_gcups"], w["e2e_over_device"], w["launches_per_step"], w["parity_sample_ok"])
print(d["sharded_stream"]["e2e_gcups"], d["sustained"]["gcups"], d["cpu_baseline"]["value"])
PY
