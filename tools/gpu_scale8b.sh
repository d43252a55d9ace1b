#!/bin/bash
# the shorter multi-GPU evidence run (under gpurun --gpus N): multi-device tests, bench.py at 1 and N ranks, a 1 M-region sharded
# stream at N ranks, ONE process driving all N GPUs.  Usage: bash tools/gpu_scale8b.sh <tag> <N>
TAG=${1:-r02_scale}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpus.csv; nproc >> gpurun_out/${TAG}_gpus.csv
timeout 900 python -m pytest tests -m gpu -x -q -k "several_devices or torchrun or window_and_devices or marginalisation" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
: > gpurun_out/${TAG}_bench.jsonl
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-workloads >> gpurun_out/${TAG}_bench.jsonl 2>> gpurun_out/${TAG}_bench.err; echo "bench n=1 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 5 --warmup 3 2>> gpurun_out/${TAG}_bench.err | grep '^{' >> gpurun_out/${TAG}_bench.jsonl; echo "bench n=$N exit $?"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $N --steps 2 --warmup 3 --regions 1024 --sustain 0 --stream-regions 1048576 2>> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_stream1M.json; echo "1M stream exit $?"
timeout 900 python tools/multi_gpu_stream.py 1048576 0.5 all > gpurun_out/${TAG}_one_process.json 2>> gpurun_out/${TAG}_bench.err; echo "one process exit $?"
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_bench.jsonl"):
    d = json.loads(ln); s = d.get("sharded_stream") or {}
    print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "| sharded stream e2e", s.get("e2e_gcups"), "single", s.get("single_gpu_e2e_gcups"), "identical", s.get("identical_to_single_gpu"))
d = json.loads(open("gpurun_out/${TAG}_stream1M.json").read()); print("1M:", json.dumps(d["sharded_stream"]))
print(open("gpurun_out/${TAG}_one_process.json").read())
PY
