#!/bin/bash
# smoke + GPU parity tests + a short bench (no ncu).  Usage under gpurun: bash tools/gpu_check.sh <tag> [bench args]
TAG=${1:-chk}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.csv; nproc >> gpurun_out/${TAG}_gpu.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
