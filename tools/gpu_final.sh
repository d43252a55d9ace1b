#!/bin/bash
# round-end evidence on one GPU: the whole GPU suite, the default bench line, the synchronous-caller table (+ the lone caller's
# launch list), then the ncu passes of tools/gpu_profile.sh
TAG=${1:-r02final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference arm rc=$?"
bash tools/gpu_callers.sh ${TAG}_callers > /dev/null 2>&1; cut -c1-230 gpurun_out/${TAG}_callers.txt
PHMM_B200_CALLER_RUNS=0 tools/caller_bench --threads 1 --regions 4096 2>&1 | cut -c1-230
bash tools/gpu_profile.sh ${TAG} > gpurun_out/${TAG}_profile.log 2>&1; grep "exit" gpurun_out/${TAG}_profile.log
tools/caller_bench --threads 1 --regions 40 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_lone_launches.csv tools/caller_bench --threads 1 --regions 8 > gpurun_out/${TAG}_lone_ncu.log 2>&1
tail -6 gpurun_out/${TAG}_lone_launches.csv | cut -d, -f5,13-
