#!/bin/bash
# ncu evidence (one gpurun call, one GPU): plain run, launch list (device time per launch) and one --set full capture of the
# forward kernel on the headline workload; --set full captures of the forward kernels on a cfg5 stream (12 rows per lane) and of
# the fast FP64 recompute kernel.  Usage under gpurun:  bash tools/gpu_profile.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads --sustain 0 --stream-regions 0"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.csv
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/${TAG}_plain2.json 2>> gpurun_out/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:phmm_fwd -s 3 -c 2 -f -o gpurun_out/${TAG}_fwd $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "forward (cfg2) capture exit $?"
CMD2="python tools/resident_bench.py cfg5 8192 0.5 1"
$CMD2 > gpurun_out/${TAG}_cfg5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:phmm_fwd -s 4 -c 4 -f -o gpurun_out/${TAG}_fwd_cfg5 $CMD2 > gpurun_out/${TAG}_ncu3.log 2>&1
echo "forward (cfg5) capture exit $?"
export FP64_ALWAYS=1
CMD3="python tools/resident_bench.py cfg3 48 0 1"
$CMD3 > gpurun_out/${TAG}_fp64_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:phmm_pair64 -s 1 -c 1 -f -o gpurun_out/${TAG}_fp64 $CMD3 > gpurun_out/${TAG}_ncu4.log 2>&1
echo "fp64 capture exit $?"
unset FP64_ALWAYS
cat gpurun_out/${TAG}_plain.json | cut -c1-300
ls -la gpurun_out/ | grep ${TAG}
