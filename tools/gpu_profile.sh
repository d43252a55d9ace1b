#!/bin/bash
# ncu evidence for one bench command: plain run, launch list (device time per launch), one --set full capture of
# the forward wavefront kernel.  Usage under gpurun:  bash tools/gpu_profile.sh <tag> [regions]
TAG=${1:-r01}
REG=${2:-1024}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --regions $REG --no-cpu-baseline"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.csv
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/${TAG}_plain2.json 2>> gpurun_out/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:phmm_fwd -s 3 -c 2 -f -o gpurun_out/${TAG}_fwd $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:phmm_epilogue -s 3 -c 1 -f -o gpurun_out/${TAG}_epi $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "epilogue capture exit $?"
cat gpurun_out/${TAG}_plain.json
ls -la gpurun_out/
