#!/bin/bash
# scratch experiments on the GPU box (output: gpurun_out/<tag>.txt)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-exp}.txt
: > $OUT
echo "## chunk length (PHMM_B200_STREAM_COLS) against shared-memory-limited occupancy, final round-2 kernels" >> $OUT
for spec in "cfg5 16384 0.0" "cfg5 16384 0.5" "cfg5 32768 1.0" "cfg3 96 0.0" "cfg4 192 0.0"; do
  for cols in 2400 2200 1600 1000; do
    PHMM_B200_STREAM_COLS=$cols python tools/resident_bench.py $spec 3 >> $OUT 2>&1
  done
done
cat $OUT | cut -c1-200
