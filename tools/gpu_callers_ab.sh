#!/bin/bash
# synchronous callers: the path where the batch runs on the caller's own thread (and small batches take the short device path)
# against the hand-over to a worker it replaces; the GPU tests that walk the ticket path; one ncu launch list of a lone caller
TAG=${1:-r02_callers_ab}
mkdir -p gpurun_out
g++ -O2 -std=c++17 -Iinclude tools/caller_bench.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -lpthread -o tools/caller_bench || exit 1
OUT=gpurun_out/${TAG}.txt
: > $OUT
run() {  # label, env...
  echo "## $1" >> $OUT; shift
  for t in 1 2 4 8 16 32 64; do
    env "$@" PHMM_B200_PHASES=1 tools/caller_bench --threads $t --regions $((4096 / t > 64 ? 4096 / t : 64)) >> $OUT 2>&1
  done
}
run "the caller's thread runs the batch" X=1
run "hand-over to a worker" PHMM_B200_CALLER_RUNS=0
echo "## windows (16 threads)" >> $OUT
for w in 4 16; do tools/caller_bench --threads 16 --regions 512 --window $w >> $OUT 2>&1; done
cat $OUT
true
true
tools/caller_bench --threads 1 --regions 40 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_lone_launches.csv tools/caller_bench --threads 1 --regions 8 > gpurun_out/${TAG}_ncu.log 2>&1
tail -12 gpurun_out/${TAG}_lone_launches.csv | cut -d, -f5,13- 
