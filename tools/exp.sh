#!/bin/bash
# scratch: synchronous callers, pipeline depth x pinned layout
mkdir -p gpurun_out
OUT=gpurun_out/${1:-exp}.txt
: > $OUT
g++ -O2 -std=c++17 -Iinclude tools/caller_bench.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -lpthread -o tools/caller_bench || exit 1
for t in 16 32 64; do
  for d in 3 6 8 12; do
    for lanes in 0 8 16; do
      echo "threads $t depth $d lanes $lanes" >> $OUT
      tools/caller_bench --threads $t --regions $((8192 / t)) --lanes $lanes --depth $d >> $OUT 2>&1
    done
  done
done
cut -c1-200 $OUT
