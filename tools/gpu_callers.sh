#!/bin/bash
# the reference's calling pattern (many synchronous callers) against the C ABI; see tools/caller_bench.cpp
mkdir -p gpurun_out
# always against the current header: the structs of include/phmm_b200.h are the ABI
g++ -O2 -std=c++17 -Iinclude tools/caller_bench.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -lpthread -o tools/caller_bench || exit 1
OUT=gpurun_out/${1:-callers}.txt
: > $OUT
tools/caller_bench --threads 1 --regions 512 >> $OUT 2>&1
for t in 8 16 32 64 128 256 512; do tools/caller_bench --threads $t --regions $((8192 / t > 64 ? 8192 / t : 64)) >> $OUT 2>&1; done
for w in 4 16 64; do tools/caller_bench --threads 16 --regions 512 --window $w >> $OUT 2>&1; done
cat $OUT
