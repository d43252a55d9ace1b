#!/bin/bash
# the reference's calling pattern (many synchronous callers) against the C ABI; see tools/caller_bench.cpp
mkdir -p gpurun_out
OUT=gpurun_out/${1:-callers}.txt
: > $OUT
for t in 8 16 32 64 128 256; do tools/caller_bench --threads $t --regions $((4096 / t > 16 ? 4096 / t : 16)) >> $OUT 2>&1; done
for w in 4 16 64; do tools/caller_bench --threads 16 --regions 256 --window $w >> $OUT 2>&1; done
cat $OUT
