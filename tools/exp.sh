#!/bin/bash
# scratch experiments on the GPU box (output: gpurun_out/exp.txt)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-exp}.txt
: > $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${1:-exp}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT; tail -3 gpurun_out/${1:-exp}_pytest.log >> $OUT
echo "## paired reads on / off (PHMM_B200_NO_DUO), device-resident" >> $OUT
for spec in "cfg5 16384 0.0" "cfg5 16384 0.5" "cfg5 65536 1.0" "cfg3 96 0.0" "cfg4 192 0.0" "cfg2 4096 0.0"; do
  python tools/resident_bench.py $spec >> $OUT 2>&1
  PHMM_B200_NO_DUO=1 python tools/resident_bench.py $spec >> $OUT 2>&1
done
echo "## allocator A/B on cfg4 (PHMM_B200_SYNC_ALLOC = cudaMalloc / cudaFree)" >> $OUT
for k in 1 2; do
  python tools/resident_bench.py cfg4 192 0.0 >> $OUT 2>&1
  PHMM_B200_SYNC_ALLOC=1 python tools/resident_bench.py cfg4 192 0.0 >> $OUT 2>&1
done
echo "## callers" >> $OUT
g++ -O2 -std=c++17 -Iinclude tools/caller_bench.cpp -Lrovaca_b200 -lphmm_b200 -Wl,-rpath,'$ORIGIN/../rovaca_b200' -lpthread -o tools/caller_bench
for t in 1 2 4 32 64; do tools/caller_bench --threads $t --regions $((4096 / t > 64 ? 4096 / t : 64)) >> $OUT 2>&1; done
cat $OUT | cut -c1-260
