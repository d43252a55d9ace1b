#!/bin/bash
# scratch: stream word fetched a step ahead in the throughput kernels (library variants built with -DPHMM_EXP_WORD_PREFETCH=1|2)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-exp}.txt
: > $OUT
for spec in "cfg2 4096 0.0" "cfg5 16384 0.5" "cfg4 192 0.0"; do
  python tools/resident_bench.py $spec 3 >> $OUT 2>&1
  for v in 1 2; do
    PHMM_B200_LIB=$PWD/rovaca_b200/libphmm_b200_exp$v.so python tools/resident_bench.py $spec 3 >> $OUT 2>&1
  done
done
cut -c1-220 $OUT
