for r in 1 7 9 11 13 15 17; do n=$((4096*19/r)); if [ $r -eq 1 ]; then n=40000; fi; PHMM_B200_RPT=$r python tools/resident_bench.py uniform:64x8x$((8*r-1))x300 $n; done
# 4-lane reads at 12 and 19 rows per lane: does the per-step overhead depend on lanes per read?
PHMM_B200_RPT=12 python tools/resident_bench.py uniform:64x8x47x300 12970
PHMM_B200_RPT=19 python tools/resident_bench.py uniform:64x8x75x300 8192
PHMM_B200_RPT=12 python tools/resident_bench.py uniform:64x8x191x300 3200
for r in 11 12 13; do PHMM_B200_RPT=$r python tools/resident_bench.py cfg5 16384 0.5; done
for r in 11 12 13; do PHMM_B200_RPT=$r python tools/resident_bench.py cfg3 96; done
