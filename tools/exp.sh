for i in 1 2; do
python tools/resident_bench.py cfg5 32768 1.0
PHMM_B200_CTA_RULE=old python tools/resident_bench.py cfg5 32768 1.0
done
python tools/resident_bench.py cfg5 16384 0.5
PHMM_B200_CTA_RULE=old python tools/resident_bench.py cfg5 16384 0.5
python tools/resident_bench.py cfg3 96
PHMM_B200_CTA_RULE=old python tools/resident_bench.py cfg3 96
python tools/resident_bench.py cfg5 16384
PHMM_B200_CTA_RULE=old python tools/resident_bench.py cfg5 16384
