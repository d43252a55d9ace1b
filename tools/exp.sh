for c in 600 1000 1400 2000 2400; do PHMM_B200_STREAM_COLS=$c python tools/resident_bench.py cfg4 192; done
for c in 600 1000 1400 2000 2400; do PHMM_B200_STREAM_COLS=$c python tools/resident_bench.py cfg5 16384; done
for c in 600 1000 2400; do PHMM_B200_STREAM_COLS=$c python tools/resident_bench.py cfg3 96; done
for c in 600 1000 2400; do PHMM_B200_STREAM_COLS=$c python tools/resident_bench.py cfg2 4096; done
