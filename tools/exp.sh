lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" 
for w in 4 6 8 12 16; do PHMM_B200_WORKERS=$w python tools/stream_bench.py 131072 1.0 | cut -c1-330; done
for w in 6 8 12 16; do PHMM_B200_WORKERS=$w python tools/stream_bench.py 131072 0.5 | cut -c1-330; done
