python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do python tools/resident_bench.py cfg2 4096; done
python tools/resident_bench.py cfg3 96
python tools/resident_bench.py cfg4 192
python tools/resident_bench.py cfg5 16384
