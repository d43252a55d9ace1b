python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "fp64" 2>&1 | tail -2
python tools/resident_bench.py cfg3 96
FP64_ALWAYS=1 python tools/resident_bench.py cfg3 96
