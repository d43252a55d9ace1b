python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu -k "fp64 or fast_fp64 or kernel_info" 2>&1 | tail -5
python tools/resident_bench.py cfg3 96
FP64_ALWAYS=1 python tools/resident_bench.py cfg3 96
FP64_ALWAYS=1 PHMM_B200_FP64_GENERAL=1 python tools/resident_bench.py cfg3 96
python tools/resident_bench.py cfg5 16384
FP64_ALWAYS=1 python tools/resident_bench.py cfg5 16384
FP64_ALWAYS=1 PHMM_B200_FP64_GENERAL=1 python tools/resident_bench.py cfg5 16384
bash tools/gpu_callers.sh r02_callers 2>&1 | tail -14
