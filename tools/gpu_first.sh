#!/bin/bash
# first contact with the B200: probes, smoke, parity tests, micro-benchmarks, a short bench
mkdir -p gpurun_out
{
echo "== host"; nproc; lscpu | grep -E "Model name|^CPU\(s\)" ; lscpu | grep -o -E "avx512[a-z_0-9]*" | sort -u | tr '\n' ' '; echo
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,memory.total --format=csv
ls /root/reference 2>&1 | head -2
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== ubench"; timeout 120 python - <<'PY'
from rovaca_b200 import PairHmmEngine
e = PairHmmEngine()
for mode, name in enumerate(["FFMA scalar", "FFMA2 packed", "FFMA2+LOP3/FSEL/FADD", "mode2+SHFL"]):
    ms, rate = e.ubench(mode, 8192)
    print(f"ubench mode {mode} {name:24s} {ms:8.3f} ms  {rate*2/1e12:7.2f} TFLOP/s (FMA flops)")
for n in (15, 31, 47, 63, 79, 95, 111, 127, 143, 150, 175, 191, 223, 255, 287, 319, 351, 383):
    print("read_len", n, e.kernel_info(n))
e.close()
PY
} > gpurun_out/first.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/first.log
tail -15 gpurun_out/pytest_gpu.log >> gpurun_out/first.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench exit $?" >> gpurun_out/first.log
cat gpurun_out/bench_first.json >> gpurun_out/first.log; tail -5 gpurun_out/bench_first.err >> gpurun_out/first.log
cat gpurun_out/first.log
