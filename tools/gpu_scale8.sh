#!/bin/bash
# 8-GPU weak-scaling check of bench.py (cfg2, 4096 regions per rank): device step and end-to-end GCUPS, host phases
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node ${3:-8} --master-addr 127.0.0.1 --master-port $1 bench.py --gpus ${3:-8} --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$2', d['n_gpus'], d['value'], e['value'], e['ms_per_step'], e['host_plan_ms_per_step'], e['host_epilogue_ms_per_step'])"; }
run 29521 default 8
run 29522 default 8
nproc
