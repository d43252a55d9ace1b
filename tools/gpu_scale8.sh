#!/bin/bash
# 8-GPU end-to-end experiments: default workers, then overrides
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$2', d['value'], e['value'], e['ms_per_step'], e['host_plan_ms_per_step'], e['host_epilogue_ms_per_step'])"; }
run 29521 default
PHMM_B200_WORKERS=3 run 29522 workers3
PHMM_B200_WORKERS=8 run 29523 workers8
