#!/bin/bash
# N-GPU A/B runs of bench.py under different data-parallel knobs (tools/, not part of the bench contract)
N=${1:-2}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 100)) bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/b${N}_$name.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N $name', round(d['value']), round(d['ms_per_step'],3), d['loss'])
except Exception as e: print('N=$N $name', 'FAILED', e)"
}
run sum_prescale MAMBA_B200_NCCL_AVG=0
run avg MAMBA_B200_NCCL_AVG=1
