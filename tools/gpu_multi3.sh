#!/bin/bash
# tools/gpu_multi3.sh -- repeatability of the default multi-GPU line (cfg2, weak scaling)
N=${N:-8}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}" 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['steps'], round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms'],3) for k,v in d['legs'].items()}, round(d['e2e']['value']))"; }
run 29531 --steps 5 --warmup 3
run 29532 --steps 5 --warmup 3
run 29533 --steps 20 --warmup 3
