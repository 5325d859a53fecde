#!/bin/bash
# tools/gpu_multi_cfg3.sh -- the icon batch (cfg3) sharded by image index over N GPUs
N=${N:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --workload cfg3 > gpurun_out/multi_cfg3_n$N.log 2> gpurun_out/multi_cfg3_n$N.err
grep '^{' gpurun_out/multi_cfg3_n$N.log | cut -c1-1500
