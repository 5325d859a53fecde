#!/bin/bash
mkdir -p gpurun_out
for nw in 1 0; do
SQOA_BENCH_QOI_NOWAIT=$nw SQOA_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 8 --steps 20 --warmup 3 --only cfg3 > gpurun_out/r2aj_cfg3_n8_$nw.json 2> gpurun_out/r2aj_cfg3_n8_$nw.err
echo "nowait=$nw rc=$?"; grep "cfg3 rank" gpurun_out/r2aj_cfg3_n8_$nw.err | sort
done
