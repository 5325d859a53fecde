#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded or shard" 2>&1 | tail -15
echo "== N=2 cfg4"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 5 --warmup 3 --only cfg4 > gpurun_out/r2y_cfg4_n2.json 2> gpurun_out/r2y_cfg4_n2.err
echo "rc=$?"; tail -c 3000 gpurun_out/r2y_cfg4_n2.json; tail -5 gpurun_out/r2y_cfg4_n2.err
