#!/bin/bash
# N=8: the full bench under torchrun, then the reference arm (rank 0 only works)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err
echo "bench n8 rc=$?"
tail -c 6000 gpurun_out/r2p_bench_n8.json
grep '^\[bench\]' gpurun_out/r2p_bench_n8.err | tail -30
