#!/bin/bash
# tools/gpu_multi.sh -- multi-GPU legs (run under gpurun --gpus N): NCCL exchange of shard summaries (cfg4),
# icon batch sharded by index (cfg3), and the default weak-scaling line.
N=${N:-2}
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 3 --warmup 3 > gpurun_out/multi_cfg2_n$N.log 2> gpurun_out/multi_cfg2_n$N.err; echo "cfg2 rc=$?"
run 29512 --steps 3 --warmup 3 --workload cfg3 --images ${IMAGES:-100000} > gpurun_out/multi_cfg3_n$N.log 2> gpurun_out/multi_cfg3_n$N.err; echo "cfg3 rc=$?"
run 29513 --steps 2 --warmup 3 --workload cfg4 ${CFG4_ARGS:-} > gpurun_out/multi_cfg4_n$N.log 2> gpurun_out/multi_cfg4_n$N.err; echo "cfg4 rc=$?"
for f in gpurun_out/multi_cfg*_n$N.log; do echo "== $f"; grep '^{' $f | cut -c1-900; done
for f in gpurun_out/multi_cfg*_n$N.err; do echo "== $f"; grep -v -i "warn" $f | tail -5; done
