#!/bin/bash
# tools/gpu_multi2.sh -- multi-GPU lines for the default workload (cfg2, weak) and the transcode corpus (cfg5, strong)
N=${N:-8}
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29521 --steps 3 --warmup 3 > gpurun_out/multi_cfg2_n$N.log 2> gpurun_out/multi_cfg2_n$N.err; echo "cfg2 rc=$?"
run 29522 --steps 3 --warmup 3 --workload cfg5 --scale ${SCALE:-1.0} > gpurun_out/multi_cfg5_n$N.log 2> gpurun_out/multi_cfg5_n$N.err; echo "cfg5 rc=$?"
run 29523 --steps 3 --warmup 3 --workload cfg3 > gpurun_out/multi_cfg3_n$N.log 2> gpurun_out/multi_cfg3_n$N.err; echo "cfg3 rc=$?"
for f in gpurun_out/multi_cfg2_n$N.log gpurun_out/multi_cfg5_n$N.log gpurun_out/multi_cfg3_n$N.log; do echo "== $f"; grep '^{' $f | cut -c1-200; done
for f in gpurun_out/multi_cfg2_n$N.err gpurun_out/multi_cfg5_n$N.err; do echo "== $f"; grep -v -i "warn" $f | tail -3; done
