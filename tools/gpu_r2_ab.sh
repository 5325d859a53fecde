#!/bin/bash
mkdir -p gpurun_out
for nw in 0 1; do
echo "== nowait=$nw images=12500"
SQOA_BENCH_QOI_NOWAIT=$nw timeout 300 python bench.py --only cfg3 --images 12500 --steps 10 --warmup 3 > gpurun_out/r2ab_$nw.json 2> gpurun_out/r2ab_$nw.err
tail -c 1500 gpurun_out/r2ab_$nw.json; tail -5 gpurun_out/r2ab_$nw.err
done
timeout 300 python tools/gpu_shard_passes.py --rows 5000 --world 2 2>&1 | tee gpurun_out/r2ab_passes.log
timeout 300 python tools/gpu_shard_passes.py --rows 5000 --world 8 2>&1 | tee -a gpurun_out/r2ab_passes.log
