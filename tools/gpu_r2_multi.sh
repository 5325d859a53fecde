#!/bin/bash
# tools/gpu_r2_multi.sh N -- the full bench line at N GPUs (as the driver launches it) into gpurun_out/r02_bench_nN.json
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","gpu_launches","n_gpus")}, d["e2e"]["value"])
        for c,v in d.get("configs",{}).items():
            print(c, {l:round(x.get("ms",0),3) for l,x in v.get("legs",{}).items()}, v.get("value"), v.get("parity"))
PY
