#!/bin/bash
mkdir -p gpurun_out
SQOA_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 \
  bench.py --gpus 8 --steps 10 --warmup 3 --only cfg5 > gpurun_out/r2ao_cfg5_n8.json 2> gpurun_out/r2ao_cfg5_n8.err
echo "rc=$?"; grep -o "cfg5 rank.*" gpurun_out/r2ao_cfg5_n8.err | sort
python - <<'PY'
import json
for l in open("gpurun_out/r2ao_cfg5_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); c=d.get("configs",{}).get("cfg5",d); print({k:round(v["ms"],3) for k,v in c["legs"].items()}, c.get("value"), c.get("parity"))
PY
