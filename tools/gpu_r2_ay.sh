#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "transcode" 2>&1 | tail -3
bash tools/gpu_r2_n1.sh 2>&1 | tail -2
python - <<'PY'
import json
for l in open("gpurun_out/r02_bench_n1.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
        for c,v in d.get("configs",{}).items():
            print(c, {l:round(x.get("ms",0),3) for l,x in v.get("legs",{}).items()}, v.get("value"), v.get("parity"))
PY
