#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r2ag_bench.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d["roofline"]["per_leg_frac"])
        for c,v in d.get("configs",{}).items():
            print(c, {l:round(x.get("ms",0),3) for l,x in v.get("legs",{}).items()}, v.get("value"), v.get("parity"), v.get("e2e"), v.get("latency_us"))
PY
