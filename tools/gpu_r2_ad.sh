#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2ad_ref.json 2> gpurun_out/r2ad_ref.err ) 2>&1 | grep real
( time timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2ad_ref.json; tail -3 gpurun_out/r2ad_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r2ad_bench.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["per_leg_frac"])
        for c,v in d.get("configs",{}).items():
            print(c, {l:round(x.get("ms",0),3) for l,x in v.get("legs",{}).items()}, v.get("value"), v.get("parity"), v.get("wall_s"))
PY
