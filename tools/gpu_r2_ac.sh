#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
VARIANTS="default" SHAPES="4k3 4k4 big4 1080p4" LEGS=sqoa_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2ac_sfx.log
for nw in 0 1; do
echo "== nowait=$nw images=12500"
SQOA_BENCH_QOI_NOWAIT=$nw timeout 300 python bench.py --only cfg3 --images 12500 --steps 10 --warmup 3 > gpurun_out/r2ac_$nw.json 2> gpurun_out/r2ac_$nw.err
python - <<PY
import json
for l in open("gpurun_out/r2ac_$nw.json"):
    if l.startswith("{"):
        d=json.loads(l); c=d.get("configs",{}).get("cfg3",d)
        print({k:round(v["ms"],3) for k,v in c["legs"].items()}, c.get("ms_per_step"), c.get("parity"))
PY
done
