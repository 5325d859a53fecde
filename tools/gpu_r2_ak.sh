#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for n in 12500 100000; do
echo "== images=$n"
timeout 300 python bench.py --only cfg3 --images $n --steps 10 --warmup 3 > gpurun_out/r2ak_$n.json 2> gpurun_out/r2ak_$n.err
python - <<PY
import json
for l in open("gpurun_out/r2ak_$n.json"):
    if l.startswith("{"):
        d=json.loads(l); c=d.get("configs",{}).get("cfg3",d)
        print({k:round(v["ms"],3) for k,v in c["legs"].items()}, c.get("ms_per_step"), c.get("parity"))
PY
done
timeout 300 python bench.py --only cfg5 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d.get('configs',{}).get('cfg5',d); print({k:round(v['ms'],3) for k,v in c['legs'].items()}, c.get('parity'))
"
