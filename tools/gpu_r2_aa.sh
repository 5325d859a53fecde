#!/bin/bash
mkdir -p gpurun_out
for nw in 0 1; do for n in 12500 100000; do
echo "== nowait=$nw images=$n"
SQOA_BENCH_QOI_NOWAIT=$nw timeout 300 python bench.py --only cfg3 --images $n --steps 10 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); c=d.get('configs',{}).get('cfg3',d)
        print({k:round(v['ms'],3) for k,v in c['legs'].items()}, c.get('ms_per_step'), c.get('parity'))
"
done; done 2>&1 | tee gpurun_out/r2aa_cfg3.log
