#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/final_default_bench.json 2> gpurun_out/final_default_bench.err; echo "bench rc=$?"; python -c "
import json
for l in open('gpurun_out/final_default_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], [ (k, v.get('parity')) for k,v in d['configs'].items()])
"
