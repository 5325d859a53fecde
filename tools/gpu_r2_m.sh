#!/bin/bash
mkdir -p gpurun_out
SHAPES="4k3 4k4 big4" LEGS=sqoa_encode,qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2m_variants.log
timeout 300 python bench.py --skip-configs --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print(k, round(v['ms'],4),'ms', round(v['frac_of_measured_hbm'],4))
print('value', d['value'], 'e2e', d['e2e']['value'], d['parity_spot_check'])"
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 3 --warmup 3 $2 $3 > gpurun_out/n2_$4.log 2> gpurun_out/n2_$4.err; echo "$4 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n2_$4.log').read().strip().splitlines()[-1])
    print('$4', 'parity', d.get('parity'), d.get('parity_against'))
    for k,v in d.get('legs',{}).items(): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','frac_of_measured_hbm','gpus','unsupported','round_trip_ok')})
except Exception as e: print('no line', e)
PY
grep -i "error\|Traceback" gpurun_out/n2_$4.err | head -3; }
run2 29522 --only cfg3 cfg3
run2 29523 --only cfg4 cfg4
