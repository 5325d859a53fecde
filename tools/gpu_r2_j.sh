#!/bin/bash
mkdir -p gpurun_out
make -s -C tools
timeout 900 python -m pytest tests/test_tools.py tests/test_gpu_parity.py -q -m gpu -p timeout --timeout=600 --timeout-method=thread 2>&1 | tail -6
timeout 120 tools/bin/sqoabench_b200 3 --synth cfg1 --synth cfg2 --reference oracle/_ref/libsqoa_ref.so > gpurun_out/r02_sqoabench.txt 2>&1; echo "sqoabench rc=$?"; cat gpurun_out/r02_sqoabench.txt
timeout 600 python bench.py --only cfg5 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print('cfg5',k, round(v['ms'],3),'ms')
print('parity', d['parity'], 'value', d['value'])"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err ); echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.log').read().strip().splitlines()[-1])
    for k,v in d['legs'].items(): print(k, round(v['ms'],4),'ms', round(v['frac_of_measured_hbm'],4))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_spot_check'])
    for name,c in d.get('configs',{}).items():
        print(name, 'parity', c.get('parity'), c.get('parity_against'), c.get('error'))
        for k,v in c.get('legs',{}).items(): print('   ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','frac_of_measured_hbm','gpus','mpx_s','unsupported','round_trip_ok')})
except Exception as e:
    print("no line", e)
PY
