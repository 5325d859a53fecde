#!/bin/bash
# full GPU test-suite + the new bench.py at N=1 (all configs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p timeout --timeout=600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err ); echo "bench rc=$?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    for k,v in d['legs'].items(): print(k, round(v['ms'],4),'ms', round(v['mpx_s']),'Mpx/s', round(v['gb_s']),'GB/s', round(v['frac_of_measured_hbm'],4))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_spot_check'], 'launches', d['gpu_launches'])
    print('cpu', json.dumps(d.get('cpu_baseline'))[:300])
    for name,c in d.get('configs',{}).items():
        print(name, json.dumps(c)[:1500])
except Exception as e:
    print("no line", e)
PY
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1 ); tail -c 600 gpurun_out/bench_ref.log
