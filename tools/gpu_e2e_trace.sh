for thr in 8 12; do
echo "== threads $thr"
SQOA_B200_COPY_THREADS=$thr SQOA_B200_TRACE=1 timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_trace_$thr.log 2> gpurun_out/bench_trace_$thr.err; tail -10 gpurun_out/bench_trace_$thr.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_trace_$thr.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])"
done
