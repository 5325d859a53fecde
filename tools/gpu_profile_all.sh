#!/bin/bash
# tools/gpu_profile_all.sh -- the round's evidence in one gpurun call: plain bench, its ncu launch list,
# and one ncu --set full capture per leg (cfg2 image; each capture only after the plain command exited 0).
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for leg in sqoa_encode sqoa_decode qoi_encode qoi_decode; do
  python tools/prof_legs.py --legs $leg --reps 2 > gpurun_out/prof_plain_$leg.log 2>&1 || exit 1
  case $leg in
    sqoa_encode) rx='encode_block'; skip=1; cnt=1;;   # launches: <3,0> <3,0> <3,1>
    qoi_encode)  rx='encode_block'; skip=2; cnt=1;;   # launches: <3,0> <3,1> <3,1>
    sqoa_decode) rx='sqoa_decode_kernel'; skip=1; cnt=1;;
    qoi_decode)  rx='qoi_'; skip=1; cnt=1;;            # launches: qoi_rows_kernel x2 (one launch per decode)
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f \
      -o gpurun_out/final_$leg python tools/prof_legs.py --legs $leg --reps 2 > gpurun_out/ncu_final_$leg.log 2>&1
  tail -1 gpurun_out/ncu_final_$leg.log
done
cut -c1-600 gpurun_out/bench_final.log
