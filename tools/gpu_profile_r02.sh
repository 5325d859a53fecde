#!/bin/bash
# tools/gpu_profile_r02.sh -- round-2 profile captures: launch list of the bench command, one ncu --set full capture
# per hot kernel on the bench's own launch shape (16 cfg2 images per launch).
mkdir -p gpurun_out
CMD="python bench.py --skip-configs --steps 2 --warmup 3"
timeout 300 $CMD > gpurun_out/r02_bench_plain.log 2> gpurun_out/r02_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"; grep -c "encode_block\|decode_kernel\|rows_kernel" gpurun_out/r02_bench_launches.csv
for leg in sqoa_encode qoi_encode sqoa_decode qoi_decode; do
  case $leg in sqoa_encode|qoi_encode) rx=encode_block; skip=1;; sqoa_decode) rx=sqoa_decode_kernel; skip=1;; qoi_decode) rx=qoi_rows_kernel; skip=1;; esac
  timeout 200 python tools/prof_bench_legs.py --legs $leg --reps 3 > gpurun_out/r02_prof_plain_$leg.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r02_$leg -f python tools/prof_bench_legs.py --legs $leg --reps 3 > gpurun_out/r02_ncu_$leg.log 2>&1
  echo "$leg rc=$?"; tail -1 gpurun_out/r02_ncu_$leg.log
done
