#!/bin/bash
# tools/gpu_rows_quick.sh -- timing + one ncu capture of qoi_rows_kernel (no pytest)
mkdir -p gpurun_out
for shape in ${SHAPES:-4k3 big3}; do
  timeout 200 python tools/time_legs.py --shape $shape --legs qoi_decode > gpurun_out/rows_time_$shape.log 2>&1
  cat gpurun_out/rows_time_$shape.log
done
timeout 300 python tools/prof_legs.py --legs qoi_decode --reps 2 > gpurun_out/prof_plain_rows.log 2>&1 || { cat gpurun_out/prof_plain_rows.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qoi_rows" -s 1 -c 1 -f \
    -o gpurun_out/${NAME:-r01_rows} python tools/prof_legs.py --legs qoi_decode --reps 2 > gpurun_out/ncu_rows.log 2>&1
tail -2 gpurun_out/ncu_rows.log
