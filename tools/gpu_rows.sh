#!/bin/bash
# tools/gpu_rows.sh -- parity + timing + one ncu capture of the one-launch QOI decoder (qoi_rows_kernel).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p timeout --timeout=150 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for shape in 4k3 big3 4k4; do
  timeout 200 python tools/time_legs.py --shape $shape --legs qoi_decode,sqoa_decode > gpurun_out/rows_time_$shape.log 2>&1
  cat gpurun_out/rows_time_$shape.log
done
timeout 300 python tools/prof_legs.py --legs qoi_decode --reps 2 > gpurun_out/prof_plain_rows.log 2>&1 || { cat gpurun_out/prof_plain_rows.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:qoi_rows" -s 1 -c 1 -f \
    -o gpurun_out/r01_rows python tools/prof_legs.py --legs qoi_decode --reps 2 > gpurun_out/ncu_rows.log 2>&1
tail -2 gpurun_out/ncu_rows.log
