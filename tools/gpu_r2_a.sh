#!/bin/bash
# tools/gpu_r2_a.sh -- round 2, first GPU call: parity tests, per-leg timings on three shapes, ncu capture of the SQOA encoder
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu -p timeout --timeout=200 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for shape in 4k3 4k4 big4 big3; do
  timeout 200 python tools/time_legs.py --shape $shape --legs sqoa_encode,qoi_encode > gpurun_out/r2a_time_$shape.log 2>&1; echo "time $shape rc=$?"; cat gpurun_out/r2a_time_$shape.log
done
timeout 200 python tools/prof_legs.py --legs sqoa_encode,qoi_encode --reps 4 > gpurun_out/r2a_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:encode_block -s 2 -c 1 -o gpurun_out/r2a_sqoa_encode -f python tools/prof_legs.py --legs sqoa_encode --reps 4 > gpurun_out/r2a_ncu1.log 2>&1
tail -2 gpurun_out/r2a_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:encode_block -s 6 -c 1 -o gpurun_out/r2a_qoi_encode -f python tools/prof_legs.py --legs sqoa_encode,qoi_encode --reps 4 > gpurun_out/r2a_ncu2.log 2>&1
tail -2 gpurun_out/r2a_ncu2.log
