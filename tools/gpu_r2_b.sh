#!/bin/bash
# tools/gpu_r2_b.sh -- encoder timing over variants + ncu capture of the default build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p timeout --timeout=200 --timeout-method=thread -k "encode or kat or roundtrip" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
VARIANTS="${VARIANTS:-default c4}" SHAPES="${SHAPES:-4k3 4k4 big4}" LEGS=sqoa_encode,qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2b_variants.log
timeout 200 python tools/prof_legs.py --legs sqoa_encode --reps 4 > gpurun_out/r2b_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:encode_block -s 2 -c 1 -o gpurun_out/r2b_sqoa_encode -f python tools/prof_legs.py --legs sqoa_encode --reps 4 > gpurun_out/r2b_ncu1.log 2>&1
tail -2 gpurun_out/r2b_ncu1.log
