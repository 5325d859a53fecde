#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "encode or qoi or roundtrip" 2>&1 | tail -3
SHAPES="4k3 4k4 big4" LEGS=qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2r_bits.log
