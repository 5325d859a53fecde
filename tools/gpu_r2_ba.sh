#!/bin/bash
SQOA_B200_LIB=$PWD/gpurun_variants/libsqoa_b200_fence1.so timeout 70 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -4
