#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "many_files or write_and_read or several_threads" 2>&1 | tail -4
timeout 600 python tools/gpu_many_files.py --icons 4096 2>&1 | tail -4 | tee gpurun_out/r02_many_files.txt
timeout 600 python tools/gpu_many_files.py --icons 32768 2>&1 | tail -4 | tee -a gpurun_out/r02_many_files.txt
