#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/t_e2e.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 3 -o gpurun_out/prof_gemm python tools/profile_step.py 1 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 8 gpurun_out/t_e2e.log
