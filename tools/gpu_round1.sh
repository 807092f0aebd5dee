#!/bin/bash
# First GPU bring-up: kernel parity, module parity, smoke, bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm" -x 2>&1 | tail -40 > gpurun_out/t_gemm.log
echo "gemm exit ${PIPESTATUS[0]}" >> gpurun_out/t_gemm.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not gemm" 2>&1 | tail -80 > gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s 2>&1 | tail -120 > gpurun_out/t_e2e.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/t_gemm.log gpurun_out/t_kernels.log gpurun_out/t_e2e.log gpurun_out/smoke.log gpurun_out/bench.log
