#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "mha" 2>&1 | tail -25 > gpurun_out/t_mha.log
tail -5 gpurun_out/t_mha.log
if grep -q "passed" gpurun_out/t_mha.log && ! grep -q "failed" gpurun_out/t_mha.log; then
  timeout 120 python tools/mha_probe.py
  CFFM_MHA_LEGACY=1 timeout 120 python tools/mha_probe.py
  timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x 2>&1 | tail -3
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err
  grep -o "\"value\": [0-9.]*" gpurun_out/bench.log | head -2
fi
