#!/bin/bash
mkdir -p gpurun_out
T=${1:-}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "patch_embed" 2>&1 | tail -12 | tee gpurun_out/pe_tests.log
if grep -q "failed\|error" gpurun_out/pe_tests.log; then exit 1; fi
if [ -n "$T" ]; then bash tools/gpu_iter.sh $T; fi
