#!/bin/bash
# MHA kernel iteration: parity cases, stage-1 timeline + timing; with a tag as $1 also the full suite + bench (tools/gpu_iter.sh).
mkdir -p gpurun_out
T=${1:-}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k mha 2>&1 | tail -8 | tee gpurun_out/mha_tests.log
if grep -q "failed\|error" gpurun_out/mha_tests.log; then exit 1; fi
timeout 300 python tools/mha_timeline.py 2>&1 | head -12
TIME=1 timeout 300 python tools/one_kernel.py mha 2>&1 | tail -1
if [ -n "$T" ]; then bash tools/gpu_iter.sh $T; fi
