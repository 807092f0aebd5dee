#!/bin/bash
mkdir -p gpurun_out
T=${1:-cfm}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cfm or cffa_norm" -s 2>&1 | tail -12
TIME=1 timeout 300 python tools/one_kernel.py cfm 2>&1 | tail -1
timeout 300 python tools/cfm_timeline.py 2>&1 | head -14
