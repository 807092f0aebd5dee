#!/bin/bash
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cfm_attention_vs_oracle" -s 2>&1 | grep -E "passed|failed|Error|error|rel err" | head -10
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cfm_attention_vs_oracle and 1-21-28" 2>&1 | grep -E "=========|Invalid|at 0x|by thread|Address|passed|failed" | head -30
