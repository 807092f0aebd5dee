#!/bin/bash
# CFM kernel bring-up: parity of the new kernel in both key-row layouts, then the whole GPU suite and a bench.
mkdir -p gpurun_out
T=${1:-cfm}
K='cfm or cffa_norm'
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$K" -s 2>&1 | tail -25 > gpurun_out/${T}_tight.log
CFFM_CFM_ALIGNED=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$K" -s 2>&1 | tail -25 > gpurun_out/${T}_aligned.log
echo "== tight"; cat gpurun_out/${T}_tight.log; echo "== aligned"; cat gpurun_out/${T}_aligned.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/${T}_tests.log
echo "== all"; cat gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench.json; tail -n 5 gpurun_out/${T}_bench.err
