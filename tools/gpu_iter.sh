#!/bin/bash
# One optimisation iteration: variant timing, GPU parity tests, bench.  Small outputs only.
mkdir -p gpurun_out
T=${1:-it}
[ -f tools/tune.py ] && timeout 600 python tools/tune.py > gpurun_out/${T}_tune.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_tune.log 2>/dev/null; tail -n 8 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_bench.json; tail -n 5 gpurun_out/${T}_bench.err
