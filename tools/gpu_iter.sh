#!/bin/bash
# One optimisation iteration: GPU parity tests, CFM kernel timing, bench.  Small outputs only.
mkdir -p gpurun_out
T=${1:-it}
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
TIME=1 timeout 300 python tools/one_kernel.py cfm 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["launch_mode"])
print("kernel sum ms", d["kernel_time_sum_ms_per_step"])
for k,v in d["kernel_breakdown"].items(): print("  ", k, v)
print("gemm roofline", d["roofline"]["frac"], "cfm", d["roofline_cfm_attention"]["frac"], d["roofline_cfm_attention"]["launch_ms"])
PY
tail -n 3 gpurun_out/${T}_bench.err
