#!/bin/bash
# A/B bench runs of environment toggles: tools/gpu_ab.sh "VAR=val ..." "VAR=val ..." ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  env $cfg timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$i.json"))
print("$cfg", "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["launch_mode"])
PY
  i=$((i+1))
done
