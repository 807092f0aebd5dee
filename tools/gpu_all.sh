#!/bin/bash
# Full GPU suite + headline bench + the other BASELINE configs (bench lines saved under gpurun_out/ with tag T).
mkdir -p gpurun_out
T=${1:-all}
timeout 1800 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "rel err|passed|failed|Error|error" | tail -40 > gpurun_out/${T}_tests.log
cat gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --variant b0 --clips 1 --T 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_cfg0_b0_T2.json 2> gpurun_out/${T}_cfg0.err; echo "cfg0 rc=$?"; tail -n 3 gpurun_out/${T}_cfg0.err
timeout 600 python bench.py --variant b2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_cfg3_b2.json 2> gpurun_out/${T}_cfg3.err; echo "cfg3 rc=$?"; tail -n 3 gpurun_out/${T}_cfg3.err
timeout 600 python bench.py --kind cffmpp --protos 64 --steps 20 --warmup 5 > gpurun_out/${T}_bench_cfg4_cffmpp.json 2> gpurun_out/${T}_cfg4.err; echo "cfg4 rc=$?"; tail -n 3 gpurun_out/${T}_cfg4.err
python - <<PY
import json
for f in ("bench", "bench_cfg0_b0_T2", "bench_cfg3_b2", "bench_cfg4_cffmpp"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/${T}_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], "e2e", d["e2e"]["value"], "mmseg", (d.get("e2e_mmseg_call") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
              "gemm frac", d["roofline"]["frac"], "cfm", (d.get("roofline_cfm_attention") or {}).get("frac"))
    except Exception as e:
        print(f, "no json", e)
PY
