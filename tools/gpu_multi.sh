#!/bin/bash
# 2-GPU checks: frame-sharded bit identity, then the bench line with its frame_shard record.
mkdir -p gpurun_out
N=${1:-2}
T=${2:-m}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_frame_shard.py 2>&1 | grep -E "FRAME_SHARD|rank|Error|error" | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${T}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("value", d["value"], "n_gpus", d["n_gpus"], "e2e", d["e2e"]["value"])
    print("frame_shard", json.dumps(d.get("frame_shard"), indent=1))
except Exception as e:
    print("no json:", e)
PY
tail -n 6 gpurun_out/${T}_bench_${N}gpu.err
