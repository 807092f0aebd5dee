#!/bin/bash
# 2-GPU validation of the frame-sharded path + both sharding modes of bench.py
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_frame_shard.py 4 > gpurun_out/shard_check.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2gpu_clips.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --shard frames > gpurun_out/bench_2gpu_frames.log 2>&1
grep -h "rank\|FRAME_SHARD\|Error\|error" gpurun_out/shard_check.log | tail -8
for f in bench_2gpu_clips bench_2gpu_frames; do grep -o '"value": [0-9.]*, "unit": "clip-frames/s", "n_gpus": [0-9]*\|"ms_per_step": [0-9.]*' gpurun_out/$f.log | head -3; tail -n 3 gpurun_out/$f.log | grep -i "error\|Traceback" ; done
