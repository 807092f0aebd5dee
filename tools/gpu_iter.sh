#!/bin/bash
# One optimisation iteration on the GPU: parity, bench, launch list.  Small outputs only.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/t_e2e.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/ncu_launch.log 2>&1
for f in t_kernels t_e2e bench bench_eager; do echo "=== $f"; tail -n 12 gpurun_out/$f.log; done; tail -n 5 gpurun_out/bench.err
