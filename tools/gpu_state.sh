#!/bin/bash
# Full state check on one B200: GPU tests, smoke, both bench arms, ncu launch list + DRAM metric pass.
mkdir -p gpurun_out
T=${1:-v8}
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 2 > gpurun_out/${T}_ncu_launch.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_kernels.csv python tools/profile_step.py 2 > gpurun_out/${T}_ncu_kern.log 2>&1
for f in tests smoke; do echo "=== $f"; tail -n 8 gpurun_out/${T}_$f.log; done; cat gpurun_out/${T}_bench.json gpurun_out/${T}_bench_ref.json; tail -n 5 gpurun_out/${T}_bench.err
