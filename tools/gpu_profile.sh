#!/bin/bash
# Profile evidence of one build on one B200: ncu launch list + DRAM metric pass of two eager steps of the bench workload,
# and ncu --set full captures (with source) of the hot kernels on production shapes.  Outputs under gpurun_out/<tag>_*.
mkdir -p gpurun_out
T=${1:-r02}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py 2 > gpurun_out/${T}_ncu_launch.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_kernels.csv python tools/profile_step.py 2 > gpurun_out/${T}_ncu_kern.log 2>&1
tail -n 2 gpurun_out/${T}_ncu_launch.log gpurun_out/${T}_ncu_kern.log
shift
bash tools/gpu_ncu_full.sh $T "$@"
