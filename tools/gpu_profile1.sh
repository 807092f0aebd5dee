#!/bin/bash
# ncu evidence for round 1.  Only small CSV / one small .ncu-rep are left in gpurun_out (64 MiB cap).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cffa" 2>&1 | tail -5 > gpurun_out/t_cffa.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py 2 > gpurun_out/ncu_launch.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/kernels_r1.csv python tools/profile_step.py 2 > gpurun_out/ncu_kern.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cfm_attention -s 2 -c 2 -o gpurun_out/prof_cfm_r1 python tools/profile_step.py 2 > gpurun_out/ncu_cfm.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
