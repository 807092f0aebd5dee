#!/bin/bash
mkdir -p gpurun_out
T=${1:-p}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cfm_att -s 2 -c 1 -f -o gpurun_out/${T}_cfm python tools/one_kernel.py cfm > gpurun_out/${T}_cfm.log 2>&1
tail -n 3 gpurun_out/${T}_cfm.log
ls -la gpurun_out/${T}_cfm*
