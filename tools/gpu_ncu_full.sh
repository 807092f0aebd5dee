#!/bin/bash
# ncu --set full capture of single hot kernels (third launch of each), reports into gpurun_out/.
mkdir -p gpurun_out
T=${1:-r01}
shift
for k in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|dwconv|head_fuse|argmax|mha_|layernorm|cfm_att|mixffn|patch_embed' -s 2 -c 1 -f -o gpurun_out/${T}_${k} python tools/one_kernel.py $k > gpurun_out/${T}_${k}.log 2>&1
  tail -n 2 gpurun_out/${T}_${k}.log
done
ls -la gpurun_out | tail -n 12
