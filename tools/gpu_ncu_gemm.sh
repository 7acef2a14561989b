#!/bin/bash
mkdir -p gpurun_out
for spec in "qkv:21" "gate:1" "ffw1:27"; do
  name=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s $skip -c 1 \
     -o gpurun_out/prof_gemm_$name -f python tools/profile_block.py > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
done
ls -la gpurun_out/*.ncu-rep
