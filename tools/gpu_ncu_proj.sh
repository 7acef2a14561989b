#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:proj_tc -s 8 -c 1 \
   -o gpurun_out/prof_proj_ff -f python tools/profile_block.py > gpurun_out/ncu_proj.log 2>&1; echo "ncu proj ff rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:proj_tc -s 3 -c 1 \
   -o gpurun_out/prof_proj_outer -f python tools/profile_block.py > gpurun_out/ncu_proj2.log 2>&1; echo "ncu proj outer rc=$?"
