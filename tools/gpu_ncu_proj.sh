#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:proj_tc -s 6 -c 1 \
   -o gpurun_out/prof_proj_attn -f python tools/profile_block.py > gpurun_out/ncu_proj2.log 2>&1; echo "ncu proj attn rc=$?"
