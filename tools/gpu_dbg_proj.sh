#!/bin/bash
mkdir -p gpurun_out
AF2_PROJ_DBG=744 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:proj_tc -s 6 -c 1 \
   -o gpurun_out/prof_proj_dbg744 -f python tools/profile_block.py > gpurun_out/ncu_proj_dbg.log 2>&1; echo "rc=$?"
