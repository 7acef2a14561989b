#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pair_bias -s 1 -c 1 \
   -o gpurun_out/prof_pair_bias -f python tools/profile_block.py > gpurun_out/ncu_pb.log 2>&1; echo "ncu pair_bias rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:chan_to_token -s 1 -c 1 \
   -o gpurun_out/prof_c2t -f python tools/profile_block.py > gpurun_out/ncu_c2t.log 2>&1; echo "ncu c2t rc=$?"
