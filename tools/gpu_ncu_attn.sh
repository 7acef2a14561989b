#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -s 2 -c 1 \
   -o gpurun_out/prof_attention -f python tools/profile_block.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
