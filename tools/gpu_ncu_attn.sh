#!/bin/bash
# full ncu capture (with source) of one triangle-attention launch of a C2 block
# usage: tools/gpu_ncu_attn.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -s 2 -c 1 \
   -o gpurun_out/prof_attention_${TAG} -f python tools/profile_block.py > gpurun_out/ncu_attn_${TAG}.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/prof_attention_${TAG}.ncu-rep
