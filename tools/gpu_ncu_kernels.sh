#!/bin/bash
# full ncu captures (with source) of selected launches of one C2 block
# usage: tools/gpu_ncu_kernels.sh <tag> <name:regex:skip> ...      e.g.  p31 attn:attention_tc:2 projff:proj_tc:8
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name regex skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${regex} -s ${skip} -c 1 \
     -o gpurun_out/prof_${name}_${TAG} -f python tools/profile_block.py > gpurun_out/ncu_${name}_${TAG}.log 2>&1; echo "ncu ${name} rc=$?"
done
ls -la gpurun_out/*_${TAG}.ncu-rep
