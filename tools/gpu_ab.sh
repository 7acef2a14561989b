#!/bin/bash
# A/B of build variants and runtime knobs on ONE box (box-to-box variation of the same binary is ~3-4 %):
#   tools/gpu_ab.sh <tag>  -> gpurun_out/ab_<tag>.log
TAG=${1:-ab}
mkdir -p gpurun_out
L=gpurun_out/ab_${TAG}.log
: > $L
run_bench() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
  python - "$label" <<'PY' >> $L
import json, sys
try:
    d = json.load(open('gpurun_out/ab_tmp.json'))
    print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'], 3), 'e2e_ms', round(d['e2e']['ms_per_step'], 3),
          [(k['name'][:10], round(k['ms_per_step'], 3)) for k in d['kernel_classes'][:3]])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run_bench base_1 AF2_X=0
run_bench norcp AF2_LIB_PATH=$PWD/alphafold2_b200/csrc/libaf2b200_norcp.so
run_bench l2pf0 AF2_PROJ_L2PF=0
run_bench base_2 AF2_X=0
for cfg in "384 512" "512 1024"; do
  set -- $cfg
  for g in 1 0; do
    AF2_N=$1 AF2_S=$2 AF2_ATTN_GROUP=$g AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1
  done
done
AF2_N=256 AF2_S=128 AF2_PRECISION_BLOCK=strict AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1
cat $L
