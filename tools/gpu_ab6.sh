#!/bin/bash
# A/B of a second library build on one box: tools/gpu_ab6.sh <tag> <other.so> -> C2 forward (x2 each) + one C3 / C4 block per build
TAG=${1:-ab6}; OTHER=$PWD/alphafold2_b200/csrc/${2:-libaf2b200_onemma.so}
mkdir -p gpurun_out
L=gpurun_out/ab_${TAG}.log
: > $L
run_bench() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
  tail -2 gpurun_out/ab_tmp.err >> $L
  python - "$label" <<'PY' >> $L
import json, sys
try:
    d = json.loads([l for l in open('gpurun_out/ab_tmp.json') if l.startswith('{')][-1])
    print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'], 3), 'e2e_ms', round(d['e2e']['ms_per_step'], 3),
          [(k['name'][:10], round(k['ms_per_step'], 3)) for k in d['kernel_classes']])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run_bench new_a AF2_X=0
run_bench other_a AF2_LIB_PATH=$OTHER
run_bench new_b AF2_X=0
run_bench other_b AF2_LIB_PATH=$OTHER
for cfg in "512 1024" "384 512"; do set -- $cfg; AF2_N=$1 AF2_S=$2 AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1; AF2_LIB_PATH=$OTHER AF2_N=$1 AF2_S=$2 AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1; done
cat $L
