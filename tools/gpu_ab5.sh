#!/bin/bash
# A/B of one environment knob on one box: tools/gpu_ab5.sh <tag> <VAR> -> C2 forward (x2 each) + one C4 block per value
TAG=${1:-ab5}; VAR=${2:-AF2_ATTN_L2PF}
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
run_bench ${VAR}=1_a ${VAR}=1
run_bench ${VAR}=0_a ${VAR}=0
run_bench ${VAR}=1_b ${VAR}=1
run_bench ${VAR}=0_b ${VAR}=0
for v in 1 0; do env ${VAR}=$v AF2_N=512 AF2_S=1024 AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1; done
for v in 1 0; do env ${VAR}=$v AF2_N=384 AF2_S=512 AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1; done
cat $L
