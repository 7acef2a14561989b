#!/bin/bash
# second A/B pass on one box: L2 residency experiments for the fp32 pair stream
TAG=${1:-ab2}
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
    d = json.load(open('gpurun_out/ab_tmp.json'))
    print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'], 3), 'e2e_ms', round(d['e2e']['ms_per_step'], 3),
          [(k['name'][:10], round(k['ms_per_step'], 3)) for k in d['kernel_classes']])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
run_bench base_1 AF2_X=0
run_bench l2persist AF2_L2_PERSIST=1
run_bench evict_last AF2_X_EVICT_LAST=1
run_bench both AF2_X_EVICT_LAST=1 AF2_L2_PERSIST=1
run_bench base_2 AF2_X=0
cat $L
