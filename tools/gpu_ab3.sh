#!/bin/bash
# A/B of programmatic dependent launch on one box
TAG=${1:-ab3}
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
run_bench pdl1_a AF2_PDL=1
run_bench pdl0_a AF2_PDL=0
run_bench pdl1_b AF2_PDL=1
run_bench pdl0_b AF2_PDL=0
run_bench pdl1_l2pf0 AF2_PDL=1 AF2_PROJ_L2PF=0
cat $L
