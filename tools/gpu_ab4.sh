#!/bin/bash
# A/B of the polynomial softmax exponentials on one box (C2 forward + one C4 block)
TAG=${1:-ab4}
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
NP=$PWD/alphafold2_b200/csrc/libaf2b200_nopoly.so
run_bench poly_a AF2_X=0
run_bench nopoly_a AF2_LIB_PATH=$NP
run_bench poly_b AF2_X=0
run_bench nopoly_b AF2_LIB_PATH=$NP
AF2_N=512 AF2_S=1024 AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1
AF2_LIB_PATH=$NP AF2_N=512 AF2_S=1024 AF2_ITERS=5 timeout 300 python tools/time_block.py >> $L 2>&1
cat $L
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_configs.py tests/test_gpu_next_rows.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | tail -4
