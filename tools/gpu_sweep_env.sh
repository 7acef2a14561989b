#!/bin/bash
# bench.py under several values of one environment variable, same box: tools/gpu_sweep_env.sh VAR v1 v2 ...
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${VAR}_${v}.json 2> gpurun_out/bench_${VAR}_${v}.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${VAR}_${v}.json'))
print('${VAR}=${v}', 'ms_per_step', round(d['ms_per_step'],3), [ (k['name'][:12], round(k['ms_per_step'],3)) for k in d.get('kernel_classes',[])][:3])
PY
done
