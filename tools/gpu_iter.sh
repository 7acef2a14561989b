#!/bin/bash
# quick iteration: new-kernel tests first (under a hard timeout), then the full parity suite and the bench line
# usage: tools/gpu_iter.sh [tag]
TAG=${1:-iter}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_proj.py -q -m gpu --timeout 120 -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/test_proj_${TAG}.log
echo "== proj tests exit ${PIPESTATUS[0]}"; tail -25 gpurun_out/test_proj_${TAG}.log
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x --deselect tests/test_gpu_proj.py 2>&1 | tail -30 > gpurun_out/test_gpu_${TAG}.log
echo "== all tests exit ${PIPESTATUS[0]}"; tail -12 gpurun_out/test_gpu_${TAG}.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${TAG}.json'))
    print('ms_per_step',d['ms_per_step'],'ms_per_block',d['ms_per_block'],'value',d['value'],'e2e',d['e2e']['value'])
    for c in d['kernel_classes']: print(c)
except Exception as e: print('no bench', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_block_${TAG}.csv \
   python tools/profile_block.py > /dev/null 2>&1; echo "ncu block rc=$?"
