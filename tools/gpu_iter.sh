#!/bin/bash
# quick iteration: parity tests, bench line, per-launch list of one block
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
for f in test_gpu_kernels test_gpu_modules; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 120 -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/$f.log
  echo "== $f exit ${PIPESTATUS[0]}"; tail -15 gpurun_out/$f.log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms_per_step',d['ms_per_step'],'ms_per_block',d['ms_per_block'],'value',d['value'],'e2e',d['e2e']['value'])
print('roofline',d['roofline'])
for c in d['kernel_classes']: print(c)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_block.csv \
   python tools/profile_block.py > /dev/null 2>&1; echo "ncu block rc=$?"
