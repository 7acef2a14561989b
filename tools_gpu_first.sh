#!/bin/bash
# first GPU bring-up: each test file separately under a timeout so a hang cannot eat the session
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/parity_report.jsonl
for f in test_gpu_kernels test_gpu_modules; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 120 -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/$f.log
  echo "== $f exit ${PIPESTATUS[0]}"; tail -40 gpurun_out/$f.log
done
