#!/bin/bash
# One measuring pass on a 1-GPU box: parity suite, bench (both arms), ncu launch list of the bench command, `ncu --set full`
# of one block at C2 and at C3 (BASELINE config 3), sanitizer.  Everything lands in gpurun_out/ under <tag>.
# usage: tools/gpu_round.sh <tag> [tests|bench|ncu|san ...]   (default: all stages)
TAG=${1:-r02}; shift
STAGES=${@:-tests bench ncu san}
mkdir -p gpurun_out
has() { [[ " $STAGES " == *" $1 "* ]]; }
if has tests; then
  rm -f gpurun_out/parity_report.jsonl
  timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/test_gpu_${TAG}.log
  echo "tests rc=${PIPESTATUS[0]}"; grep -E 'FAILED|ERROR|passed|failed' gpurun_out/test_gpu_${TAG}.log | tail -30
  cp gpurun_out/parity_report.jsonl gpurun_out/parity_report_${TAG}.jsonl 2>/dev/null
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "bench ref rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${TAG}.json'))
    print('ms_per_step',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'roofline',d['roofline']['frac'],'whole',d['roofline_whole_step']['frac'])
    print('cpu',d.get('cpu_baseline')); print('gpu_torch',d.get('gpu_torch_baseline'))
    for c in d['kernel_classes']: print(c)
    print(open('gpurun_out/bench_ref_${TAG}.json').read()[:600])
except Exception as e: print('no bench', e)
PY
fi
if has ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "ncu bench list rc=$?"
  for CFGN in C2:256:128 C3:384:512; do
    IFS=: read -r name N S <<< "$CFGN"
    AF2_N=$N AF2_S=$S timeout 900 ncu --set full --clock-control none --profile-from-start off \
       -o /tmp/prof_block_${name}_${TAG} -f python tools/profile_block.py > gpurun_out/ncu_block_${name}_${TAG}.log 2>&1; echo "ncu block $name rc=$?"
    ncu -i /tmp/prof_block_${name}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_block_${name}_${TAG}_raw.csv 2>/dev/null
  done
fi
if has san; then
  bash tools/gpu_sanitize.sh ${TAG}
fi
ls gpurun_out | grep ${TAG} | head -40
