#!/bin/bash
# tests + bench + ncu launch list of the bench command + full ncu captures of one C2 block's kernels
# usage: tools/gpu_profile_full.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/test_gpu_${TAG}.log; echo "tests rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/test_gpu_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1; echo "ncu bench list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_block_${TAG}.csv \
   python tools/profile_block.py > /dev/null 2>&1; echo "ncu block list rc=$?"
timeout 900 ncu --set full --clock-control none --profile-from-start off \
   -o /tmp/prof_block_${TAG} -f python tools/profile_block.py > gpurun_out/ncu_block_${TAG}.log 2>&1; echo "ncu block full rc=$?"
ncu -i /tmp/prof_block_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_block_${TAG}_raw.csv 2>/dev/null; ls -la /tmp/prof_block_${TAG}.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -s 2 -c 1 \
   -o gpurun_out/prof_attention_${TAG} -f python tools/profile_block.py > /dev/null 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out | head -40
