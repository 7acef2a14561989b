#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM pipelines: racecheck (shared-memory hazards between
# the warp roles), synccheck (barrier misuse) and memcheck on the C1-shape Evoformer block and on the n > 256 streamed-bias
# attention path.  Slow (10-100x): small shapes only, every tool under its own timeout.
# usage: tools/gpu_sanitize.sh [tag]        -> gpurun_out/sanitize_<tool>_<tag>.log + one summary line per tool
TAG=${1:-r02}
mkdir -p gpurun_out
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck; do
  AF2_SAN_CASE=all timeout ${SAN_TIMEOUT:-300} $SAN --tool $tool --print-limit 20 --error-exitcode 9 \
      python tools/sanitize_cases.py > gpurun_out/sanitize_${tool}_${TAG}.log 2>&1
  rc=$?
  echo "sanitize $tool rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK SUMMARY' gpurun_out/sanitize_${tool}_${TAG}.log | tail -1)"
done
