#!/bin/bash
# N-GPU check of the sharded trunk + scaling bench (an N-GPU gpurun call is charged N x its wall time: keep the timeouts tight)
# usage: [ONLY_CHECK=1] tools/gpu_parallel.sh <N> <tag> [C4]
N=${1:-2}; TAG=${2:-r02}; BIG=${3:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29551 tests/parallel_check_multi_gpu.py > gpurun_out/parallel_check_${N}gpu_${TAG}.log 2>&1; echo "parallel_check rc=$?"
grep -E '^\{' gpurun_out/parallel_check_${N}gpu_${TAG}.log | cut -c1-330 | head -6; grep -v '^{' gpurun_out/parallel_check_${N}gpu_${TAG}.log | tail -4
[ -n "$ONLY_CHECK" ] && exit 0
timeout 200 $TR --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_C2_${N}gpu_${TAG}.json 2> gpurun_out/bench_C2_${N}gpu_${TAG}.err; echo "bench C2 N=$N rc=$?"; tail -2 gpurun_out/bench_C2_${N}gpu_${TAG}.err
if [ "$N" -le 2 ] || [ -n "$AB_NCCL" ]; then
  AF2_PEER_EXCHANGE=0 timeout 240 $TR --master-port 29555 tests/parallel_check_multi_gpu.py > gpurun_out/parallel_check_${N}gpu_nccl_${TAG}.log 2>&1; echo "parallel_check (NCCL all_to_all) rc=$?"
  AF2_PEER_EXCHANGE=0 timeout 200 $TR --master-port 29556 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_${N}gpu_nccl_${TAG}.json 2> /dev/null; echo "bench C2 N=$N (NCCL all_to_all) rc=$?"
fi
if [ "$N" -le 2 ] && [ -z "$SKIP_PIECES" ]; then
  AF2_GATHER_FUSED=0 timeout 200 $TR --master-port 29553 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_C2_${N}gpu_pieces_${TAG}.json 2> /dev/null; echo "bench C2 N=$N (per-piece launches) rc=$?"
fi
if [ -n "$BIG" ]; then
  timeout 300 $TR --master-port 29554 bench.py --gpus $N --steps 5 --warmup 3 --workload C4 --no-cpu-baseline > gpurun_out/bench_C4_${N}gpu_${TAG}.json 2> gpurun_out/bench_C4_${N}gpu_${TAG}.err; echo "bench C4 N=$N rc=$?"; tail -2 gpurun_out/bench_C4_${N}gpu_${TAG}.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_C*_${N}gpu*_${TAG}.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms_per_step', round(d['ms_per_step'], 3), 'value', round(d['value']), 'launches', d['gpu_launches'], d.get('schedule'), (d.get('exchange') or {}).get('all_to_all', '')[:10], 'replicas', (d.get('replicas') or {}).get('value'),
              [(k['name'][:10], k['launches_per_step'], round(k['ms_per_step'], 2)) for k in d['kernel_classes']])
    except Exception as e:
        print(f, 'unreadable', e)
PY
