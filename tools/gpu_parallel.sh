#!/bin/bash
# 2-GPU check of the sharded trunk + scaling bench at N=2 (a 2-GPU gpurun call is charged twice its wall time: keep the timeouts tight)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x --timeout 200 -p no:cacheprovider 2>&1 | tail -8
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/parallel_check_multi_gpu.py > gpurun_out/parallel_check.log 2>&1; echo "parallel_check rc=$?"; grep -E '^\{' gpurun_out/parallel_check.log | head -12; tail -5 gpurun_out/parallel_check.log | grep -v '^{'
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -c 1500 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
