#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_proj.py -q -m gpu --timeout 120 -p no:cacheprovider -x 2>&1 | tail -4
for d in 0 8; do
  AF2_PROJ_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_dbg_$d.csv \
     python tools/profile_block.py > /dev/null 2>&1; echo "dbg $d rc=$?"
done
