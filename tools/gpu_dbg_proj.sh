#!/bin/bash
mkdir -p gpurun_out
for d in 0 256 264; do
  AF2_PROJ_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_dbg_$d.csv \
     python tools/profile_block.py > /dev/null 2>&1; echo "dbg $d rc=$?"
done
