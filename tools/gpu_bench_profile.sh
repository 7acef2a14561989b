#!/bin/bash
# bench line + ncu launch list of the same command + full captures of the three tensor-core kernels
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_block.csv \
   python tools/profile_block.py > /dev/null 2>&1; echo "ncu block rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernelILi256ELi4ELb0 -s 25 -c 1 \
   -o gpurun_out/prof_gemm_ffw1 -f python tools/profile_block.py > /dev/null 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernelILi256ELi4ELb0 -s 13 -c 1 \
   -o gpurun_out/prof_gemm_chan -f python tools/profile_block.py > /dev/null 2>&1; echo "ncu chan rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc -s 2 -c 1 \
   -o gpurun_out/prof_attention -f python tools/profile_block.py > /dev/null 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out
