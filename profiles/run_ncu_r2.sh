#!/bin/bash
# Round-2 captures of the shipped kernels (run under gpurun on one B200).  $1 = tag
#  1. launch list of the default bench command (serialised, cold-cache: compare SHARES of the step, not absolutes)
#  2. one --set full capture of each fused kernel at the bench size (8M particles), right after a physical sort
TAG=${1:-r2a}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_f[1-4]|k_n[1-3]' -s 14 -c 7 -o gpurun_out/fused_$TAG -f \
    --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_op_red.sum,smsp__inst_executed_op_global_red.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/fused_$TAG.ncu-rep --page raw --csv > gpurun_out/fused_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
