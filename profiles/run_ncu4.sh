#!/bin/bash
# round-1 final captures: launch list of the default bench + one full capture of each fused kernel
TAG=${1:-r1d}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_f[1-4]|k_n[1-3]' -s 14 -c 7 -o gpurun_out/fused_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncell 64 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -4
