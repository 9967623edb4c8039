#!/bin/bash
TAG=${1:-r1c}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_f4_pipe|k_f2' -s 4 -c 2 -o gpurun_out/fused_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncell 64 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -3
