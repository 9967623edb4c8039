#!/bin/bash
TAG=${1:-r1e}
mkdir -p gpurun_out
MPMGPU_NVCC_DEFS="-DTILE_W=8" python nairn_mpm_fea_b200/build.py -f > /dev/null
ncu --set full --clock-control none --import-source on -k regex:'k_f[1-4]' -s 8 -c 4 -o gpurun_out/fused_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncell 64 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -3
