#!/bin/bash
# sort interval / tile width on the moving block8m workload
mkdir -p gpurun_out
OUT=gpurun_out/tune_sort_r1.txt
: > $OUT
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$2', round(d['value']/1e9,3), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['task_ms'].items() if v>0})" >> $OUT; }
for si in 5 10 15 25 50; do run "TILE_W=8" "--sort-interval $si"; done
MPMGPU_NVCC_DEFS="-DTILE_W=10" python nairn_mpm_fea_b200/build.py -f > /dev/null
for si in 10 25; do run "TILE_W=10" "--sort-interval $si"; done
MPMGPU_NVCC_DEFS="-DTILE_W=12" python nairn_mpm_fea_b200/build.py -f > /dev/null
for si in 10 25; do run "TILE_W=12" "--sort-interval $si"; done
python nairn_mpm_fea_b200/build.py -f > /dev/null
cat $OUT
