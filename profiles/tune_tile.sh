#!/bin/bash
# sweep of the node-tile width / launch bounds on the GPU box (rebuilds libmpmgpu with -D flags, runs the 8M bench)
mkdir -p gpurun_out
OUT=gpurun_out/tune_tile_r1.txt
: > $OUT
for cfg in "" "-DTILE_W=8" "-DTILE_W=12" "-DF2_MINB=6" "-DF4_MINB=7" "-DF4_MINB=5" "-DF3_MINB=6"; do
  MPMGPU_NVCC_DEFS="$cfg" python nairn_mpm_fea_b200/build.py -f > /dev/null
  echo "== [$cfg]" >> $OUT
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['task_ms'].items() if v>0})" >> $OUT
done
python nairn_mpm_fea_b200/build.py -f > /dev/null
cat $OUT
