#!/bin/bash
# launch-bounds sweep on the GPU box (rebuilds libmpmgpu with -D flags, runs the 8M bench, prints task times)
mkdir -p gpurun_out
OUT=gpurun_out/tune_bounds_r1.txt
: > $OUT
for cfg in "-DF2_MINB=4 -DF4_MINB=4" "-DF2_MINB=5 -DF4_MINB=5" "-DF2_MINB=6 -DF4_MINB=6" "-DF2_MINB=5 -DF4_MINB=8" "-DF2_MINB=3 -DF4_MINB=3 -DF3_MINB=6"; do
  MPMGPU_NVCC_DEFS="$cfg" python nairn_mpm_fea_b200/build.py -f > /dev/null
  for P in 0 1; do
    echo "== $cfg PIPE=$P" >> $OUT
    MPMGPU_PIPE=$P python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['task_ms'].items() if v>0})" >> $OUT
  done
done
cat $OUT
