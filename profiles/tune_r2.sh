#!/bin/bash
# round-2 sweep on the GPU box: row-buffer width, launch bounds and run length of the sliding-window scatter
# (rebuilds libmpmgpu with -D flags, runs the 8M bench, prints task times)
mkdir -p gpurun_out
OUT=gpurun_out/tune_r2.txt
: > $OUT
run() {
  echo "== DEFS='$1' ENV='$2'" >> $OUT
  MPMGPU_NVCC_DEFS="$1" python nairn_mpm_fea_b200/build.py -f > /dev/null
  env $2 python bench.py --steps 24 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['task_ms'].items() if v>0})" >> $OUT 2>&1
}
for cfg in "$@"; do
  run "${cfg%%|*}" "${cfg#*|}"
done
python nairn_mpm_fea_b200/build.py -f > /dev/null
cat $OUT
