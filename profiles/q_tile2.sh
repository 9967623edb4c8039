python -m pytest tests -m gpu -x -q -k "whole_steps or slab" 2>&1 | tail -2
for cfg in "-DTILE_W=8" "-DTILE_W=7" "-DTILE_W=6"; do
  MPMGPU_NVCC_DEFS="$cfg" python nairn_mpm_fea_b200/build.py -f > /dev/null
  echo "== [$cfg]"
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['task_ms'].items() if v>0})"
done
