#!/bin/bash
# sort-interval sweep at ~100 steps (round 2): the step gets ~5 % slower as the compression wave develops
mkdir -p gpurun_out; : > gpurun_out/tune_sort_r2.txt
for si in "$@"; do
  python bench.py --steps 96 --warmup 5 --no-cpu-baseline --sort-interval $si 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sort interval', sys.argv[1], round(d['value']/1e9,3), 'G p-s/s', round(d['ms_per_step'],3), 'ms', {k:round(v,3) for k,v in d['roofline']['task_ms'].items() if v>0.1})" $si >> gpurun_out/tune_sort_r2.txt
done
cat gpurun_out/tune_sort_r2.txt
