"""Tuning helper (GPU box): event counters of the sliding-window scatter on the bench workload.
Needs a library built with -DSLIDE_STATS:  MPMGPU_NVCC_DEFS=-DSLIDE_STATS python nairn_mpm_fea_b200/build.py -f"""
import ctypes
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from nairn_mpm_fea_b200 import MpmGpu, capi

wl = sys.argv[1] if len(sys.argv) > 1 else "block1m"
pr, _ = bench.make_problem(wl)
sim = MpmGpu(pr, device=0)
lib = capi.load_library()
out = (ctypes.c_ulonglong * 8)()
names = ["groups", "same-cell", "retire1", "retire2", "restarts", "strays", "flushes", "flushed cols"]
n = pr.particles["pos"].shape[-1] if hasattr(pr, "particles") else 0
for step in range(1, 14):
    sim.step(1)
    lib.mpmgpu_debug_slide_stats(out, 1)
    print("step %2d " % step + "  ".join("%s %d" % (a, b) for a, b in zip(names, out)), flush=True)
sim.close()
