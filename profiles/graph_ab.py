import os, sys, json
sys.path.insert(0, os.getcwd())
import bench
for wl, st, wu in (("disks2d", 400, 40), ("block1m", 100, 10)):
    r = bench.time_workload(wl, 0, st, wu)
    print("MPMGPU_GRAPHS=%s" % os.environ.get("MPMGPU_GRAPHS", "default"), wl, "%.4f ms/step" % r["ms_per_step"], "%.3f G" % (r["value"] / 1e9))
