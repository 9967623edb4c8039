"""Archive steps of the drop-in driver (reference driver + libmpmgpu): full download + the reference's own writer
(`-hostoutput`, what round 1 did) against records packed on the device (default).  Block of ncell^3 cells x 8 particles,
uGIMP, USAVG+, `-fused`; an archive every `every` steps.  Prints the reference's "Elapsed Time" of the analysis for both,
the adapter's GPU ARCHIVES line (bytes per archive = D2H traffic, time per archive including the file write) and checks
that both runs wrote the same archives.

    python profiles/dropin_archive.py [ncell=100] [steps=40] [every=10]
"""
import glob
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import inputs  # noqa: E402

GPU = os.path.join(ROOT, "nairn_mpm_fea_b200", "host", "_build", "NairnMPM_gpu")


def run(extra, xml):
    d = tempfile.mkdtemp(prefix="arch_")
    path = os.path.join(d, "in.fmcmd")
    open(path, "w").write(xml)
    t0 = time.perf_counter()
    p = subprocess.run([GPU, *extra, path], cwd=d, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    exe = re.search(r"Elapsed Time:\s*([0-9.eE+-]+)", p.stdout)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("GPU ARCHIVES")]
    return d, wall, float(exe.group(1)), line[0] if line else "(no device archives)"


def main():
    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    every = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    from nairn_mpm_fea_b200 import problem
    dt = problem.block3d(ncell=4, margin=7).dt * 1e3          # ms; same CFL rule as the reference
    xml = inputs.block3d(ncell=ncell, margin=7, maxtime=(nsteps - 0.5) * dt).replace(
        '<ArchiveTime units="ms">1000</ArchiveTime>', '<ArchiveTime units="ms">%r</ArchiveTime>' % ((every - 0.01) * dt))
    res = {}
    for name, extra in (("device-packed records", ("-fused",)), ("full download + host writer", ("-fused", "-hostoutput"))):
        d, wall, secs, line = run(extra, xml)
        files = sorted(glob.glob(os.path.join(d, "res", "blk.[0-9]*")))
        res[name] = (d, files)
        print("%-30s %d particles, %d steps, %d archives: process wall %.1f s, analysis %.2f s   %s" % (
            name, ncell ** 3 * 8, nsteps, len(files), wall, secs, line), flush=True)
    (da, fa), (db, fb) = res.values()
    assert [os.path.basename(f) for f in fa] == [os.path.basename(f) for f in fb] and len(fa) >= 2
    # two RUNS are compared here (FP64 atomics add in a different order from run to run, so the states differ in the last
    # bits); that the device packer writes the same BYTES as the host writer from one state is tests/test_zzz_archive_gpu.py
    import numpy as np
    n = ncell ** 3 * 8
    worst = 0.0
    for x, y in zip(fa, fb):
        a, b = np.fromfile(x, np.uint8), np.fromfile(y, np.uint8)
        assert a.size == b.size and np.array_equal(a[:64], b[:64]), "size or header of %s differs" % os.path.basename(x)
        ra, rb = a[64:].reshape(n, -1), b[64:].reshape(n, -1)
        nd = (ra.shape[1] - 16) // 8
        assert np.array_equal(ra[:, :4], rb[:, :4]) and np.array_equal(ra[:, 12:16], rb[:, 12:16]) and np.array_equal(ra[:, 16 + 8 * nd:], rb[:, 16 + 8 * nd:])
        da, db_ = ra[:, 16:16 + 8 * nd].copy().view(np.float64), rb[:, 16:16 + 8 * nd].copy().view(np.float64)
        scale = np.maximum(np.max(np.abs(db_), axis=0), 1e-300)
        worst = max(worst, float(np.max(np.abs(da - db_) / scale)))
    print("archives of the two runs: same headers, element ids, materials and crossing counters; doubles agree to %.1e of the column max" % worst)
    assert worst < 1e-9


if __name__ == "__main__":
    main()
