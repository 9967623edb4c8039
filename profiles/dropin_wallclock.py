"""Wall clock of the reference CLI against the drop-in driver (reference driver + libmpmgpu) on the same XML input:
config 2 (50^3-cell block, 1,000,000 particles, uGIMP, USAVG+), N steps, one archive at the end.  Both binaries run the
reference's own XML reader, generators, set-up and archiver; only the step tasks differ.  The step time is read from
the line the reference itself prints ("Elapsed Time per Step")."""
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import inputs  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "NairnMPM")
GPU = os.path.join(ROOT, "nairn_mpm_fea_b200", "host", "_build", "NairnMPM_gpu")


def run(binary, extra, xml):
    d = tempfile.mkdtemp(prefix="wall_")
    path = os.path.join(d, "in.fmcmd")
    open(path, "w").write(xml)
    t0 = time.perf_counter()
    p = subprocess.run([binary, *extra, path], cwd=d, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    m = re.search(r"Elapsed Time per Step:\s*([0-9.eE+-]+)\s*(\w+)", p.stdout)
    exe = re.search(r"Elapsed Time:\s*([0-9.eE+-]+)", p.stdout)
    steps = re.search(r"Calculation Steps:\s*(\d+)", p.stdout)
    return dt, (m.group(1) + " " + m.group(2)) if m else "?", (exe.group(1) if exe else "?") + " s over " + (steps.group(1) if steps else "?") + " steps", float(exe.group(1)), int(steps.group(1))


def main():
    ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    # maxtime (ms) so that nsteps steps run: the host generator applies the reference's CFL rule (tests/test_host_cpu.py)
    from nairn_mpm_fea_b200 import problem
    dt = problem.block3d(ncell=4, margin=7).dt          # seconds; same CFL rule
    ncpu = os.cpu_count() or 1
    for name, binary, extra, mult in (("reference CLI (-np %d)" % ncpu, REF, ("-np", str(ncpu)), (1,)),
                                      ("drop-in, per-task entry points", GPU, (), (10, 100)),
                                      ("drop-in, -fused", GPU, ("-fused",), (10, 100))):
        res = []
        for k in mult:
            xml = inputs.block3d(ncell=ncell, margin=7, maxtime=(k * nsteps - 0.5) * dt * 1e3)
            wall, per_step, exe, secs, steps = run(binary, extra, xml)
            res.append((secs, steps))
            print("%-34s process wall %.2f s   analysis %s   per step %s" % (name, wall, exe, per_step), flush=True)
        if len(res) == 2:       # marginal cost of a step: fixed costs (CUDA start-up, upload, archives) cancel
            print("%-34s marginal %.3f ms per step" % (name, 1e3 * (res[1][0] - res[0][0]) / (res[1][1] - res[0][1])), flush=True)


if __name__ == "__main__":
    main()
