"""The combinatorial parity sweep of tests/test_sweep_cpu.py on the GPU, through the C ABI: the same seeded random combinations,
run by the unmodified reference (oracle/_ref, live) and by libmpmgpu -- on the per-task kernels, and on the fused dual-cell path
whenever the combination is eligible for it (3D uGIMP, no large-rotation material).  Runs last: it was written after the
round-1 GPU budget was spent, so its first execution is the driver's."""
import numpy as np
import pytest

from oracle import refharness
from tests.parity import TOL_1STEP, TOL_100STEP, TOL_ITERATIVE, TOL_LR3D, compare_nodes, compare_particles, xpic_for_step
from tests.test_sweep_cpu import NSTEPS, make_config

pytestmark = pytest.mark.gpu

NCONFIG = 40


@pytest.mark.parametrize("seed", range(NCONFIG))
def test_random_combination_on_the_gpu_matches_the_live_reference(seed):
    if not refharness.available():
        pytest.skip("oracle/_ref not built")
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    xml, (ja, va), lr3d, desc = make_config(1000 + seed)
    try:
        z = refharness.run_reference(xml, snaps=(1, NSTEPS), per_task_steps=0, nprocs=1, jitter_amp=ja, vel_amp=va)
    except RuntimeError as e:
        if "could not be bracketed" in str(e) or "position nan" in str(e):
            pytest.skip("the reference itself aborts on this combination: %s" % desc)
        raise
    prob = from_reference_dump(z)
    lr = any(m["p"][7] != 0.0 or m["kind"] == 8 or (m["kind"] == 9 and m["p"][16] > 1.0) for m in prob.materials)          # extended law dispatch: per-task kernels only
    mirrored = any(m["kind"] == 11 and m["p"][9] != 0.0 for m in prob.materials)
    fused_ok = prob.is3d and prob.shape == 1 and not lr and not mirrored
    for kernel_path in (1, 2) if fused_ok else (1,):
        sim = MpmGpu(prob, device=0, kernel_path=kernel_path, sort_interval=5 if kernel_path == 2 else 0)
        done = 0
        for s in (1, NSTEPS):
            while done < s:
                x = xpic_for_step(z, done + 1)
                if x:
                    sim.set_xpic(*x)
                sim.step(1)
                done += 1
            tol = TOL_LR3D if lr3d else (TOL_ITERATIVE if "<Hardening>Nonlinear" in xml or "<Hardening>JohnsonCook" in xml else (TOL_1STEP if s == 1 else TOL_100STEP))
            got = sim.download()
            errs, bad = compare_particles(got, z, "p%d" % s, tol)
            assert not bad, "[%s, kernel_path %d] after %d steps: particles %s" % (desc, kernel_path, s, bad)
            assert np.array_equal(got["in_elem"], z["p%d/inElem" % s]), desc
            assert np.array_equal(got["crossings"], z["p%d/crossings" % s]), desc
            errs, bad = compare_nodes(sim.download_nodes(), z, "n%d" % s, tol)
            assert not bad, "[%s, kernel_path %d] after %d steps: nodes %s" % (desc, kernel_path, s, bad)
        sim.close()
