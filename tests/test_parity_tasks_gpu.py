"""GPU parity, through the C ABI, against golden dumps of the UNMODIFIED reference
(tests/golden/*.npz, made by tests/golden/make_golden.py with oracle/_ref).

1. every task of step 1 in isolation: node fields after each grid task, particle fields after each
   particle task, to 1e-10 of the field's max; element ids bit-exact;
2. whole steps: 1 step to 1e-10, N<=100 steps to 1e-7 (BASELINE.json north_star tolerances).
"""
import numpy as np
import pytest

from tests.parity import (TASK_MAP, TOL_1STEP, TOL_100STEP, compare_nodes, compare_particles, load_golden, per_task_steps,
                          xpic_for_step)

pytestmark = pytest.mark.gpu

CASES = ["trac3d_pressure_b2gimp", "trac2d_disks_b2cpdi", "trac3d_pressure_shear_ugimp", "trac3d_pressure_lcpdi_usl", "trac2d_disks_normal_tangent", "block3d_scgl", "disks2d_scgl_planestrain", "disks2d_symmetry_planes", "disks2d_rigid_plate", "block3d_rigid_mirrored", "block3d_lcpdi_rigid_wall", "disks2d_isoplastic_planestress", "block3d_pic", "block3d_fmpm1", "block3d_usavg_minus", "block3d_usavg_minus_xpic2", "block3d_usl_minus_fmpm2", "block3d_usf_fmpm2",
         "block3d_neohookean_av", "block3d_isoplastic_av", "block3d_rigid_wall", "block3d_rigid_piston", "block3d_rigid_linear_xpic2", "block3d_lcpdi_neo_xpic2", "block3d_lcpdi_rcrit", "disks2d_lcpdi", "disks2d_qcpdi", "block3d_xpic3", "block3d_fmpm2", "disks2d_fmpm3_neo", "block3d_neohookean", "block3d_neohookean_uj1", "block3d_isoplastic", "disks2d_neohookean", "disks2d_isoplastic",
         "disks2d_ugimp_planestrain", "disks2d_linear_planestress", "block3d_jitter", "block3d_ugimp_usavg", "block3d_fast_crossings", "block3d_gravity_damping", "block3d_linear_usl",
         "block3d_ugimp_usf"]


# the fused path: 3D uGIMP, any of the materials, FLIP/PIC and XPIC(k)/FMPM(k), rigid-BC particles
FUSED_CASES = [c for c in CASES if "linear" not in c and "2d" not in c and "cpdi" not in c and "mirrored" not in c and "scgl" not in c and "b2" not in c]


def make_sim(z, kernel_path=1, sort_interval=0):
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    prob = from_reference_dump(z)
    return MpmGpu(prob, device=0, kernel_path=kernel_path, sort_interval=sort_interval), prob


@pytest.mark.parametrize("case", CASES)
def test_each_task_of_step_one(case):
    z = load_golden(case)
    sim, prob = make_sim(z)
    names = [str(s) for s in z["task_names"]]
    for step in range(1, per_task_steps(z) + 1):
        x = xpic_for_step(z, step)
        if x:
            sim.set_xpic(*x)
        TOL = TOL_1STEP if step == 1 else 1.0e-8       # see tests/test_oracle_cpu.py
        for i, nm in enumerate(names):
            if TASK_MAP[nm] is None:
                continue
            sim.run_task(TASK_MAP[nm])
            pre = "s%d/t%d" % (step, i)
            nodes = sim.download_nodes()
            errs, bad = compare_nodes(nodes, z, pre + "/nodes", TOL)
            assert not bad, "%s step %d after task %d (%s): node fields %s" % (case, step, i, nm, bad)
            assert np.array_equal(nodes["number_points"] > 0, z[pre + "/nodes/numberPoints"] > 0), "active node set differs"
            got = sim.download()
            errs, bad = compare_particles(got, z, pre + "/p", TOL)
            assert not bad, "%s step %d after task %d (%s): particle fields %s" % (case, step, i, nm, bad)
            assert np.array_equal(got["in_elem"], z[pre + "/p/inElem"]), "element ids differ after %s" % nm
            assert np.array_equal(got["crossings"], z[pre + "/p/crossings"])
    sim.close()


@pytest.mark.parametrize("case,kernel_path,sort_interval",
                         [(c, 1, 0) for c in CASES] + [(c, 2, 0) for c in FUSED_CASES] +
                         [("block3d_jitter", 2, 1), ("block3d_fast_crossings", 2, 3), ("trac3d_pressure_shear_ugimp", 2, 1)])
def test_whole_steps(case, kernel_path, sort_interval):
    """kernel_path 1 = per-task kernels, 2 = fused dual-cell path (with its periodic physical sort)."""
    z = load_golden(case)
    sim, prob = make_sim(z, kernel_path, sort_interval)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        while done < s:
            x = xpic_for_step(z, done + 1)
            if x:
                sim.set_xpic(*x)
            sim.step(1)
            done += 1
        tol = TOL_1STEP if s == 1 else TOL_100STEP
        got = sim.download()
        errs, bad = compare_particles(got, z, "p%d" % s, tol)
        assert not bad, "%s after %d steps: %s (all: %s)" % (case, s, bad, errs)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s]), "%s: element ids differ after %d steps" % (case, s)
        assert np.array_equal(got["crossings"], z["p%d/crossings" % s])
        nodes = sim.download_nodes()
        errs, bad = compare_nodes(nodes, z, "n%d" % s, tol)
        assert not bad, "%s after %d steps: nodes %s" % (case, s, bad)
    st = sim.status()
    assert st["mstep"] == snaps[-1]
    sim.close()
