"""Pins the C restatement (oracle/mpm_oracle.c) against the golden dumps of the UNMODIFIED reference:
every task of step 1 (nodes + particles) and whole-step snapshots, same tolerances as the GPU tests.
Runs on CPU."""
import numpy as np
import pytest

from tests.parity import (LR3D_CASES, TASK_MAP, TOL_1STEP, compare_nodes, compare_particles, load_golden, per_task_steps,
                          tolerances, xpic_for_step)

CASES = ["disks2d_symmetry_planes", "block3d_isoplastic_nonlinear", "block3d_isoplastic_nonlinear2_soft", "block3d_johnsoncook", "disks2d_johnsoncook_planestress", "disks2d_nonlinear_planestrain_lr", "disks2d_nonlinear2_planestress", "block3d_b2spline", "block3d_b2gimp", "block3d_b2gimp_rigid_wall", "disks2d_b2spline", "disks2d_b2gimp_planestress", "block3d_b2cpdi", "disks2d_b2cpdi", "block3d_mooney", "block3d_mooney_uj2", "disks2d_mooney_planestrain", "disks2d_mooney_planestress", "disks2d_mooney_planestress_uj0", "block3d_free_ugimp", "block3d_free_lcpdi_xpic2", "disks2d_neo_planestress", "disks2d_neo_planestress_av", "block3d_isoplastic_softening", "block3d_material_pdamping", "block3d_isotropic_lr", "block3d_isoplastic_lr", "disks2d_lr_planestrain", "disks2d_lr_planestress", "disks2d_rigid_plate", "block3d_rigid_mirrored", "block3d_lcpdi_rigid_wall", "disks2d_isoplastic_planestress", "block3d_pic", "block3d_fmpm1", "block3d_usavg_minus", "block3d_usavg_minus_xpic2", "block3d_usl_minus_fmpm2", "block3d_usf_fmpm2",
         "block3d_neohookean_av", "block3d_isoplastic_av", "block3d_rigid_wall", "block3d_rigid_piston", "block3d_rigid_linear_xpic2", "block3d_lcpdi_neo_xpic2", "block3d_lcpdi_rcrit", "disks2d_lcpdi", "disks2d_qcpdi", "block3d_xpic3", "block3d_fmpm2", "disks2d_fmpm3_neo", "block3d_neohookean", "block3d_neohookean_uj1", "block3d_isoplastic", "disks2d_neohookean", "disks2d_isoplastic",
         "disks2d_ugimp_planestrain", "disks2d_linear_planestress", "block3d_jitter", "block3d_ugimp_usavg", "block3d_fast_crossings", "block3d_gravity_damping", "block3d_linear_usl",
         "block3d_ugimp_usf"]
TASK_INDEX = {"initialization": 0, "mass_and_momentum": 1, "post_extrapolation": 2, "update_strains_first": 3, "grid_forces": 4,
              "post_forces": 5, "update_momenta": 6, "update_particles": 7, "update_strains_last": 8, "reset_elements": 9,
              "project_rigid_bcs": 10}


def make(z):
    from nairn_mpm_fea_b200.problem import from_reference_dump
    from oracle.port import PortOracle
    return PortOracle(from_reference_dump(z))


@pytest.mark.parametrize("case", CASES)
def test_port_tasks_match_reference(case):
    z = load_golden(case)
    o = make(z)
    for step in range(1, per_task_steps(z) + 1):
        x = xpic_for_step(z, step)
        if x:
            o.set_xpic(*x)
        # 1e-10 holds for the first step; in later steps FMPM/XPIC iterations amplify round-off at nearly
        # massless edge nodes, so the per-task check of step 2 uses 1e-8 (the 100-step bound is 1e-7)
        TOL = tolerances(case)[0] if step == 1 else tolerances(case)[1]
        for i, nm in enumerate(str(s) for s in z["task_names"]):
            if TASK_MAP[nm] is None:
                continue
            o.run_task(TASK_INDEX[TASK_MAP[nm]])
            pre = "s%d/t%d" % (step, i)
            errs, bad = compare_nodes(o.download_nodes(), z, pre + "/nodes", TOL)
            assert not bad, "%s step %d task %d (%s): nodes %s" % (case, step, i, nm, bad)
            got = o.download()
            errs, bad = compare_particles(got, z, pre + "/p", TOL)
            assert not bad, "%s step %d task %d (%s): particles %s" % (case, step, i, nm, bad)
            assert np.array_equal(got["in_elem"], z[pre + "/p/inElem"])
    o.close()


@pytest.mark.parametrize("case", CASES)
def test_port_whole_steps_match_reference(case):
    z = load_golden(case)
    o = make(z)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        while done < s:
            x = xpic_for_step(z, done + 1)
            if x:
                o.set_xpic(*x)
            o.step(1)
            done += 1
        got = o.download()
        errs, bad = compare_particles(got, z, "p%d" % s, tolerances(case)[0] if s == 1 else tolerances(case)[2])
        assert not bad, "%s after %d steps: %s" % (case, s, bad)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
        assert np.array_equal(got["crossings"], z["p%d/crossings" % s])
    o.close()


@pytest.mark.parametrize("case", LR3D_CASES)
def test_large_rotation_3d_is_ill_conditioned_in_the_reference_algorithm(case):
    """Why the 3D large-rotation cases carry TOL_LR3D instead of 1e-10: move every particle velocity by ONE ulp and the
    restated reference algorithm (trigonometric eigenvalues inside the polar decomposition) answers with stresses that
    differ by ~1e-7..1e-6 after a single step -- four orders of magnitude above the standard tolerance and the same size as
    the restatement's own distance from the golden dump.  The small-rotation law on the same input moves by ~1e-16."""
    z = load_golden(case)
    from nairn_mpm_fea_b200.problem import from_reference_dump
    from oracle.port import PortOracle

    def one_step(bump, small_rotation=False):
        pr = from_reference_dump(z)
        if bump:
            pr.particles["vel"] = np.nextafter(pr.particles["vel"], np.inf)
        if small_rotation:
            for m in pr.materials:
                m["p"][7] = 0.0
        o = PortOracle(pr)
        o.step(1)
        g = o.download()
        o.close()
        return g

    def moved(a, b, k):
        scale = float(np.max(np.abs(a[k])))
        return float(np.max(np.abs(a[k] - b[k]))) / scale if scale > 0.0 else 0.0

    lr = max(moved(one_step(False), one_step(True), k) for k in ("sp", "pressure"))
    sr = max(moved(one_step(False, True), one_step(True, True), k) for k in ("sp", "pressure"))
    assert lr > 1.0e3 * TOL_1STEP, lr
    assert sr < TOL_1STEP * 1.0e-3, sr
