"""GPU parity, through the C ABI, against golden dumps of the unmodified reference for the cases added late in round 1:
the large-rotation hypoelastic laws (Elastic::useLargeRotation: IsotropicMat::LRConstitutiveLaw and IsoPlasticity with
LRGetStrainIncrement; verified on a B200 before the round's GPU budget ran out) and four cases that reach branches of the
laws no earlier golden did (Neo-Hookean in plane stress with the three U(J) options, 2D artificial viscosity, softening down
to a minimum yield stress, a material's own particle damping).  Same checks as tests/test_parity_tasks_gpu.py.

2D cases: the standard tolerances (1e-10 after one step, 1e-7 after 100).  3D cases: TOL_LR3D -- the reference's own
polar decomposition is ill-conditioned for small strain increments, see tests/parity.py and
tests/test_oracle_cpu.py::test_large_rotation_3d_is_ill_conditioned_in_the_reference_algorithm."""
import numpy as np
import pytest

from tests.parity import TASK_MAP, compare_nodes, compare_particles, load_golden, tolerances

pytestmark = pytest.mark.gpu

LR_CASES = ["block3d_isotropic_lr", "block3d_isoplastic_lr", "disks2d_lr_planestrain", "disks2d_lr_planestress"]
BRANCH_CASES = ["disks2d_neo_planestress", "disks2d_neo_planestress_av", "block3d_isoplastic_softening", "block3d_material_pdamping",
                "block3d_free_ugimp", "block3d_free_lcpdi_xpic2",          # free flight, no grid BCs
                "block3d_mooney", "block3d_mooney_uj2", "disks2d_mooney_planestrain", "disks2d_mooney_planestress", "disks2d_mooney_planestress_uj0",
                # hardening laws returned numerically
                "block3d_isoplastic_nonlinear", "block3d_isoplastic_nonlinear2_soft", "block3d_johnsoncook", "disks2d_johnsoncook_planestress",
                "disks2d_nonlinear_planestrain_lr", "disks2d_nonlinear2_planestress",
                # quadratic B-spline shape functions
                "block3d_b2spline", "block3d_b2gimp", "block3d_b2gimp_rigid_wall", "disks2d_b2spline", "disks2d_b2gimp_planestress",
                "block3d_b2cpdi", "disks2d_b2cpdi"]
CASES = LR_CASES + BRANCH_CASES
FUSED_CASES = ["block3d_isoplastic_softening", "block3d_material_pdamping", "block3d_free_ugimp"]          # 3D uGIMP without large rotation


def make_sim(z, kernel_path=0):
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    return MpmGpu(from_reference_dump(z), device=0, kernel_path=kernel_path)


@pytest.mark.parametrize("case", CASES)
def test_each_task_of_step_one(case):
    z = load_golden(case)
    from tests.parity import xpic_for_step
    sim = make_sim(z, 1)
    tol = tolerances(case)[0]
    x = xpic_for_step(z, 1)
    if x:
        sim.set_xpic(*x)
    for i, nm in enumerate(str(s) for s in z["task_names"]):
        if TASK_MAP[nm] is None:
            continue
        sim.run_task(TASK_MAP[nm])
        pre = "s1/t%d" % i
        errs, bad = compare_nodes(sim.download_nodes(), z, pre + "/nodes", tol)
        assert not bad, "%s after task %d (%s): node fields %s" % (case, i, nm, bad)
        got = sim.download()
        errs, bad = compare_particles(got, z, pre + "/p", tol)
        assert not bad, "%s after task %d (%s): particle fields %s" % (case, i, nm, bad)
        assert np.array_equal(got["in_elem"], z[pre + "/p/inElem"])
    sim.close()


@pytest.mark.parametrize("case,kernel_path", [(c, 0) for c in LR_CASES] + [(c, 1) for c in BRANCH_CASES] + [(c, 2) for c in FUSED_CASES])
def test_whole_steps(case, kernel_path):
    """kernel_path 0 (auto): a large-rotation material sends the run to the per-task kernels; 1 = per-task, 2 = fused."""
    z = load_golden(case)
    sim = make_sim(z, kernel_path)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    from tests.parity import xpic_for_step
    done = 0
    for s in snaps:
        while done < s:
            x = xpic_for_step(z, done + 1)
            if x:
                sim.set_xpic(*x)
            sim.step(1)
            done += 1
        tol = tolerances(case)[0] if s == 1 else tolerances(case)[2]
        got = sim.download()
        errs, bad = compare_particles(got, z, "p%d" % s, tol)
        assert not bad, "%s after %d steps: %s (all: %s)" % (case, s, bad, errs)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
        assert np.array_equal(got["crossings"], z["p%d/crossings" % s])
        errs, bad = compare_nodes(sim.download_nodes(), z, "n%d" % s, tol)
        assert not bad, "%s after %d steps: nodes %s" % (case, s, bad)
    sim.close()


def test_fused_path_refuses_large_rotation():
    from nairn_mpm_fea_b200.capi import MpmGpuError
    z = load_golden("block3d_isotropic_lr")
    with pytest.raises(MpmGpuError):
        make_sim(z, 2)


# ---- CPDI kernels with the corners' contributions merged per node (opt-in MPMGPU_CPDI_MERGE=1, shape.cuh) ----------------
CPDI_CASES = ["block3d_free_lcpdi_xpic2", "block3d_lcpdi_neo_xpic2", "block3d_lcpdi_rcrit", "block3d_lcpdi_rigid_wall", "disks2d_lcpdi", "disks2d_qcpdi"]


@pytest.mark.parametrize("case", CPDI_CASES)
def test_cpdi_merged_whole_steps(case, monkeypatch):
    """Same goldens and tolerances as the plain CPDI kernels (tests/test_parity_tasks_gpu.py): merging only changes the order in
    which a node's corner contributions are summed."""
    from tests.parity import xpic_for_step
    monkeypatch.setenv("MPMGPU_CPDI_MERGE", "1")
    z = load_golden(case)
    sim = make_sim(z, 1)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        while done < s:
            x = xpic_for_step(z, done + 1)
            if x:
                sim.set_xpic(*x)
            sim.step(1)
            done += 1
        tol = tolerances(case)[0] if s == 1 else tolerances(case)[2]
        got = sim.download()
        errs, bad = compare_particles(got, z, "p%d" % s, tol)
        assert not bad, "%s (merged CPDI) after %d steps: %s" % (case, s, bad)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
        errs, bad = compare_nodes(sim.download_nodes(), z, "n%d" % s, tol)
        assert not bad, "%s (merged CPDI) after %d steps: nodes %s" % (case, s, bad)
    sim.close()
