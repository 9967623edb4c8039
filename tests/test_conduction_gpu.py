"""Heat conduction on the GPU through the C ABI (mpmgpu_set_conduction): nodal transport field and particle temperatures against
golden dumps of the unmodified reference -- every task of the first two steps (1e-10 / 1e-8), whole runs (1e-7) -- with one
velocity field and with material velocity fields + contact; conservation of heat; the refusals."""
import numpy as np
import pytest

from tests.parity import COND_CASES, THERMAL_COND_CASES, THERMAL_OFFSET_CASES, check_multimaterial_run, check_multimaterial_tasks, load_golden

pytestmark = pytest.mark.gpu


def make_sim(z, **kw):
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    prob = from_reference_dump(z)
    return MpmGpu(prob, device=0, **kw), prob


@pytest.mark.parametrize("case", COND_CASES)
def test_each_task_of_the_first_two_steps(case):
    z = load_golden(case)
    sim, _ = make_sim(z)
    check_multimaterial_tasks(sim, z, case, require="transport_value")
    sim.close()


@pytest.mark.parametrize("case", COND_CASES)
def test_whole_runs(case):
    z = load_golden(case)
    sim, _ = make_sim(z)
    check_multimaterial_run(sim, z, case)
    sim.close()


@pytest.mark.parametrize("case", THERMAL_COND_CASES + THERMAL_OFFSET_CASES)
def test_thermal_strains_match_reference(case):
    """Thermal strains in the laws (conduction with expanding materials; a start off the stress-free temperature): every task of
    two steps and the whole run against the reference, residual energy included."""
    z = load_golden(case)
    sim, _ = make_sim(z)
    check_multimaterial_tasks(sim, z, case, require="transport_value" if case in THERMAL_COND_CASES else "mass")
    sim.close()
    sim, _ = make_sim(z)
    check_multimaterial_run(sim, z, case)
    sim.close()


def test_insulated_bodies_keep_their_heat():
    """sum_p mp Cv T_p changes only by round-off: conduction moves heat between particles, the grid update conserves it
    (FLIP update: dT_p = dt sum_i S_ip rate_i, and sum_p mp Cv S_ip = gVCT_i)."""
    z = load_golden("cond2d_disks_usavg")
    sim, prob = make_sim(z)
    mp = np.asarray(prob.particles["mp"])
    cv = np.array([m["p"][1] for m in prob.materials])[np.asarray(prob.particles["matnum"]) - 1]
    h0 = float(np.sum(mp * cv * prob.particles["temperature"]))
    sim.step(50)
    got = sim.download()
    h1 = float(np.sum(mp * cv * got["temperature"]))
    assert abs(h1 - h0) <= 1e-10 * abs(h0)
    assert np.max(np.abs(got["temperature"] - prob.particles["temperature"])) > 1.0
    sim.close()


def test_refusals():
    from nairn_mpm_fea_b200 import MpmGpu, MpmGpuError
    from nairn_mpm_fea_b200.problem import from_reference_dump
    z = load_golden("cond2d_disks_usavg")
    prob = from_reference_dump(z)
    prob.materials[0]["p"][17] = 6.0e-5           # thermal expansion on the large-rotation IsotropicMat is the one combination not built
    prob.materials[0]["p"][7] = 1.0
    with pytest.raises(MpmGpuError, match="thermal expansion"):
        MpmGpu(prob, device=0)
    prob = from_reference_dump(z)
    with pytest.raises(MpmGpuError, match="per-task"):
        MpmGpu(prob, device=0, kernel_path=2)
    # (a mechanical FMPM(2) update with conduction runs: goldens cond3d_block_fmpm2_temperature_bcs, cond2d_disks_xpic2_usl)
