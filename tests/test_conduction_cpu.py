"""Heat conduction (<Thermal><Conduction/></Thermal>), the first transport task on the
scatter/gather skeleton of the step.

The DEVICE SOURCE (csrc/kernels_task.cuh: k_p2g_temperature, k_transport_nodal_value, k_transport_gradients, k_p2g_conduction,
k_transport_update, k_update_temperature) compiled for the host and run thread by thread against golden dumps of the unmodified
reference: nodal transport field and particle temperatures after every task of the first two steps, whole runs -- with one velocity
field (2D uGIMP USAVG, 2D lCPDI USL Neo-Hookean plane stress) and with material velocity fields + frictional contact in 3D."""
import numpy as np
import pytest

from nairn_mpm_fea_b200.problem import from_reference_dump
from tests.parity import COND_CASES, THERMAL_COND_CASES, THERMAL_OFFSET_CASES, check_multimaterial_run, check_multimaterial_tasks, load_golden
from tests.test_device_step_cpu import EmuSim, lib  # noqa: F401


@pytest.mark.parametrize("case", COND_CASES)
def test_device_source_tasks_match_reference(lib, case):  # noqa: F811
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z))
    check_multimaterial_tasks(sim, z, case, require="transport_value")
    sim.close()


@pytest.mark.parametrize("case", COND_CASES)
def test_device_source_whole_runs_match_reference(lib, case):  # noqa: F811
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z))
    check_multimaterial_run(sim, z, case)
    got = sim.download()
    last = max(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/temperature"))
    t0, t1 = z["p0/temperature"], z["p%d/temperature" % last]
    assert np.max(np.abs(t1 - t0)) > 1.0, "no heat moved"
    assert np.max(np.abs(got["temperature"] - t1)) <= 1e-9 * np.max(np.abs(t1))
    sim.close()


@pytest.mark.parametrize("case", THERMAL_COND_CASES + THERMAL_OFFSET_CASES)
def test_thermal_strains_match_reference(lib, case):  # noqa: F811
    """The temperature change of a step reaches the laws as ResidualStrains::dT (scaled per pass in USAVG): IsotropicMat (3D,
    plane strain, plane stress), IsoPlasticity (3D USL and plane stress, yielding), Neohookean (3D, plane stress) and Mooney
    (plane stress) with thermal expansion -- under conduction, and after a start off the stress-free temperature without any
    transport task.  Every task of two steps and the whole run; the residual energy is part of the compared energies."""
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z))
    check_multimaterial_tasks(sim, z, case, require="transport_value" if case in THERMAL_COND_CASES else "mass")
    sim.close()
    sim = EmuSim(lib, from_reference_dump(z))
    check_multimaterial_run(sim, z, case)
    last = max(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/energies"))
    if case != "th3d_adiabatic_johnsoncook":     # (without conduction the adiabatic rise moves both temperatures and never reaches res.dT)
        assert np.max(np.abs(z["p%d/energies" % last][1])) > 0.0, "no residual energy: the case does not exercise thermal strains"
    else:
        assert np.max(z["p%d/temperature" % last]) > 400.0, "no adiabatic heating"
    sim.close()


def test_heat_is_conserved_by_the_grid_update():
    """sum_i gVCT_i * rate_i = sum_i gQ_i = 0 for insulated bodies: the conduction flows of a particle sum to zero over its nodes
    (the shape-function gradients do).  Checked on the reference's own dump and therefore on everything that matches it."""
    z = load_golden("cond2d_disks_usavg")
    names = [str(s) for s in z["task_names"]]
    i = names.index("Extrapolate Grid Forces")
    q = z["s2/t%d/nodes/gQ" % i]
    assert abs(q.sum()) <= 1e-9 * np.abs(q).sum()


def test_problem_carries_the_conduction_settings():
    pr = from_reference_dump(load_golden("cond3d_blocks_multimaterial"))
    assert pr.conduction is not None and pr.multimaterial is not None
    assert np.all(pr.conduction["kcond"][:2] > 0) and set(np.unique(pr.particles["temperature"])) == {400.0, 280.0}
    assert np.all(pr.particles["energies"][5] == 300.0)        # pPreviousTemperature starts at the stress-free temperature
