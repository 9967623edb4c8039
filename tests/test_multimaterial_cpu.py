"""Multimaterial mode (<MultiMaterialMode>): material velocity fields + material contact.

The DEVICE SOURCE of the path (csrc/kernels_task.cuh: field offsets in every particle<->grid kernel, k_p2g_contact_terms,
k_material_contact) compiled for the host and run thread by thread (tests/devlaws/host_step.cpp) against golden dumps of the
unmodified reference: every task of the first two steps for every material velocity field, and whole runs.  The compiled
kernels and capi.cu's orchestration are checked by tests/test_multimaterial_gpu.py on the GPU."""
import numpy as np
import pytest

from nairn_mpm_fea_b200.problem import from_reference_dump
from tests.parity import MM_CASES, check_multimaterial_run, check_multimaterial_tasks, load_golden
from tests.test_device_step_cpu import EmuSim, lib  # noqa: F401  (lib is the fixture that builds the host-compiled device source)


@pytest.mark.parametrize("case", MM_CASES)
def test_device_source_tasks_match_reference(lib, case):  # noqa: F811
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z))
    check_multimaterial_tasks(sim, z, case)
    sim.close()


@pytest.mark.parametrize("case", MM_CASES)
def test_device_source_whole_runs_match_reference(lib, case):  # noqa: F811
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z))
    shared = check_multimaterial_run(sim, z, case)
    assert shared >= 10, "the bodies never met"
    f = sim.flags()
    assert f["nan"] == 0 and f["cpdi_left"] == 0
    sim.close()


def test_problem_carries_the_reference_settings():
    """from_reference_dump turns the reference's per-material-pair law table into the per-field table of mpmgpu_multimaterial."""
    pr = from_reference_dump(load_golden("mm2d_friction_sn_powerlaw"))
    mm = pr.multimaterial
    assert mm["n_fields"] == 2 and mm["normal_method"] == 4 and mm["by_displacements"] == 0 and mm["position_cutoff"] == -0.7
    assert mm["law_kind"][0, 1] == 3 and mm["law_kind"][1, 0] == 3 and mm["law_friction"][0, 1] == 0.5
    assert abs(np.linalg.norm(mm["contact_normal"]) - 1.0) < 1e-12
    assert [m["kind"] for m in pr.materials] == [1, 1, 0]           # the contact law keeps its place in the materials list
    pr = from_reference_dump(load_golden("mm2d_ignore_lcpdi_usf"))
    assert pr.multimaterial["law_kind"][0, 1] == 0
    pr = from_reference_dump(load_golden("mm2d_stick_maxv_linear_usl"))
    assert pr.multimaterial["law_kind"][0, 1] == 1 and pr.multimaterial["normal_method"] == 1
