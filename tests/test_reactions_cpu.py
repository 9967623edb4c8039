"""Reaction forces of the velocity BCs (NodalVelBC::freaction, the "reactionx/y/z" global quantities): the device source run on
the host (tests/devlaws/host_step.cpp) against the reference's NodalVelBC::TotalReactionForce per BC id, after every task-by-task
step and at the snapshots -- grid BCs with ids (fixed, moving, skewed), rigid-particle BCs (id = material), FMPM(2) and XPIC(2)
(lumped additions of the particle-update pass), and multimaterial mode (the fields of a node add into one BC)."""
import numpy as np
import pytest

from nairn_mpm_fea_b200.problem import from_reference_dump
from tests.parity import REACTION_CASES, check_reaction_run, load_golden
from tests.test_device_step_cpu import EmuSim, lib  # noqa: F401


@pytest.mark.parametrize("case", REACTION_CASES)
def test_reaction_forces_match_reference(lib, case):  # noqa: F811
    z = load_golden(case)
    prob = from_reference_dump(z)
    sim = EmuSim(lib, prob)
    check_reaction_run(sim, prob, z, case)
    sim.close()


def test_reaction_goldens_exercise_every_kind():
    ids = {c: [int(i) for i in load_golden(c)["reaction_ids"]] for c in REACTION_CASES}
    assert ids["react3d_walls_ugimp"][:4] == [-4, -3, -2, -1]
    z = load_golden("react3d_rigid_piston_fmpm2")
    r = z["reaction30"]
    assert np.abs(r[2]).max() > 0 and np.abs(r[0] - r[2]).max() > 0        # the rigid material's BCs and the grid BCs both react
    assert np.abs(load_golden("react2d_multimaterial_wall")["reaction40"]).max() > 0
