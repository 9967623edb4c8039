"""Reaction forces of the velocity BCs through the C ABI (mpmgpu_track_reactions / mpmgpu_download_reactions) against the
reference's NodalVelBC::TotalReactionForce per BC id (goldens react*): per-task kernels for every case, the fused 3D uGIMP path
where it applies (its N2 node sweep holds the grid-forces BC pass)."""
import pytest

from tests.parity import REACTION_CASES, check_reaction_run, load_golden

pytestmark = pytest.mark.gpu

FUSED = ["react3d_walls_ugimp", "react3d_rigid_piston_fmpm2", "react3d_rigid_wall_xpic2"]


@pytest.mark.parametrize("case,kernel_path", [(c, 1) for c in REACTION_CASES] + [(c, 2) for c in FUSED])
def test_reaction_forces_match_reference(case, kernel_path):
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    z = load_golden(case)
    prob = from_reference_dump(z)
    sim = MpmGpu(prob, device=0, kernel_path=kernel_path)
    sim.track_reactions()
    check_reaction_run(sim, prob, z, case)
    sim.close()


def test_reactions_need_tracking_and_refuse_slabs():
    from nairn_mpm_fea_b200 import MpmGpu, MpmGpuError
    from nairn_mpm_fea_b200.problem import from_reference_dump
    z = load_golden("react3d_walls_ugimp")
    prob = from_reference_dump(z)
    sim = MpmGpu(prob, device=0)
    sim.step(1)
    with pytest.raises(MpmGpuError, match="mpmgpu_track_reactions"):
        sim.reactions()
    sim.close()
