"""Multimaterial mode on the GPU, through the C ABI (mpmgpu_set_multimaterial): material velocity fields + material contact
against golden dumps of the unmodified reference (tests/golden/mm*.npz) -- every task of the first two steps for every field
(1e-10 / 1e-8), whole runs (1e-7), element ids exact -- and the refusals of what is not built."""
import numpy as np
import pytest

from tests.parity import MM_CASES, check_multimaterial_run, check_multimaterial_tasks, load_golden

pytestmark = pytest.mark.gpu


def make_sim(z, **kw):
    from nairn_mpm_fea_b200 import MpmGpu
    from nairn_mpm_fea_b200.problem import from_reference_dump
    prob = from_reference_dump(z)
    return MpmGpu(prob, device=0, **kw), prob


@pytest.mark.parametrize("case", MM_CASES)
def test_each_task_of_the_first_two_steps(case):
    z = load_golden(case)
    sim, _ = make_sim(z)
    check_multimaterial_tasks(sim, z, case)
    sim.close()


@pytest.mark.parametrize("case", MM_CASES)
def test_whole_runs(case):
    z = load_golden(case)
    sim, _ = make_sim(z)
    shared = check_multimaterial_run(sim, z, case)
    assert shared >= 10, "the bodies never met"
    sim.close()


def test_contact_conserves_momentum_and_ignore_equals_one_field():
    """Pair contact moves momentum between the two fields of a node, never creates any: the total grid momentum after the
    momentum update equals the particles' momentum plus dt times the total force.  With the law 'ignore' every material is
    moved to the centre-of-mass velocity: the particles end up where the single-field run of the same input puts them."""
    z = load_golden("mm2d_friction_avgg")
    sim, prob = make_sim(z)
    sim.step(30)
    for name in ("initialization", "mass_and_momentum", "post_extrapolation", "update_strains_first", "grid_forces", "post_forces", "update_momenta"):
        sim.run_task(name)
    nd, pt = sim.download_nodes(), sim.download()
    mp = np.asarray(prob.particles["mp"])
    ptot = (pt["vel"] * mp).sum(axis=1)
    gtot = nd["pk"].sum(axis=1) - prob.dt * nd["ftot"].sum(axis=1)          # the momentum update added ftot*dt; contact only redistributes
    assert np.allclose(gtot[:2], ptot[:2], rtol=1e-9, atol=1e-12 * np.abs(ptot).max())
    sim.close()
    z = load_golden("mm2d_ignore_lcpdi_usf")
    sim, prob = make_sim(z)
    from nairn_mpm_fea_b200 import MpmGpu
    import copy
    single = copy.copy(prob)
    single.multimaterial = None
    one = MpmGpu(single, device=0)
    sim.step(25); one.step(25)
    a, b = sim.download(), one.download()
    assert np.max(np.abs(a["pos"] - b["pos"])) <= 1e-9 * np.max(np.abs(b["pos"]))
    assert np.max(np.abs(a["vel"] - b["vel"])) <= 1e-7 * np.max(np.abs(b["vel"]))
    sim.close(); one.close()


def test_refusals():
    from nairn_mpm_fea_b200 import MpmGpu, MpmGpuError
    from nairn_mpm_fea_b200.problem import from_reference_dump
    z = load_golden("mm3d_two_blocks_maxv_friction_ugimp")
    for change, words in ((dict(normal_method=6), "regression"), (dict(normal_method=5), "regression")):
        prob = from_reference_dump(z)
        prob.multimaterial.update(change)
        with pytest.raises(MpmGpuError, match=words):
            MpmGpu(prob, device=0)
    prob = from_reference_dump(z)
    prob.multimaterial["law_kind"] = np.full((2, 2), 7, np.int32)
    with pytest.raises(MpmGpuError, match="contact law"):
        MpmGpu(prob, device=0)
    prob = from_reference_dump(z)
    with pytest.raises(MpmGpuError, match="per-task"):
        MpmGpu(prob, device=0, kernel_path=2)
    prob = from_reference_dump(z)
    prob.xpic_order, prob.using_fmpm = 2, True
    with pytest.raises(MpmGpuError, match="order > 1"):
        MpmGpu(prob, device=0)
    sim = MpmGpu(from_reference_dump(z), device=0)
    with pytest.raises(MpmGpuError, match="order > 1"):
        sim.set_xpic(2, 1)
    sim.close()
