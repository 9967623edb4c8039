"""The drop-in driver (reference driver + GpuTasks.cpp) must refuse, loudly and before touching the device, every input whose
replaced CPU tasks would have done something libmpmgpu does not do -- never a silently different run.  These checks
happen before mpmgpu_create, so they run without a GPU; an eligible input then fails here with "no CUDA device"
(there is no CPU fallback) and runs on the GPU box (tests/test_dropin_gpu.py)."""
import os
import subprocess
import tempfile

import pytest

from tests import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU = os.path.join(ROOT, "nairn_mpm_fea_b200", "host", "_build", "NairnMPM_gpu")

LOAD_BC = ('<ParticleBCs><BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="3" zmax="20"><LoadBC dir="3" style="1" load="10"/>'
           '</BCBox></ParticleBCs>')

CASES = {
    "archiving custom task": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, custom_tasks='<CustomTasks><Schedule name="VTKArchive">'
                                             '<Parameter name="mass"/></Schedule></CustomTasks>'), "custom tasks other than PeriodicXPIC"),
    "feedback damping": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, extra_header="<FeedbackDamping>10</FeedbackDamping>"),
                         "time-dependent or feedback damping"),
    # (load BCs that are functions of time alone run on the device since round 2: tests/test_dropin_gpu.py)
    "silent particle loads": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace("</GridBCs>", "</GridBCs>" + LOAD_BC.replace('style="1" load="10"', 'style="5"')),
                              "silent particle load BCs"),
    "particle loads by function": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace("</GridBCs>", "</GridBCs>" + LOAD_BC.replace('style="1" load="10"', 'style="6" function="10*x*t"')),
                                   "particle load BCs set by a function"),
    "unsupported material": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, material='<Material Type="2" Name="Blk"><rho>1</rho><EA>1000</EA>'
                                            '<ET>500</ET><GA>300</GA><nuT>0.3</nuT><nuA>0.25</nuA><alphaA>0</alphaA><alphaT>0</alphaT></Material>'), "material type"),
    "other hardening law": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, material='<Material Type="9" Name="Blk"><rho>8.9</rho><E>100000</E><nu>0.33</nu>'
                                           '<alpha>20</alpha><Hardening>SL</Hardening><yield>120</yield><GPpG0>0.01</GPpG0><betahard>36</betahard><nhard>0.45</nhard>'
                                           '<yieldMax>640</yieldMax></Material>'), "hardening law other than"),
    "ideal rubber": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, material='<Material Type="8" Name="Blk"><rho>1</rho><G1>30</G1><G2>0</G2><K>100</K>'
                                    '<alpha>0</alpha><IdealRubber/></Material>'), "IdealRubber"),
    "unsupported shape functions": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, gimp="Finite"), "shape functions"),
    # failure handling of the replaced tasks that libmpmgpu does not do
    "time-step restarts": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, method=3, extra_header="<RestartScaling>0.5</RestartScaling>"),
                           "time-step restarts"),
    # (with fewer than 100 particles the reference's DEFAULT LeaveLimit, 1 % of the particles, rounds to 0 and lands in the delete
    # branch too -- NairnMPM.cpp:814-832 -- which is why every input of this file has at least 200 particles)
    "deleting leavers": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, extra_header="<LeaveLimit>-5</LeaveLimit>"), "deleting particles that leave the grid"),
    "deleting nan particles": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, extra_header="<DeleteLimit>5</DeleteLimit>"), "deleting nan particles"),
    # silent-divergence holes closed in round 2: the replaced tasks would add position/time dependent grid forces, an adiabatic
    # temperature rise, or a thermal strain from particles that start away from the stress-free temperature
    "grid body force function": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace(
        "</JANFEAInput>", "<Gravity><GridBodyXForce>10*x</GridBodyXForce></Gravity></JANFEAInput>"), "grid body force functions"),
    # (adiabatic coupling, <EnergyCoupling/>, runs on the device: tests/test_dropin_gpu.py)
    # (particles that start off the stress-free temperature run on the device since the laws carry thermal strains: tests/test_dropin_gpu.py)
    "thermal expansion with large rotation": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, extra_header="<StressFreeTemp>300</StressFreeTemp>")
                                              .replace("<alpha>0</alpha>", "<alpha>50</alpha><largeRotation>1</largeRotation>").replace('<Body ', '<Body temp="350" ', 1),
                                              None),
    # multimaterial mode runs on the device (tests/test_dropin_gpu.py) except for what mpmgpu_set_multimaterial does not cover
    "regression contact normals": (inputs.oblique_disks(inputs.disks2d(analysis=10, maxtime=0.5, extra_header="<MultiMaterialMode/>")),
                                   "contact normals by linear or logistic regression"),
    "own-gradient contact normals": (inputs.oblique_disks(inputs.disks2d(analysis=10, maxtime=0.5, extra_header=inputs.multimaterial(3, 0.3))),
                                     "each material's own normal"),
    "imperfect interface": (inputs.oblique_disks(inputs.disks2d(analysis=10, maxtime=0.5,
                                                 extra_header='<MultiMaterialMode Normals="2"><Friction Dn="1000" Dt="500">11</Friction></MultiMaterialMode>')),
                            "imperfect interfaces"),
    "multimaterial fmpm2": (inputs.oblique_disks(inputs.disks2d(analysis=10, maxtime=0.5, extra_header=inputs.multimaterial(2, 0.3)))
                            .replace("</JANFEAInput>", inputs.periodic_xpic(2, True, 1) + "</JANFEAInput>"), "order > 1 in multimaterial mode"),
    # conduction runs on the device (tests/test_dropin_gpu.py); its BCs, other transport tasks and thermal expansion do not
    # (nodal temperature BCs run on the device: tests/test_dropin_gpu.py)
    "heat flux BCs": (inputs.conduction(inputs.block3d(ncell=3, margin=2, maxtime=0.003), (350.0,), (2000.0,), (800.0,))
                      .replace("</GridBCs>", '</GridBCs><ParticleBCs><BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="4" zmax="20"><HeatFluxBC dir="2" face="1" style="6" function="10*(t-300)"/></BCBox></ParticleBCs>'),
                      "particle heat-flux BCs that are silent, coupled or set by a function of position"),
    "diffusion": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace("</MPMHeader>", '<Diffusion reference="0"/></MPMHeader>'),
                  "transport tasks other than conduction"),
    # global quantities the reference reads from its nodes / BC objects would be silently zero: the replaced tasks no longer fill them
    # (reaction forces are kept on the device: tests/test_dropin_gpu.py)
    "traction function": (inputs.particle_bcs(inputs.block3d(ncell=3, margin=2, maxtime=0.003), [
        ('<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="4" zmax="20">', '<TractionBC dir="11" face="6" style="6" function="-5*t*(1+x)"/>')]), "particle traction BCs set by a function of position"),
    "contact force quantity without multimaterial mode": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace(
        "</MPMHeader>", '<GlobalArchiveTime units="ms">0.001</GlobalArchiveTime><GlobalArchive type="contactz"/></MPMHeader>'), "contact-force global quantities outside multimaterial mode"),
    "grid kinetic energy quantity on several GPUs": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace(
        "</MPMHeader>", '<GlobalArchiveTime units="ms">0.001</GlobalArchiveTime><GlobalArchive type="Grid Kinetic Energy"/></MPMHeader>'), "grid kinetic energy in multimaterial mode or with -gpus N"),
    "reaction force quantity on several GPUs": (inputs.block3d(ncell=3, margin=2, maxtime=0.003).replace(
        "</MPMHeader>", '<GlobalArchiveTime units="ms">0.001</GlobalArchiveTime><GlobalArchive type="reactionz"/></MPMHeader>'), "reaction-force global quantities with -gpus N"),
    "more exponential terms": (inputs.block3d(ncell=3, margin=2, maxtime=0.003, material=inputs.neohookean_material(),
                                              extra_header="<DefGradTerms>3</DefGradTerms>"), "<DefGradTerms> other than the default"),
}


def run(xml, extra=()):
    d = tempfile.mkdtemp(prefix="dropin_cpu_")
    path = os.path.join(d, "in.fmcmd")
    open(path, "w").write(xml)
    return subprocess.run([GPU, *extra, path], cwd=d, capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("case", sorted(CASES))
def test_ineligible_inputs_are_refused_with_the_reason(case):
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    xml, reason = CASES[case]
    p = run(xml, ("-gpus", "2") if "several GPUs" in case else ())
    assert p.returncode == 2, (p.returncode, p.stderr[-500:], p.stdout[-300:])
    if reason is None:          # refused by the library after mpmgpu_create: needs the device (tests/test_dropin_gpu.py); here it stops at "no CUDA device"
        assert "cannot run this input on libmpmgpu" in p.stderr
        return
    assert "cannot run this input on libmpmgpu" in p.stderr and reason in p.stderr, p.stderr[-500:]


def test_eligible_input_reaches_the_device_and_has_no_cpu_fallback():
    import torch
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_dropin_gpu.py")
    p = run(inputs.block3d(ncell=3, margin=2, maxtime=0.003))
    assert p.returncode == 2 and "no CUDA device" in p.stderr, (p.returncode, p.stderr[-500:])


def test_options_that_do_not_matter_to_the_built_laws_are_accepted():
    """<DefGradTerms> only enters laws that exponentiate du (Neohookean, large rotation): a small-rotation IsotropicMat run stays
    eligible; so does <LeaveLimit> > 0 (push back, the default behaviour with another threshold)."""
    import torch
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_dropin_gpu.py")
    p = run(inputs.block3d(ncell=3, margin=2, maxtime=0.003, extra_header="<DefGradTerms>3</DefGradTerms><LeaveLimit>20</LeaveLimit>"))
    assert p.returncode == 2 and "no CUDA device" in p.stderr, (p.returncode, p.stderr[-500:])


def test_cpu_switch_runs_the_reference_tasks_unchanged():
    """`-cpu` leaves the reference's own tasks in place: the same binary is then the reference."""
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    p = run(inputs.block3d(ncell=3, margin=2, maxtime=0.003), ("-cpu",))
    assert p.returncode == 0 and "GPU TASKS" not in p.stdout and "Calculation Steps" in p.stdout
