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
    "archiving custom task": (inputs.block3d(ncell=2, margin=2, maxtime=0.003, custom_tasks='<CustomTasks><Schedule name="VTKArchive">'
                                             '<Parameter name="mass"/></Schedule></CustomTasks>'), "custom tasks other than PeriodicXPIC"),
    "feedback damping": (inputs.block3d(ncell=2, margin=2, maxtime=0.003, extra_header="<FeedbackDamping>10</FeedbackDamping>"),
                         "time-dependent or feedback damping"),
    "particle loads": (inputs.block3d(ncell=2, margin=2, maxtime=0.003).replace("</GridBCs>", "</GridBCs>" + LOAD_BC), "particle load BCs"),
    "unsupported material": (inputs.block3d(ncell=2, margin=2, maxtime=0.003, material='<Material Type="8" Name="Blk"><rho>1</rho><G1>30</G1>'
                                            '<G2>0</G2><K>100</K><alpha>0</alpha></Material>'), "material type"),
    "unsupported shape functions": (inputs.block3d(ncell=2, margin=2, maxtime=0.003, gimp="B2GIMP"), "shape functions"),
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
    p = run(xml)
    assert p.returncode == 2, (p.returncode, p.stderr[-500:], p.stdout[-300:])
    assert "cannot run this input on libmpmgpu" in p.stderr and reason in p.stderr, p.stderr[-500:]


def test_eligible_input_reaches_the_device_and_has_no_cpu_fallback():
    import torch
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_dropin_gpu.py")
    p = run(inputs.block3d(ncell=2, margin=2, maxtime=0.003))
    assert p.returncode == 2 and "no CUDA device" in p.stderr, (p.returncode, p.stderr[-500:])


def test_cpu_switch_runs_the_reference_tasks_unchanged():
    """`-cpu` leaves the reference's own tasks in place: the same binary is then the reference."""
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    p = run(inputs.block3d(ncell=2, margin=2, maxtime=0.003), ("-cpu",))
    assert p.returncode == 0 and "GPU TASKS" not in p.stdout and "Calculation Steps" in p.stdout
