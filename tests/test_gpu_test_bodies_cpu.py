"""Dry run of GPU test BODIES on the CPU: `MpmGpu` is replaced by the host-run device source (tests/test_device_step_cpu.py::EmuSim,
same methods), so the code of the late GPU tests -- written when no GPU minutes were left -- is itself exercised before it meets a
GPU.  It checks the tests, not the kernels: a sample of cases is enough."""
import ctypes as C
import os

import pytest

import nairn_mpm_fea_b200
from tests.test_device_step_cpu import EmuSim, lib  # noqa: F401


def _fake_class(emulib):
    class FakeGpu(EmuSim):
        def __init__(self, prob, device=0, kernel_path=0, max_particles=0, sort_interval=0, upload=True):
            from nairn_mpm_fea_b200.capi import MpmGpuError
            extended = any(m["p"][7] != 0.0 or m["kind"] == 8 or (m["kind"] == 9 and m["p"][16] > 1.0) for m in prob.materials)
            if kernel_path == 2 and (extended or prob.shape != 1 or not prob.is3d):
                raise MpmGpuError(-1, "kernel_path=2 (fused) is not eligible")
            EmuSim.__init__(self, emulib, prob, merged_cpdi=os.environ.get("MPMGPU_CPDI_MERGE") == "1")

        def status(self):
            return dict(mstep=0)
    return FakeGpu


@pytest.fixture
def fake_gpu(lib, monkeypatch):  # noqa: F811
    monkeypatch.setattr(nairn_mpm_fea_b200, "MpmGpu", _fake_class(lib))


@pytest.mark.parametrize("case", ["block3d_isotropic_lr", "disks2d_neo_planestress", "block3d_free_lcpdi_xpic2", "block3d_mooney", "block3d_johnsoncook",
                                  "block3d_b2gimp", "disks2d_b2cpdi"])
def test_late_golden_bodies(fake_gpu, case):
    import tests.test_zz_late_goldens_gpu as T
    T.test_each_task_of_step_one(case)
    T.test_whole_steps(case, 0 if case in T.LR_CASES else 1)


def test_fused_refusal_and_merged_cpdi_bodies(fake_gpu, monkeypatch):
    import tests.test_zz_late_goldens_gpu as T
    T.test_fused_path_refuses_large_rotation()
    T.test_whole_steps("block3d_material_pdamping", 2)
    T.test_cpdi_merged_whole_steps("disks2d_qcpdi", monkeypatch)


@pytest.mark.parametrize("seed", [0, 7, 19, 33])
def test_gpu_sweep_bodies(fake_gpu, seed):
    from oracle import refharness
    if not refharness.available():
        pytest.skip("oracle/_ref not built")
    import tests.test_zzzz_sweep_gpu as S
    S.test_random_combination_on_the_gpu_matches_the_live_reference(seed)
