"""Dry run of GPU test BODIES on the CPU: `MpmGpu` is replaced by the host-run device source (tests/test_device_step_cpu.py::EmuSim,
same methods), so the code of the late GPU tests -- written when no GPU minutes were left -- is itself exercised before it meets a
GPU.  It checks the tests, not the kernels: a sample of cases is enough."""
import ctypes as C
import os

import pytest

import nairn_mpm_fea_b200
from tests.test_device_laws_cpu import libs  # noqa: F401
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

        # ---- output side: the device functions of csrc/archive.cuh on the emulation's state (tests/devlaws/host_laws.cpp) ----
        def _device_state(self):
            import numpy as np
            from nairn_mpm_fea_b200 import materials as M
            n = self.n
            st = dict(pos=np.zeros((3, n)), vel=np.zeros((3, n)), mp=np.zeros(n), F=np.zeros((9, n)), sp=np.zeros((6, n)), pressure=np.zeros(n),
                      eplast=np.zeros((6, n)), energies=np.zeros((6, n)), hist=np.zeros((M.MAX_HISTORY, n)), elem=np.zeros(n, np.int32),
                      mat0=np.zeros(n, np.int32), cross=np.zeros(n, np.int32))
            dp, ip = (lambda a: a.ctypes.data_as(C.POINTER(C.c_double))), (lambda a: a.ctypes.data_as(C.POINTER(C.c_int)))
            self.lib.emu_get_device_state(self.h, dp(st["pos"]), dp(st["vel"]), dp(st["mp"]), dp(st["F"]), dp(st["sp"]), dp(st["pressure"]), dp(st["eplast"]),
                                          dp(st["energies"]), dp(st["hist"]), ip(st["elem"]), ip(st["mat0"]), ip(st["cross"]))
            kinds = np.ascontiguousarray([m["kind"] for m in self.prob.materials], dtype=np.int32)
            params = np.ascontiguousarray(np.stack([m["p"] for m in self.prob.materials]), dtype=np.float64)
            args = [dp(st["pos"]), dp(st["vel"]), dp(st["mp"]), dp(st["F"]), dp(st["sp"]), dp(st["pressure"]), dp(st["eplast"]), dp(st["energies"]), dp(st["hist"]),
                    ip(st["elem"]), ip(st["mat0"]), ip(st["cross"]), len(kinds), ip(kinds), dp(params)]
            return st, kinds, params, args

        def set_archive_origin(self, origpos=None, angles0=None, thickness=1.0):
            import numpy as np
            c = np.ascontiguousarray
            self._arch = (c(self.prob.particles["pos"] if origpos is None else origpos, dtype=np.float64),
                          None if angles0 is None else c(angles0, dtype=np.float64), float(thickness))

        def _dev(self):
            from tests.test_device_laws_cpu import LIBDEV
            return C.CDLL(LIBDEV)

        def archive_record_size(self, order):
            st, kinds, params, args = self._device_state()
            return self._dev().devarch_records(3 if self.prob.is3d else 2, self.n, order.encode(), *args, None, None, C.c_double(1.0), None)

        def pack_archive(self, order, out=None):
            import numpy as np
            from nairn_mpm_fea_b200.capi import MpmGpuError
            rec = self.archive_record_size(order)
            if rec < 0:
                raise MpmGpuError(-1, "unsupported archive item")
            buf = np.zeros(rec * self.n, np.uint8) if out is None else out
            if buf.nbytes < rec * self.n:
                raise MpmGpuError(-1, "buffer too small")
            st, kinds, params, args = self._device_state()
            orig, ang, thick = self._arch
            dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))          # noqa: E731
            self._dev().devarch_records(3 if self.prob.is3d else 2, self.n, order.encode(), *args, dp(orig), dp(ang), C.c_double(thick),
                                        buf.ctypes.data_as(C.POINTER(C.c_ubyte)))
            return buf.tobytes() if out is None else buf

        def global_sums(self):
            import numpy as np
            from nairn_mpm_fea_b200.capi import GS_NSUMS
            nn = int(self.prob.particles.get("n_nonrigid", self.n))
            st, kinds, params, _ = self._device_state()
            sub = {k: np.ascontiguousarray(v[..., :nn]) for k, v in st.items()}
            dp, ip = (lambda a: a.ctypes.data_as(C.POINTER(C.c_double))), (lambda a: a.ctypes.data_as(C.POINTER(C.c_int)))
            out = np.zeros((len(kinds), GS_NSUMS))
            self._dev().devarch_global_sums(3 if self.prob.is3d else 2, nn, dp(sub["pos"]), dp(sub["vel"]), dp(sub["mp"]), dp(sub["F"]), dp(sub["sp"]),
                                            dp(sub["pressure"]), dp(sub["eplast"]), dp(sub["energies"]), dp(sub["hist"]), ip(sub["elem"]), ip(sub["mat0"]),
                                            ip(sub["cross"]), len(kinds), ip(kinds), dp(params), dp(out))
            return out
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


@pytest.mark.parametrize("case,kernel_path,order", [("block3d_jitter", 2, "iYYYYYNYYYNNYYYYYY"), ("disks2d_isoplastic", 1, "iYYYYYNYYYNNYYYYYY"),
                                                    ("block3d_rigid_wall", 2, "iYYYYNNYNNNNNYNNYN")])
def test_archive_gpu_bodies(fake_gpu, libs, case, kernel_path, order):  # noqa: F811
    import tests.test_zzz_archive_gpu as A
    A.test_device_records_equal_host_writer_on_download(case, kernel_path, order)
    A.test_global_sums_equal_numpy_over_download(case, kernel_path)


def test_archive_refusal_body(fake_gpu, libs):  # noqa: F811
    import tests.test_zzz_archive_gpu as A
    A.test_unsupported_archive_items_are_refused()
