"""CPU-side checks (no GPU): the host set-up reproduces the reference's set-up arithmetic bit for bit
(against the golden dumps), and the C-ABI library is built and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.parity import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nairn_mpm_fea_b200 import build, capi
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "mpmgpu.h")).read()
    declared = set(re.findall(r"\b(mpmgpu_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for sym in sorted(declared):
        assert hasattr(lib, sym), "libmpmgpu.so does not export %s" % sym
    assert set(capi.EXPORTS) == declared, (set(capi.EXPORTS) ^ declared)
    assert lib.mpmgpu_abi_version() == capi.ABI_VERSION


def test_no_device_fails_loudly():
    """On a box without CUDA the product must refuse, not fall back (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nairn_mpm_fea_b200 import MpmGpu, MpmGpuError, problem
    prob = problem.block3d(ncell=2, margin=2)
    with pytest.raises(MpmGpuError) as ei:
        MpmGpu(prob)
    assert ei.value.code == -2


def test_block3d_generator_matches_reference_setup():
    from nairn_mpm_fea_b200 import problem
    z = load_golden("block3d_ugimp_usavg")
    a = problem.from_reference_dump(z)
    b = problem.block3d(ncell=4, margin=2)
    for k in ("np", "horiz", "vert", "depth", "grid", "shape", "method", "dt", "dt_strain_first", "dt_strain_last"):
        assert getattr(a, k) == getattr(b, k), k
    for k in ("xpts", "ypts", "zpts"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for k in ("pos", "vel", "mp", "lp", "in_elem", "matnum"):
        assert np.array_equal(np.asarray(a.particles[k]), np.asarray(b.particles[k])), k
    assert np.array_equal(a.bc_node, b.bc_node) and np.array_equal(a.bc_norm, b.bc_norm)
    assert np.array_equal(a.materials[0]["p"], b.materials[0]["p"])


def test_taylor_bar_generator_matches_reference_setup():
    """IsoPlasticity block on a plate of rigid-BC particles (BASELINE config 4 family): the host generator gives
    the particles, their order (rigid ones last) and both material blocks exactly as the reference sets them up."""
    from nairn_mpm_fea_b200 import materials as M, problem
    z = load_golden("block3d_rigid_wall_lattice")
    a = problem.from_reference_dump(z)
    u = M.xml_units(E=2000.0, rho=2.0, yld=20.0, Ep=100.0)
    mat = M.isoplasticity(u["E"], 0.33, u["rho"], u["yld"], u["Ep"], aI=20.0)
    b = problem.block3d(ncell=2, margin=3, velocity=(0.0, 0.0, -4.0e4), bottom_bc=False, material=mat,
                        rigid_wall=dict(set_direction=4))
    for k in ("np", "horiz", "vert", "depth", "grid", "shape", "method", "dt", "dt_strain_first", "dt_strain_last"):
        assert getattr(a, k) == getattr(b, k), k
    for k in ("pos", "vel", "mp", "lp", "in_elem", "matnum"):
        assert np.array_equal(np.asarray(a.particles[k]), np.asarray(b.particles[k])), k
    assert a.particles["n_nonrigid"] == b.particles["n_nonrigid"] == 64 and b.nparticles == 192
    for i in range(2):
        assert a.materials[i]["kind"] == b.materials[i]["kind"] and np.array_equal(a.materials[i]["p"], b.materials[i]["p"])
    assert len(b.bc_node) == 0


def test_material_block_matches_reference_properties():
    from nairn_mpm_fea_b200 import problem
    for case in ("block3d_ugimp_usavg", "block3d_fast_crossings"):
        z = load_golden(case)
        a = problem.from_reference_dump(z)
        q = z["mat_params"][0]
        p = a.materials[0]["p"]
        assert p[8] == q[14] and p[9] == q[15] and p[16] == q[16], (p[8:17], q[14:17])
        assert p[20] == q[12]


def test_neohookean_and_isoplasticity_blocks_match_reference_properties():
    """materials.py reproduces what the reference's VerifyAndLoadProperties computed (values read from the live
    reference objects by oracle/ref_harness.cpp): Neohookean Gsp, Ksp, Lamesp, gamma0; IsoPlasticity Gred, Kred,
    yldred, Epred, alphaMax, yldredMin, gamma0; rigid direction bits; artificial viscosity coefficients."""
    from nairn_mpm_fea_b200 import materials as M, problem
    seen = set()
    for case in ("block3d_neohookean", "block3d_neohookean_uj1", "disks2d_neohookean", "block3d_isoplastic", "disks2d_isoplastic",
                 "disks2d_isoplastic_planestress", "block3d_rigid_piston", "block3d_neohookean_av", "block3d_isoplastic_av"):
        z = load_golden(case)
        a = problem.from_reference_dump(z)
        for mid, q, m in zip(z["mat_ids"], z["mat_params"], a.materials):
            p = m["p"]
            assert p[0] == q[0] and p[1] == q[1]
            assert (p[3] != 0.0) == bool(q[5])
            if q[5]:
                assert p[4] == q[6] and p[5] == q[7]
            if mid == M.NEOHOOKEAN:
                assert p[8] == q[11] and p[9] == q[12] and p[10] == q[13] and p[11] == q[14], (case, p[8:14], q[8:17])
                assert abs(p[13] - q[16]) <= 1e-15 * max(abs(q[16]), 1e-300)
                seen.add("neo")
            elif mid == M.ISOPLASTICITY:
                assert p[8] == q[13] and p[9] == q[14] and p[10] == q[17] and p[11] == q[18], (case, p[8:16], q[8:22])
                assert p[14] == q[19] and p[15] == q[20]
                assert abs(p[13] - q[12]) <= 1e-15 * max(abs(q[12]), 1e-300)
                seen.add("isoplastic")
            elif mid == M.RIGIDBC:
                assert p[8] == q[8]
                seen.add("rigid")
    assert seen == {"neo", "isoplastic", "rigid"}
