"""The DEVICE constitutive laws (nairn_mpm_fea_b200/csrc/materials.cuh) compiled for the host (tests/devlaws) against the C
restatement of the reference's laws (oracle/mpm_oracle.c, itself pinned to the reference by tests/test_oracle_cpu.py), on
random deformed states: checks the CUDA source of every law and analysis type without a GPU.  The compiled kernels are
checked by the GPU parity tests."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from nairn_mpm_fea_b200 import materials as M
from tests.parity import TOL_LR3D

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEV = os.path.join(HERE, "devlaws")
LIBDEV = os.path.join(DEV, "_build", "libdevlaws.so")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def libs():
    src = os.path.join(DEV, "host_laws.cpp")
    csrc = os.path.join(ROOT, "nairn_mpm_fea_b200", "csrc")
    deps = [src, os.path.join(csrc, "materials.cuh"), os.path.join(csrc, "mpm_types.cuh"), os.path.join(csrc, "archive.cuh"), os.path.join(csrc, "shape.cuh"), os.path.join(DEV, "stub", "cuda_runtime.h")]
    if not os.path.exists(LIBDEV) or any(os.path.getmtime(d) > os.path.getmtime(LIBDEV) for d in deps):
        os.makedirs(os.path.dirname(LIBDEV), exist_ok=True)
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-std=c++17", "-I" + os.path.join(DEV, "stub"),
                        "-I" + csrc, src, "-o", LIBDEV], check=True)
    from oracle import port
    port.build()
    return C.CDLL(LIBDEV), C.CDLL(port.LIB)


def _material_struct(m):
    from nairn_mpm_fea_b200.capi import Material
    s = Material()
    s.kind, s.n_history = m["kind"], m.get("n_history", 0)
    for j, v in enumerate(m["p"]):
        s.p[j] = v
    return s


def _random_state(kind, dim, n, rng, modulus):
    F = np.zeros((9, n))
    g = 0.05 * rng.standard_normal((9, n))
    if dim == 2:
        g[[2, 5, 6, 7]] = 0.0
    F[:] = g
    F[[0, 4, 8]] += 1.0
    sp = 1.0e-3 * modulus * rng.standard_normal((6, n))
    eplast = 5.0e-3 * rng.standard_normal((6, n))
    if dim == 2:
        sp[[3, 4]] = 0.0
        eplast[[3, 4]] = 0.0
    pressure = 1.0e-3 * modulus * rng.standard_normal(n)
    hist = np.zeros((M.MAX_HISTORY, n))
    if kind in (M.NEOHOOKEAN, M.MOONEY):
        # elastic left Cauchy-Green tensor B = F F^T (xx,yy,zz,yz,xz,xy), J = det F, Jres = 1
        Fm = F.T.reshape(n, 3, 3)
        B = Fm @ np.transpose(Fm, (0, 2, 1))
        eplast = np.stack([B[:, 0, 0], B[:, 1, 1], B[:, 2, 2], B[:, 1, 2], B[:, 0, 2], B[:, 0, 1]])
        hist[0] = np.linalg.det(Fm)
        hist[1] = 1.0
    elif kind == M.ISOPLASTICITY:
        hist[0] = 0.02 * rng.random(n)
    energies = np.zeros((6, n))
    energies[:5] = 1.0e-6 * modulus * rng.standard_normal((5, n))
    energies[5] = 300.0                      # previous temperature: entropy stays finite
    du = 2.0e-3 * rng.standard_normal((n, 9))
    if dim == 2:
        du[:, [2, 5, 6, 7]] = 0.0
        du[:, 8] = 0.0                       # plane strain / plane stress: no out-of-plane velocity gradient
    return dict(F=F, sp=sp, eplast=eplast, pressure=pressure, hist=hist, energies=energies, du=du)


def _F_to_ep_wrot(F, dim):
    ep = np.zeros((6, F.shape[1]))
    wrot = np.zeros((3, F.shape[1]))
    ep[0], ep[1], ep[2] = F[0] - 1.0, F[4] - 1.0, F[8] - 1.0
    ep[5], wrot[0] = F[3] + F[1], F[3] - F[1]
    if dim == 3:
        ep[4], wrot[1] = F[6] + F[2], F[6] - F[2]
        ep[3], wrot[2] = F[7] + F[5], F[7] - F[5]
    return ep, wrot


U = M.xml_units(E=1000.0, G=40.0, K=200.0, rho=1.0, yld=2.0, Ep=100.0)
NPS = {"3d": (3, M.THREED_MPM), "planestrain": (2, M.PLANE_STRAIN_MPM), "planestress": (2, M.PLANE_STRESS_MPM)}


def _mat(name, np_):
    if name == "isotropic":
        return M.isotropic(U["E"], 0.3, U["rho"], aI=40.0, np_=np_)
    if name == "isotropic_lr":
        return M.isotropic(U["E"], 0.3, U["rho"], aI=40.0, np_=np_, large_rotation=True)
    if name.startswith("neohookean"):
        return M.neohookean(U["G"], U["K"], U["rho"], aI=40.0, UofJOption=int(name[-1]), av=(0.3, 1.5) if name[-1] == "0" else None)
    if name.startswith("mooney"):
        return M.mooney(0.75 * U["G"], 0.25 * U["G"], U["K"], U["rho"], aI=40.0, UofJOption=int(name[-1]), av=(0.3, 1.5) if name[-1] == "1" else None)
    if name in ("nonlinear", "nonlinear2"):
        return M.isoplasticity(U["E"], 0.3, U["rho"], U["yld"], None, 0.0, aI=20.0, np_=np_, hardening=(name, 8.0, 0.4))
    if name == "johnsoncook":
        return M.isoplasticity(U["E"], 0.3, U["rho"], U["yld"], None, 0.0, aI=20.0, np_=np_,
                               hardening=("johnsoncook", dict(B=1.5 * U["yld"], n=0.5, C=0.02, ep0=1.0, D=0.01, n2=2.0, Tm=1600.0, m=1.1, Tref=250.0)))
    if name == "isoplasticity":
        return M.isoplasticity(U["E"], 0.3, U["rho"], U["yld"], U["Ep"], aI=20.0, np_=np_, av=(0.2, 2.0))
    if name == "isoplasticity_lr":
        return M.isoplasticity(U["E"], 0.3, U["rho"], U["yld"], U["Ep"], aI=20.0, np_=np_, large_rotation=True)
    raise KeyError(name)


LAWS = ["isotropic", "isotropic_lr", "neohookean0", "neohookean1", "neohookean2", "mooney0", "mooney1", "mooney2", "isoplasticity", "isoplasticity_lr",
        "nonlinear", "nonlinear2", "johnsoncook"]


@pytest.mark.parametrize("analysis", list(NPS))
@pytest.mark.parametrize("law", LAWS)
def test_device_law_source_matches_oracle(libs, law, analysis):
    dev, orc = libs
    dim, np_ = NPS[analysis]
    m = _mat(law, np_)
    m["p"][6] = 1.0                          # average cell size: the library fills this slot (mpmgpu_set_materials)
    n = 400
    rng = np.random.default_rng(zlib.crc32((law + analysis).encode()))
    modulus = U["E"] / U["rho"]
    st = _random_state(m["kind"], dim, n, rng, modulus)
    delTime = 1.0e-7
    # device source
    d = {k: v.copy() for k, v in st.items()}
    p = np.ascontiguousarray(m["p"], dtype=np.float64)
    rc = dev.devlaws_batch(dim, np_, m["kind"], m.get("n_history", 0), _dp(p), n, _dp(d["F"]), _dp(d["sp"]), _dp(d["pressure"]),
                           _dp(d["eplast"]), _dp(d["energies"]), _dp(d["hist"]), _dp(d["du"]), C.c_double(delTime))
    assert rc == 0
    # oracle
    o = {k: v.copy() for k, v in st.items()}
    ep, wrot = _F_to_ep_wrot(o["F"], dim)
    ms = _material_struct(m)
    rc = orc.oracle_law_batch(np_, C.c_double(1.0), C.c_double(1.0), C.c_double(1.0 if dim == 3 else 0.0), C.byref(ms), n,
                              _dp(o["sp"]), _dp(o["pressure"]), _dp(ep), _dp(wrot), _dp(o["eplast"]), _dp(o["energies"]), _dp(o["hist"]),
                              _dp(o["du"]), C.c_double(delTime))
    assert rc == 0
    dep, dwrot = _F_to_ep_wrot(d["F"], dim)
    # 3D large rotation: the reference's polar decomposition is reproducible to ~1e-6 of the increment only (tests/parity.py)
    tol = TOL_LR3D if (dim == 3 and law.endswith("_lr")) else 1.0e-12
    comp = [0, 1, 2, 5] if dim == 2 else slice(None)

    def check(name, a, b, scale=None):
        scale = max(float(np.max(np.abs(b))), 1e-300) if scale is None else scale
        err = float(np.max(np.abs(a - b))) / scale
        assert err <= tol, "%s %s: %s differs by %.3e" % (law, analysis, name, err)

    Fscale = float(np.max(np.abs(ep)))
    check("ep", dep, ep, Fscale)
    check("wrot", dwrot, wrot, Fscale)
    check("sp", d["sp"][comp], o["sp"][comp])
    check("pressure", d["pressure"], o["pressure"])
    check("eplast", d["eplast"][comp], o["eplast"][comp])
    for i, nm in enumerate(["work", "res", "heat", "entropy", "plast"]):
        check(nm, d["energies"][i], o["energies"][i])
    check("history", d["hist"], o["hist"])
    if law.startswith("isoplasticity") or law in ("nonlinear", "nonlinear2", "johnsoncook"):
        assert np.count_nonzero(d["hist"][0] != st["hist"][0]) > n // 10, "the sample should yield on a good part of the particles"
        if law.startswith("isoplasticity"):
            assert np.count_nonzero(d["hist"][0] == st["hist"][0]) > 0, "and stay elastic on some"


@pytest.mark.parametrize("law", ["isotropic", "neohookean0", "isoplasticity"])
def test_plain_dispatch_equals_lr_dispatch_without_large_rotation(libs, law):
    """k_update_strains_lr goes through constitutive_law_lr, every other kernel through constitutive_law: the two must be
    the same function when no material asks for large rotation."""
    dev, _ = libs
    dim, np_ = NPS["3d"]
    m = _mat(law, np_)
    m["p"][6] = 1.0
    rng = np.random.default_rng(7)
    st = _random_state(m["kind"], dim, 16, rng, U["E"] / U["rho"])
    p = np.ascontiguousarray(m["p"], dtype=np.float64)
    a = {k: v.copy() for k, v in st.items()}
    dev.devlaws_batch(dim, np_, m["kind"], 0, _dp(p), 16, _dp(a["F"]), _dp(a["sp"]), _dp(a["pressure"]), _dp(a["eplast"]),
                      _dp(a["energies"]), _dp(a["hist"]), _dp(a["du"]), C.c_double(1.0e-7))
    for q in range(16):
        F, sp, el = st["F"][:, q].copy(), st["sp"][:, q].copy(), st["eplast"][:, q].copy()
        pr = np.array([st["pressure"][q]])
        en, hi, du = st["energies"][:, q].copy(), st["hist"][:, q].copy(), st["du"][q].copy()
        dev.devlaws_plain_one(dim, np_, m["kind"], _dp(p), _dp(F), _dp(sp), _dp(pr), _dp(el), _dp(en), _dp(hi), _dp(du), C.c_double(1.0e-7))
        assert np.array_equal(F, a["F"][:, q]) and np.array_equal(sp, a["sp"][:, q]) and np.array_equal(el, a["eplast"][:, q])
        assert pr[0] == a["pressure"][q] and np.array_equal(en[:5], a["energies"][:5, q]) and np.array_equal(hi, a["hist"][:, q])
