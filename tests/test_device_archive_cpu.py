"""The DEVICE source of the output side (nairn_mpm_fea_b200/csrc/archive.cuh: archive records, global-quantity summands),
compiled for the host (tests/devlaws), against nairn_mpm_fea_b200/archive.py -- itself byte-identical to the files of the
reference's CLI (tests/test_archive_cpu.py) -- and against plain numpy sums, on golden states.  No GPU needed; the compiled
kernels are checked by tests/test_zzz_archive_gpu.py."""
import ctypes as C

import numpy as np
import pytest

from nairn_mpm_fea_b200 import archive, problem
from nairn_mpm_fea_b200 import materials as M
from nairn_mpm_fea_b200.capi import GS_NSUMS
from tests.parity import load_golden
from tests.test_device_laws_cpu import libs  # noqa: F401  (fixture: builds tests/devlaws)

CASES = [("block3d_ugimp_usavg", "iYYYYNNNNNNNNNNNNN"), ("block3d_neohookean_uj1", "iYYYYYNYYYNNYYYYYY"), ("disks2d_isoplastic", "iYYYYYNYYYNNYYYYYY"),
         ("disks2d_neohookean", "iYYYYYNYYYNNYCYYYY"), ("block3d_rigid_wall", "iYYYYNNYNNNNNYNNYN"), ("block3d_isoplastic", "iNYNYNNYNYNNYANNNY"),
         ("block3d_mooney", "iYYYYYNYYYNNYCYYYY")]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def device_view(z, prob, step):
    """The device's SoA view of the golden state after `step` steps, and the dict MpmGpu.download() would return for it."""
    pre = "p%d/" % step
    n = prob.nparticles
    dim = 3 if prob.is3d else 2
    ep, wrot = z[pre + "ep"], z[pre + "wrot"]
    F = np.zeros((9, n))
    F[0], F[4], F[8] = 1.0 + ep[0], 1.0 + ep[1], 1.0 + ep[2]
    F[1], F[3] = 0.5 * (ep[5] - wrot[0]), 0.5 * (ep[5] + wrot[0])
    if dim == 3:
        F[2], F[6] = 0.5 * (ep[4] - wrot[1]), 0.5 * (ep[4] + wrot[1])
        F[5], F[7] = 0.5 * (ep[3] - wrot[2]), 0.5 * (ep[3] + wrot[2])
    # what the download kernel makes of F (k_F_to_epwrot)
    ep2, wrot2 = np.zeros((6, n)), np.zeros((3, n))
    ep2[0], ep2[1], ep2[2], ep2[5], wrot2[0] = F[0] - 1.0, F[4] - 1.0, F[8] - 1.0, F[3] + F[1], F[3] - F[1]
    if dim == 3:
        ep2[4], ep2[3], wrot2[1], wrot2[2] = F[6] + F[2], F[7] + F[5], F[6] - F[2], F[7] - F[5]
    c = np.ascontiguousarray
    dev = dict(pos=c(z[pre + "pos"], dtype=np.float64), vel=c(z[pre + "vel"], dtype=np.float64), mp=c(prob.particles["mp"], dtype=np.float64), F=F,
               sp=c(z[pre + "sp"], dtype=np.float64), pressure=c(z[pre + "pressure"], dtype=np.float64), eplast=c(z[pre + "eplast"], dtype=np.float64),
               energies=c(z[pre + "energies"], dtype=np.float64), hist=np.zeros((M.MAX_HISTORY, n)),
               elem=c(z[pre + "inElem"], dtype=np.int32), mat0=c(np.asarray(prob.particles["matnum"]) - 1, dtype=np.int32),
               cross=c(z[pre + "crossings"], dtype=np.int32))
    h = z[pre + "hist"]
    dev["hist"][:min(h.shape[0], M.MAX_HISTORY)] = h[:M.MAX_HISTORY]
    state = dict(pos=dev["pos"], vel=dev["vel"], sp=dev["sp"], pressure=dev["pressure"], ep=ep2, wrot=wrot2, eplast=dev["eplast"],
                 energies=dev["energies"], history=dev["hist"], in_elem=dev["elem"], crossings=dev["cross"])
    kinds = np.array([m["kind"] for m in prob.materials], np.int32)
    params = np.ascontiguousarray(np.stack([m["p"] for m in prob.materials]), dtype=np.float64)
    return dim, n, dev, state, kinds, params


def _args(dim, n, dev, kinds, params):
    return [dim, n], [_dp(dev["pos"]), _dp(dev["vel"]), _dp(dev["mp"]), _dp(dev["F"]), _dp(dev["sp"]), _dp(dev["pressure"]), _dp(dev["eplast"]),
                      _dp(dev["energies"]), _dp(dev["hist"]), _ip(dev["elem"]), _ip(dev["mat0"]), _ip(dev["cross"]), len(kinds), _ip(kinds), _dp(params)]


@pytest.mark.parametrize("case,order", CASES)
def test_device_record_source_is_byte_identical_to_the_host_writer(libs, case, order):  # noqa: F811
    dev_lib, _ = libs
    z = load_golden(case)
    prob = problem.from_reference_dump(z)
    step = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")[-1]
    dim, n, dev, state, kinds, params = device_view(z, prob, step)
    rng = np.random.default_rng(3)
    angles0 = 0.3 * rng.standard_normal((3, n))
    origpos = np.ascontiguousarray(prob.particles["pos"], dtype=np.float64)
    want, recsize = archive.records(prob, state, order, origpos=origpos, thickness=np.full(n, 1.25), angles0=angles0)
    head, tail = _args(dim, n, dev, kinds, params)
    out = np.zeros(n * recsize, np.uint8)
    rec = dev_lib.devarch_records(*head, order.encode(), *tail, _dp(origpos), _dp(angles0), C.c_double(1.25), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    assert rec == recsize
    if out.tobytes() != want:
        a, b = out.reshape(n, recsize), np.frombuffer(want, np.uint8).reshape(n, recsize)
        cols = np.nonzero(np.any(a != b, axis=0))[0]
        raise AssertionError("%s: record bytes differ at offsets %s (record size %d)" % (case, cols[:24].tolist(), recsize))


def test_record_size_follows_the_order_string(libs):  # noqa: F811
    dev_lib, _ = libs
    z = load_golden("block3d_ugimp_usavg")
    prob = problem.from_reference_dump(z)
    dim, n, dev, state, kinds, params = device_view(z, prob, 1)
    head, tail = _args(dim, n, dev, kinds, params)
    for order in ("i", "iY", "iYYYY", "iYYYYYNYYYNNYYYYYY", "iNNNNNNNNNNNNONNNN"):
        _, recsize = archive.records(prob, state, order)
        assert dev_lib.devarch_records(*head, order.encode(), *tail, None, None, C.c_double(1.0), None) == recsize, order
    # items this path does not produce are refused, as by archive.py
    assert dev_lib.devarch_records(*head, b"iYYYYNNNNNNYNNNNNN", *tail, None, None, C.c_double(1.0), None) == -1


def reference_sums(prob, state, dim):
    """The GlobalQuantity sums written out with numpy (GlobalQuantity.cpp:394-1075), per material."""
    n = prob.nparticles
    nn = int(prob.particles.get("n_nonrigid", n))
    mp = np.asarray(prob.particles["mp"], np.float64)
    mat0 = np.asarray(prob.particles["matnum"]) - 1
    out = np.zeros((len(prob.materials), GS_NSUMS))
    ep, wrot = state["ep"], state["wrot"]
    F = np.zeros((9, n))
    F[0], F[4], F[8] = 1.0 + ep[0], 1.0 + ep[1], 1.0 + ep[2]
    F[1], F[3] = 0.5 * (ep[5] - wrot[0]), 0.5 * (ep[5] + wrot[0])
    F[2], F[6] = 0.5 * (ep[4] - wrot[1]), 0.5 * (ep[4] + wrot[1])
    F[5], F[7] = 0.5 * (ep[3] - wrot[2]), 0.5 * (ep[3] + wrot[2])
    mag = np.zeros_like(out)            # sum of |terms|: the scale round-off is measured against

    def put(m, k, terms):
        out[m, k] = terms.sum()
        mag[m, k] = np.abs(terms).sum()

    for m, mat in enumerate(prob.materials):
        sel = np.nonzero(mat0[:nn] == m)[0]
        if mat["kind"] == M.RIGIDBC or sel.size == 0:
            continue
        J = state["history"][0][sel] if mat["kind"] in (M.NEOHOOKEAN, M.MOONEY) else 1.0
        Vp = J * mp[sel] / mat["rho"]
        v = state["vel"][:, sel]
        en = state["energies"][:, sel]
        pr = state["pressure"][sel] if mat["kind"] in (M.NEOHOOKEAN, M.ISOPLASTICITY, M.MOONEY) else 0.0
        put(m, 0, mp[sel])
        put(m, 1, Vp * np.ones(sel.size))
        for c in range(dim):
            put(m, 2 + c, mp[sel] * v[c])
            put(m, 17 + c, Vp * v[c])
        put(m, 5, 0.5 * mp[sel] * (v[:dim] ** 2).sum(axis=0))
        for k, terms in ((6, en[0]), (7, en[0] - en[1]), (8, en[2]), (9, en[3]), (10, en[4])):
            put(m, k, mp[sel] * terms)
        for c in range(6):
            put(m, 11 + c, mp[sel] * (state["sp"][c][sel] - (pr if c < 3 else 0.0)))
        for i in range(9):
            put(m, 20 + i, Vp * F[i][sel])
    return out, mag


def sums_close(got, want_and_mag, tol=1e-12):
    want, mag = want_and_mag
    both_nan = np.isnan(got) & np.isnan(want)          # entropy is 0/0 in the reference when the temperature is 0
    return bool(np.all(both_nan | (np.abs(got - want) <= tol * mag)))


@pytest.mark.parametrize("case", ["block3d_neohookean", "disks2d_isoplastic", "block3d_rigid_wall"])
def test_device_global_summands_match_numpy(libs, case):  # noqa: F811
    dev_lib, _ = libs
    z = load_golden(case)
    prob = problem.from_reference_dump(z)
    step = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")[-1]
    dim, n, dev, state, kinds, params = device_view(z, prob, step)
    nn = int(prob.particles.get("n_nonrigid", n))
    head, tail = _args(dim, nn, dev, kinds, params)          # the library sums the non-rigid particles
    # component-major arrays: a view of the first nn particles needs its own copy
    dev_nr = {k: np.ascontiguousarray(v[..., :nn]) for k, v in dev.items()}
    head, tail = _args(dim, nn, dev_nr, kinds, params)
    sums = np.zeros((len(kinds), GS_NSUMS))
    assert dev_lib.devarch_global_sums(*head, *tail, _dp(sums)) == GS_NSUMS
    want = reference_sums(prob, state, dim)
    assert sums_close(sums, want), np.abs(sums - want[0]).max(axis=0)
    assert want[0][0][0] > 0 and np.isfinite(want[0][0][5])


# ---- the sums' DEFINITIONS against the reference itself: its CLI writes the global quantities to <root>.global ----------
GLOBAL_TYPES = ["Kinetic Energy", "Work Energy", "Strain Energy", "Heat Energy", "Plastic Energy", "sxx", "syy", "szz", "sxy", "sxz", "syz",
                "velx", "vely", "velz", "Fxx", "Fxy", "Fxz", "Fyx", "Fyy", "Fyz", "Fzx", "Fzy", "Fzz"]


def quantities_from_sums(sums):
    """What GlobalQuantity::AppendQuantity prints (Legacy units), from the per-material raw sums (all non-rigid materials)."""
    from nairn_mpm_fea_b200 import capi as K
    t = np.nansum(sums, axis=0)
    vol = t[K.GS_VOLUME]
    out = {"Kinetic Energy": 1e-9 * t[K.GS_KINETIC], "Work Energy": 1e-9 * t[K.GS_WORK], "Strain Energy": 1e-9 * t[K.GS_STRAIN_ENERGY],
           "Heat Energy": 1e-9 * t[K.GS_HEAT], "Plastic Energy": 1e-9 * t[K.GS_PLASTIC]}
    for name, c in (("sxx", 0), ("syy", 1), ("szz", 2), ("syz", 3), ("sxz", 4), ("sxy", 5)):
        out[name] = 1e-6 * t[K.GS_STRESS + c] / vol
    for c, name in enumerate(("velx", "vely", "velz")):
        out[name] = t[K.GS_VOL_VEL + c] / vol
    for i, name in enumerate(("Fxx", "Fxy", "Fxz", "Fyx", "Fyy", "Fyz", "Fzx", "Fzy", "Fzz")):
        out[name] = 100.0 * t[K.GS_VOL_F + i] / vol
    return out


# cases whose golden state comes from the XML alone (no harness-side jitter), so the CLI run reproduces it
@pytest.mark.parametrize("case", ["block3d_neohookean_uj1", "disks2d_isoplastic", "block3d_rigid_wall_lattice", "disks2d_neohookean", "disks2d_mooney_planestrain"])
def test_global_sums_reproduce_the_reference_cli_global_file(libs, case):  # noqa: F811
    import os
    import re
    import subprocess
    import tempfile
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "NairnMPM")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/NairnMPM not built")
    dev_lib, _ = libs
    z = load_golden(case)
    prob = problem.from_reference_dump(z)
    xml = str(z["xml"])
    dt_ms = prob.dt * 1.0e3
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    glob = '<GlobalArchiveTime units="ms">%r</GlobalArchiveTime>' % (0.999 * dt_ms) + "".join('<GlobalArchive type="%s"/>' % t for t in GLOBAL_TYPES)
    xml = xml.replace("</MPMHeader>", glob + "</MPMHeader>")
    tmax = (snaps[-1] + 0.5) * dt_ms
    xml = re.sub(r'<MaxTime units="ms">[^<]*</MaxTime>', '<MaxTime units="ms">%r</MaxTime>' % tmax, xml)
    d = tempfile.mkdtemp(prefix="glob_")
    open(os.path.join(d, "in.fmcmd"), "w").write(xml)
    p = subprocess.run([ref, "-np", "1", "in.fmcmd"], cwd=d, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-1000:] + p.stderr[-1000:]
    root = re.search(r"<ArchiveRoot>([^<]*)</ArchiveRoot>", xml).group(1).rstrip(".")
    rows = [ln.split("\t") for ln in open(os.path.join(d, root + ".global")) if ln and not ln.startswith("#")]
    names = None
    for ln in open(os.path.join(d, root + ".global")):
        if ln.startswith("#setName"):
            names = [s.strip().strip('"') for s in ln.rstrip("\n").split("\t")[1:]]
    table = np.array([[float(x) for x in r] for r in rows if len(r) > 1])
    dim = 3 if prob.is3d else 2
    checked = 0
    for step in snaps:
        row = table[np.argmin(np.abs(table[:, 0] - step * dt_ms))]
        assert abs(row[0] - step * dt_ms) < 0.01 * dt_ms
        _, n, dev, state, kinds, params = device_view(z, prob, step)
        nn = int(prob.particles.get("n_nonrigid", n))
        dev_nr = {k: np.ascontiguousarray(v[..., :nn]) for k, v in dev.items()}
        head, tail = _args(dim, nn, dev_nr, kinds, params)
        sums = np.zeros((len(kinds), GS_NSUMS))
        assert dev_lib.devarch_global_sums(*head, *tail, _dp(sums)) == GS_NSUMS
        got = quantities_from_sums(sums)
        big = max(abs(v) for k, v in got.items() if k[0] == "s")
        for k, nm in enumerate(names):
            want = row[1 + k]
            scale = {"s": big, "v": float(np.abs(state["vel"]).max()), "F": 100.0}.get(nm[0], abs(want))
            # the file carries 7 significant digits
            assert abs(got[nm] - want) <= 2.0e-6 * max(scale, abs(want)) + 1e-300, (case, step, nm, got[nm], want)
            checked += 1
    assert checked >= len(GLOBAL_TYPES)
