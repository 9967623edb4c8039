"""CPDI node iteration of the DEVICE source (csrc/shape.cuh) compiled for the host (tests/devlaws): the merged variant
(for_each_node_cpdi_merged, one call per touched node; opt-in MPMGPU_CPDI_MERGE=1) must hand every node the same total
weights as the plain one (one call per corner node), with fewer calls -- on domains inside one element, straddling two
elements per axis, stretched beyond the three-node window (fallback path) and with corners exactly on element faces
(zero shape function values are skipped)."""
import ctypes as C

import numpy as np
import pytest

from tests.test_device_laws_cpu import libs  # noqa: F401

LCPDI, QCPDI = 10, 11


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def synthetic_domains(dim, nc, n, horiz, vert, depth, rng):
    """cpElem [nc][n] (1-based), cpXi [3*nc][n], cpWg [3*nc][n] for four kinds of particle domain."""
    elem = np.zeros((nc, n), np.int32)
    xi = rng.uniform(-1.0, 1.0, (3 * nc, n))
    wg = rng.standard_normal((3 * nc, n))
    kind = np.arange(n) % 4            # 0 one element, 1 two elements per axis, 2 stretched (offsets up to 3), 3 corners on faces
    bi = rng.integers(1, horiz - 4, n)
    bj = rng.integers(1, vert - 4, n)
    bk = rng.integers(1, depth - 4, n) if dim == 3 else np.zeros(n, np.int64)
    for c in range(nc):
        span = np.where(kind == 0, 1, np.where(kind == 2, 4, 2))
        oi, oj = rng.integers(0, 4, n) % span, rng.integers(0, 4, n) % span
        ok = rng.integers(0, 4, n) % span if dim == 3 else np.zeros(n, np.int64)
        elem[c] = 1 + (bi + oi) + horiz * ((bj + oj) + vert * (bk + ok))
    face = kind == 3
    for r in range(3 * nc):
        snap = face & (rng.random(n) < 0.5)
        xi[r] = np.where(snap, np.sign(xi[r]), xi[r])
    if dim == 2:
        xi[2::3] = 0.0
        wg[2::3] = 0.0
    return elem, np.ascontiguousarray(xi), np.ascontiguousarray(wg), kind


def _real_domains(n, horiz, vert, depth, rng, stretch, cell=(1.0, 0.5, 0.25), origin=(-2.0, 1.0, 0.125)):
    """Particles with deformation gradients around the identity (rotation + stretch up to `stretch`) well inside a grid with unequal
    cell sizes per axis: the input of the 3D lCPDI iterations (corners are derived from pos, F, lp)."""
    xp, yp, zp = (origin[d] + cell[d] * np.arange(m + 1) for d, m in enumerate((horiz, vert, depth)))
    ijk = np.stack([rng.integers(2, m - 2, n) for m in (horiz, vert, depth)])
    frac = rng.uniform(0.0, 1.0, (3, n))
    pos = np.stack([origin[d] + cell[d] * (ijk[d] + frac[d]) for d in range(3)])
    elem = (1 + ijk[0] + horiz * (ijk[1] + vert * ijk[2])).astype(np.int32)
    F = np.zeros((9, n))
    for p in range(n):
        a = rng.standard_normal((3, 3)) * 0.25 * stretch
        m = np.eye(3) + a
        if np.linalg.det(m) < 0.2:
            m = np.eye(3)
        F[:, p] = m.reshape(9)
    lp = np.full((3, n), 0.5)
    return xp, yp, zp, np.ascontiguousarray(pos), np.ascontiguousarray(F), lp, elem


@pytest.mark.parametrize("stretch,rcrit", [(0.3, -1.0), (1.0, -1.0), (2.5, 0.6)])
def test_register_merged_lcpdi3_gives_every_node_the_same_weights(libs, stretch, rcrit):  # noqa: F811
    """3D lCPDI: the merged iteration (hat functions in grid units, signed corner sums, one call per node) against the plain walk of
    the stored corners, on real domains: mildly deformed, strongly deformed (some span three cells: corner-walk fallback), and with
    the rcrit rescaling of ScaleSemiSideVectorsForCPDI."""
    dev, _ = libs
    horiz, vert, depth, n = 14, 13, 12, 1500
    rng = np.random.default_rng(int(10 * stretch) + 3)
    xp, yp, zp, pos, F, lp, elem = _real_domains(n, horiz, vert, depth, rng, stretch)
    nnodes = (horiz + 1) * (vert + 1) * (depth + 1)
    plain, mer = np.zeros((4, nnodes)), np.zeros((4, nnodes))
    cp, cm = np.zeros(nnodes, np.int32), np.zeros(nnodes, np.int32)
    fb = C.c_int(0)
    rc = dev.devshape_cpdi3_particles(horiz, vert, depth, _dp(xp), _dp(yp), _dp(zp), C.c_double(rcrit), n, _dp(pos), _dp(F), _dp(lp), _ip(elem),
                                      _dp(plain), _ip(cp), _dp(mer), _ip(cm), C.byref(fb))
    assert rc == 0
    scale = np.abs(plain).max(axis=1, keepdims=True)
    assert np.all(np.abs(mer - plain) <= 1e-12 * scale), float((np.abs(mer - plain) / scale).max())
    assert abs(plain[0].sum() - n) < 1e-9 * n and abs(mer[0].sum() - n) < 1e-9 * n          # partition of unity
    assert np.all(np.abs(mer[1:].sum(axis=1)) < 1e-9 * n)                                       # gradients sum to zero
    assert cm.sum() < 0.5 * cp.sum(), (cm.sum(), cp.sum())
    if stretch < 0.5:
        assert fb.value == 0
    if stretch > 2.0 and rcrit < 0:
        assert fb.value > 0, "some strongly stretched domains should take the corner-walk fallback"


@pytest.mark.parametrize("dim,shape,nc", [(2, LCPDI, 4), (2, QCPDI, 9)])
def test_merged_cpdi_iteration_gives_every_node_the_same_weights(libs, dim, shape, nc):  # noqa: F811
    dev, _ = libs
    horiz, vert, depth = 12, 11, 10
    n = 2000
    rng = np.random.default_rng(5 + dim + shape)
    elem, xi, wg, kind = synthetic_domains(dim, nc, n, horiz, vert, depth, rng)
    nnodes = (horiz + 1) * (vert + 1) * ((depth + 1) if dim == 3 else 1)
    res = []
    for merged in (0, 1):
        out, calls = np.zeros((4, nnodes)), np.zeros(nnodes, np.int32)
        assert dev.devshape_cpdi_nodes(dim, shape, merged, horiz, vert, depth, n, _ip(elem), _dp(xi), _dp(wg), _dp(out), _ip(calls)) == 0
        res.append((out, calls))
    (plain, cp), (mer, cm) = res
    assert plain[0].sum() > 0
    scale = np.abs(plain).max(axis=1, keepdims=True)
    assert np.all(np.abs(mer - plain) <= 1e-13 * scale), float((np.abs(mer - plain) / scale).max())
    assert np.array_equal(cp > 0, cm > 0), "the same set of nodes is touched"
    assert cm.sum() < (0.5 if dim == 3 else 0.8) * cp.sum(), (cm.sum(), cp.sum())
    # partition of unity: the S weights of one particle sum to 1 whatever the corner positions
    assert abs(plain[0].sum() - n) < 1e-9 * n and abs(mer[0].sum() - n) < 1e-9 * n


def test_merged_iteration_calls_once_per_node_inside_one_element(libs):  # noqa: F811
    """Undeformed domains that fit inside their own element: 8 calls instead of 64."""
    dev, _ = libs
    horiz, vert, depth, n = 8, 8, 8, 50
    rng = np.random.default_rng(1)
    xp, yp, zp, pos, F, lp, elem = _real_domains(n, horiz, vert, depth, rng, 0.0)
    cell = np.array([xp[1] - xp[0], yp[1] - yp[0], zp[1] - zp[0]])
    org = np.array([xp[0], yp[0], zp[0]])
    ijk = np.floor((pos - org[:, None]) / cell[:, None])
    pos = org[:, None] + cell[:, None] * (ijk + rng.uniform(0.3, 0.7, (3, n)))        # corners at +-0.25 cell stay inside
    pos = np.ascontiguousarray(pos)
    nnodes = 9 * 9 * 9
    plain, mer = np.zeros((4, nnodes)), np.zeros((4, nnodes))
    cp, cm = np.zeros(nnodes, np.int32), np.zeros(nnodes, np.int32)
    fb = C.c_int(0)
    assert dev.devshape_cpdi3_particles(horiz, vert, depth, _dp(xp), _dp(yp), _dp(zp), C.c_double(-1.0), n, _dp(pos), _dp(F), _dp(lp), _ip(elem),
                                        _dp(plain), _ip(cp), _dp(mer), _ip(cm), C.byref(fb)) == 0
    assert [int(cp.sum()), int(cm.sum())] == [64 * n, 8 * n]


# ---- Linear / uGIMP shape functions and the element search: device source against the C restatement ---------------------
def _grid(dim, horiz=9, vert=8, depth=7, cell=(1.0, 0.5, 0.25), origin=(-2.0, 1.0, 0.125)):
    from nairn_mpm_fea_b200.capi import ABI_VERSION, Config
    xp = origin[0] + cell[0] * np.arange(horiz + 1)
    yp = origin[1] + cell[1] * np.arange(vert + 1)
    zp = origin[2] + cell[2] * np.arange(depth + 1)
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.np = 12 if dim == 3 else 10
    cfg.horiz, cfg.vert, cfg.depth = horiz, vert, depth if dim == 3 else 0
    cfg.xpts, cfg.ypts, cfg.zpts = _dp(xp), _dp(yp), (_dp(zp) if dim == 3 else None)
    cfg.gridx, cfg.gridy, cfg.gridz = cell[0], cell[1], (cell[2] if dim == 3 else 0.0)
    return cfg, (xp, yp, zp), (horiz, vert, depth)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shape", [0, 1])          # Linear ("Classic"), uGIMP
def test_shape_functions_of_the_device_source_are_bitwise_those_of_the_oracle(libs, dim, shape):  # noqa: F811
    dev, orc = libs
    cfg, (xp, yp, zp), (horiz, vert, depth) = _grid(dim)
    cfg.shape = shape
    n = 3000
    rng = np.random.default_rng(17 + 2 * dim + shape)
    ei, ej = rng.integers(1, horiz - 1, n), rng.integers(1, vert - 1, n)
    ek = rng.integers(1, depth - 1, n) if dim == 3 else np.zeros(n, np.int64)
    in_elem = np.ascontiguousarray(1 + ei + horiz * (ej + vert * ek), dtype=np.int32)
    ncpos = rng.uniform(-1.0, 1.0, (3, n))
    # the branch points of the GIMP weight: on the element faces, at |xi - xi_node| = lp and 2 - lp
    special = np.array([-1.0, 1.0, 0.0, -0.5, 0.5, -0.25, 0.75])
    pick = rng.random((3, n)) < 0.3
    ncpos = np.where(pick, special[rng.integers(0, special.size, (3, n))], ncpos)
    lp = np.where(rng.random((3, n)) < 0.5, 0.5, rng.uniform(0.2, 1.0, (3, n)))
    if dim == 2:
        ncpos[2] = 0.0
    ncpos, lp = np.ascontiguousarray(ncpos), np.ascontiguousarray(lp)
    out = []
    for which in ("oracle", "device"):
        count = np.zeros(n, np.int32)
        nds = np.full((n, 64), -1, np.int32)
        fn, xd, yd, zd = (np.zeros((n, 64)) for _ in range(4))
        if which == "oracle":
            rc = orc.oracle_shape_batch(C.byref(cfg), n, _ip(in_elem), _dp(ncpos), _dp(lp), 1, _ip(count), _ip(nds), _dp(fn), _dp(xd), _dp(yd), _dp(zd))
        else:
            rc = dev.devshape_nodes(dim, cfg.np, shape, horiz, vert, depth, _dp(xp), _dp(yp), _dp(zp), C.c_double(cfg.gridx), C.c_double(cfg.gridy),
                                    C.c_double(cfg.gridz), n, _ip(in_elem), _dp(ncpos), _dp(lp), _ip(count), _ip(nds), _dp(fn), _dp(xd), _dp(yd), _dp(zd))
        assert rc == 0
        out.append((count, nds, fn, xd, yd, zd))
    (co, no, fo, xo, yo, zo), (cd, nd_, fd, xdd, ydd, zdd) = out
    assert np.array_equal(co, cd), "same number of nodes per particle"
    assert co.min() >= (4 if dim == 2 else 8) and co.max() <= (16 if dim == 2 else 64)
    for p in range(n):
        k = co[p]
        # the device walks its own node order; the values per node must be the same bits
        io, idv = np.argsort(no[p, :k], kind="stable"), np.argsort(nd_[p, :k], kind="stable")
        assert np.array_equal(no[p, :k][io], nd_[p, :k][idv]), p
        for a, b, nm in ((fo, fd, "S"), (xo, xdd, "dS/dx"), (yo, ydd, "dS/dy"), (zo, zdd, "dS/dz")):
            assert np.array_equal(a[p, :k][io], b[p, :k][idv]), (p, nm, a[p, :k][io] - b[p, :k][idv])


@pytest.mark.parametrize("dim", [2, 3])
def test_element_search_of_the_device_source_is_the_oracles(libs, dim):  # noqa: F811
    """MeshInfo::FindElementFromPoint on points inside, on cell faces (the integer must be the reference's), on the far grid
    faces (the reference assigns them to the last element) and off the grid."""
    dev, orc = libs
    cfg, (xp, yp, zp), (horiz, vert, depth) = _grid(dim)
    rng = np.random.default_rng(23 + dim)
    n = 6000
    lo = np.array([xp[0], yp[0], zp[0]])
    hi = np.array([xp[-1], yp[-1], zp[-1]])
    x = lo + (hi - lo) * rng.uniform(-0.05, 1.05, (n, 3))
    faces = [xp, yp, zp]
    for c in range(3):
        on_face = rng.random(n) < 0.3
        x[:, c] = np.where(on_face, faces[c][rng.integers(0, faces[c].size, n)], x[:, c])
        nudge = rng.random(n) < 0.15
        x[:, c] = np.where(nudge, np.nextafter(x[:, c], np.where(rng.random(n) < 0.5, -np.inf, np.inf)), x[:, c])
    if dim == 2:
        x[:, 2] = 0.0
    x = np.ascontiguousarray(x)
    eo, ed = np.zeros(n, np.int32), np.zeros(n, np.int32)
    xi, inside = np.zeros((n, 3)), np.zeros(n, np.int32)
    assert orc.oracle_find_element_batch(C.byref(cfg), n, _dp(x), _ip(eo)) == 0
    assert dev.devshape_find_element(dim, cfg.np, horiz, vert, depth, _dp(xp), _dp(yp), _dp(zp), C.c_double(cfg.gridx), C.c_double(cfg.gridy),
                                     C.c_double(cfg.gridz), n, _dp(x), _ip(ed), _dp(xi), _ip(inside)) == 0
    assert np.array_equal(eo, ed)
    assert 0.05 < np.mean(eo == 0) < 0.6 and np.count_nonzero(eo) > n // 3
    # natural coordinates of a point of the grid box lie in [-1, 1] in the element found for it, and PtInElement agrees except
    # on the far faces (the (int) truncation also "finds" points up to one cell below the low faces -- the reference's behaviour,
    # reproduced by both)
    d = 3 if dim == 3 else 2
    # (a point one ulp inside a far face rounds to column = horiz in (x - xmin)/dx and is declared off the grid by the reference's
    # arithmetic -- also reproduced by both -- so the sanity check stays 1e-9 away from the far faces)
    box = np.all((x[:, :d] >= lo[:d]) & (x[:, :d] <= hi[:d] - 1e-9), axis=1)
    assert np.all(ed[box] > 0)
    assert np.all(np.abs(xi[box][:, :d]) <= 1.0 + 1e-12)          # a point one ulp under a face can land in the cell above it
    assert np.mean(inside[box] == 1) > 0.95
