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


@pytest.mark.parametrize("dim,shape,nc", [(3, LCPDI, 8), (2, LCPDI, 4), (2, QCPDI, 9)])
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
    """Undeformed lattice: all 8 corners in the particle's own element -> 8 calls instead of 64."""
    dev, _ = libs
    horiz, vert, depth, n, nc = 8, 8, 8, 50, 8
    rng = np.random.default_rng(1)
    elem = np.zeros((nc, n), np.int32)
    base = 1 + rng.integers(1, 6, n) + horiz * (rng.integers(1, 6, n) + vert * rng.integers(1, 6, n))
    elem[:] = base
    xi = np.ascontiguousarray(rng.uniform(-0.9, 0.9, (3 * nc, n)))
    wg = np.ascontiguousarray(rng.standard_normal((3 * nc, n)))
    nnodes = 9 * 9 * 9
    tot = []
    for merged in (0, 1):
        out, calls = np.zeros((4, nnodes)), np.zeros(nnodes, np.int32)
        assert dev.devshape_cpdi_nodes(3, LCPDI, merged, horiz, vert, depth, n, _ip(elem), _dp(xi), _dp(wg), _dp(out), _ip(calls)) == 0
        tot.append(int(calls.sum()))
    assert tot == [64 * n, 8 * n]
