"""Particle archives in the reference's binary format, written from the state libmpmgpu hands back.

The reference writes one file per archive time: a 64-byte header (ArchiveData::SetArchiveHeader,
NairnMPM/src/System/ArchiveData.cpp:464-489) followed by one fixed-size record per material point
(ArchiveData::ArchiveResults, :806-1100; record size CalcArchiveSize :328-396).  This module packs the same
records from the structure-of-arrays state of `MpmGpu.download()` so that a host which does not run inside the
reference driver (the Python host of this package) still produces files the reference's tools read.  Inside the
reference driver the adapter refreshes `mpm[]` and the reference's own archiver writes the files
(INTEGRATION.md); this is the same format from the other side of the boundary.

Supported order flags (`<MPMArchiveOrder>`, ArchiveData.hpp:22-32): velocity, stress, strain, plastic strain, work
energy, temperature, plastic energy, strain energy, history 1-4, concentration (zeros: no
transport on this path), heat energy, element crossings, initial rotation angles.  Anything else (shear components need
the polar decomposition of F, damage normals, spin, history 5-19, particle size) raises.
"""
import struct

import numpy as np

from . import materials as M

HEADER_LENGTH = 64
ARCH_MAXMPMITEMS = 24
(ARCH_Velocity, ARCH_Stress, ARCH_Strain, ARCH_PlasticStrain, ARCH_OldOrigPosition, ARCH_WorkEnergy, ARCH_DeltaTemp,
 ARCH_PlasticEnergy, ARCH_ver2Empty, ARCH_ShearComponents, ARCH_StrainEnergy, ARCH_History, ARCH_Concentration,
 ARCH_HeatEnergy, ARCH_ElementCrossings, ARCH_RotStrain, ARCH_DamageNormal, ARCH_SpinMomentum, ARCH_SpinVelocity,
 ARCH_History59, ARCH_History1014, ARCH_History1519, ARCH_Size) = range(2, 25)
UNSUPPORTED = (ARCH_ShearComponents, ARCH_DamageNormal, ARCH_SpinMomentum, ARCH_SpinVelocity, ARCH_History59, ARCH_History1014, ARCH_History1519, ARCH_Size)
DEFAULT_CRACK_ORDER = "iYNNNNN"


def normalise_order(order):
    """Pad an <MPMArchiveOrder> string to the reference's full length; byte 0 says little-endian ('i')."""
    order = "i" + order[1:]
    return order + "N" * (ARCH_MAXMPMITEMS - len(order))


def header(order, crack_order, three_d, time_s, structured=True):
    """ArchiveData.cpp:464-489: "ver6", the two order strings each preceded by its length, '3'/'2', '1'/'0', float32 time in ms."""
    h = b"ver6" + bytes([len(order)]) + order.encode("latin-1") + bytes([len(crack_order)]) + crack_order.encode("latin-1")
    h += (b"3" if three_d else b"2") + (b"1" if structured else b"0")
    h += struct.pack("<f", np.float32(time_s * 1.0e3))
    assert len(h) <= HEADER_LENGTH
    return h + b"\0" * (HEADER_LENGTH - len(h))


def _history_bits(ch):
    if ch == "Y":
        return [1]
    if ch == "N":
        return []
    return [k + 1 for k in range(4) if ord(ch) & (1 << k)]


def records(prob, state, order, origpos=None, thickness=None, angles0=None, temperature=None):
    """The record block as bytes.  state: dict as returned by MpmGpu.download() (pos, vel, sp, pressure, ep, wrot, eplast,
    energies[work, res, heat, entropy, plast, prevT], history, in_elem, crossings).  origpos defaults to the initial
    positions in `prob`, thickness (2D) to the grid thickness, angles0 (initial material angles, radians, [3][n]: z, y, x) to 0."""
    order = normalise_order(order)
    for bit in UNSUPPORTED:
        if bit < len(order) and order[bit] != "N":
            raise NotImplementedError("archive item %d is not produced on this path" % bit)
    three_d = prob.is3d
    pt = prob.particles
    n = int(np.asarray(pt["mp"]).shape[0])
    mp = np.asarray(pt["mp"], np.float64)
    matnum = np.asarray(pt.get("matnum", np.ones(n, np.int32)), np.int32)
    origpos = np.asarray(pt["pos"] if origpos is None else origpos, np.float64)
    angles0 = np.zeros((3, n)) if angles0 is None else np.asarray(angles0, np.float64)
    wrot = np.asarray(state["wrot"], np.float64)
    kinds = np.array([prob.materials[m - 1]["kind"] for m in matnum])
    rho0 = np.array([prob.materials[m - 1]["rho"] for m in matnum])
    hist = np.asarray(state["history"], np.float64)
    # current density: rho0 / GetCurrentRelativeVolume (1 unless the material tracks J: Neohookean.cpp:374-376)
    relvol = np.where((kinds == M.NEOHOOKEAN) | (kinds == M.MOONEY), hist[0], 1.0)
    rho = rho0 / relvol
    # total stress: materials that keep the pressure apart add it back (MaterialBase::GetStressPandDev, MaterialBaseMPM.cpp:1635-1641)
    sp = np.array(state["sp"], np.float64, copy=True)
    pand = (kinds == M.NEOHOOKEAN) | (kinds == M.ISOPLASTICITY) | (kinds == M.MOONEY)
    for c in range(3):
        sp[c] = np.where(pand, sp[c] - np.asarray(state["pressure"]), sp[c])
    en = np.asarray(state["energies"], np.float64)
    cols = []           # list of (dtype, array) in record order

    def d(a):
        cols.append(("<f8", np.asarray(a, np.float64)))

    cols.append(("<i4", np.asarray(state["in_elem"], np.int32)))
    d(mp)
    cols.append(("<i2", matnum.astype(np.int16)))
    cols.append(("<i2", np.zeros(n, np.int16)))         # two zero bytes for alignment
    pi = 3.141592653589793          # PI_CONSTANT; same operation order as MPMBase.cpp:601-619
    if three_d:
        d(180.0 * (angles0[0] - 0.5 * wrot[0]) / pi); d(180.0 * (angles0[1] + 0.5 * wrot[1]) / pi); d(180.0 * (angles0[2] - 0.5 * wrot[2]) / pi)
    else:
        d(180.0 * (angles0[0] - 0.5 * wrot[0]) / pi)
        d(np.full(n, prob.thickness) if thickness is None else thickness)
    dim = 3 if three_d else 2
    for c in range(dim):
        d(state["pos"][c])
    for c in range(dim):
        d(origpos[c])
    if order[ARCH_Velocity] == "Y":
        for c in range(dim):
            d(state["vel"][c])
    tens = (0, 1, 2, 5, 4, 3) if three_d else (0, 1, 2, 5)          # xx yy zz xy [xz yz]; state tensors are xx yy zz yz xz xy
    if order[ARCH_Stress] == "Y":
        for c in tens:
            d(rho * sp[c])
    if order[ARCH_Strain] == "Y":
        for c in tens:
            d(state["ep"][c])
    if order[ARCH_PlasticStrain] == "Y":
        for c in tens:
            d(state["eplast"][c])
    if order[ARCH_WorkEnergy] == "Y":
        d(1.0e-9 * mp * en[0])
    if order[ARCH_DeltaTemp] == "Y":
        d(en[5] if temperature is None else temperature)
    if order[ARCH_PlasticEnergy] == "Y":
        d(1.0e-9 * mp * en[4])
    if order[ARCH_StrainEnergy] == "Y":
        d(1.0e-9 * mp * (en[0] - en[1]))
    for k in _history_bits(order[ARCH_History]):
        d(hist[k - 1])
    if order[ARCH_Concentration] == "Y":
        for _ in range(dim + 1):
            d(np.zeros(n))
    if order[ARCH_HeatEnergy] == "Y":
        d(1.0e-9 * mp * en[2])
    if order[ARCH_ElementCrossings] == "Y":
        cols.append(("<i4", np.abs(np.asarray(state["crossings"], np.int32))))
    if order[ARCH_RotStrain] == "Y":
        for c in range(3 if three_d else 1):
            d(180.0 * angles0[c] / pi)
    rec = np.dtype([("f%d" % i, t) for i, (t, _) in enumerate(cols)])
    out = np.zeros(n, rec)
    for i, (_, a) in enumerate(cols):
        out["f%d" % i] = a
    return out.tobytes(), rec.itemsize


def write_archive(path, prob, state, order, time_s, crack_order=DEFAULT_CRACK_ORDER, **kw):
    """Write one archive file; returns the record size in bytes."""
    body, recsize = records(prob, state, order, **kw)
    with open(path, "wb") as f:
        f.write(header(normalise_order(order), crack_order, prob.is3d, time_s))
        f.write(body)
    return recsize
