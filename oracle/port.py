"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/_build/libmpm_oracle.so (oracle/mpm_oracle.c, the plain-C
restatement of the reference step).  Only tests/, smoke() and bench.py's cpu_baseline may import this."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from nairn_mpm_fea_b200.capi import (Config, Material, NodesView, ParticlesView, ABI_VERSION, MAX_HISTORY,  # noqa: E402
                                     _c32, _c64, _d, _i)

LIB = os.path.join(HERE, "_build", "libmpm_oracle.so")


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return LIB


class PortOracle:
    """One run of the C restatement (process-global state: one at a time)."""

    def __init__(self, prob):
        if not os.path.exists(LIB):
            build()
        self.lib = C.CDLL(LIB)
        self.prob = prob
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.np = prob.np
        cfg.horiz, cfg.vert, cfg.depth = prob.horiz, prob.vert, prob.depth
        self._keep = [_c64(prob.xpts), _c64(prob.ypts), _c64(prob.zpts) if prob.is3d else None]
        cfg.xpts, cfg.ypts, cfg.zpts = [_d(a) for a in self._keep]
        cfg.gridx, cfg.gridy, cfg.gridz = prob.grid
        cfg.shape, cfg.method = prob.shape, prob.method
        cfg.cpdi_rcrit = prob.rcrit
        cfg.skip_post_extrapolation = int(prob.skip_post_extrapolation)
        cfg.fraction_usf = prob.fraction_usf
        cfg.xpic_order, cfg.using_fmpm = prob.xpic_order, int(prob.using_fmpm)
        cfg.grid_damping, cfg.particle_damping = prob.grid_damping, prob.particle_damping
        cfg.gravity = (C.c_double * 3)(*prob.gravity)
        mats = (Material * len(prob.materials))()
        for k, m in enumerate(prob.materials):
            mats[k].kind, mats[k].n_history = m["kind"], m.get("n_history", 0)
            for j, v in enumerate(m["p"]):
                mats[k].p[j] = v
        pt = prob.particles
        v = ParticlesView()
        self.n = int(np.asarray(pt["mp"]).shape[0])
        v.n, v.n_nonrigid = self.n, int(pt.get("n_nonrigid", self.n))
        keep = {}
        for k in ("pos", "vel", "mp", "lp", "sp", "pressure", "ep", "wrot", "eplast", "energies", "history", "pfext"):
            keep[k] = _c64(pt.get(k))
            setattr(v, k, _d(keep[k]))
        for k in ("in_elem", "matnum", "crossings"):
            keep[k] = _c32(pt.get(k))
            setattr(v, k, _i(keep[k]))
        bn, bnorm, bval = _c32(prob.bc_node), _c64(prob.bc_norm), _c64(prob.bc_value)
        bact, bsym = _c32(prob.bc_active), _c32(prob.bc_symdir)
        self.lib.oracle_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(Material), C.POINTER(ParticlesView), C.c_int,
                                           C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double]
        rc = self.lib.oracle_create(C.byref(cfg), len(prob.materials), mats, C.byref(v), len(bn), _i(bn), _d(bnorm), _d(bval),
                                    _i(bact), _i(bsym), prob.dt, prob.dt_strain_first, prob.dt_strain_last)
        assert rc == 0
        if getattr(prob, "bc_reflected", None) is not None:
            refl, ratio = _c32(prob.bc_reflected), _c64(prob.bc_ratio)
            self.lib.oracle_set_bc_reflections.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
            assert self.lib.oracle_set_bc_reflections(len(refl), _i(refl), _d(ratio)) == 0

    def set_xpic(self, order, using_fmpm):
        assert self.lib.oracle_set_xpic(int(order), int(using_fmpm)) == 0

    def step(self, n=1):
        assert self.lib.oracle_step(int(n)) == 0

    def run_task(self, t):
        assert self.lib.oracle_task(int(t)) == 0

    def download(self):
        n = self.n
        out = dict(pos=np.zeros((3, n)), vel=np.zeros((3, n)), sp=np.zeros((6, n)), pressure=np.zeros(n), ep=np.zeros((6, n)),
                   wrot=np.zeros((3, n)), eplast=np.zeros((6, n)), energies=np.zeros((6, n)), history=np.zeros((MAX_HISTORY, n)),
                   acc=np.zeros((3, n)), in_elem=np.zeros(n, np.int32), crossings=np.zeros(n, np.int32))
        v = ParticlesView()
        for k in ("pos", "vel", "sp", "pressure", "ep", "wrot", "eplast", "energies", "history", "acc"):
            setattr(v, k, _d(out[k]))
        v.in_elem, v.crossings = _i(out["in_elem"]), _i(out["crossings"])
        self.lib.oracle_get_particles(C.byref(v))
        return out

    def download_nodes(self):
        n = self.prob.nnodes
        out = dict(number_points=np.zeros(n, np.int32), mass=np.zeros(n), pk=np.zeros((3, n)), ftot=np.zeros((3, n)),
                   vk=np.zeros((3, n)), pk_copy=np.zeros((3, n)))
        v = NodesView()
        v.number_points = _i(out["number_points"])
        for k in ("mass", "pk", "ftot", "vk", "pk_copy"):
            setattr(v, k, _d(out[k]))
        self.lib.oracle_get_nodes(C.byref(v))
        return out

    def close(self):
        self.lib.oracle_destroy()
