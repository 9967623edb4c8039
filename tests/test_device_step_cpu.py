"""The per-task kernels of libmpmgpu (csrc/kernels_task.cuh + shape.cuh + materials.cuh) compiled for the host and run one CUDA
thread after the other (tests/devlaws/host_step.cpp), in the reference's task order, against the golden dumps of the unmodified
reference: every task of step 1 and whole runs, same tolerances as the GPU parity tests, for EVERY golden case (grid velocity
BCs, rigid-BC particles, XPIC/FMPM, CPDI, every material and analysis type).

This checks the CUDA SOURCE of the general path on a machine without a GPU.  It is test infrastructure -- the product has no
CPU path (tests/test_host_cpu.py::test_no_device_fails_loudly)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nairn_mpm_fea_b200 import materials as M
from nairn_mpm_fea_b200.problem import from_reference_dump
from tests.parity import TASK_MAP, compare_nodes, compare_particles, load_golden, per_task_steps, tolerances, xpic_for_step

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEV = os.path.join(HERE, "devlaws")
LIB = os.path.join(DEV, "_build", "libdevstep.so")

# (the multimaterial goldens mm* and the conduction goldens cond* have their own checks: tests/test_multimaterial_cpu.py, tests/test_conduction_cpu.py)
CASES = sorted(os.path.basename(f)[:-4] for f in os.listdir(os.path.join(HERE, "golden")) if f.endswith(".npz") and not f.startswith(("mm", "cond", "th", "react2d_multimaterial", "trac2d_multimaterial")))
TASK_INDEX = {"initialization": 0, "mass_and_momentum": 1, "post_extrapolation": 2, "update_strains_first": 3, "grid_forces": 4,
              "post_forces": 5, "update_momenta": 6, "update_particles": 7, "update_strains_last": 8, "reset_elements": 9,
              "project_rigid_bcs": 10}
MERGED = {10: 20, 11: 21}          # SHAPE_LCPDI -> SHAPE_LCPDI_MERGED, SHAPE_QCPDI -> SHAPE_QCPDI_MERGED


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(DEV, "host_step.cpp")
    csrc = os.path.join(ROOT, "nairn_mpm_fea_b200", "csrc")
    deps = [src, os.path.join(DEV, "stub", "cuda_runtime.h")] + [os.path.join(csrc, f) for f in ("kernels_task.cuh", "shape.cuh", "materials.cuh", "mpm_types.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-std=c++17", "-I" + os.path.join(DEV, "stub"), "-I" + csrc,
                        src, "-o", LIB], check=True)
    lib = C.CDLL(LIB)
    lib.emu_create.restype = C.c_void_p
    for f in ("emu_set_multimaterial", "emu_get_contact", "emu_set_conduction", "emu_get_transport", "emu_set_temperature_bcs", "emu_set_energy_coupling", "emu_get_reactions", "emu_set_tractions", "emu_set_heat_fluxes", "emu_set_bcs", "emu_set_bc_reflections", "emu_set_xpic", "emu_task", "emu_step", "emu_flags", "emu_get_particles", "emu_get_nodes", "emu_get_device_state", "emu_destroy"):
        getattr(lib, f).argtypes = None
    return lib


class EmuSim:
    def __init__(self, lib, prob, merged_cpdi=False):
        self.lib, self.prob = lib, prob
        c = np.ascontiguousarray
        pt = prob.particles
        self.n = n = prob.nparticles
        f64 = lambda k, shape: c(pt[k], dtype=np.float64) if pt.get(k) is not None else np.zeros(shape)          # noqa: E731
        keep = dict(pos=f64("pos", (3, n)), vel=f64("vel", (3, n)), mp=f64("mp", n), lp=f64("lp", (3, n)), sp=f64("sp", (6, n)),
                    pressure=f64("pressure", n), ep=f64("ep", (6, n)), wrot=f64("wrot", (3, n)), eplast=f64("eplast", (6, n)),
                    energies=f64("energies", (6, n)))
        hist = np.zeros((M.MAX_HISTORY, n))
        h = np.asarray(pt.get("history", hist))
        hist[:min(h.shape[0], M.MAX_HISTORY)] = h[:M.MAX_HISTORY]
        elem, matnum = c(pt["in_elem"], dtype=np.int32), c(pt["matnum"], dtype=np.int32)
        cross = c(pt.get("crossings", np.zeros(n)), dtype=np.int32)
        kinds = c([m["kind"] for m in prob.materials], dtype=np.int32)
        nhist = c([m.get("n_history", 0) for m in prob.materials], dtype=np.int32)
        params = c(np.stack([m["p"] for m in prob.materials]), dtype=np.float64)
        xp, yp = c(prob.xpts, dtype=np.float64), c(prob.ypts, dtype=np.float64)
        zp = c(prob.zpts, dtype=np.float64) if prob.is3d else None
        grav = c(prob.gravity, dtype=np.float64)
        shape = MERGED.get(prob.shape, prob.shape) if merged_cpdi else prob.shape
        d = C.c_double
        self.h = C.c_void_p(lib.emu_create(
            prob.np, prob.horiz, prob.vert, prob.depth, _dp(xp), _dp(yp), _dp(zp), d(prob.grid[0]), d(prob.grid[1]), d(prob.grid[2]),
            shape, d(prob.rcrit), prob.method, int(prob.skip_post_extrapolation), d(prob.fraction_usf), prob.xpic_order, int(prob.using_fmpm),
            d(prob.grid_damping), d(prob.particle_damping), _dp(grav), d(prob.dt), d(prob.dt_strain_first), d(prob.dt_strain_last),
            len(prob.materials), _ip(kinds), _ip(nhist), _dp(params),
            n, int(pt.get("n_nonrigid", n)), _dp(keep["pos"]), _dp(keep["vel"]), _dp(keep["mp"]), _dp(keep["lp"]), _ip(elem), _ip(matnum), _dp(keep["sp"]), _dp(keep["pressure"]),
            _dp(keep["ep"]), _dp(keep["wrot"]), _dp(keep["eplast"]), _dp(keep["energies"]), _dp(hist), _ip(cross)))
        nb = int(np.asarray(prob.bc_node).size)
        if nb:
            self._bc = [c(prob.bc_node, dtype=np.int32), c(prob.bc_norm, dtype=np.float64), c(prob.bc_value, dtype=np.float64),
                        c(prob.bc_active, dtype=np.int32), c(prob.bc_symdir, dtype=np.int32)]
            lib.emu_set_bcs(self.h, nb, _ip(self._bc[0]), _dp(self._bc[1]), _dp(self._bc[2]), _ip(self._bc[3]), _ip(self._bc[4]))
            if getattr(prob, "bc_reflected", None) is not None:
                self._bc += [c(prob.bc_reflected, dtype=np.int32), c(prob.bc_ratio, dtype=np.float64)]
                lib.emu_set_bc_reflections(self.h, nb, _ip(self._bc[5]), _dp(self._bc[6]))
        self.nnodes = (prob.horiz + 1) * (prob.vert + 1) * ((prob.depth + 1) if prob.is3d else 1)
        self.n_fields = 0
        self.conduction = getattr(prob, "conduction", None) is not None
        self.real_nodes = self.nnodes
        adiabatic = bool(getattr(prob, "adiabatic", False))
        if adiabatic:
            lib.emu_set_energy_coupling(self.h, 1)
        self.thermal = self.conduction or adiabatic or pt.get("temperature") is not None
        if self.conduction:
            self._cond = [c(prob.conduction["kcond"], dtype=np.float64), c(pt["temperature"], dtype=np.float64)]
            lib.emu_set_conduction(self.h, _dp(self._cond[0]), _dp(self._cond[1]))
            if prob.conduction.get("tbc_node") is not None:
                self._cond += [c(prob.conduction["tbc_node"], dtype=np.int32), c(prob.conduction["tbc_value"], dtype=np.float64)]
                lib.emu_set_temperature_bcs(self.h, len(self._cond[2]), _ip(self._cond[2]), _dp(self._cond[3]))
        elif self.thermal:
            self._cond = [c(pt["temperature"] if pt.get("temperature") is not None else keep["energies"][5], dtype=np.float64)]
            lib.emu_set_conduction(self.h, None, _dp(self._cond[0]))
        mm = getattr(prob, "multimaterial", None)
        if mm is not None:
            nf = int(mm["n_fields"])
            orig = c(prob.origpos if prob.origpos is not None else keep["pos"], dtype=np.float64)
            self._mm = [c(mm["field_of_material"], dtype=np.int32), c(mm["contact_normal"], dtype=np.float64), c(np.asarray(mm["law_kind"]).reshape(-1), dtype=np.int32),
                        c(np.asarray(mm["law_friction"]).reshape(-1), dtype=np.float64), c(np.asarray(mm["law_static"]).reshape(-1), dtype=np.float64), orig]
            lib.emu_set_multimaterial(self.h, nf, _ip(self._mm[0]), int(mm["normal_method"]), int(mm["by_displacements"]), d(mm["position_cutoff"]),
                                      _dp(self._mm[1]), _ip(self._mm[2]), _dp(self._mm[3]), _dp(self._mm[4]), _dp(self._mm[5]), d(mm.get("rigid_gradient_bias", 1.0)))
            self.nnodes *= nf
            self.n_fields = nf
        self._set_tractions()
        self._set_heat_fluxes()

    def _set_tractions(self):
        tr = getattr(self.prob, "tractions", None)
        if tr is None or not len(tr["particle"]):
            return
        c = np.ascontiguousarray
        self._tr = [c(tr["particle"], dtype=np.int32), c(tr["face"], dtype=np.int32), c(tr["direction"], dtype=np.int32), c(tr["value"], dtype=np.float64)]
        self.lib.emu_set_tractions(self.h, len(self._tr[0]), _ip(self._tr[0]), _ip(self._tr[1]), _ip(self._tr[2]), _dp(self._tr[3]), C.c_double(float(self.prob.thickness)))

    def _set_heat_fluxes(self):
        hf = getattr(self.prob, "heat_fluxes", None)
        if hf is None or not len(hf["particle"]):
            return
        c = np.ascontiguousarray
        self._hf = [c(hf["particle"], dtype=np.int32), c(hf["face"], dtype=np.int32), c(hf["value"], dtype=np.float64)]
        self.lib.emu_set_heat_fluxes(self.h, len(self._hf[0]), _ip(self._hf[0]), _ip(self._hf[1]), _dp(self._hf[2]), C.c_double(float(self.prob.thickness)))

    def set_xpic(self, order, fmpm):
        self.lib.emu_set_xpic(self.h, int(order), int(fmpm))

    def run_task(self, name):
        self.lib.emu_task(self.h, TASK_INDEX[name])

    def step(self, nsteps=1):
        self.lib.emu_step(self.h, int(nsteps))

    def download(self):
        n = self.n
        o = dict(pos=np.zeros((3, n)), vel=np.zeros((3, n)), sp=np.zeros((6, n)), pressure=np.zeros(n), ep=np.zeros((6, n)), wrot=np.zeros((3, n)),
                 eplast=np.zeros((6, n)), energies=np.zeros((6, n)), history=np.zeros((M.MAX_HISTORY, n)), acc=np.zeros((3, n)),
                 in_elem=np.zeros(n, np.int32), crossings=np.zeros(n, np.int32))
        self.lib.emu_get_particles(self.h, _dp(o["pos"]), _dp(o["vel"]), _dp(o["sp"]), _dp(o["pressure"]), _dp(o["ep"]), _dp(o["wrot"]),
                                   _dp(o["eplast"]), _dp(o["energies"]), _dp(o["history"]), _dp(o["acc"]), _ip(o["in_elem"]), _ip(o["crossings"]))
        if self.thermal:
            o["temperature"] = np.zeros(n)
            t = [np.zeros(self.real_nodes) for _ in range(3)]
            self.lib.emu_get_transport(self.h, _dp(t[0]), _dp(t[1]), _dp(t[2]), _dp(o["temperature"]))
        return o

    def download_nodes(self):
        nn = self.nnodes
        o = dict(number_points=np.zeros(nn, np.int32), mass=np.zeros(nn), pk=np.zeros((3, nn)), ftot=np.zeros((3, nn)), vk=np.zeros((3, nn)),
                 pk_copy=np.zeros((3, nn)))
        self.lib.emu_get_nodes(self.h, _ip(o["number_points"]), _dp(o["mass"]), _dp(o["pk"]), _dp(o["ftot"]), _dp(o["vk"]), _dp(o["pk_copy"]))
        if self.conduction:
            o.update(transport_value=np.zeros(self.real_nodes), transport_capacity=np.zeros(self.real_nodes), transport_rate=np.zeros(self.real_nodes))
            self.lib.emu_get_transport(self.h, _dp(o["transport_value"]), _dp(o["transport_capacity"]), _dp(o["transport_rate"]), _dp(np.zeros(self.n)))
        if self.n_fields:
            o.update(contact_volume=np.zeros(nn), contact_gradient=np.zeros((3, nn)), contact_disp=np.zeros((3, nn)))
            self.lib.emu_get_contact(self.h, _dp(o["contact_volume"]), _dp(o["contact_gradient"]), _dp(o["contact_disp"]))
        return o

    def reactions(self):
        """(per grid BC of the list [n,3], per material of the rigid-BC particles [nmat,3]): mirror of MpmGpu.reactions"""
        bc = np.zeros((len(self._bc[0]) if self.prob.bc_node is not None and len(self.prob.bc_node) else 0, 3))
        rigid = np.zeros((len(self.prob.materials), 3))
        self.lib.emu_get_reactions(self.h, _dp(bc) if len(bc) else None, _dp(rigid))
        return bc, rigid

    def flags(self):
        cr, lg, nan, cp = C.c_longlong(0), C.c_longlong(0), C.c_int(0), C.c_int(0)
        self.lib.emu_flags(self.h, C.byref(cr), C.byref(lg), C.byref(nan), C.byref(cp))
        return dict(crossings=cr.value, left_grid=lg.value & 0xffffffff, nan=nan.value, cpdi_left=cp.value)

    def close(self):
        self.lib.emu_destroy(self.h)


@pytest.mark.parametrize("case", CASES)
def test_device_source_tasks_match_reference(lib, case):
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z))
    names = [str(s) for s in z["task_names"]]
    for step in range(1, per_task_steps(z) + 1):
        x = xpic_for_step(z, step)
        if x:
            sim.set_xpic(*x)
        tol = tolerances(case)[0] if step == 1 else tolerances(case)[1]
        for i, nm in enumerate(names):
            if TASK_MAP[nm] is None:
                continue
            sim.run_task(TASK_MAP[nm])
            pre = "s%d/t%d" % (step, i)
            nodes = sim.download_nodes()
            errs, bad = compare_nodes(nodes, z, pre + "/nodes", tol)
            assert not bad, "%s step %d after task %d (%s): node fields %s" % (case, step, i, nm, bad)
            assert np.array_equal(nodes["number_points"] > 0, z[pre + "/nodes/numberPoints"] > 0)
            got = sim.download()
            errs, bad = compare_particles(got, z, pre + "/p", tol)
            assert not bad, "%s step %d after task %d (%s): particle fields %s" % (case, step, i, nm, bad)
            assert np.array_equal(got["in_elem"], z[pre + "/p/inElem"])
    sim.close()


@pytest.mark.parametrize("case,merged", [(c, False) for c in CASES] + [(c, True) for c in CASES if "cpdi" in c])
def test_device_source_whole_steps_match_reference(lib, case, merged):
    """merged: the CPDI kernels with the corners' contributions merged per node (MPMGPU_CPDI_MERGE=1 in the library)."""
    z = load_golden(case)
    sim = EmuSim(lib, from_reference_dump(z), merged_cpdi=merged)
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        while done < s:
            x = xpic_for_step(z, done + 1)
            if x:
                sim.set_xpic(*x)
            sim.step(1)
            done += 1
        tol = tolerances(case)[0] if s == 1 else tolerances(case)[2]
        got = sim.download()
        errs, bad = compare_particles(got, z, "p%d" % s, tol)
        assert not bad, "%s after %d steps: %s" % (case, s, bad)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
        assert np.array_equal(got["crossings"], z["p%d/crossings" % s])
        errs, bad = compare_nodes(sim.download_nodes(), z, "n%d" % s, tol)
        assert not bad, "%s after %d steps: nodes %s" % (case, s, bad)
    f = sim.flags()
    assert f["nan"] == 0 and f["cpdi_left"] == 0
    sim.close()
