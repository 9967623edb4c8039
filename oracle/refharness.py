"""TEST INFRASTRUCTURE (oracle) -- ctypes access to oracle/_ref/libnairnmpm_ref.so.

The library is the reference NairnMPM, unmodified, plus oracle/ref_harness.cpp.  The reference keeps all
state in process globals, so one process can open ONE input: use `run_reference()` (spawns a worker
process per input) from tests, or `RefRun` directly inside a dedicated process.

Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline arm may import this.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libnairnmpm_ref.so")
BIN = os.path.join(HERE, "_ref", "NairnMPM")

INT_KEYS = ["np", "is3D", "nmpms", "nmpmsNR", "nmpmsRB", "nmpmsRC", "nnodes", "nelems", "horiz", "vert",
            "depth", "mpmApproach", "useGimp", "skipPostExtrapolation", "XPICOrder", "usingFMPM", "mstep",
            "nmat", "maxShapeNodes", "incrementalDefGradTerms", "adiabatic", "conduction", "numPatches",
            "hasGravity", "useDamping", "usePDamping"]
DBL_KEYS = ["xmin", "ymin", "zmin", "gridx", "gridy", "gridz", "timestep", "strainTimestepFirst",
            "strainTimestepLast", "fractionUSF", "mtime", "damping", "pdamping", "gx", "gy", "gz", "rcrit",
            "thickness", "maxtime"]


def available():
    return os.path.exists(LIB)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class RefRun:
    """One reference run in THIS process (cannot be reopened)."""

    def __init__(self, xml_path, nprocs=1, log=None):
        self.lib = C.CDLL(LIB)
        self.lib.ref_last_error.restype = C.c_char_p
        self.lib.ref_task_name.restype = C.c_char_p
        rv = self.lib.ref_open(xml_path.encode(), int(nprocs), (log or "").encode())
        if rv != 0:
            raise RuntimeError("ref_open failed: %s" % self.lib.ref_last_error().decode())
        self.info = self.get_info()

    def get_info(self):
        iv = np.zeros(32, dtype=np.int32)
        dv = np.zeros(32, dtype=np.float64)
        self.lib.ref_info(_ip(iv), _dp(dv))
        d = {k: int(iv[i]) for i, k in enumerate(INT_KEYS)}
        d.update({k: float(dv[i]) for i, k in enumerate(DBL_KEYS)})
        return d

    def set_particles(self, pos=None, vel=None):
        """Overwrite positions/velocities ([3][n] arrays); returns number of particles off the grid."""
        pos = None if pos is None else np.ascontiguousarray(pos, dtype=np.float64)
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        bad = self.lib.ref_set_particles(_dp(pos), _dp(vel))
        self.info = self.get_info()
        return bad

    def step(self, n=1):
        if self.lib.ref_step(int(n)) != 0:
            raise RuntimeError("ref_step failed: %s" % self.lib.ref_last_error().decode())

    def task_names(self):
        return [self.lib.ref_task_name(i).decode() for i in range(self.lib.ref_num_tasks())]

    def run_task(self, i):
        if self.lib.ref_run_task(int(i)) != 0:
            raise RuntimeError("ref_run_task failed: %s" % self.lib.ref_last_error().decode())

    def particles(self, nhist=4):
        n = self.info["nmpms"]
        out = dict(pos=np.zeros((3, n)), vel=np.zeros((3, n)), mp=np.zeros(n), lp=np.zeros((3, n)),
                   inElem=np.zeros(n, np.int32), matnum=np.zeros(n, np.int32), sp=np.zeros((6, n)),
                   pressure=np.zeros(n), ep=np.zeros((6, n)), wrot=np.zeros((3, n)), eplast=np.zeros((6, n)),
                   energies=np.zeros((6, n)), hist=np.zeros((nhist, n)), pFext=np.zeros((3, n)),
                   origpos=np.zeros((3, n)), crossings=np.zeros(n, np.int32), ncpos=np.zeros((3, n)),
                   acc=np.zeros((3, n)))
        o = out
        self.lib.ref_get_particles(_dp(o["pos"]), _dp(o["vel"]), _dp(o["mp"]), _dp(o["lp"]), _ip(o["inElem"]),
                                   _ip(o["matnum"]), _dp(o["sp"]), _dp(o["pressure"]), _dp(o["ep"]),
                                   _dp(o["wrot"]), _dp(o["eplast"]), _dp(o["energies"]), _dp(o["hist"]),
                                   C.c_int(nhist), _dp(o["pFext"]), _dp(o["origpos"]), _ip(o["crossings"]),
                                   _dp(o["ncpos"]), _dp(o["acc"]))
        # pTemperature always (a start off the stress-free temperature strains the particles also without conduction);
        # the temperature gradient only exists with the conduction task
        out.update(temperature=np.zeros(n), tgrad=np.zeros((3, n)))
        self.lib.ref_get_temperatures(_dp(out["temperature"]), _dp(out["tgrad"]))
        if not self.lib.ref_conduction_on():
            del out["tgrad"]
        return out

    def multimaterial(self):
        """<MultiMaterialMode> settings (None when off): fields, normal method, contact laws per field pair."""
        nf = self.lib.ref_num_fields()
        nm = self.info["nmat"]
        out = np.zeros(8, np.int32)
        field = np.zeros(nm, np.int32)
        law = np.zeros((nm, nm, 4))
        normal = np.zeros(5)
        self.lib.ref_get_multimaterial(_ip(out), _ip(field), _dp(law), _dp(normal))
        if not out[3]:
            return None
        return dict(nfields=nf, normal_method=int(out[0]), by_displacements=int(out[1]), field=field, law=law,
                    position_cutoff=float(normal[3]), contact_normal=normal[:3].copy(), rigid_gradient_bias=float(normal[4]))

    def conduction(self):
        """Conduction settings (None when the task is off): kcond per material, counts of the BCs this repo does not build."""
        out = np.zeros(8, np.int32)
        k = np.zeros(self.info["nmat"])
        self.lib.ref_get_conduction(_ip(out), _dp(k))
        if not out[0]:
            return None
        d = dict(kcond=k, adiabatic=int(out[1]), n_temp_bcs=int(out[2]), n_flux_bcs=int(out[3]), contact_heating=int(out[4]))
        nb = self.lib.ref_get_temp_bcs(None, None)
        if nb:
            d["tbc_node"], d["tbc_value"] = np.zeros(nb, np.int32), np.zeros(nb)
            self.lib.ref_get_temp_bcs(_ip(d["tbc_node"]), _dp(d["tbc_value"]))
        return d

    def _transport(self, o):
        if self.lib.ref_conduction_on():
            n = self.info["nnodes"]
            o.update(gT=np.zeros(n), gVCT=np.zeros(n), gQ=np.zeros(n))
            self.lib.ref_get_node_transport(_dp(o["gT"]), _dp(o["gVCT"]), _dp(o["gQ"]))
        return o

    def nodes(self):
        return self._transport(self._nodes())

    def _nodes(self):
        nf = self.lib.ref_num_fields()
        if nf > 1 or self.lib.ref_multimaterial_on():
            # multimaterial mode: every array is field-major, [field][node] flattened (vectors [3][field*nnodes + node])
            n = self.info["nnodes"]
            N = nf * n
            o = dict(numberPoints=np.zeros(N, np.int32), mass=np.zeros(N), pk=np.zeros((3, N)), ftot=np.zeros((3, N)),
                     vk0=np.zeros((3, N)), pkcopy=np.zeros((3, N)), cvolume=np.zeros(N), cgrad=np.zeros((3, N)), cdisp=np.zeros((3, N)))
            for f in range(nf):
                t = dict(numberPoints=np.zeros(n, np.int32), mass=np.zeros(n), pk=np.zeros((3, n)), ftot=np.zeros((3, n)),
                         vk0=np.zeros((3, n)), pkcopy=np.zeros((3, n)), cvolume=np.zeros(n), cgrad=np.zeros((3, n)), cdisp=np.zeros((3, n)))
                self.lib.ref_get_nodes_field(f, _ip(t["numberPoints"]), _dp(t["mass"]), _dp(t["pk"]), _dp(t["ftot"]), _dp(t["vk0"]),
                                             _dp(t["pkcopy"]), _dp(t["cvolume"]), _dp(t["cgrad"]), _dp(t["cdisp"]))
                for k, v in t.items():
                    o[k][..., f * n:(f + 1) * n] = v
            fixed = np.zeros(n, np.int32)
            dummy = [np.zeros(n, np.int32), np.zeros(n)] + [np.zeros((3, n)) for _ in range(4)]
            self.lib.ref_get_nodes(_ip(dummy[0]), _dp(dummy[1]), _dp(dummy[2]), _dp(dummy[3]), _dp(dummy[4]), _dp(dummy[5]), _ip(fixed))
            o["fixedDirection"] = fixed
            return o
        n = self.info["nnodes"]
        o = dict(numberPoints=np.zeros(n, np.int32), mass=np.zeros(n), pk=np.zeros((3, n)),
                 ftot=np.zeros((3, n)), vk0=np.zeros((3, n)), pkcopy=np.zeros((3, n)),
                 fixedDirection=np.zeros(n, np.int32))
        self.lib.ref_get_nodes(_ip(o["numberPoints"]), _dp(o["mass"]), _dp(o["pk"]), _dp(o["ftot"]),
                               _dp(o["vk0"]), _dp(o["pkcopy"]), _ip(o["fixedDirection"]))
        return o

    def node_coords(self):
        xyz = np.zeros((self.info["nnodes"], 3))
        self.lib.ref_get_node_coords(_dp(xyz))
        return xyz

    def element_extents(self):
        ext = np.zeros((self.info["nelems"], 6))
        self.lib.ref_get_element_extents(_dp(ext))
        return ext

    def velbcs(self):
        n = self.lib.ref_num_velbcs()
        o = dict(node=np.zeros(n, np.int32), dir=np.zeros(n, np.int32), style=np.zeros(n, np.int32),
                 norm=np.zeros((n, 3)), value=np.zeros(n), ftime=np.zeros(n), offset=np.zeros(n),
                 currentValue=np.zeros(n), reflected=np.full(n, -1, np.int32), ratio=np.ones(n), id=np.zeros(n, np.int32))
        if n:
            self.lib.ref_get_velbcs(_ip(o["node"]), _ip(o["dir"]), _ip(o["style"]), _dp(o["norm"]),
                                    _dp(o["value"]), _dp(o["ftime"]), _dp(o["offset"]), _dp(o["currentValue"]))
            self.lib.ref_get_velbc_reflections(_ip(o["reflected"]), _dp(o["ratio"]))
            self.lib.ref_get_velbc_ids(_ip(o["id"]))
        return o

    def tractions(self):
        """MatPtTractionBC list: 0-based particle, face, direction, style, BCValue at the current time (None when empty)."""
        n = self.lib.ref_num_tractions()
        if n == 0:
            return None
        o = dict(particle=np.zeros(n, np.int32), face=np.zeros(n, np.int32), direction=np.zeros(n, np.int32), style=np.zeros(n, np.int32), value=np.zeros(n))
        self.lib.ref_get_tractions(_ip(o["particle"]), _ip(o["face"]), _ip(o["direction"]), _ip(o["style"]), _dp(o["value"]))
        o["particle"] -= 1
        return o

    def heat_fluxes(self):
        """MatPtHeatFluxBC list: 0-based particle, face, direction (1 external), style, BCValue at the current time (None when empty)."""
        n = self.lib.ref_num_heat_fluxes()
        if n == 0:
            return None
        o = dict(particle=np.zeros(n, np.int32), face=np.zeros(n, np.int32), direction=np.zeros(n, np.int32), style=np.zeros(n, np.int32), value=np.zeros(n))
        self.lib.ref_get_heat_fluxes(_ip(o["particle"]), _ip(o["face"]), _ip(o["direction"]), _ip(o["style"]), _dp(o["value"]))
        o["particle"] -= 1
        return o

    def reactions(self, ids):
        """NodalVelBC::TotalReactionForce for each BC id (0 = every BC; rigid-particle BCs carry their material number)."""
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros((len(ids), 3))
        self.lib.ref_reaction_forces(len(ids), _ip(ids), _dp(out))
        return out

    def materials(self):
        nm = self.info["nmat"]
        ids = np.zeros(nm, np.int32)
        par = np.zeros((nm, 32))
        self.lib.ref_get_materials(_ip(ids), _dp(par))
        return ids, par

    def close(self):
        self.lib.ref_close()


def _flatten(prefix, d, out):
    for k, v in d.items():
        out["%s/%s" % (prefix, k)] = np.asarray(v)


from nairn_mpm_fea_b200.problem import jitter  # noqa: E402  (test infrastructure may import the product)


def _worker(xml, out_npz, nprocs, snaps, per_task_steps, jitter_amp=0.0, vel_amp=0.0):
    """snaps: sorted step counts at which to snapshot particles (+nodes); per_task_steps: number of
    initial steps run task-by-task with node+particle dumps after every task.  jitter_amp/vel_amp:
    hash-jitter the initial positions (length units) and velocities before the first step."""
    r = RefRun(xml, nprocs)
    if jitter_amp > 0.0 or vel_amp > 0.0:
        p0 = r.particles()
        newpos = jitter(p0["pos"], jitter_amp, 12345) if jitter_amp > 0.0 else None
        newvel = jitter(p0["vel"], vel_amp, 777) if vel_amp > 0.0 else None
        bad = r.set_particles(newpos, newvel)
        assert bad == 0, "jitter pushed %d particles off the grid" % bad
    out = {}
    _flatten("info", r.info, out)
    out["node_coords"] = r.node_coords()
    out["element_extents"] = r.element_extents()
    vb = r.velbcs()
    _flatten("velbcs", vb, out)
    react_ids = np.array(sorted(set([0] + [int(i) for i in vb["id"]] + list(range(1, r.info["nmat"] + 1)))), np.int32)
    out["reaction_ids"] = react_ids
    ids, par = r.materials()
    out["mat_ids"], out["mat_params"] = ids, par
    mm = r.multimaterial()
    if mm is not None:
        _flatten("mm", mm, out)
    cond = r.conduction()
    if cond is not None:
        _flatten("conduction", cond, out)
    trac = r.tractions()
    if trac is not None:
        _flatten("tractions", trac, out)
    hflux = r.heat_fluxes()
    if hflux is not None:
        _flatten("heatflux", hflux, out)
    names = r.task_names()
    out["task_names"] = np.array(names)
    _flatten("p0", r.particles(), out)
    done = 0
    xpic = []          # (order, usingFMPM) in force DURING each step (PeriodicXPIC changes it between steps)

    def note_xpic():
        inf = r.get_info()
        xpic.append((inf["XPICOrder"], inf["usingFMPM"]))

    for s in range(per_task_steps):
        note_xpic()
        for i, nm in enumerate(names):
            r.run_task(i)
            _flatten("s%d/t%d/nodes" % (s + 1, i), r.nodes(), out)
            _flatten("s%d/t%d/p" % (s + 1, i), r.particles(), out)
        done += 1
        out["reaction%d" % done] = r.reactions(react_ids)
        if done in snaps:
            _flatten("p%d" % done, r.particles(), out)
            _flatten("n%d" % done, r.nodes(), out)
    for s in snaps:
        if s > done:
            while done < s:
                note_xpic()
                r.step(1)
                done += 1
            _flatten("p%d" % done, r.particles(), out)
            _flatten("n%d" % done, r.nodes(), out)
            out["reaction%d" % done] = r.reactions(react_ids)
    out["xpic_by_step"] = np.array(xpic, dtype=np.int32).reshape(-1, 2)
    _flatten("info_end", r.get_info(), out)
    r.close()
    np.savez_compressed(out_npz, **out)


def run_reference(xml_text_or_path, snaps=(1,), per_task_steps=0, nprocs=1, workdir=None, jitter_amp=0.0, vel_amp=0.0):
    """Run the reference on an XML input in a fresh process; returns dict of arrays (see _worker)."""
    if not available():
        raise RuntimeError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    tmp = workdir or tempfile.mkdtemp(prefix="mpmref_")
    if os.path.exists(xml_text_or_path):
        xml = os.path.abspath(xml_text_or_path)
    else:
        xml = os.path.join(tmp, "input.fmcmd")
        with open(xml, "w") as f:
            f.write(xml_text_or_path)
    out = os.path.join(tmp, "ref_out.npz")
    cmd = [sys.executable, os.path.abspath(__file__), xml, out, str(nprocs),
           ",".join(str(s) for s in sorted(snaps)), str(per_task_steps), repr(jitter_amp), repr(vel_amp)]
    p = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("reference worker failed (rc %d):\n%s\n%s" % (p.returncode, p.stdout[-2000:], p.stderr[-2000:]))
    with np.load(out) as z:
        return {k: z[k] for k in z.files}


if __name__ == "__main__":
    _xml, _out, _np, _snaps, _pt = sys.argv[1:6]
    _ja = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    _va = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0
    _worker(_xml, _out, int(_np), [int(s) for s in _snaps.split(",") if s], int(_pt), _ja, _va)
