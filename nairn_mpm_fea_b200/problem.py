"""Host-side problem description: what the reference driver holds after ReadFile + CMPreparations
(grid, particle lattice, materials, BC list, time step) in the flat form libmpmgpu takes.

Generators here reproduce the reference's own set-up arithmetic so the same XML gives the same
numbers:  grid + border cells  Read_MPM/Generators.cpp:1760-1935;  particle lattice
Elements/MoreMPMElementBase.cpp:348-441 (MPMPoints) in element-major order Generators.cpp:1486-1588;
mass and time step  NairnMPM_Class/NairnMPM.cpp:589-730, :1207-1240.
"""
import numpy as np

from . import materials as M

USF, USAVG, USL = 0, 2, 3
POINT_GIMP, UNIFORM_GIMP, LINEAR_CPDI, QUADRATIC_CPDI = 0, 1, 10, 11
X_DIRECTION, Y_DIRECTION, Z_DIRECTION = 1, 2, 4


class Problem:
    """Plain data; see capi.MpmGpu for how it is handed to the library."""

    def __init__(self):
        self.np = M.THREED_MPM
        self.horiz = self.vert = self.depth = 0
        self.xpts = self.ypts = self.zpts = None
        self.grid = (0.0, 0.0, 0.0)
        self.thickness = 1.0
        self.shape = UNIFORM_GIMP
        self.rcrit = -1.0
        self.method = USAVG
        self.skip_post_extrapolation = False
        self.fraction_usf = 0.5
        self.xpic_order = 0
        self.using_fmpm = False
        self.grid_damping = 0.0
        self.particle_damping = 0.0
        self.gravity = (0.0, 0.0, 0.0)
        self.materials = []
        self.dt = self.dt_strain_first = self.dt_strain_last = 0.0
        self.maxtime = 0.0
        self.bc_node = np.zeros(0, np.int32)
        self.bc_norm = np.zeros((0, 3))
        self.bc_value = np.zeros(0)
        self.bc_active = np.zeros(0, np.int32)
        self.bc_symdir = np.zeros(0, np.int32)
        self.bc_reflected = None        # per BC: 1-based node across a symmetry plane (NodalVelBC::reflectedNode) or -1; None = no such BCs
        self.bc_ratio = None            # per BC: NodalVelBC::reflectRatio
        # <MultiMaterialMode>: None, or dict(n_fields, field_of_material [nmat], normal_method, by_displacements, position_cutoff,
        # contact_normal [3], law_kind / law_friction / law_static [n_fields][n_fields]) -- mpmgpu_multimaterial in include/mpmgpu.h
        self.multimaterial = None
        self.origpos = None             # [3][n] MPMBase::origpos when it differs from the uploaded positions
        # <Thermal><Conduction/>: None, or dict(kcond [nmat] = conductivity / rho); particles["temperature"] = pTemperature
        self.conduction = None
        self.adiabatic = False          # <EnergyCoupling/> (ConductionTask::adiabatic)
        self.particles = {}

    @property
    def is3d(self):
        return self.np == M.THREED_MPM

    @property
    def nnodes(self):
        n = (self.horiz + 1) * (self.vert + 1)
        return n * (self.depth + 1) if self.is3d else n

    @property
    def nelems(self):
        return self.horiz * self.vert * (self.depth if self.is3d else 1)

    @property
    def nparticles(self):
        return int(np.asarray(self.particles["mp"]).shape[0])

    def set_time_step(self, dt):
        """CFLTimeStep tail (NairnMPM.cpp:1228-1236)."""
        self.dt = dt
        if self.method == USAVG:
            self.dt_strain_first = self.fraction_usf * dt
            self.dt_strain_last = dt - self.dt_strain_first
        else:
            self.dt_strain_first = self.dt_strain_last = dt


def jitter(arr, amplitude, seed=12345, id_offset=0):
    """Deterministic integer-hash jitter (same on every platform):
    arr[c][p] += amplitude * (hash(p, c) / 2^32 - 0.5)."""
    arr = np.array(arr, dtype=np.float64, copy=True)
    n = arr.shape[1]
    idx = np.arange(n, dtype=np.uint64) + np.uint64(id_offset)
    for c in range(arr.shape[0]):
        h = (idx * np.uint64(2654435761) + np.uint64(seed + 7919 * c)) & np.uint64(0xFFFFFFFF)
        h ^= h >> np.uint64(16)
        h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
        h ^= h >> np.uint64(13)
        h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
        h ^= h >> np.uint64(16)
        arr[c] += amplitude * (h.astype(np.float64) / 4294967296.0 - 0.5)
    return arr


def structured_axis(lo, hi, cell):
    """One axis of <Grid>: returns (ncells incl. border, node coordinates) as Generators.cpp:1769-1833."""
    n = int((hi - lo) / cell + 0.5)          # Nhoriz from cellsize (MPMReadHandler Horiz cellsize)
    c = (hi - lo) / float(n)
    n += 2
    lo2 = lo - c
    hi2 = hi + c
    delta = (hi2 - lo2) / float(n)
    pts = np.array([lo2 + float(i) * delta for i in range(n + 1)])
    return n, pts, delta


def lattice_points(xpts, ypts, zpts, elems_ijk, pts_per_side=2):
    """Particle positions of ElementBase::MPMPoints for the listed elements, in the reference's
    order (z slowest, x fastest within an element) -- positions come from the element's linear
    shape functions applied to its node coordinates, summed in node order."""
    is3d = zpts is not None
    n = pts_per_side
    gap = 1.0 / float(n)
    xi1 = np.array([-1.0 + (2.0 * j + 1.0) * gap for j in range(n)])
    if is3d:
        zz, yy, xx = np.meshgrid(xi1, xi1, xi1, indexing="ij")
        nat = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)          # (n^3, 3)
        sx = np.array([-1., 1., 1., -1., -1., 1., 1., -1.])
        sy = np.array([-1., -1., 1., 1., -1., -1., 1., 1.])
        sz = np.array([-1., -1., -1., -1., 1., 1., 1., 1.])
    else:
        yy, xx = np.meshgrid(xi1, xi1, indexing="ij")
        nat = np.stack([xx.ravel(), yy.ravel(), 0.0 * xx.ravel()], axis=1)
        sx = np.array([-1., 1., 1., -1.])
        sy = np.array([-1., -1., 1., 1.])
        sz = None
    ei, ej, ek = elems_ijk
    ne = len(ei)
    npp = nat.shape[0]
    pos = np.zeros((3, ne, npp))
    for a in range(len(sx)):
        if is3d:
            fxn = 0.125 * (1.0 + sx[a] * nat[:, 0]) * (1.0 + sy[a] * nat[:, 1]) * (1.0 + sz[a] * nat[:, 2])
            nz = zpts[ek + (1 if sz[a] > 0 else 0)]
        else:
            fxn = 0.25 * (1.0 + sx[a] * nat[:, 0]) * (1.0 + sy[a] * nat[:, 1])
        nx = xpts[ei + (1 if sx[a] > 0 else 0)]
        ny = ypts[ej + (1 if sy[a] > 0 else 0)]
        pos[0] += nx[:, None] * fxn[None, :]
        pos[1] += ny[:, None] * fxn[None, :]
        if is3d:
            pos[2] += nz[:, None] * fxn[None, :]
    return pos.reshape(3, ne * npp), gap


def block3d(ncell=50, margin=7, cell=1.0, E=1000.0, nu=0.3, rho=1.0, velocity=(0.0, 0.0, -1000.0), cfl=0.4,
            step_ms=1e-3, method=USAVG, shape=UNIFORM_GIMP, bottom_bc=True, gravity=None, velocity_fn=None,
            pts_per_side=2, ncell_xyz=None, jitter_amp=0.0, cells_z=None, material=None, rigid_wall=None):
    """BASELINE.json config 2 family: block of ncell^3 cells (pts_per_side^3 particles per cell) of
    IsotropicMat inside a (ncell+2*margin)^3-cell grid (+1 border cell per side), initial velocity,
    bottom plane z<=margin held in z.  Numbers are XML (Legacy) units: mm, MPa, g/cm^3, mm/s, ms.
    Same problem as tests/inputs.py::block3d(ncell, margin) fed to the reference.
    cells_z=(lo, hi): generate only the particles of block cell planes [lo, hi) (one slab of a
    multi-GPU run); the grid, BCs and time step are still those of the whole problem and the dict gets
    'ids' = the particles' indices in the whole problem.
    material: a materials.py block (internal units) replacing the IsotropicMat.
    rigid_wall = dict(set_direction=4, velocity=(0, 0, 0), overhang=1): BASELINE config 4 family (Taylor bar
    on a rigid wall) -- a one-cell-thick plate of rigid-BC particles (RigidMaterial, rho 1 so mp = volume)
    under the block, `overhang` cells wider on each side; they follow the nonrigid particles as in the
    reference's ordering (NairnMPM.cpp:1121-1190) and every slab of a multi-GPU run gets all of them."""
    pr = Problem()
    pr.np = M.THREED_MPM
    pr.method = method
    pr.shape = shape
    ncx, ncy, ncz = ncell_xyz or (ncell, ncell, ncell)
    ext = [n + 2 * margin for n in (ncx, ncy, ncz)]
    pr.horiz, pr.xpts, gx = structured_axis(0.0, ext[0] * cell, cell)
    pr.vert, pr.ypts, gy = structured_axis(0.0, ext[1] * cell, cell)
    pr.depth, pr.zpts, gz = structured_axis(0.0, ext[2] * cell, cell)
    pr.grid = (gx, gy, gz)
    u = M.xml_units(E=E, rho=rho)
    mat = material or M.isotropic(u["E"], nu, u["rho"], 0.0, M.DEFAULT_CV, M.THREED_MPM)
    pr.materials = [mat]
    # filled elements: cells [margin, margin+nc) of the user grid = element index +1 (border) per axis
    ii = np.arange(margin + 1, margin + 1 + ncx)
    jj = np.arange(margin + 1, margin + 1 + ncy)
    kz0, kz1 = cells_z if cells_z is not None else (0, ncz)
    kk = np.arange(margin + 1 + kz0, margin + 1 + kz1)
    id_offset = kz0 * ncx * ncy * pts_per_side ** 3
    K, J, I = np.meshgrid(kk, jj, ii, indexing="ij")
    ei, ej, ek = I.ravel(), J.ravel(), K.ravel()
    pos, gap = lattice_points(pr.xpts, pr.ypts, pr.zpts, (ei, ej, ek), pts_per_side)
    npp = pts_per_side ** 3
    n = pos.shape[1]
    elem = (pr.horiz * (ek * pr.vert + ej) + ei + 1).astype(np.int32)
    in_elem = np.repeat(elem, npp)
    lp = np.full((3, n), gap)
    if jitter_amp > 0.0:
        # off-lattice start (so the uGIMP stencils are the generic 27-node ones); elements re-found as
        # MeshInfo::FindElementFromPoint does (MeshInfo.cpp:593-633)
        pos = jitter(pos, jitter_amp * cell, 12345, id_offset)
        col = ((pos[0] - pr.xpts[0]) / gx).astype(np.int64)
        row = ((pos[1] - pr.ypts[0]) / gy).astype(np.int64)
        zrow = ((pos[2] - pr.zpts[0]) / gz).astype(np.int64)
        in_elem = (pr.horiz * (zrow * pr.vert + row) + col + 1).astype(np.int32)
    # mp = rho * 8 * psize.x*psize.y*psize.z, psize = 0.5*lp*cell extent (NairnMPM.cpp:624-637, MatPoint3D.cpp:219-225)
    dxe = (pr.xpts[ei + 1] - pr.xpts[ei])
    dye = (pr.ypts[ej + 1] - pr.ypts[ej])
    dze = (pr.zpts[ek + 1] - pr.zpts[ek])
    psx, psy, psz = dxe * (0.5 * gap), dye * (0.5 * gap), dze * (0.5 * gap)
    mp = np.repeat(mat["rho"] * (8.0 * psx * psy * psz), npp)
    vel = np.zeros((3, n))
    if velocity_fn is not None:
        vel[:] = velocity_fn(pos)
    else:
        for c in range(3):
            vel[c] = velocity[c]
    energies = np.zeros((6, n))
    energies[5] = 1.0            # pPreviousTemperature: thermal.reference default... set by caller if needed
    pr.particles = dict(pos=pos, vel=vel, mp=mp, lp=lp, in_elem=in_elem, matnum=np.ones(n, np.int32),
                        n_nonrigid=n, energies=energies)
    if cells_z is not None:
        pr.particles["ids"] = (np.arange(n, dtype=np.int64) + id_offset).astype(np.int32)
    if mat.get("init_history") is not None:
        hist = np.zeros((M.MAX_HISTORY, n))
        for i, v in enumerate(mat["init_history"]):
            hist[i] = v
        pr.particles["history"] = hist
    if mat.get("init_eplast") is not None:
        pr.particles["eplast"] = np.tile(np.asarray(mat["init_eplast"], float)[:, None], (1, n))
    if rigid_wall is not None:
        oh = int(rigid_wall.get("overhang", 1))
        pr.materials.append(M.rigid_bc(int(rigid_wall.get("set_direction", 4))))
        ri = np.arange(margin + 1 - oh, margin + 1 + ncx + oh)
        rj = np.arange(margin + 1 - oh, margin + 1 + ncy + oh)
        RK, RJ, RI = np.meshgrid(np.array([margin]), rj, ri, indexing="ij")
        ri, rj, rk = RI.ravel(), RJ.ravel(), RK.ravel()
        rpos, _ = lattice_points(pr.xpts, pr.ypts, pr.zpts, (ri, rj, rk), pts_per_side)
        nr = rpos.shape[1]
        relem = np.repeat((pr.horiz * (rk * pr.vert + rj) + ri + 1).astype(np.int32), npp)
        rmp = np.repeat(8.0 * ((pr.xpts[ri + 1] - pr.xpts[ri]) * (0.5 * gap)) * ((pr.ypts[rj + 1] - pr.ypts[rj]) * (0.5 * gap))
                        * ((pr.zpts[rk + 1] - pr.zpts[rk]) * (0.5 * gap)), npp)
        rvel = np.tile(np.asarray(rigid_wall.get("velocity", (0.0, 0.0, 0.0)), float)[:, None], (1, nr))
        n_total_nr = ncx * ncy * ncz * npp
        add = dict(pos=rpos, vel=rvel, mp=rmp, lp=np.full((3, nr), gap), in_elem=relem, matnum=np.full(nr, 2, np.int32),
                   energies=np.concatenate([np.zeros((5, nr)), np.ones((1, nr))]))
        if "ids" in pr.particles:
            add["ids"] = (n_total_nr + np.arange(nr)).astype(np.int32)
        if "history" in pr.particles:
            add["history"] = np.zeros((M.MAX_HISTORY, nr))
        if "eplast" in pr.particles:
            add["eplast"] = np.zeros((6, nr))
        for key, v in add.items():
            pr.particles[key] = np.concatenate([pr.particles[key], v], axis=-1)
        pr.n_rigid = nr
    # time step (NairnMPM.cpp:695-699, :1207-1227): dcell = grid.x (cubic grid, MeshInfo.cpp:1555-1563)
    dt_cfl = cfl * (gx / mat["wave_speed"])
    pr.set_time_step(min(step_ms * 1.0e-3, dt_cfl))
    if bottom_bc:
        # BCBox zmax = margin + 0.01: every node with z <= that, in node order, dir 3 (z), value 0
        zsel = np.nonzero(pr.zpts <= margin * cell + 0.01)[0]
        nx1, ny1 = pr.horiz + 1, pr.vert + 1
        nodes = (zsel[:, None] * (nx1 * ny1) + np.arange(nx1 * ny1)[None, :] + 1).ravel()
        nb = len(nodes)
        pr.bc_node = nodes.astype(np.int32)
        pr.bc_norm = np.tile(np.array([0.0, 0.0, 1.0]), (nb, 1))
        pr.bc_value = np.zeros(nb)
        pr.bc_active = np.ones(nb, np.int32)
        pr.bc_symdir = np.zeros(nb, np.int32)
    if gravity is not None:
        pr.gravity = tuple(gravity)
    return pr


def from_reference_dump(z, snapshot="p0"):
    """Problem from an oracle/refharness.py dump: the state the reference itself holds after
    ReadFile + CMPreparations.  Used by the parity tests (same inputs on both sides)."""
    info = {k[5:]: z[k].item() for k in z if k.startswith("info/")}
    pr = Problem()
    pr.np = int(info["np"])
    pr.horiz, pr.vert, pr.depth = int(info["horiz"]), int(info["vert"]), int(info["depth"])
    xyz = z["node_coords"]
    nx1, ny1 = pr.horiz + 1, pr.vert + 1
    pr.xpts = xyz[:nx1, 0].copy()
    pr.ypts = xyz[0:nx1 * ny1:nx1, 1].copy()
    if pr.is3d:
        pr.zpts = xyz[0::nx1 * ny1, 2].copy()
    else:
        pr.depth = 0
    pr.grid = (info["gridx"], info["gridy"], info["gridz"])
    pr.thickness = info["thickness"]
    pr.shape = int(info["useGimp"])
    pr.rcrit = info["rcrit"]
    pr.method = int(info["mpmApproach"])
    pr.skip_post_extrapolation = bool(info["skipPostExtrapolation"])
    pr.fraction_usf = info["fractionUSF"]
    pr.xpic_order = int(info["XPICOrder"])
    pr.using_fmpm = bool(info["usingFMPM"])
    pr.grid_damping = info["damping"]
    pr.particle_damping = info["pdamping"]
    pr.gravity = (info["gx"], info["gy"], info["gz"]) if info["hasGravity"] else (0.0, 0.0, 0.0)
    pr.dt = info["timestep"]
    pr.dt_strain_first = info["strainTimestepFirst"]
    pr.dt_strain_last = info["strainTimestepLast"]
    pr.maxtime = info["maxtime"]
    pr.materials = []
    for mid, q in zip(z["mat_ids"], z["mat_params"]):
        pd = None if q[3] < 0 else q[3]
        av = (q[6], q[7]) if q[5] else None
        if av is not None and mid not in (M.NEOHOOKEAN, M.ISOPLASTICITY, M.MOONEY):
            raise NotImplementedError("artificial viscosity on material id %d" % mid)
        if mid == M.ISOTROPIC:
            m = M.isotropic(q[8], q[9], q[0], q[11] * 1.0e6, q[1], pr.np, pd, large_rotation=bool(q[13]))
        elif mid == M.NEOHOOKEAN:
            m = M.neohookean(q[8], q[9], q[0], q[15] * 1.0e6, q[1], int(q[14]), pd, av)
        elif mid == M.MOONEY:
            if q[17] != 0:
                raise NotImplementedError("Mooney with the IdealRubber option")
            m = M.mooney(q[8], q[9], q[10], q[0], q[15] * 1.0e6, q[1], int(q[14]), pd, av)
        elif mid == M.ISOPLASTICITY:
            law = int(q[23]) if len(q) > 23 else 1
            lr = bool(q[22]) if len(q) > 22 else False
            if law in (0, M.HARD_LINEAR):
                m = M.isoplasticity(q[8], q[9], q[0], q[15], q[16] if q[16] >= 0 else None, q[21], q[11] * 1.0e6, q[1], pr.np, pd,
                                    q[20] * q[0], av, large_rotation=lr)
            elif law in (M.HARD_NONLINEAR, M.HARD_NONLINEAR2):
                m = M.isoplasticity(q[8], q[9], q[0], q[15], None, 0.0, q[11] * 1.0e6, q[1], pr.np, pd, q[20] * q[0], av, large_rotation=lr,
                                    hardening=("nonlinear" if law == M.HARD_NONLINEAR else "nonlinear2", q[24], q[25]))
            elif law == M.HARD_JOHNSONCOOK:
                m = M.isoplasticity(q[8], q[9], q[0], q[15], None, 0.0, q[11] * 1.0e6, q[1], pr.np, pd, q[20] * q[0], av, large_rotation=lr,
                                    hardening=("johnsoncook", dict(B=q[24] * q[0], n=q[25], C=q[26], ep0=q[27], D=q[28], n2=q[29], Tm=q[30],
                                                                   m=q[31], Tref=q[16])))
            elif law == M.HARD_SCGL:
                m = M.isoplasticity(q[8], q[9], q[0], q[15], None, 0.0, q[11] * 1.0e6, q[1], pr.np, pd, q[20] * q[0], av, large_rotation=lr,
                                    hardening=("scgl", dict(beta=q[24], n=q[25], yld_max=q[26] * q[0], GPp=q[27] / q[0], GTp=q[28], Tref=q[16])))
            else:
                raise NotImplementedError("hardening law %d" % law)
        elif mid in (M.CONTACT_LAW, M.COULOMB_FRICTION_LAW) and "mm/nfields" in z:
            m = M.contact_law_placeholder()
        elif mid == M.RIGIDBC and int(q[8]) == 8:
            if q[10] != 0 or q[11] != 0:
                raise NotImplementedError("rigid contact material with setting functions (host-evaluated: update_rigid_velocities) / temperature or concentration")
            m = M.rigid_contact()
        elif mid == M.RIGIDBC:
            if q[10] != 0 or int(q[11]) & 2 or (len(q) > 12 and q[12] != 0):
                raise NotImplementedError("rigid material with setting functions (host-evaluated: update_rigid_velocities) / concentration")
            m = M.rigid_bc(int(q[8]), int(q[9]), sets_temperature=bool(int(q[11]) & 1))
        else:
            raise NotImplementedError("material id %d" % mid)
        pr.materials.append(m)
    s = snapshot
    n = z[s + "/mp"].shape[0]
    pr.particles = dict(pos=z[s + "/pos"], vel=z[s + "/vel"], mp=z[s + "/mp"], lp=z[s + "/lp"],
                        in_elem=z[s + "/inElem"], matnum=z[s + "/matnum"], sp=z[s + "/sp"],
                        pressure=z[s + "/pressure"], ep=z[s + "/ep"], wrot=z[s + "/wrot"], eplast=z[s + "/eplast"],
                        energies=z[s + "/energies"], history=z[s + "/hist"], crossings=z[s + "/crossings"],
                        n_nonrigid=int(info["nmpmsNR"]))
    if np.any(z[s + "/pFext"] != 0.0):
        pr.particles["pfext"] = z[s + "/pFext"]
    pr.adiabatic = bool(info.get("adiabatic", 0))
    if "conduction/kcond" in z:
        if int(z["conduction/contact_heating"]):
            raise NotImplementedError("conduction with contact_heating")
        if int(z["conduction/n_flux_bcs"]) and "heatflux/particle" not in z:
            raise NotImplementedError("conduction with heat-flux BCs the dump does not list")
        pr.conduction = dict(kcond=np.asarray(z["conduction/kcond"], float))
        if "conduction/tbc_node" in z:      # nodal temperature BCs (constant values in the goldens; a host re-evaluates others every step)
            pr.conduction["tbc_node"] = np.asarray(z["conduction/tbc_node"], np.int32)
            pr.conduction["tbc_value"] = np.asarray(z["conduction/tbc_value"], float)
        pr.particles["temperature"] = np.asarray(z[s + "/temperature"], float)
    elif (s + "/temperature") in z and np.any(np.asarray(z[s + "/temperature"]) != np.asarray(z[s + "/energies"])[5]):
        # no transport task, but particles that start off the temperature of their last strain update: the first particle
        # update hands the laws that difference (UpdateParticlesTask.cpp:246-251)
        pr.particles["temperature"] = np.asarray(z[s + "/temperature"], float)
    if "mm/nfields" in z:
        # multimaterial mode: the reference's table is by material pair; the device wants it by velocity-field pair
        nf = int(z["mm/nfields"])
        field = np.asarray(z["mm/field"], np.int32)
        law = np.asarray(z["mm/law"])
        kind = np.zeros((nf, nf), np.int32)
        fric = np.zeros((nf, nf))
        stat = np.full((nf, nf), -1.0)
        for i, fi in enumerate(field):
            for j, fj in enumerate(field):
                if fi < 0 or fj < 0 or fi == fj:
                    continue
                if law[i, j, 0] < 0:
                    raise NotImplementedError("contact law between materials %d and %d (imperfect interface, adhesion, ...)" % (i + 1, j + 1))
                kind[fi, fj], fric[fi, fj], stat[fi, fj] = int(law[i, j, 0]), law[i, j, 1], law[i, j, 2]
        if int(z["mm/normal_method"]) > 4:
            raise NotImplementedError("contact normals by linear / logistic regression")
        pr.multimaterial = dict(n_fields=nf, field_of_material=np.where(field < 0, 0, field).astype(np.int32),
                                normal_method=int(z["mm/normal_method"]), by_displacements=int(z["mm/by_displacements"]),
                                position_cutoff=float(z["mm/position_cutoff"]), contact_normal=np.asarray(z["mm/contact_normal"], float),
                                law_kind=kind, law_friction=fric, law_static=stat,
                                rigid_gradient_bias=float(z["mm/rigid_gradient_bias"]) if "mm/rigid_gradient_bias" in z else 1.0)
        pr.origpos = np.asarray(z[s + "/origpos"], float) if (s + "/origpos") in z else None
    nb = z["velbcs/node"].shape[0]
    pr.bc_node = z["velbcs/node"].astype(np.int32)
    pr.bc_norm = z["velbcs/norm"]
    pr.bc_value = z["velbcs/value"].copy()            # constant-style BCs: value; others evaluated by the host
    pr.bc_active = np.ones(nb, np.int32)
    pr.bc_symdir = np.zeros(nb, np.int32)
    pr.bc_style = z["velbcs/style"]
    if "heatflux/particle" in z:        # MatPtHeatFluxBC list (external fluxes of constant style)
        if np.any(z["heatflux/direction"] != 1) or np.any(z["heatflux/style"] == 5):
            raise NotImplementedError("coupled or silent heat-flux BCs")
        pr.heat_fluxes = {k: np.asarray(z["heatflux/" + k]) for k in ("particle", "face", "value")}
    if "tractions/particle" in z:       # MatPtTractionBC list (constant-style values; others are the host's to re-evaluate)
        pr.tractions = {k: np.asarray(z["tractions/" + k]) for k in ("particle", "face", "direction", "value")}
    pr.bc_id = z["velbcs/id"].astype(np.int32) if "velbcs/id" in z else np.zeros(nb, np.int32)        # BoundaryCondition::bcID
    pr.bc_ftime = z["velbcs/ftime"]
    if "velbcs/reflected" in z and np.any(z["velbcs/reflected"] > 0):
        # symmetry planes (<Horiz symmin=...>): the plane's nodes carry the symmetry bits of NodalPoint::fixedDirection
        # (32|64|128, ADJUST_COPIED_PK), their outer neighbours reflect the inner ones (Generators.cpp:2150-2260)
        pr.bc_reflected = z["velbcs/reflected"].astype(np.int32)
        pr.bc_ratio = z["velbcs/ratio"].astype(np.float64)
        fixed = z["n1/fixedDirection"] if "n1/fixedDirection" in z else z["s1/t0/nodes/fixedDirection"]
        pr.bc_symdir = (fixed[pr.bc_node - 1] & (32 | 64 | 128)).astype(np.int32)
    return pr
