"""ctypes binding of libmpmgpu (include/mpmgpu.h) -- the call a Python user makes.

`MpmGpu` mirrors the life cycle of the reference driver around its MPMTask list
(NairnMPM_Class/NairnMPM.cpp:135-200): create from a `Problem`, upload, step (whole steps or task by
task), download.  There is no CPU path: if the library or a CUDA device is missing this raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmpmgpu.so")

ABI_VERSION = 1
MAT_NPARAMS = 32
MAX_HISTORY = 4

F_POS, F_VEL, F_STRESS, F_STRAIN, F_EPLAST, F_ENERGY, F_HISTORY, F_ELEM, F_ACC = (
    0x001, 0x002, 0x004, 0x008, 0x010, 0x020, 0x040, 0x080, 0x100)
F_TEMPERATURE = 0x200
F_ALL = 0x1FF

TASKS = ["initialization", "mass_and_momentum", "post_extrapolation", "update_strains_first", "grid_forces",
         "post_forces", "update_momenta", "update_particles", "update_strains_last", "reset_elements"]

# every symbol include/mpmgpu.h declares (tests check the library exports all of them)
EXPORTS = ["mpmgpu_abi_version", "mpmgpu_create", "mpmgpu_destroy", "mpmgpu_last_error", "mpmgpu_set_materials",
           "mpmgpu_set_multimaterial", "mpmgpu_set_conduction", "mpmgpu_set_energy_coupling", "mpmgpu_set_temperature_bcs", "mpmgpu_upload_particles",
           "mpmgpu_track_reactions", "mpmgpu_download_reactions", "mpmgpu_contact_forces", "mpmgpu_update_rigid_temperatures", "mpmgpu_set_particle_tractions", "mpmgpu_update_particle_traction_values",
           "mpmgpu_set_particle_heat_fluxes", "mpmgpu_update_particle_heat_flux_values", "mpmgpu_set_time_step", "mpmgpu_set_xpic", "mpmgpu_set_velocity_bcs",
           "mpmgpu_update_velocity_bc_values", "mpmgpu_set_velocity_bc_reflections", "mpmgpu_update_particle_loads", "mpmgpu_update_rigid_velocities", "mpmgpu_step", "mpmgpu_set_poll_interval"] + ["mpmgpu_task_" + t for t in TASKS] + [
    "mpmgpu_task_project_rigid_bcs",
    "mpmgpu_download_particles", "mpmgpu_download_nodes", "mpmgpu_synchronize", "mpmgpu_get_status",
    "mpmgpu_launch_count", "mpmgpu_stream", "mpmgpu_set_profiling", "mpmgpu_task_times",
    "mpmgpu_slab_configure", "mpmgpu_slab_halo_buffers", "mpmgpu_slab_step_phase", "mpmgpu_slab_set_halo_callback", "mpmgpu_slab_migration_counts",
    "mpmgpu_slab_migration_buffers", "mpmgpu_slab_pack_migrants", "mpmgpu_slab_finish_migration",
    "mpmgpu_num_particles", "mpmgpu_set_stream",
    "mpmgpu_left_grid_counts",
    "mpmgpu_nccl_unique_ids", "mpmgpu_slab_connect", "mpmgpu_slab_step", "mpmgpu_slab_migrated",
    "mpmgpu_archive_record_size", "mpmgpu_set_archive_origin", "mpmgpu_pack_archive", "mpmgpu_global_sums", "mpmgpu_download_ids"]

HALO_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)     # mpmgpu_halo_fn
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("device", C.c_int), ("np", C.c_int),
                ("horiz", C.c_int), ("vert", C.c_int), ("depth", C.c_int),
                ("xpts", _dp), ("ypts", _dp), ("zpts", _dp),
                ("gridx", C.c_double), ("gridy", C.c_double), ("gridz", C.c_double),
                ("thickness", C.c_double), ("shape", C.c_int), ("cpdi_rcrit", C.c_double),
                ("method", C.c_int), ("skip_post_extrapolation", C.c_int), ("fraction_usf", C.c_double),
                ("xpic_order", C.c_int), ("using_fmpm", C.c_int),
                ("grid_damping", C.c_double), ("particle_damping", C.c_double), ("gravity", C.c_double * 3),
                ("max_particles", C.c_int), ("sort_interval", C.c_int), ("kernel_path", C.c_int)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_history", C.c_int), ("p", C.c_double * MAT_NPARAMS)]


class ParticlesView(C.Structure):
    _fields_ = [("n", C.c_int), ("n_nonrigid", C.c_int),
                ("pos", _dp), ("vel", _dp), ("mp", _dp), ("lp", _dp), ("in_elem", _ip), ("matnum", _ip),
                ("sp", _dp), ("pressure", _dp), ("ep", _dp), ("wrot", _dp), ("eplast", _dp),
                ("energies", _dp), ("history", _dp), ("pfext", _dp), ("crossings", _ip), ("acc", _dp), ("ids", _ip), ("temperature", _dp)]


class NodesView(C.Structure):
    _fields_ = [("nnodes", C.c_int), ("number_points", _ip), ("mass", _dp), ("pk", _dp), ("ftot", _dp),
                ("vk", _dp), ("pk_copy", _dp), ("contact_volume", _dp), ("contact_gradient", _dp), ("contact_disp", _dp),
                ("transport_value", _dp), ("transport_capacity", _dp), ("transport_rate", _dp)]


class MultiMaterial(C.Structure):
    _fields_ = [("n_fields", C.c_int), ("field_of_material", _ip), ("normal_method", C.c_int), ("contact_by_displacements", C.c_int),
                ("position_cutoff", C.c_double), ("contact_normal", C.c_double * 3), ("law_kind", _ip), ("law_friction", _dp),
                ("law_static", _dp), ("rigid_gradient_bias", C.c_double)]


GS_NSUMS = 29
(GS_MASS, GS_VOLUME, GS_LINMOM, GS_KINETIC, GS_WORK, GS_STRAIN_ENERGY, GS_HEAT, GS_ENTROPY, GS_PLASTIC, GS_STRESS, GS_VOL_VEL,
 GS_VOL_F) = (0, 1, 2, 5, 6, 7, 8, 9, 10, 11, 17, 20)


class MpmGpuError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libmpmgpu error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen libmpmgpu.so and declare prototypes.  Fails loudly if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise MpmGpuError(-2, "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.mpmgpu_abi_version.restype = C.c_int
    lib.mpmgpu_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.mpmgpu_destroy.argtypes = [vp]
    lib.mpmgpu_last_error.argtypes = [vp]
    lib.mpmgpu_last_error.restype = C.c_char_p
    lib.mpmgpu_set_materials.argtypes = [vp, C.c_int, C.POINTER(Material)]
    lib.mpmgpu_upload_particles.argtypes = [vp, C.POINTER(ParticlesView)]
    lib.mpmgpu_set_multimaterial.argtypes = [vp, C.POINTER(MultiMaterial)]
    lib.mpmgpu_set_conduction.argtypes = [vp, C.c_int, _dp]
    lib.mpmgpu_set_temperature_bcs.argtypes = [vp, C.c_int, _ip, _dp, _ip]
    lib.mpmgpu_set_particle_tractions.argtypes = [vp, C.c_int, _ip, _ip, _ip, _dp]
    lib.mpmgpu_update_particle_traction_values.argtypes = [vp, C.c_int, _dp]
    lib.mpmgpu_set_particle_heat_fluxes.argtypes = [vp, C.c_int, _ip, _ip, _dp]
    lib.mpmgpu_update_particle_heat_flux_values.argtypes = [vp, C.c_int, _dp]
    lib.mpmgpu_contact_forces.argtypes = [vp, C.c_int, _dp]
    lib.mpmgpu_update_rigid_temperatures.argtypes = [vp, C.c_int, _dp]
    lib.mpmgpu_track_reactions.argtypes = [vp, C.c_int]
    lib.mpmgpu_download_reactions.argtypes = [vp, C.c_int, _dp, _dp]
    lib.mpmgpu_set_time_step.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.mpmgpu_set_xpic.argtypes = [vp, C.c_int, C.c_int]
    lib.mpmgpu_set_velocity_bcs.argtypes = [vp, C.c_int, _ip, _dp, _dp, _ip, _ip]
    lib.mpmgpu_update_velocity_bc_values.argtypes = [vp, C.c_int, _dp, _ip]
    lib.mpmgpu_set_velocity_bc_reflections.argtypes = [vp, C.c_int, _ip, _dp]
    lib.mpmgpu_update_particle_loads.argtypes = [vp, C.c_int, _ip, _dp]
    lib.mpmgpu_update_rigid_velocities.argtypes = [vp, C.c_int, _dp]
    lib.mpmgpu_step.argtypes = [vp, C.c_int]
    for t in TASKS + ["project_rigid_bcs"]:
        getattr(lib, "mpmgpu_task_" + t).argtypes = [vp]
    lib.mpmgpu_download_particles.argtypes = [vp, C.POINTER(ParticlesView), C.c_uint]
    lib.mpmgpu_download_nodes.argtypes = [vp, C.POINTER(NodesView)]
    lib.mpmgpu_synchronize.argtypes = [vp]
    lib.mpmgpu_get_status.argtypes = [vp, C.POINTER(C.c_longlong), _dp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.mpmgpu_launch_count.argtypes = [vp]
    lib.mpmgpu_launch_count.restype = C.c_longlong
    lib.mpmgpu_stream.argtypes = [vp]
    lib.mpmgpu_stream.restype = vp
    lib.mpmgpu_set_profiling.argtypes = [vp, C.c_int]
    lib.mpmgpu_task_times.argtypes = [vp, _dp, C.POINTER(C.c_longlong)]
    pvp = C.POINTER(vp)
    lib.mpmgpu_slab_configure.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.mpmgpu_slab_halo_buffers.argtypes = [vp, pvp, pvp, pvp, pvp, C.POINTER(C.c_longlong)]
    lib.mpmgpu_slab_step_phase.argtypes = [vp, C.c_int]
    lib.mpmgpu_slab_set_halo_callback.argtypes = [vp, HALO_FN, vp]
    lib.mpmgpu_slab_migration_counts.argtypes = [vp, _ip, _ip]
    lib.mpmgpu_slab_migration_buffers.argtypes = [vp, pvp, pvp, pvp, pvp, _ip, _ip]
    lib.mpmgpu_slab_pack_migrants.argtypes = [vp]
    lib.mpmgpu_slab_finish_migration.argtypes = [vp, C.c_int, C.c_int]
    lib.mpmgpu_num_particles.argtypes = [vp]
    lib.mpmgpu_set_stream.argtypes = [vp, vp]
    lib.mpmgpu_left_grid_counts.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.mpmgpu_archive_record_size.argtypes = [vp, C.c_char_p]
    lib.mpmgpu_set_archive_origin.argtypes = [vp, _dp, _dp, C.c_double]
    lib.mpmgpu_pack_archive.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    lib.mpmgpu_global_sums.argtypes = [vp, _dp]
    lib.mpmgpu_download_ids.argtypes = [vp, _ip, C.c_int]
    lib.mpmgpu_nccl_unique_ids.argtypes = [vp]
    lib.mpmgpu_slab_connect.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.mpmgpu_slab_step.argtypes = [vp, C.c_int]
    lib.mpmgpu_slab_migrated.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    if lib.mpmgpu_abi_version() != ABI_VERSION:
        raise MpmGpuError(-1, "libmpmgpu ABI %d, binding expects %d" % (lib.mpmgpu_abi_version(), ABI_VERSION))
    if path == LIB_PATH:
        _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _c64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _c32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


class MpmGpu:
    """One libmpmgpu context bound to one GPU, created from a `problem.Problem`."""

    def __init__(self, prob, device=0, kernel_path=0, max_particles=0, sort_interval=0, upload=True):
        self.lib = load_library()
        self.prob = prob
        self._keep = []
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.device = device
        cfg.np = prob.np
        cfg.horiz, cfg.vert, cfg.depth = prob.horiz, prob.vert, prob.depth
        xp, yp = _c64(prob.xpts), _c64(prob.ypts)
        zp = _c64(prob.zpts) if prob.is3d else None
        self._keep += [xp, yp, zp]
        cfg.xpts, cfg.ypts, cfg.zpts = _d(xp), _d(yp), _d(zp)
        cfg.gridx, cfg.gridy, cfg.gridz = prob.grid
        cfg.thickness = prob.thickness
        cfg.shape = prob.shape
        cfg.cpdi_rcrit = prob.rcrit
        cfg.method = prob.method
        cfg.skip_post_extrapolation = int(prob.skip_post_extrapolation)
        cfg.fraction_usf = prob.fraction_usf
        cfg.xpic_order = prob.xpic_order
        cfg.using_fmpm = int(prob.using_fmpm)
        cfg.grid_damping = prob.grid_damping
        cfg.particle_damping = prob.particle_damping
        cfg.gravity = (C.c_double * 3)(*prob.gravity)
        cfg.max_particles = max_particles
        cfg.sort_interval = sort_interval
        cfg.kernel_path = kernel_path
        self.ctx = C.c_void_p()
        rc = self.lib.mpmgpu_create(C.byref(cfg), C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.mpmgpu_last_error(None).decode()
            self.ctx = None
            raise MpmGpuError(rc, msg)
        self.nnodes = prob.nnodes
        mats = (Material * len(prob.materials))()
        for k, m in enumerate(prob.materials):
            mats[k].kind = m["kind"]
            mats[k].n_history = m.get("n_history", 0)
            for j, v in enumerate(m["p"]):
                mats[k].p[j] = v
        self._check(self.lib.mpmgpu_set_materials(self.ctx, len(prob.materials), mats))
        self._check(self.lib.mpmgpu_set_time_step(self.ctx, prob.dt, prob.dt_strain_first, prob.dt_strain_last))
        mm = getattr(prob, "multimaterial", None)
        if mm is not None:
            self.set_multimaterial(mm)
        self.adiabatic = bool(getattr(prob, "adiabatic", False))
        if self.adiabatic:
            self._check(self.lib.mpmgpu_set_energy_coupling(self.ctx, 1))
        self.conduction = getattr(prob, "conduction", None) is not None
        if self.conduction:
            k = _c64(prob.conduction["kcond"])
            self._check(self.lib.mpmgpu_set_conduction(self.ctx, len(prob.materials), _d(k)))
            if prob.conduction.get("tbc_node") is not None:
                self.set_temperature_bcs(prob.conduction["tbc_node"], prob.conduction["tbc_value"])
        self.set_velocity_bcs(prob.bc_node, prob.bc_norm, prob.bc_value, prob.bc_active, prob.bc_symdir)
        if getattr(prob, "bc_reflected", None) is not None:
            self.set_velocity_bc_reflections(prob.bc_reflected, prob.bc_ratio)
        if upload:
            self.upload(prob.particles)
            if mm is not None and getattr(prob, "origpos", None) is not None:
                self.set_archive_origin(origpos=prob.origpos)        # contact by displacements: MPMBase::origpos
            tr = getattr(prob, "tractions", None)
            if tr is not None and len(tr["particle"]):
                self.set_particle_tractions(tr["particle"], tr["face"], tr["direction"], tr["value"])
            hf = getattr(prob, "heat_fluxes", None)
            if hf is not None and len(hf["particle"]):
                self.set_particle_heat_fluxes(hf["particle"], hf["face"], hf["value"])

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise MpmGpuError(rc, self.lib.mpmgpu_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.mpmgpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host -> device -----------------------------------------------------------------------
    def upload(self, pt):
        """pt: dict of host arrays in the layout of mpmgpu_particles (component-major)."""
        v = ParticlesView()
        n = int(np.asarray(pt["mp"]).shape[0])
        v.n = n
        v.n_nonrigid = int(pt.get("n_nonrigid", n))
        keep = {}
        for k in ("pos", "vel", "mp", "lp", "sp", "pressure", "ep", "wrot", "eplast", "energies", "history", "pfext", "temperature"):
            keep[k] = _c64(pt.get(k))
            setattr(v, k, _d(keep[k]))
        for k in ("in_elem", "matnum", "crossings", "ids"):
            keep[k] = _c32(pt.get(k))
            setattr(v, k, _i(keep[k]))
        self.n = n
        self.thermal = bool(getattr(self, "conduction", False)) or bool(getattr(self, "adiabatic", False)) or pt.get("temperature") is not None
        self._check(self.lib.mpmgpu_upload_particles(self.ctx, C.byref(v)))

    def set_multimaterial(self, mm):
        """<MultiMaterialMode> settings (Problem.multimaterial): one velocity field per material field + material contact."""
        v = MultiMaterial()
        nf = int(mm["n_fields"])
        keep = [_c32(mm["field_of_material"]), _c32(np.asarray(mm["law_kind"]).reshape(nf * nf)),
                _c64(np.asarray(mm["law_friction"], float).reshape(nf * nf)), _c64(np.asarray(mm["law_static"], float).reshape(nf * nf))]
        v.n_fields = nf
        v.field_of_material = _i(keep[0])
        v.normal_method = int(mm["normal_method"])
        v.contact_by_displacements = int(mm["by_displacements"])
        v.position_cutoff = float(mm["position_cutoff"])
        v.contact_normal = (C.c_double * 3)(*[float(x) for x in mm["contact_normal"]])
        v.law_kind, v.law_friction, v.law_static = _i(keep[1]), _d(keep[2]), _d(keep[3])
        v.rigid_gradient_bias = float(mm.get("rigid_gradient_bias", 1.0))
        self._check(self.lib.mpmgpu_set_multimaterial(self.ctx, C.byref(v)))
        self.nnodes = self.prob.nnodes * nf          # node arrays are field-major from now on
        self.n_fields = nf

    def set_temperature_bcs(self, node, value, active=None):
        """Nodal temperature BCs in list order (1-based nodes, values at this step's time); call again when values change."""
        node, value, active = _c32(node), _c64(value), _c32(active)
        self._check(self.lib.mpmgpu_set_temperature_bcs(self.ctx, 0 if node is None else len(node), _i(node), _d(value), _i(active)))

    def set_velocity_bcs(self, node, norm, value, active=None, symdir=None):
        n = 0 if node is None else len(node)
        node, norm, value = _c32(node), _c64(norm), _c64(value)
        active, symdir = _c32(active), _c32(symdir)
        self._check(self.lib.mpmgpu_set_velocity_bcs(self.ctx, n, _i(node), _d(norm), _d(value), _i(active), _i(symdir)))

    def set_particle_tractions(self, particle, face, direction, value):
        """Traction BCs in list order: 0-based particle, face, direction (1, 2, 3, 11 normal, 12 tangent), stress at this step's time."""
        particle, face, direction, value = _c32(particle), _c32(face), _c32(direction), _c64(value)
        self._check(self.lib.mpmgpu_set_particle_tractions(self.ctx, len(particle), _i(particle), _i(face), _i(direction), _d(value)))

    def set_particle_heat_fluxes(self, particle, face, value):
        """External heat-flux BCs in list order: 0-based particle, face, flux at this step's time (conduction must be on)."""
        particle, face, value = _c32(particle), _c32(face), _c64(value)
        self._check(self.lib.mpmgpu_set_particle_heat_fluxes(self.ctx, len(particle), _i(particle), _i(face), _d(value)))

    def update_particle_traction_values(self, value):
        value = _c64(value)
        self._check(self.lib.mpmgpu_update_particle_traction_values(self.ctx, len(value), _d(value)))

    def contact_forces(self, clear=True):
        """[n_fields, 3]: momentum the contacts of each rigid material field gave the other materials since the last clearing."""
        out = np.zeros((self.n_fields, 3))
        self._check(self.lib.mpmgpu_contact_forces(self.ctx, 1 if clear else 0, _d(out)))
        return out

    def track_reactions(self, on=True):
        """Keep the reaction force of every velocity BC (NodalVelBC::freaction); read them with reactions()."""
        self._check(self.lib.mpmgpu_track_reactions(self.ctx, 1 if on else 0))

    def reactions(self):
        """(bc [n,3] in the order of the velocity-BC list, rigid [nmat,3] per rigid-BC material) of the last step."""
        n = 0 if self.prob.bc_node is None else len(self.prob.bc_node)
        bc = np.zeros((n, 3))
        rigid = np.zeros((len(self.prob.materials), 3))
        self._check(self.lib.mpmgpu_download_reactions(self.ctx, n, _d(bc), _d(rigid)))
        return bc, rigid

    def set_velocity_bc_reflections(self, reflected_node, ratio):
        """Symmetry-plane BCs: per BC of the list the 1-based node it reflects (<= 0: plain BC) and the cell-size ratio."""
        reflected_node, ratio = _c32(reflected_node), _c64(ratio)
        self._check(self.lib.mpmgpu_set_velocity_bc_reflections(self.ctx, len(reflected_node), _i(reflected_node), _d(ratio)))

    def update_velocity_bc_values(self, value, active=None):
        value, active = _c64(value), _c32(active)
        self._check(self.lib.mpmgpu_update_velocity_bc_values(self.ctx, len(value), _d(value), _i(active)))

    def update_particle_loads(self, fext, particle=None):
        """fext [3][n_loaded]: this step's external forces on the loaded particles (MatPtLoadBC evaluated by the host);
        particle: their 0-based indices, on the first call."""
        fext, particle = _c64(fext), _c32(particle)
        self._check(self.lib.mpmgpu_update_particle_loads(self.ctx, int(fext.shape[-1]), _i(particle), _d(fext)))

    def update_rigid_velocities(self, vel):
        """vel [3][n_rigid]: this step's velocities of the rigid-BC particles (setting functions evaluated by the host)."""
        vel = _c64(vel)
        self._check(self.lib.mpmgpu_update_rigid_velocities(self.ctx, int(vel.shape[-1]), _d(vel)))

    def set_time_step(self, dt, dt_first, dt_last):
        self._check(self.lib.mpmgpu_set_time_step(self.ctx, dt, dt_first, dt_last))

    def set_xpic(self, order, using_fmpm):
        self._check(self.lib.mpmgpu_set_xpic(self.ctx, order, int(using_fmpm)))

    # -- the step -----------------------------------------------------------------------------
    def set_poll_interval(self, k):
        """Read the status word (one stream synchronisation) every k-th step() call only; errors surface up to k-1 steps late."""
        self._check(self.lib.mpmgpu_set_poll_interval(self.ctx, int(k)))

    def step(self, nsteps=1):
        self._check(self.lib.mpmgpu_step(self.ctx, int(nsteps)))

    def run_task(self, name_or_index):
        name = TASKS[name_or_index] if isinstance(name_or_index, int) else name_or_index
        self._check(getattr(self.lib, "mpmgpu_task_" + name)(self.ctx))

    def synchronize(self):
        self._check(self.lib.mpmgpu_synchronize(self.ctx))

    # -- device -> host -----------------------------------------------------------------------
    def download(self, mask=F_ALL, out=None):
        """Returns dict of host arrays.  `out`: reuse caller-provided (e.g. pinned) arrays of the right shape."""
        n = self.num_particles()
        if out is not None:
            return self._download_into(out, mask)
        out = dict(pos=np.zeros((3, n)), vel=np.zeros((3, n)), sp=np.zeros((6, n)), pressure=np.zeros(n),
                   ep=np.zeros((6, n)), wrot=np.zeros((3, n)), eplast=np.zeros((6, n)), energies=np.zeros((6, n)),
                   history=np.zeros((MAX_HISTORY, n)), acc=np.zeros((3, n)),
                   in_elem=np.zeros(n, np.int32), crossings=np.zeros(n, np.int32), ids=np.zeros(n, np.int32))
        v = ParticlesView()
        v.ids = _i(out["ids"])
        for k in ("pos", "vel", "sp", "pressure", "ep", "wrot", "eplast", "energies", "history", "acc"):
            setattr(v, k, _d(out[k]))
        v.in_elem, v.crossings = _i(out["in_elem"]), _i(out["crossings"])
        if getattr(self, "thermal", False):
            out["temperature"] = np.zeros(n)
            v.temperature = _d(out["temperature"])
            mask |= F_TEMPERATURE
        self._check(self.lib.mpmgpu_download_particles(self.ctx, C.byref(v), mask))
        return out

    def _download_into(self, out, mask):
        v = ParticlesView()
        for k in ("pos", "vel", "sp", "pressure", "ep", "wrot", "eplast", "energies", "history", "acc"):
            if k in out:
                setattr(v, k, _d(out[k]))
        for k in ("in_elem", "crossings", "ids"):
            if k in out:
                setattr(v, k, _i(out[k]))
        self._check(self.lib.mpmgpu_download_particles(self.ctx, C.byref(v), mask))
        return out

    @staticmethod
    def pinned_download_buffers(n):
        """Page-locked host arrays for download(out=...) (D2H at full PCIe rate)."""
        import torch
        shapes = dict(pos=(3, n), vel=(3, n), sp=(6, n), pressure=(n,), ep=(6, n), wrot=(3, n), eplast=(6, n),
                      energies=(6, n), history=(MAX_HISTORY, n), acc=(3, n))
        out = {k: torch.zeros(sh, dtype=torch.float64).pin_memory().numpy() for k, sh in shapes.items()}
        for k in ("in_elem", "crossings", "ids"):
            out[k] = torch.zeros(n, dtype=torch.int32).pin_memory().numpy()
        return out

    def download_nodes(self):
        n = self.nnodes
        out = dict(number_points=np.zeros(n, np.int32), mass=np.zeros(n), pk=np.zeros((3, n)), ftot=np.zeros((3, n)),
                   vk=np.zeros((3, n)), pk_copy=np.zeros((3, n)))
        v = NodesView()
        v.number_points = _i(out["number_points"])
        for k in ("mass", "pk", "ftot", "vk", "pk_copy"):
            setattr(v, k, _d(out[k]))
        if getattr(self, "conduction", False):
            nr = self.prob.nnodes
            out.update(transport_value=np.zeros(nr), transport_capacity=np.zeros(nr), transport_rate=np.zeros(nr))
            for k in ("transport_value", "transport_capacity", "transport_rate"):
                setattr(v, k, _d(out[k]))
        if getattr(self, "n_fields", 0):
            out.update(contact_volume=np.zeros(n), contact_gradient=np.zeros((3, n)), contact_disp=np.zeros((3, n)))
            for k in ("contact_volume", "contact_gradient", "contact_disp"):
                setattr(v, k, _d(out[k]))
        self._check(self.lib.mpmgpu_download_nodes(self.ctx, C.byref(v)))
        return out

    # ---- output side on the device ----
    def set_archive_origin(self, origpos=None, angles0=None, thickness=None):
        """Constants of the archive records: original positions [3][n], initial material angles [3][n] (z, y, x; radians),
        2D thickness.  Without this call (or for None): the positions at upload, zero angles, the problem's thickness."""
        self._keep_arch = [_c64(origpos), _c64(angles0)]
        self._check(self.lib.mpmgpu_set_archive_origin(self.ctx, _d(self._keep_arch[0]), _d(self._keep_arch[1]),
                                                       float(self.prob.thickness if thickness is None else thickness)))

    def archive_record_size(self, order):
        return int(self.lib.mpmgpu_archive_record_size(self.ctx, order.encode("latin-1")))

    def pack_archive(self, order, out=None):
        """The record block of one particle archive (reference binary format, caller's particle order), packed on the
        device: bytes (or the filled `out`, a uint8 array that may be pinned).  archive.header() makes the file header."""
        rec = self.archive_record_size(order)
        if rec < 0:
            raise MpmGpuError(-1, "archive order %r asks for an item this path does not produce" % order)
        nbytes = rec * self.num_particles()
        buf = np.empty(nbytes, np.uint8) if out is None else out
        self._check(self.lib.mpmgpu_pack_archive(self.ctx, order.encode("latin-1"), buf.ctypes.data_as(C.c_void_p), buf.nbytes))
        return buf[:nbytes].tobytes() if out is None else buf

    # ---- NCCL inside the library ----
    def nccl_unique_ids(self):
        """256 bytes: the ids of the two NCCL communicators of a slab run (rank 0 makes them, every rank gets them)."""
        buf = C.create_string_buffer(256)
        if self.lib.mpmgpu_nccl_unique_ids(buf) != 0:
            raise MpmGpuError(-1, "mpmgpu_nccl_unique_ids failed (libnccl.so.2 not found?)")
        return buf.raw

    def slab_connect(self, rank, world, ids):
        buf = C.create_string_buffer(bytes(ids), 256)
        self._check(self.lib.mpmgpu_slab_connect(self.ctx, int(rank), int(world), buf))

    def slab_step(self, nsteps=1):
        self._check(self.lib.mpmgpu_slab_step(self.ctx, int(nsteps)))

    def slab_migrated(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._check(self.lib.mpmgpu_slab_migrated(self.ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def download_ids(self):
        """Particle ids in device order (slab mode: global ids, the order of pack_archive's records there)."""
        ids = np.zeros(self.num_particles(), np.int32)
        self._check(self.lib.mpmgpu_download_ids(self.ctx, _i(ids), len(ids)))
        return ids

    def global_sums(self):
        """[nmat][GS_NSUMS] raw sums behind the reference's GlobalQuantity rows (see include/mpmgpu.h MPMGPU_GS_*)."""
        out = np.zeros((len(self.prob.materials), GS_NSUMS))
        self._check(self.lib.mpmgpu_global_sums(self.ctx, _d(out)))
        return out

    def num_particles(self):
        return int(self.lib.mpmgpu_num_particles(self.ctx))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.mpmgpu_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)))

    # -- slab mode (see slab.py) ----------------------------------------------------------------
    def slab_configure(self, cell_lo, cell_hi, has_lower, has_upper, migration_capacity=0):
        self._check(self.lib.mpmgpu_slab_configure(self.ctx, cell_lo, cell_hi, int(has_lower), int(has_upper), migration_capacity))

    def slab_halo_buffers(self):
        p = [C.c_void_p() for _ in range(4)]
        pn = C.c_longlong()
        self._check(self.lib.mpmgpu_slab_halo_buffers(self.ctx, *[C.byref(x) for x in p], C.byref(pn)))
        return [x.value for x in p], int(pn.value)

    def slab_migration_buffers(self):
        p = [C.c_void_p() for _ in range(4)]
        row, cap = C.c_int(), C.c_int()
        self._check(self.lib.mpmgpu_slab_migration_buffers(self.ctx, *[C.byref(x) for x in p], C.byref(row), C.byref(cap)))
        return [x.value for x in p], int(row.value), int(cap.value)

    def slab_set_halo_callback(self, fn):
        """fn(which): enqueue the swap of halo kind `which` with the neighbours (called from inside a phase when the
        XPIC/FMPM order is > 1).  An exception raised by fn is kept and re-raised when the phase returns."""
        self._halo_error = None

        def tramp(_user, which):
            try:
                fn(int(which))
            except BaseException as e:      # must not unwind through the C frames
                self._halo_error = e
        self._halo_cb = HALO_FN(tramp)      # keep the thunk alive
        self._check(self.lib.mpmgpu_slab_set_halo_callback(self.ctx, self._halo_cb, None))

    def slab_phase(self, phase):
        rc = self.lib.mpmgpu_slab_step_phase(self.ctx, phase)
        err = getattr(self, "_halo_error", None)
        if err is not None:
            self._halo_error = None
            raise err
        self._check(rc)

    def slab_migration_counts(self):
        a, b = C.c_int(), C.c_int()
        self._check(self.lib.mpmgpu_slab_migration_counts(self.ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def slab_pack_migrants(self):
        self._check(self.lib.mpmgpu_slab_pack_migrants(self.ctx))

    def slab_finish_migration(self, n_from_lo, n_from_hi):
        self._check(self.lib.mpmgpu_slab_finish_migration(self.ctx, n_from_lo, n_from_hi))

    def status(self):
        ms, cr, lg = C.c_longlong(), C.c_longlong(), C.c_longlong()
        mt = C.c_double()
        self._check(self.lib.mpmgpu_get_status(self.ctx, C.byref(ms), C.byref(mt), C.byref(cr), C.byref(lg)))
        return dict(mstep=ms.value, mtime=mt.value, crossings=cr.value, left_grid=lg.value)

    def left_grid_counts(self):
        """(push-backs, particles that left the grid for the first time) since the upload."""
        ex, pt = C.c_longlong(0), C.c_longlong(0)
        self._check(self.lib.mpmgpu_left_grid_counts(self.ctx, C.byref(ex), C.byref(pt)))
        return ex.value, pt.value

    def launch_count(self):
        return int(self.lib.mpmgpu_launch_count(self.ctx))

    def stream(self):
        return self.lib.mpmgpu_stream(self.ctx)

    def set_profiling(self, on):
        self._check(self.lib.mpmgpu_set_profiling(self.ctx, int(on)))

    def task_times(self):
        ms = np.zeros(len(TASKS))
        calls = np.zeros(len(TASKS), np.int64)
        self._check(self.lib.mpmgpu_task_times(self.ctx, _d(ms), calls.ctypes.data_as(C.POINTER(C.c_longlong))))
        return {t: (float(ms[i]), int(calls[i])) for i, t in enumerate(TASKS)}
