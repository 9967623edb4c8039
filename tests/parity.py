"""Shared helpers for the parity tests: compare libmpmgpu output with reference dumps."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# reference task name -> libmpmgpu task entry point
TASK_MAP = {
    "Initialize": "initialization",
    "Extrapolate Mass and Momentum": "mass_and_momentum",
    "Rigid BCs by Projection": "project_rigid_bcs",
    "Post Extrapolation Tasks": "post_extrapolation",
    "Update Strains First": "update_strains_first",
    "Extrapolate Grid Forces": "grid_forces",
    "Post Force Extrapolation Tasks": "post_forces",
    "Update Momenta": "update_momenta",
    "Update Particles": "update_particles",
    "Update Strains Last with Extrapolation": "update_strains_last",
    "Update Strains Last": "update_strains_last",
    "Reset Elements": "reset_elements",
    "Decipher Crack and Material Fields": None,      # InitVelocityFieldsTask: the device knows every particle's field from its material
    "Set Rigid Contact Velocities": None,            # SetRigidContactVelTask: the host evaluates the setting functions (update_rigid_velocities)
    "Run Custom Tasks": None,        # host-side: PeriodicXPIC only changes the XPIC order for the next step
}


def xpic_for_step(z, step):
    """(order, usingFMPM) the reference used during 1-based `step` (recorded by the harness), or None."""
    if "xpic_by_step" not in z:
        return None
    x = z["xpic_by_step"]
    if step - 1 < len(x):
        return int(x[step - 1][0]), int(x[step - 1][1])
    return int(x[-1][0]), int(x[-1][1])


def per_task_steps(z):
    n = 0
    while ("s%d/t0/nodes/mass" % (n + 1)) in z:
        n += 1
    return n

TOL_1STEP = 1.0e-10      # BASELINE.json north_star: relative 1e-10 after 1 step
TOL_100STEP = 1.0e-7     # and 1e-7 after 100 steps (FP64), relative to the field's max magnitude

# Elastic::useLargeRotation in 3D.  The reference finds the rotation of F through the trigonometric eigenvalues of F^T F
# (Matrix3::Eigenvalues, Common/System/Matrix3.cpp:464-520).  For the strain increments of an explicit step the cubic's
# discriminant is pure cancellation noise, the eigenvalues are wrong by O(strain) and the strain increment by ~1e-6 of
# itself -- as a function of the last bits of the input.  The reference therefore does not reproduce ITSELF to better than
# that when its input changes by one ulp (tests/test_oracle_cpu.py::test_large_rotation_3d_is_ill_conditioned_in_the_reference_algorithm),
# and no independent implementation can agree more closely.  2D (closed-form rotation angle) keeps the standard tolerances.
LR3D_CASES = ("block3d_isotropic_lr", "block3d_isoplastic_lr")
TOL_LR3D = 1.0e-4          # observed: <= 3e-6 on the two goldens (80 steps), <= 3e-5 over the random combinations of tests/test_sweep_cpu.py


# Hardening laws returned numerically (Nonlinear, Nonlinear2, JohnsonCook): the reference stops the bracketed Newton iteration for
# lambda at |d lambda / lambda| < 1e-4 (HardeningLawBase::LambdaConverged, HardeningLawBase.cpp:386-390), so the plastic increment
# is DEFINED to 1e-4 only: a last-bit difference upstream can end the iteration one step earlier or later, or turn a Newton step
# into a bisection.  With the same libm one step agrees to round-off (tests/test_device_laws_vs_reference_cpu.py holds the law to 1e-13 on
# identical inputs); whole-step and long-run comparisons are held to the solver's own tolerance.
TOL_ITERATIVE = 1.0e-4          # observed <= 2e-5 after 100 steps


def tolerances(case):
    """(after 1 step / per task of step 1, per task of later steps, after N <= 100 steps)"""
    if case in LR3D_CASES:
        return TOL_LR3D, TOL_LR3D, TOL_LR3D
    if "johnsoncook" in case or "nonlinear" in case:
        # also for a single step: a device libm that rounds pow/log differently can flip one particle's iteration path
        return TOL_ITERATIVE, TOL_ITERATIVE, TOL_ITERATIVE
    return TOL_1STEP, 1.0e-8, TOL_100STEP


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


def dedupe_per_task(z):
    """Drop a per-task dump entry (sK/tI/p/<field>, sK/tI/nodes/<field>) when it is bit-identical to the same field after the
    previous task (most tasks touch a few fields only); restore_per_task() puts them back.  Keeps the fixtures small."""
    names = list(z["task_names"])
    nt = len(names)
    out = dict(z)
    last = {}
    step = 1
    while ("s%d/t0/nodes/mass" % step) in z:
        for t in range(nt):
            pre = "s%d/t%d/" % (step, t)
            for k in [k for k in z if k.startswith(pre)]:
                fld = k[len(pre):]
                if fld in last and _same(last[fld], z[k]):
                    del out[k]
                else:
                    last[fld] = z[k]
        step += 1
    return out


def restore_per_task(z):
    if "task_names" not in z:
        return z
    nt = len(z["task_names"])
    fields = sorted({k.split("/", 2)[2] for k in z if k.startswith("s") and k.split("/")[0][1:].isdigit() and k.count("/") >= 2 and k.split("/")[1].startswith("t")})
    last = {}
    step = 1
    while any(k.startswith("s%d/t" % step) for k in z):
        for t in range(nt):
            pre = "s%d/t%d/" % (step, t)
            for fld in fields:
                k = pre + fld
                if k in z:
                    last[fld] = z[k]
                elif fld in last:
                    z[k] = last[fld]
        step += 1
    return z


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return restore_per_task({k: z[k] for k in z.files})


def rel_err(a, b, scale_with=None):
    """max |a-b| / max |b| over the whole field; NaN must match NaN (the reference's entropy is
    0/0 when the reference temperature is 0).  scale_with: another reference array that belongs to
    the same physical quantity (wrot and ep are the two halves of the deformation gradient)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return np.inf
    if a.size == 0:
        return 0.0
    d = np.where(nb, 0.0, np.abs(a - b))
    scale = np.max(np.where(nb, 0.0, np.abs(b)))
    if scale_with is not None and not np.all(np.isnan(scale_with)):
        scale = max(scale, float(np.nanmax(np.abs(scale_with))))
    if scale == 0.0:
        return float(np.max(d))          # reference field identically zero: absolute error
    return float(np.max(d) / scale)


def _scales(ref, key, scale_prefix, prefix, extra=None):
    """other reference arrays of the same physical quantity to scale an error with (rel_err's scale_with)"""
    out = [] if extra is None else [np.ravel(extra)]
    if scale_prefix is not None:
        out.append(np.ravel(ref[key.replace(prefix, scale_prefix, 1)]))
    return np.concatenate(out) if out else None


def compare_particles(got, ref, prefix, tol, fields=None, scale_prefix=None):
    """got: MpmGpu.download() dict; ref: golden dict with keys prefix/<field>.  Returns {field: err}.
    scale_prefix: a second dump (same task of a later step) whose magnitudes also scale the errors -- for fields that are
    rounding noise in `prefix` (the strains of two bodies in uniform motion before their first contact)."""
    pairs = [("pos", "pos"), ("vel", "vel"), ("sp", "sp"), ("pressure", "pressure"), ("ep", "ep"), ("wrot", "wrot"),
             ("eplast", "eplast")]
    errs = {}
    # 2D runs: the out-of-plane shear components (yz, xz) are not part of the state -- the reference's
    # IsoPlasticity even adds an uninitialised Tensor into eplast.yz/xz there (IsoPlasticity.cpp:395-405)
    two_d = "info/np" in ref and int(ref["info/np"]) != 12
    for g, r in pairs:
        if fields and g not in fields:
            continue
        a, b = np.asarray(got[g]), np.asarray(ref[prefix + "/" + r])
        if two_d and g in ("sp", "ep", "eplast"):
            a, b = a[[0, 1, 2, 5]], b[[0, 1, 2, 5]]
        errs[g] = rel_err(a, b, _scales(ref, prefix + "/" + r, scale_prefix, prefix, ref[prefix + "/ep"] if g == "wrot" else None))
    e = ref[prefix + "/energies"]
    names = ["work", "res", "heat", "entropy", "plast"]
    for i, nm in enumerate(names):
        errs[nm] = rel_err(got["energies"][i], e[i], None if scale_prefix is None else ref[scale_prefix + "/energies"][i])
    nh = min(got["history"].shape[0], ref[prefix + "/hist"].shape[0])
    errs["history"] = rel_err(got["history"][:nh], ref[prefix + "/hist"][:nh])
    if "temperature" in got and (prefix + "/temperature") in ref:      # conduction: pTemperature
        errs["temperature"] = rel_err(got["temperature"], ref[prefix + "/temperature"])
    bad = {k: v for k, v in errs.items() if not v <= tol}
    return errs, bad


def compare_nodes(got, ref, prefix, tol, scale_prefix=None):
    errs = {}
    sc = lambda k: None if scale_prefix is None else ref[scale_prefix + "/" + k]          # noqa: E731
    errs["mass"] = rel_err(got["mass"], ref[prefix + "/mass"])
    errs["pk"] = rel_err(got["pk"], ref[prefix + "/pk"])
    errs["ftot"] = rel_err(got["ftot"], ref[prefix + "/ftot"], sc("ftot"))
    if "contact_volume" in got and (prefix + "/cvolume") in ref:
        # Pair contact hands the SECOND field of a node minus the momentum change computed for the first one
        # (CrackVelocityFieldMulti.cpp:640-645), so a field with a vanishing share of the node's mass (the tails of the
        # spline shape functions: 1e-13 of its neighbour) keeps its momentum only to an ulp of the OTHER field's momentum and
        # its velocity p/m is rounding noise -- in the reference too.  Velocities are compared where the field carries mass.
        heavy = ref[prefix + "/mass"] >= 1.0e-6 * np.max(ref[prefix + "/mass"])
        errs["vk"] = rel_err(got["vk"][:, heavy], ref[prefix + "/vk0"][:, heavy])
    else:
        errs["vk"] = rel_err(got["vk"], ref[prefix + "/vk0"])
    errs["pk_copy"] = rel_err(got["pk_copy"], ref[prefix + "/pkcopy"])
    if "transport_value" in got and (prefix + "/gT") in ref:
        # conduction: NodalPoint::gCond.  The heat flow gQ is a sum of conduction terms that cancel inside a body of uniform
        # temperature: it is scaled with the capacity-weighted temperature rate the same node could carry.
        errs["transport_value"] = rel_err(got["transport_value"], ref[prefix + "/gT"])
        errs["transport_capacity"] = rel_err(got["transport_capacity"], ref[prefix + "/gVCT"])
        errs["transport_rate"] = rel_err(got["transport_rate"], ref[prefix + "/gQ"], sc("gQ"))
    if "contact_volume" in got and (prefix + "/cvolume") in ref:
        # multimaterial mode: the contact extrapolations (field-major arrays on both sides)
        errs["contact_volume"] = rel_err(got["contact_volume"], ref[prefix + "/cvolume"])
        errs["contact_gradient"] = rel_err(got["contact_gradient"], ref[prefix + "/cgrad"])
        errs["contact_disp"] = rel_err(got["contact_disp"], ref[prefix + "/cdisp"])
    bad = {k: v for k, v in errs.items() if not v <= tol}
    return errs, bad


# ---- multimaterial mode (goldens mm*): shared by the CPU run of the device source and the GPU run through the C ABI ----------
MM_CASES = ["trac2d_multimaterial_lcpdi_planestress", "mm2d_friction_avgg", "mm2d_frictionless_maxg_position", "mm2d_stick_maxv_linear_usl", "mm2d_ignore_lcpdi_usf",
            "mm2d_friction_sn_powerlaw", "mm3d_two_blocks_avgg_position", "mm3d_two_blocks_maxg_stick_b2gimp",
            "mm3d_two_blocks_maxv_friction_ugimp", "mm2d_rigid_plate_maxg_friction", "mm3d_rigid_block_avgg_position_usl",
            "mm3d_rigid_block_maxv_stick_lcpdi"]


COND_CASES = ["cond3d_block_fmpm2_temperature_bcs", "cond2d_disks_xpic2_usl", "cond3d_heat_flux_ugimp", "cond2d_heat_flux_lcpdi_planestress", "cond3d_rigid_hot_piston", "cond3d_rigid_heater_lcpdi_usl", "cond2d_disks_usavg", "cond2d_disks_lcpdi_usl_neo", "cond3d_blocks_multimaterial", "cond3d_block_temperature_bcs"]
# thermal strains: conduction with expanding materials (th*_cond_*), bodies that start off the stress-free temperature (th*_offset_*)
THERMAL_COND_CASES = ["th2d_cond_iso_planestrain", "th2d_cond_isoplastic_neo_planestress", "th2d_adiabatic_conduction_isoplastic"]
THERMAL_OFFSET_CASES = ["th3d_offset_scgl", "th3d_offset_iso", "th3d_offset_isoplastic_usl", "th3d_offset_neohookean", "th2d_offset_mooney_iso_planestress",
                        "th3d_adiabatic_johnsoncook"]


def check_multimaterial_tasks(sim, z, case, require="contact_volume"):
    """Every task of the first two steps: node fields of every material velocity field (field-major arrays), the contact
    extrapolations, particle fields.  Until the first contact changes a field's velocities the bodies move uniformly and their
    strains and forces are rounding noise, so the errors of step 1 are also scaled with the same task's dump of step 2."""
    names = [str(s) for s in z["task_names"]]
    for step in range(1, per_task_steps(z) + 1):
        tol = TOL_1STEP if step == 1 else 1.0e-8
        for i, nm in enumerate(names):
            if TASK_MAP[nm] is None:
                continue
            sim.run_task(TASK_MAP[nm])
            pre = "s%d/t%d" % (step, i)
            later = "s2/t%d" % i if step == 1 else None
            nodes = sim.download_nodes()
            errs, bad = compare_nodes(nodes, z, pre + "/nodes", tol, later and later + "/nodes")
            assert not bad, "%s step %d after task %d (%s): node fields %s" % (case, step, i, nm, bad)
            assert require in errs
            assert np.array_equal(nodes["number_points"] > 0, z[pre + "/nodes/numberPoints"] > 0), "active (field, node) set differs"
            got = sim.download()
            errs, bad = compare_particles(got, z, pre + "/p", tol, scale_prefix=later and later + "/p")
            assert not bad, "%s step %d after task %d (%s): particle fields %s" % (case, step, i, nm, bad)
            assert require != "transport_value" or "temperature" in errs
            assert np.array_equal(got["in_elem"], z[pre + "/p/inElem"])


def check_multimaterial_run(sim, z, case):
    """Whole steps: 1 step to 1e-10 (noise fields scaled as above), N steps to 1e-7; element ids and crossings exact.
    Returns the number of nodes that carry more than one material at the end (the test wants contact to have happened)."""
    snaps = sorted(int(k[1:].split("/")[0]) for k in z if k.startswith("p") and k.endswith("/pos") and k[1] != "0")
    done = 0
    for s in snaps:
        sim.step(s - done)
        done = s
        tol = TOL_1STEP if s <= 2 else TOL_100STEP
        later = "p2" if s == 1 else None
        got = sim.download()
        errs, bad = compare_particles(got, z, "p%d" % s, tol, scale_prefix=later)
        assert not bad, "%s after %d steps: %s" % (case, s, bad)
        assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
        assert np.array_equal(got["crossings"], z["p%d/crossings" % s])
        nodes = sim.download_nodes()
        errs, bad = compare_nodes(nodes, z, "n%d" % s, tol, "n2" if s == 1 else None)
        assert not bad, "%s after %d steps: nodes %s" % (case, s, bad)
    nf = int(z["mm/nfields"]) if "mm/nfields" in z else 1
    cnt = nodes["number_points"].reshape(nf, -1) > 0
    return int(np.sum(cnt.sum(axis=0) > 1))


# reaction forces of the velocity BCs: goldens with "reaction<step>" = NodalVelBC::TotalReactionForce(id) for id in "reaction_ids"
REACTION_CASES = ["react3d_walls_ugimp", "react3d_walls_lcpdi_usl", "react3d_rigid_piston_fmpm2", "react3d_rigid_wall_xpic2", "react2d_multimaterial_wall"]


def reaction_totals(prob, ids, bc, rigid):
    """What GlobalQuantity's reactionx/y/z would read (GlobalQuantity.cpp:971-986): the BCs' freaction summed by bcID; 0 takes
    every BC, rigid-particle BCs carry their material number (ProjectRigidBCsTask.cpp:241)."""
    out = np.zeros((len(ids), 3))
    for k, i in enumerate(ids):
        if i == 0:
            out[k] = bc.sum(axis=0) + rigid.sum(axis=0)
        elif i < 0:
            out[k] = bc[prob.bc_id == i].sum(axis=0) if len(bc) else 0.0
        else:
            out[k] = rigid[i - 1]
    return out


def check_reactions(sim, prob, z, step, tol):
    """The reaction totals after `step` steps against the reference's, to tol of the largest total of that step."""
    ref = z["reaction%d" % step]
    bc, rigid = sim.reactions()
    got = reaction_totals(prob, [int(i) for i in z["reaction_ids"]], bc, rigid)
    scale = max(np.abs(ref).max(), np.abs(bc).sum(axis=0).max() if len(bc) else 0.0)
    assert scale > 0.0
    err = np.abs(got - ref).max() / scale
    assert err < tol, "reaction forces after step %d: rel err %.3g\n got %s\n ref %s" % (step, err, got, ref)
    return err


def check_reaction_run(sim, prob, z, case):
    """Whole steps with the reaction totals checked after every step the golden holds them for."""
    have = sorted(int(k[8:]) for k in z if k.startswith("reaction") and k != "reaction_ids")
    assert have and have[-1] > 2
    done = 0
    for s in have:
        while done < s:
            x = xpic_for_step(z, done + 1)
            if x:
                sim.set_xpic(*x)
            sim.step(1)
            done += 1
        check_reactions(sim, prob, z, s, TOL_1STEP if s <= 2 else TOL_100STEP)
        if ("p%d/pos" % s) in z:
            got = sim.download()
            errs, bad = compare_particles(got, z, "p%d" % s, TOL_1STEP if s <= 2 else TOL_100STEP, scale_prefix="p2" if s == 1 else None)
            assert not bad, "%s after %d steps: %s" % (case, s, bad)
            assert np.array_equal(got["in_elem"], z["p%d/inElem" % s])
