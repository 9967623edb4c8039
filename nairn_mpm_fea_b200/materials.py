"""Material parameter blocks for libmpmgpu (mpmgpu_material in include/mpmgpu.h).

Host-side mirror of what the reference's material classes compute in VerifyAndLoadProperties and hand to
MPMConstitutiveLaw through GetCopyOfMechanicalProps.  Inputs are in the reference's INTERNAL units
(Legacy units after UnitsController::ScaledPtr: modulus MPa*1e6, rho g/cm^3*1e-3, CTE ppm/K*1e-6,
Cv J/(kg-K)*1e6).  `from_xml_units` converts from the numbers written in an XML file.
"""
import numpy as np

NPARAMS = 32
MAX_HISTORY = 4            # MPMGPU_MAX_HISTORY
ISOTROPIC, MOONEY, ISOPLASTICITY, RIGIDBC, NEOHOOKEAN = 1, 8, 9, 11, 28
RIGIDCONTACT = 35           # MPMGPU_MAT_RIGIDCONTACT: RigidMaterial (MaterialID 11) with SetDirection 8, multimaterial mode only
NOT_A_PARTICLE_MATERIAL = 0      # MPMGPU_MAT_NONE: keeps the place of a contact law in the materials list (ContactLaw, MaterialID 60-63)
CONTACT_LAW, COULOMB_FRICTION_LAW = 60, 61        # Materials/ContactLaw.hpp:14 (ignore contact), Materials/CoulombFriction.hpp:14
PLANE_STRAIN_MPM, PLANE_STRESS_MPM, THREED_MPM = 10, 11, 12

DEFAULT_CV = 1.0e6          # MaterialBase.cpp:74 heatCapacity = Scaling(1.e6)


def xml_units(E=None, rho=None, alpha=None, Cv=None, G=None, K=None, yld=None, Ep=None):
    """Scale XML (Legacy) numbers to internal units.  Returns dict of the ones given."""
    out = {}
    if E is not None:
        out["E"] = E * 1.0e6            # Common/Materials/IsotropicMat.cpp:38-46
    if G is not None:
        out["G"] = G * 1.0e6
    if K is not None:
        out["K"] = K * 1.0e6
    if rho is not None:
        out["rho"] = rho * 1.0e-3       # MaterialBaseMPM.cpp (rho scaled 1e-3)
    if alpha is not None:
        out["alpha"] = alpha            # aI kept in ppm/K; 1e-6 applied in VerifyAndLoadProperties
    if Cv is not None:
        out["Cv"] = Cv * 1.0e6
    if yld is not None:
        out["yld"] = yld * 1.0e6
    if Ep is not None:
        out["Ep"] = Ep * 1.0e6
    return out


def _base(rho, Cv, pdamping, av=None, large_rotation=False):
    """av = (avA1, avA2) switches the artificial viscosity on (MaterialBaseMPM.cpp:202-216; defaults 0.2, 2.0).
    large_rotation = Elastic::useLargeRotation (<largeRotation>1</largeRotation>, Common/Materials/Elastic.cpp:27-33)."""
    p = np.zeros(NPARAMS)
    p[7] = 1.0 if large_rotation else 0.0
    p[0] = rho
    p[1] = Cv
    p[2] = -1.0 if pdamping is None else pdamping
    if av is not None:
        p[3], p[4], p[5] = 1.0, av[0], av[1]
    return p


def isotropic(E, nu, rho, aI=0.0, Cv=DEFAULT_CV, np_=THREED_MPM, pdamping=None, large_rotation=False):
    """IsotropicMat (MaterialID 1).  aI in ppm/K as in the XML <alpha>.

    Follows IsotropicMat::VerifyAndLoadProperties (Common/Materials/IsotropicMat.cpp:93-168) ->
    Elastic::SetAnalysisProps (Common/Materials/Elastic.cpp:194-329) ->
    Elastic::FillUnrotatedElasticProperties (Common/Materials/Elastic.cpp:43-140).
    """
    G = E / (2.0 * (1.0 + nu))
    alphaV = 3.0e-6 * aI
    Kbulk = E / (3.0 * (1 - 2 * nu))
    gamma0 = Kbulk * alphaV / (rho * Cv)
    e1 = e2 = e3 = E
    v12 = v13 = v23 = nu
    a1 = a2 = a3 = 1.0e-6 * aI
    v32 = v23 * e3 / e2
    v31 = v13 * e3 / e1
    v21 = v12 * e2 / e1
    p = _base(rho, Cv, pdamping, None, large_rotation)
    rrho = 1.0 / rho
    if np_ == THREED_MPM:
        xx = 1.0 - v13 * v31 - v23 * v32 - v12 * v21 - 2.0 * v13 * v32 * v21
        C11 = e1 * (1.0 - v23 * v32) / xx
        C12 = e2 * (v12 + v13 * v32) / xx
        C13 = e3 * (v13 + v12 * v23) / xx
        C22 = e2 * (1.0 - v13 * v31) / xx
        C23 = e3 * (v23 + v21 * v13) / xx
        C33 = e3 * (1.0 - v21 * v12) / xx
        C66 = C44 = C55 = G
        p[8:17] = [C11 * rrho, C12 * rrho, C13 * rrho, C22 * rrho, C23 * rrho, C33 * rrho,
                   C44 * rrho, C55 * rrho, C66 * rrho]
        p[17:20] = [a1, a2, a3]
    elif np_ == PLANE_STRAIN_MPM:
        xx = 1.0 - v13 * v31 - v23 * v32 - v12 * v21 - 2.0 * v12 * v23 * v31
        C11 = e1 * (1.0 - v23 * v32) / xx
        C12 = e2 * (v12 + v13 * v32) / xx
        C22 = e2 * (1.0 - v13 * v31) / xx
        C66 = G
        C13 = e3 * (v13 + v12 * v23) / xx
        C23 = e3 * (v23 + v21 * v13) / xx
        C33 = e3 * (1.0 - v21 * v12) / xx
        S13, S23, S33 = -v13 / e1, -v23 / e2, 1.0 / e3
        p[8], p[9], p[11], p[16] = C11 * rrho, C12 * rrho, C22 * rrho, C66 * rrho
        p[21], p[22], p[23] = C13 * rrho, C23 * rrho, C33 * rrho
        p[24] = S13 / S33
        p[17:20] = [a1 + v31 * a3, a2 + v32 * a3, a3]
    elif np_ == PLANE_STRESS_MPM:
        xx = 1.0 - v12 * v21
        C11 = e1 / xx
        C12 = e2 * v12 / xx
        C22 = e2 / xx
        C66 = G
        xx3 = 1.0 - v13 * v31 - v23 * v32 - v12 * v21 - 2.0 * v13 * v32 * v21
        C13 = -(v13 + v12 * v23) / (1.0 - v21 * v12)
        C23 = -(v23 + v21 * v13) / (1.0 - v21 * v12)
        C33 = e3 * (1.0 - v21 * v12) / xx3
        S13 = -v13 / e1
        p[8], p[9], p[11], p[16] = C11 * rrho, C12 * rrho, C22 * rrho, C66 * rrho
        p[21], p[22], p[23] = C13, C23, C33 * rrho
        p[24] = S13 * rho
        p[17:20] = [a1, a2, a3]
    else:
        raise ValueError("analysis type %r not supported" % np_)
    p[20] = gamma0
    return dict(kind=ISOTROPIC, n_history=0, p=p, rho=rho, wave_speed=float(np.sqrt(2.0 * G * (1.0 - nu) / (rho * (1.0 - 2.0 * nu)))),
                C33=C33, C66=C66)          # unreduced, as IsoPlasticity::VerifyAndLoadProperties reads them


def contact_law_placeholder():
    """The reference keeps contact laws in theMaterials[] beside the particle materials (Materials/ContactLaw.hpp); the entry
    keeps the material numbering of the particles intact.  No particle may use it."""
    return dict(kind=NOT_A_PARTICLE_MATERIAL, n_history=0, p=np.zeros(NPARAMS), rho=0.0, wave_speed=0.0)


def rigid_bc(direction_bits, mirrored=0, sets_temperature=False):
    """RigidMaterial as moving velocity BC (MaterialID 11, Materials/RigidMaterial.hpp).  mirrored = -1 / +1: the BC nodes
    reflect the velocity of the body at the minimum / maximum edge (NodalVelBC::SetMirroredVelBC).  sets_temperature
    (<SetTemperature/>): with conduction the nodes its particles touch are held at the particles' temperature."""
    p = _base(1.0, DEFAULT_CV, None)
    p[8] = float(direction_bits)
    p[9] = float(mirrored)
    p[10] = 1.0 if sets_temperature else 0.0
    return dict(kind=RIGIDBC, n_history=0, p=p, rho=1.0, wave_speed=0.0)


def rigid_contact():
    """RigidMaterial in contact mode (<SetDirection>8</SetDirection>, RIGID_MULTIMATERIAL_MODE): its particles keep their own
    velocity field, against which the other materials of a node make contact (multimaterial mode)."""
    return dict(kind=RIGIDCONTACT, n_history=0, p=_base(1.0, DEFAULT_CV, None), rho=1.0, wave_speed=0.0)


def neohookean(G, K, rho, aI=0.0, Cv=DEFAULT_CV, UofJOption=0, pdamping=None, av=None):
    """Neohookean (MaterialID 28): Neohookean::VerifyAndLoadProperties (Materials/Neohookean.cpp:88-141).
    History: J, Jres (both 1).  Particles start with elastic B = I in eplast (HyperElastic.cpp:51-60)."""
    Lame = K - 2.0 * G / 3.0
    p = _base(rho, Cv, pdamping, av)
    Gsp = G / rho
    Lamesp = Lame / rho
    Ksp = Lamesp + 2.0 * Gsp / 3.0
    gamma0 = K * (3.0e-6 * aI) / (rho * Cv)
    p[8], p[9], p[10], p[11], p[12], p[13] = Gsp, Ksp, Lamesp, float(UofJOption), 1.0e-6 * aI, gamma0
    return dict(kind=NEOHOOKEAN, n_history=2, p=p, rho=rho, wave_speed=float(np.sqrt((K + 4.0 * G / 3.0) / rho)),
                init_history=[1.0, 1.0], init_eplast=[1.0, 1.0, 1.0, 0.0, 0.0, 0.0])


def mooney(G1, G2, K, rho, aI=0.0, Cv=DEFAULT_CV, UofJOption=0, pdamping=None, av=None):
    """Mooney (MaterialID 8): Mooney::VerifyAndLoadProperties (Materials/Mooney.cpp:104-147) with G1, G2 and K given.
    History: J, Jres (both 1).  Particles start with elastic B = I in eplast (HyperElastic.cpp:51-60)."""
    p = _base(rho, Cv, pdamping, av)
    gamma0 = K * (3.0e-6 * aI) / (rho * Cv)
    p[8], p[9], p[10], p[11], p[12], p[13] = G1 / rho, G2 / rho, K / rho, float(UofJOption), 1.0e-6 * aI, gamma0
    return dict(kind=MOONEY, n_history=2, p=p, rho=rho, wave_speed=float(np.sqrt((K + 4.0 * (G1 + G2) / 3.0) / rho)),
                init_history=[1.0, 1.0], init_eplast=[1.0, 1.0, 1.0, 0.0, 0.0, 0.0])


HARD_LINEAR, HARD_NONLINEAR, HARD_JOHNSONCOOK, HARD_SCGL, HARD_NONLINEAR2 = 1, 2, 3, 4, 6        # MaterialBase::SetHardeningLaw ids


def isoplasticity(E, nu, rho, yld, Ep=None, Khard=0.0, aI=0.0, Cv=DEFAULT_CV, np_=THREED_MPM, pdamping=None, yld_min=0.0, av=None,
                  large_rotation=False, hardening=None):
    """IsoPlasticity (MaterialID 9): IsoPlasticity::VerifyAndLoadProperties (Materials/IsoPlasticity.cpp:50-70) with
    LinearHardening (LinearHardening.cpp:55-80; the default) or, through `hardening`,
      ("nonlinear", beta, n)   yield (1 + beta alpha)^n        NonlinearHardening.cpp:49-60
      ("nonlinear2", beta, n)  yield (1 + beta alpha^n)        Nonlinear2Hardening.cpp:28-39
      ("johnsoncook", dict(B=, n=, C=, ep0=, D=0, n2=1, Tm=, m=, Tref=))  JohnsonCook.cpp:106-127 (yld is A; B in the units of yld)
      ("scgl", dict(beta=, n=, yld_max=, GPp=, GTp=, Tref=))  SCGLHardening.cpp:72-91: min(yld (1 + beta alpha)^n, yld_max) times the
                   shear-modulus ratio 1 + GPp P + GTp (T - Tref), which also scales G (yld_max in the units of yld, GPp per unit of P)."""
    iso = isotropic(E, nu, rho, aI, Cv, np_, pdamping)
    p = _base(rho, Cv, pdamping, av, large_rotation)
    C66, C33 = iso["C66"], iso["C33"]
    G0red = C66 / rho
    Kred = C33 / rho - 4.0 * G0red / 3.0
    yldred = yld / rho
    if Ep is not None and Ep >= 0.0:
        beta = Ep / yld
    else:
        beta = Khard
    Epred = yldred * beta
    yldredMin = yld_min / rho
    alphaMax = 1.0e50 if beta >= 0.0 else ((yldredMin / yldred) - 1.0) / beta
    p[8], p[9], p[10], p[11] = G0red, Kred, yldred, Epred
    p[12] = iso["p"][19]
    p[13] = iso["p"][20]
    p[14], p[15] = alphaMax, yldredMin
    if hardening is not None:
        law = hardening[0]
        if law in ("nonlinear", "nonlinear2"):
            hb, hn = float(hardening[1]), float(hardening[2])
            p[16] = HARD_NONLINEAR if law == "nonlinear" else HARD_NONLINEAR2
            p[17], p[18] = hb, hn
            p[11] = 0.0
            p[14] = 1.0e50
            if hb < 0.0:
                p[14] = ((yldredMin / yldred) ** (1.0 / hn) - 1.0) / hb if law == "nonlinear" else ((yldredMin / yldred - 1.0) / hb) ** (1.0 / hn)
        elif law == "johnsoncook":
            j = hardening[1]
            C = float(j["C"])
            edot_min = float(np.exp(-0.5 / C)) if C != 0.0 else 1.0e-20
            edot_min = min(float(j["ep0"]), edot_min)
            p[16] = HARD_JOHNSONCOOK
            p[17], p[18], p[19], p[20], p[21], p[22] = j["B"] / rho, j["n"], C, j["ep0"], j.get("D", 0.0), j.get("n2", 1.0)
            p[23], p[24], p[25], p[26], p[27] = j["Tm"], j["m"], j.get("Tref", 0.0), edot_min, 1.0 + C * float(np.log(edot_min))
            p[11] = 0.0
            p[14] = 1.0e50
        elif law == "scgl":
            j = hardening[1]
            p[16] = HARD_SCGL
            p[17], p[18], p[19], p[20], p[21], p[25] = j["beta"], j["n"], j["yld_max"] / rho, j["GPp"] * rho, j["GTp"], j.get("Tref", 0.0)
            p[11] = 0.0
            p[14] = 1.0e50
        else:
            raise ValueError("hardening law %r" % (law,))
    return dict(kind=ISOPLASTICITY, n_history=1, p=p, rho=rho, wave_speed=iso["wave_speed"])
