"""Known answers that need neither the reference nor a GPU: the C restatement of the step
(oracle/mpm_oracle.c) must reproduce closed-form results on uniform deformation fields.  One USF step from the
undeformed state with a linear velocity field v = L x gives every interior particle the velocity gradient L exactly
(uGIMP reproduces linear fields away from free faces)."""
import numpy as np

from nairn_mpm_fea_b200 import materials as M
from nairn_mpm_fea_b200 import problem


def _oracle_engine(prob):
    from oracle.port import PortOracle
    return PortOracle(prob)


def _run(mat, L, nsteps=1, ncell=10, engine=_oracle_engine):
    c0 = 4.0 + ncell / 2.0

    def vel(pos):
        return L @ (pos - c0)

    prob = problem.block3d(ncell=ncell, margin=4, velocity_fn=vel, bottom_bc=False, method=problem.USF, material=mat)
    o = engine(prob)
    o.step(nsteps)
    got = o.download()
    o.close()
    pos = prob.particles["pos"]
    inner = np.all((pos > 4.0 + 3.0) & (pos < 4.0 + ncell - 3.0), axis=0)
    assert np.count_nonzero(inner) >= 64
    return prob, got, inner


def check_hookes_law(engine):
    u = M.xml_units(E=1000.0, rho=1.0)
    E, nu, rho = u["E"], 0.3, u["rho"]
    mat = M.isotropic(E, nu, rho)
    L = np.array([[40.0, 30.0, 0.0], [30.0, -10.0, 0.0], [0.0, 0.0, 25.0]])      # symmetric: no spin
    prob, got, inner = _run(mat, L, engine=engine)
    dt = prob.dt
    lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    eps = L * dt
    tr = np.trace(eps)
    want = {"xx": (lam * tr + 2 * mu * eps[0, 0]) / rho, "yy": (lam * tr + 2 * mu * eps[1, 1]) / rho,
            "zz": (lam * tr + 2 * mu * eps[2, 2]) / rho, "xy": 2 * mu * eps[0, 1] / rho}
    scale = abs(want["xx"])
    for i, k in ((0, "xx"), (1, "yy"), (2, "zz"), (5, "xy")):
        assert np.max(np.abs(got["sp"][i][inner] - want[k])) <= 1e-9 * scale, k
    assert np.max(np.abs(got["sp"][3:5][:, inner])) <= 1e-9 * scale
    # engineering shear strain and zero rotation
    assert np.max(np.abs(got["ep"][5][inner] - 2 * eps[0, 1])) <= 1e-9 * abs(eps[0, 1])
    assert np.max(np.abs(got["wrot"][:, inner])) <= 1e-9 * abs(eps[0, 1])
    # strain energy per unit mass = sigma:eps / (2 rho) by the midpoint rule from zero stress
    work = 0.5 * sum(want[k] * e for k, e in (("xx", eps[0, 0]), ("yy", eps[1, 1]), ("zz", eps[2, 2]))) + 0.5 * want["xy"] * 2 * eps[0, 1]
    assert np.max(np.abs(got["energies"][0][inner] - work)) <= 1e-8 * abs(work)


def check_neohookean_closed_form(engine):
    """Neohookean.cpp:246-300 with UJOption 0: P = -J (lambda/2)(J - 1/J) - G (tr(B)/3 - 1) (Kirchhoff, per rho0),
    s = G dev(B); one step of uniform dilation + shear from F = I."""
    u = M.xml_units(G=40.0, K=200.0, rho=1.0)
    G, K, rho = u["G"], u["K"], u["rho"]
    mat = M.neohookean(G, K, rho)
    L = np.array([[300.0, 200.0, 0.0], [0.0, -100.0, 0.0], [0.0, 0.0, 150.0]])
    prob, got, inner = _run(mat, L, engine=engine)
    F = np.eye(3) + L * prob.dt
    J = np.linalg.det(F)
    B = F @ F.T
    lame = K - 2.0 * G / 3.0
    P = -(J * 0.5 * (lame / rho) * (J - 1.0 / J) + (G / rho) * (np.trace(B) / 3.0 - 1.0))
    assert np.max(np.abs(got["pressure"][inner] - P)) <= 1e-9 * abs(P)
    dev = (G / rho) * (B - np.trace(B) / 3.0 * np.eye(3))
    scale = np.max(np.abs(dev))
    for i, (a, b) in enumerate(((0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1))):
        assert np.max(np.abs(got["sp"][i][inner] - dev[a, b])) <= 1e-9 * scale, (a, b)
    assert np.max(np.abs(got["history"][0][inner] - J)) <= 1e-12


def check_radial_return(engine):
    """Pure shear far beyond yield in one step: after the return |s| = sqrt(2/3) (yield + Ep alpha) with
    alpha = sqrt(2/3) lambda and lambda = (|s_trial| - sqrt(2/3) yield) / (2 (G + Ep/3))  (LinearHardening.cpp:124-145)."""
    u = M.xml_units(E=2000.0, rho=2.0, yld=20.0, Ep=100.0)
    E, nu, rho = u["E"], 0.33, u["rho"]
    mat = M.isoplasticity(E, nu, rho, u["yld"], u["Ep"])
    g = 9.0e4                                           # shear rate: about 15 yield strains in one step
    L = np.array([[0.0, g, 0.0], [g, 0.0, 0.0], [0.0, 0.0, 0.0]])
    prob, got, inner = _run(mat, L, engine=engine)
    Gred, yred, Epred = mat["p"][8], mat["p"][10], mat["p"][11]
    strial = np.sqrt(2.0) * Gred * (2 * g * prob.dt)     # |s| of a pure shear stress tau = G gamma: sqrt(2) tau
    lam = (strial - np.sqrt(2.0 / 3.0) * yred) / (2.0 * (Gred + Epred / 3.0))
    assert lam > 0
    alpha = np.sqrt(2.0 / 3.0) * lam
    smag = np.sqrt(2.0) * np.abs(got["sp"][5][inner])
    assert np.max(np.abs(smag - np.sqrt(2.0 / 3.0) * (yred + Epred * alpha))) <= 1e-9 * smag.max()
    assert np.max(np.abs(got["history"][0][inner] - alpha)) <= 1e-9 * alpha
    # plastic shear strain (engineering) = 2 lambda n_xy with n_xy = 1/sqrt(2)
    assert np.max(np.abs(got["eplast"][5][inner] - 2.0 * lam / np.sqrt(2.0))) <= 1e-9 * lam
    assert np.max(np.abs(got["pressure"][inner])) <= 1e-6 * yred


def test_hookes_law_and_rotation_free_shear():
    check_hookes_law(_oracle_engine)


def test_neohookean_pressure_and_deviatoric_stress_closed_form():
    check_neohookean_closed_form(_oracle_engine)


def test_radial_return_lands_on_the_hardened_yield_surface():
    check_radial_return(_oracle_engine)
