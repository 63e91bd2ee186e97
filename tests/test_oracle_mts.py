"""MTS hardening (`hardening mts`, h_type 2: mm10_setup_mts mm10_a.f:2109-2175, mm10_h/estress/
ehard/ed/dgdt/dgdh/dgdd_mts mm10_b.f:2080-2345, tangent terms JA / JB mm10_a.f:740-810) in the
oracle.  The reference has no test for it; pinned here by
  * the analytic Jacobian of the local Newton system against central differences of the residual
    (ties estress / ehard / dgdt / dgdh to h and the slip law, for Voce as well),
  * the degenerate case boltz = 0, D_0 = 0, in which MTS is algebraically the Voce law with
    tau_y = tau_a + tau_hat_y, tau_v = tau_hat_v (the Voce path is cross-checked elsewhere),
  * the stored tangent against finite differences of the stress."""
import copy

import numpy as np
import pytest

from helpers import relerr, mm10_layout


def mts_crystal(**kw):
    from cpfft_b200.problem import Crystal
    c = Crystal(slip_type=1, elastic_type=1, h_type=2, e=200000.0, nu=0.3, mu=200000.0 / 2.6, harden_n=20.0,
                theta_0=1500.0, voche_m=1.0, tau_a=20.0, tau_hat_y=180.0, g_0_y=0.4, tau_hat_v=300.0, g_0_v=1.2,
                boltzman=1.3806e-20, burgers=2.5e-7, eps_dot_0_y=1.0e10, eps_dot_0_v=1.0e10,
                mu_0=80000.0, D_0=3000.0, T_0=200.0)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.fixture(scope="module")
def Oracle(oracle_built):
    from oracle import Oracle
    return Oracle


@pytest.mark.parametrize("law", ["voce", "voce_bcc48", "voce_m2_n7p5", "voce_alter_mode", "mts", "mts_m2", "mts_bcc48"])
def test_local_jacobian_matches_central_differences(Oracle, law):
    from cpfft_b200.problem import Crystal
    rng = np.random.default_rng(5)
    if law.startswith("mts"):
        c = mts_crystal()
        if law == "mts_m2":
            c.voche_m = 2.0
        if law == "mts_bcc48":
            c.slip_type = 8
    else:
        c = Crystal(slip_type=1, e=200000.0, nu=0.3, mu=200000.0 / 2.6, harden_n=20.0, theta_0=100.0,
                    tau_v=100.0, tau_y=100.0, iD_v=1.0e-7)
        if law == "voce_bcc48":
            c.slip_type = 8
        if law == "voce_m2_n7p5":
            c.voche_m = 2.0; c.harden_n = 7.5
        if law == "voce_alter_mode":
            c.alter_mode = 1; c.eps_dot_0_y = 2.0e-3
    ang = (25.0, 40.0, 70.0)
    D = 2.0e-3 * np.array([1.0, -0.4, -0.5, 0.3, -0.2, 0.1])
    sn = 60.0 * rng.standard_normal(6)
    ttn = 230.0 if law.startswith("mts") else 115.0
    x = np.concatenate([sn + 25.0 * rng.standard_normal(6), [ttn + 4.0]])
    R, J = Oracle.mm10_residual_jacobian(c, ang, D, 1.0, x, sn, ttn)
    assert np.all(np.isfinite(R)) and np.all(np.isfinite(J))
    fd = np.zeros((7, 7))
    for j in range(7):
        h = 1e-5 * max(1.0, abs(x[j]))
        xp, xm = x.copy(), x.copy()
        xp[j] += h; xm[j] -= h
        fd[:, j] = (Oracle.mm10_residual_jacobian(c, ang, D, 1.0, xp, sn, ttn)[0] -
                    Oracle.mm10_residual_jacobian(c, ang, D, 1.0, xm, sn, ttn)[0]) / (2 * h)
    scale = np.abs(fd).max(axis=1, keepdims=True)
    diff = np.abs(J - fd)
    if law.startswith("voce"):
        # reference quirk, reproduced on purpose: mm10_ehard_voche (mm10_b.f:1962-1975) multiplies the
        # d|slip|/d(tau_tilde) term by sign(slipinc), so J22 is only exact while no system slips
        # backwards; every other entry is the exact derivative
        assert 0 < diff[6, 6] <= 1e-2
        diff[6, 6] = 0.0
    assert diff.max() <= 2e-6 * np.abs(fd).max(), diff.max() / np.abs(fd).max()
    assert (diff / scale).max() <= 2e-5          # row by row (the hardening row is much smaller)


def _drive(o, F_path, after_first_commit=None):
    o.drive_eps_sig(1, 0)
    for step, F in enumerate(F_path, start=1):
        for it in (0, 1):
            o.Fn1[:] = F
            assert o.drive_eps_sig(step, it) == 0
        o.Fn[:] = o.Fn1
        o.update()
        if step == 1 and after_first_commit:
            after_first_commit(o)


def test_degenerate_mts_is_the_voce_law(Oracle):
    """boltz = 0 (no thermal activation: tau_y = tau_hat_y, tau_v = tau_hat_v) and D_0 = 0
    (mu = mu_0): h_mts = tt_n + theta_0 sum (1 - (tt - tau_a - tau_hat_y)/tau_hat_v) |slip|, the Voce
    law with tau_y = tau_a + tau_hat_y.  The two laws start from different hardening values
    (tau_a + tau_y + 0.1 vs tau_y + 1e-5), so step 1 is kept elastic (slip ~ 1e-30) and the Voce
    run is given the MTS value of tau_tilde at the first commit.  Same equations, but not the same
    Newton iterates (the Voce J22 carries the reference's sign(slipinc) quirk, see above): both
    local solves are converged to round-off (atol 1e-9, rtol 1e-13) and the solutions compared."""
    from cpfft_b200.polycrystal import polycrystal
    pm = polycrystal(3, ngrains=6)
    pv = polycrystal(3, ngrains=6)
    tight = dict(atol=1e-9, atol1=1e-9, rtol=1e-13, rtol1=1e-13)
    pm.crystals = [mts_crystal(boltzman=0.0, D_0=0.0, tau_a=20.0, tau_hat_y=90.0, tau_hat_v=140.0, theta_0=100.0, **tight)]
    cv = copy.copy(pv.crystals[0])
    cv.theta_0 = 100.0; cv.tau_y = 110.0; cv.tau_v = 140.0; cv.voche_m = 1.0; cv.harden_n = 20.0
    for k, v in tight.items():
        setattr(cv, k, v)
    pv.crystals = [cv]
    om, ov = Oracle(pm), Oracle(pv)
    L = mm10_layout(12)
    rng = np.random.default_rng(9)
    G = rng.standard_normal((9, pm.N3))
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
    I = np.zeros((9, pm.N3)); I[[0, 4, 8]] = 1.0
    path = [I + 1e-5 * (bar + 0.3 * G)] + [I + 0.0025 * s * (bar + 0.3 * G) for s in (1, 2, 3)]
    tt_mts = {}
    _drive(om, path, lambda o: tt_mts.setdefault("tt", o.hist_n[:, L["tau_tilde"][0]].copy()))
    assert np.allclose(tt_mts["tt"], 20.0 + 90.0 + 0.1, rtol=0, atol=1e-9)     # tau_a + tau_y + init_hard
    def give_mts_value(o):
        o.hist_n[:, L["tau_tilde"][0]] = tt_mts["tt"]
    _drive(ov, path, give_mts_value)
    assert om.local_iters.sum() > 0 and ov.local_iters.sum() > 0
    assert relerr(om.urcs_n[:, :6], ov.urcs_n[:, :6]) <= 1e-10
    for name in ("stress", "Rp", "tau_tilde", "slipinc", "eps", "ep"):
        a, b = L[name]
        assert relerr(om.hist_n[:, a:b], ov.hist_n[:, a:b]) <= 1e-9, name
    # MTS keeps tau_y and mu_harden of the step in u(1:2)
    u0 = L["u"][0]
    assert np.allclose(om.hist_n[:, u0], 90.0) and np.allclose(om.hist_n[:, u0 + 1], pm.crystals[0].mu_0)


def test_mts_thermal_activation_lowers_the_thresholds(Oracle):
    """tau_y(T, rate) < tau_hat_y, mu(T) < mu_0 at 297 K (mm10_setup_mts), and the flow stress follows"""
    from cpfft_b200.polycrystal import polycrystal
    L = mm10_layout(12)
    res = {}
    for tag, kw in (("athermal", dict(boltzman=0.0, D_0=0.0)), ("thermal", {})):
        p = polycrystal(3, ngrains=6)
        p.crystals = [mts_crystal(**kw)]
        o = Oracle(p)
        I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
        bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
        _drive(o, [I + 0.003 * s * bar for s in (1, 2, 3)])
        u0 = L["u"][0]
        res[tag] = (o.hist_n[:, u0].mean(), o.hist_n[:, u0 + 1].mean(), np.abs(o.urcs_n[:, 0]).mean())
    assert res["thermal"][0] < 0.9 * res["athermal"][0] and res["thermal"][0] > 0
    assert res["thermal"][1] < res["athermal"][1]
    assert res["thermal"][2] < res["athermal"][2]


def test_mts_tangent_close_to_fd(Oracle):
    """the stored tangent (lagged Jacobian, JA from dgamma/dD, JB from ed, symmetrised) must be a
    good Newton tangent: A = dP/dF within a few per cent, as for Voce"""
    from test_oracle_material import _fd_tangent
    from cpfft_b200.polycrystal import polycrystal
    p = polycrystal(2, ngrains=3)
    p.crystals = [mts_crystal()]
    o = Oracle(p)
    rng = np.random.default_rng(4)
    Fn = np.eye(3).ravel()
    Fn1 = Fn + 4e-3 * rng.standard_normal(9)
    _, A = o.point_update(3, 1, 1, Fn, Fn1)
    fd = _fd_tangent(o, 3, 1, 1, Fn, Fn1)
    assert np.abs(A.reshape(9, 9) - fd).max() / np.abs(fd).max() <= 0.03
