"""oracle (C++) against the independent numpy restatement of the mm10 / Voce crystal update
(tests/py_mm10.py): stress, hardening variable, tangent, plastic rotation, slip increments and local
Newton iteration counts of single points, through elastic, first-yield and developed-flow sweeps."""
import numpy as np
import pytest
import scipy.linalg

import py_mm10
from helpers import mm10_layout


@pytest.fixture(scope="module")
def Oracle(oracle_built):
    from oracle import Oracle
    return Oracle


def _kinematics(Fn, Fn1):
    """drive_eps_sig.f:203-265 with scipy's polar decomposition"""
    Fn, Fn1 = Fn.reshape(3, 3), Fn1.reshape(3, 3)
    Fh, dF = 0.5 * (Fn + Fn1), Fn1 - Fn
    Rh = scipy.linalg.polar(Fh)[0]
    R = scipy.linalg.polar(Fn1)[0]
    Lm = dF @ np.linalg.inv(Fh)
    Dm = 0.5 * (Lm + Lm.T)
    d = Rh.T @ Dm @ Rh
    return R, py_mm10.sym6(d, True)


@pytest.mark.parametrize("slip_type,iD_v", [(1, 0.0), (8, 0.0), (1, 2.0e-7), (2, 0.0), (6, 0.0), (7, 0.0)])
def test_oracle_matches_numpy_restatement(Oracle, slip_type, iD_v):
    from cpfft_b200.polycrystal import polycrystal
    p = polycrystal(2, ngrains=4, slip_type=slip_type)
    c = p.crystals[0]
    c.iD_v = iD_v
    o = Oracle(p)
    b, n = Oracle.slip_table(slip_type)
    C6 = py_mm10.stiffness_isotropic(c.e, c.nu)
    nslip = len(b)
    L = mm10_layout(nslip)
    rng = np.random.default_rng(12)
    N3 = p.N3
    I = np.zeros((9, N3)); I[[0, 4, 8]] = 1.0
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = -0.35; bar[8] = -0.6; bar[1] = 0.25    # well separated stretches
    G = rng.standard_normal((9, N3))
    crys = [py_mm10.Crystal(b, n, C6, p.angles[v], c.harden_n, c.theta_0, c.tau_y, c.tau_v, c.voche_m, iD_v)
            for v in range(N3)]
    # python-side state per voxel
    state = [dict(sn=np.zeros(6), tt=c.tau_y + 1.0e-5, ttrate=0.0, Dn=np.zeros(6), Rp=np.eye(3)) for _ in range(N3)]
    o.drive_eps_sig(1, 0)
    Fn = I.copy()
    plastic_iters = 0
    for step in (1, 2, 3):
        for it, frac in ((0, 0.9), (1, 1.0)):
            F1 = I + 0.004 * (step - 1 + frac) * (bar + 0.15 * G)
            o.Fn1[:] = F1
            nfail = o.drive_eps_sig(step, it)
            flags = np.ctypeslib.as_array(o.L.orc_fail_flags(o.h), shape=(o.N3,)).copy()
            assert flags.sum() == nfail
            for v in range(N3):
                # R and the unrotated strain increment as the oracle computed them (its closed-form polar
                # decomposition is pinned separately and has a 1e-8 noise floor at small strains); the
                # scipy-based kinematics must agree with them to that floor
                h = o.hist_n1[v]
                R = h[L["R"][0]:L["R"][1]].reshape(3, 3).T
                d = h[L["D"][0]:L["D"][1]].copy()
                R2, d2 = _kinematics(Fn[:, v], F1[:, v])
                assert np.abs(R - R2).max() <= 5e-8 and np.abs(d - d2).max() <= 5e-8 * max(np.abs(d).max(), 1e-12) + 1e-10
                st = state[v]
                r = py_mm10.update(crys[v], R, d, p.tstep, st["sn"], st["tt"], st["ttrate"], st["Dn"], st["Rp"], it == 0)
                scale = max(np.abs(h[L["stress"][0]:L["stress"][1]]).max(), 1.0)
                assert np.abs(r["stress"] - h[L["stress"][0]:L["stress"][1]]).max() <= 2e-9 * scale, (step, it, v)
                assert abs(r["tt"] - h[L["tau_tilde"][0]]) <= 1e-9 * h[L["tau_tilde"][0]]
                T = h[0:36].reshape(6, 6).T                                   # column-major in the history
                assert np.abs(r["tangent"] - T).max() <= 1e-8 * np.abs(T).max(), (step, it, v)
                assert tuple(r["iters"]) == tuple(o.local_iters[v]), (step, it, v, r["iters"], o.local_iters[v])
                assert bool(r["fail"]) == bool(flags[v]), (step, it, v)      # sub-stepping exhausted: same points
                if it > 0 and not r["fail"]:
                    Rp = h[L["Rp"][0]:L["Rp"][1]].reshape(3, 3).T
                    assert np.abs(r["Rp"] - Rp).max() <= 1e-10
                    sl = h[L["slipinc"][0]:L["slipinc"][0] + nslip]
                    assert np.abs(r["slip"] - sl).max() <= 1e-9 * max(np.abs(sl).max(), 1e-12)
                    plastic_iters += r["iters"][1]
                st["last"] = (r, d)
        for v in range(N3):                                                    # commit: n <- n+1
            r, d = state[v]["last"]
            state[v].update(sn=r["stress"].copy(), tt=r["tt"], ttrate=r.get("tt_rate", 0.0), Dn=d.copy(), Rp=r["Rp"])
        Fn = F1.copy()
        o.Fn[:] = F1
        o.update()
    assert plastic_iters > 0


def test_oracle_mts_matches_numpy_restatement(Oracle):
    """the MTS law (thermally activated thresholds, mu(T), n-state flags, tangent terms JA / JB)"""
    from cpfft_b200.polycrystal import polycrystal
    from test_oracle_mts import mts_crystal
    p = polycrystal(2, ngrains=4)
    c = mts_crystal()
    p.crystals = [c]
    o = Oracle(p)
    b, n = Oracle.slip_table(1)
    C6 = py_mm10.stiffness_isotropic(c.e, c.nu)
    L = mm10_layout(12)
    N3 = p.N3
    rng = np.random.default_rng(21)
    I = np.zeros((9, N3)); I[[0, 4, 8]] = 1.0
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = -0.35; bar[8] = -0.6; bar[1] = 0.25
    G = rng.standard_normal((9, N3))
    crys = []
    for v in range(N3):
        cr = py_mm10.Crystal(b, n, C6, p.angles[v], c.harden_n, c.theta_0, 0.0, 0.0, c.voche_m, 0.0)
        cr.mts = dict(tau_a=c.tau_a, tau_hat_y=c.tau_hat_y, G_0_y=c.g_0_y, tau_hat_v=c.tau_hat_v, G_0_v=c.g_0_v,
                      p_y=c.p_y, q_y=c.q_y, p_v=c.p_v, q_v=c.q_v, boltz=c.boltzman, b=c.burgers,
                      eps_dot_0_y=c.eps_dot_0_y, eps_dot_0_v=c.eps_dot_0_v, mu_0=c.mu_0, D_0=c.D_0, T_0=c.T_0)
        crys.append(cr)
    state = [dict(sn=np.zeros(6), tt=-1.0, u1=-1.0, u2=-1.0, ttrate=0.0, Dn=np.zeros(6), Rp=np.eye(3)) for _ in range(N3)]
    o.drive_eps_sig(1, 0)
    u0 = L["u"][0]
    plastic = 0
    for step in (1, 2, 3):
        for it, frac in ((0, 0.9), (1, 1.0)):
            F1 = I + 0.003 * (step - 1 + frac) * (bar + 0.15 * G)
            o.Fn1[:] = F1
            nfail = o.drive_eps_sig(step, it)
            flags = np.ctypeslib.as_array(o.L.orc_fail_flags(o.h), shape=(o.N3,)).copy()
            for v in range(N3):
                h = o.hist_n1[v]
                R = h[L["R"][0]:L["R"][1]].reshape(3, 3).T
                d = h[L["D"][0]:L["D"][1]].copy()
                st = state[v]
                r = py_mm10.update(crys[v], R, d, p.tstep, st["sn"], st["tt"], st["ttrate"], st["Dn"], st["Rp"], it == 0,
                                   st["u1"], st["u2"])
                scale = max(np.abs(h[L["stress"][0]:L["stress"][1]]).max(), 1.0)
                assert np.abs(r["stress"] - h[L["stress"][0]:L["stress"][1]]).max() <= 2e-9 * scale, (step, it, v)
                assert abs(r["tt"] - h[L["tau_tilde"][0]]) <= 1e-9 * abs(h[L["tau_tilde"][0]]), (step, it, v)
                assert abs(r["u1"] - h[u0]) <= 1e-11 * abs(h[u0]) and abs(r["u2"] - h[u0 + 1]) <= 1e-11 * abs(h[u0 + 1])
                T = h[0:36].reshape(6, 6).T
                assert np.abs(r["tangent"] - T).max() <= 1e-8 * np.abs(T).max(), (step, it, v)
                assert tuple(r["iters"]) == tuple(o.local_iters[v]), (step, it, v, r["iters"], o.local_iters[v])
                assert bool(r["fail"]) == bool(flags[v])
                if it > 0 and not r["fail"]:
                    Rp = h[L["Rp"][0]:L["Rp"][1]].reshape(3, 3).T
                    assert np.abs(r["Rp"] - Rp).max() <= 1e-10
                    plastic += r["iters"][1]
                st["last"] = (r, d)
        for v in range(N3):
            r, d = state[v]["last"]
            state[v].update(sn=r["stress"].copy(), tt=r["tt"], u1=r["u1"], u2=r["u2"], ttrate=r.get("tt_rate", 0.0), Dn=d.copy(),
                            Rp=r["Rp"])
        o.Fn[:] = F1
        o.update()
    assert plastic > 0
