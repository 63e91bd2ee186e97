"""A second, independent restatement of the reference's crystal-plasticity update for ONE crystal with
Voce hardening, in plain numpy, written routine by routine from the Fortran (SURVEY.md 8c item 6,
"oracle vs oracle").  TEST INFRASTRUCTURE ONLY.  It shares no code with oracle/ (C++) or with the CUDA
sources: the stiffness and Schmid tensors are built with numpy linear algebra, the Jacobian is
assembled with outer products, the linear solves are numpy.linalg.solve.

  setup_mm10_rknstr      drive_eps_sig.f:537-1002   -> crystal_setup
  mm10_setup(+_voche)    mm10_a.f:830-962, 2057-2075 -> step_setup
  mm10_formR1 / R2       mm10_b.f:1065-1101, 63-113  -> residual
  mm10_formJ11/12/21/22  mm10_b.f:177-481            -> jacobian
  mm10_solve             mm10_a.f:2860-3295          -> solve
  mm10_solve_strup       mm10_a.f:2628-2845          -> update
  mm10_tangent           mm10_a.f:658-815            -> tangent (lagged Jacobian, symmetrised)
  mm10_update_rotation   mm10_a.f:3310-3414          -> Rp
  MTS hardening          mm10_a.f:2090-2175 (init / setup), mm10_b.f:2080-2345 (h, estress, ehard, ed,
                         dgdt, dgdh, dgdd), tangent terms JA / JB mm10_a.f:740-810 -> `mts=` branches
"""
import numpy as np

VOIGT = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]      # xx yy zz xy yz xz


def kocks(ang_deg):                                            # mm10_a.f:1287-1345
    psi, th, phi = np.deg2rad(ang_deg)
    return np.array([
        [-np.sin(psi) * np.sin(phi) - np.cos(psi) * np.cos(phi) * np.cos(th),
         np.cos(psi) * np.sin(phi) - np.sin(psi) * np.cos(phi) * np.cos(th), np.cos(phi) * np.sin(th)],
        [np.sin(psi) * np.cos(phi) - np.cos(psi) * np.sin(phi) * np.cos(th),
         -np.cos(psi) * np.cos(phi) - np.sin(psi) * np.sin(phi) * np.cos(th), np.sin(phi) * np.sin(th)],
        [np.cos(psi) * np.sin(th), np.sin(psi) * np.sin(th), np.cos(th)]])


def sym6(T, eng):
    """symmetric tensor -> Voigt 6-vector (engineering shear when `eng`)"""
    f = 2.0 if eng else 1.0
    return np.array([T[0, 0], T[1, 1], T[2, 2], f * T[0, 1], f * T[1, 2], f * T[0, 2]])


def ten(v, eng):
    f = 0.5 if eng else 1.0
    return np.array([[v[0], f * v[3], f * v[5]], [f * v[3], v[1], f * v[4]], [f * v[5], f * v[4], v[2]]])


def skew3(W):
    """wv = (w23, w13, w12) (mm10_a.f:1529-1531)"""
    return np.array([W[1, 2], W[0, 2], W[0, 1]])


def skewt(w):
    return np.array([[0, w[2], w[1]], [-w[2], 0, w[0]], [-w[1], -w[0], 0.0]])


def stiffness_isotropic(e, nu):
    """engineering-shear 6x6 stiffness = inverse of the compliance (mod_crystals.f:1793-1931)"""
    S = np.zeros((6, 6))
    S[:3, :3] = -nu / e
    S[np.arange(3), np.arange(3)] = 1.0 / e
    S[np.arange(3, 6), np.arange(3, 6)] = 2.0 * (1.0 + nu) / e
    return np.linalg.inv(S)


def rot6_stress(Q):
    """6x6 operator of T -> Q T Q^T on stress-type Voigt vectors (mm10_RT2RVE, mm10_a.f:1400-1447)"""
    M = np.zeros((6, 6))
    for J, (k, l) in enumerate(VOIGT):
        E = np.zeros((3, 3)); E[k, l] = E[l, k] = 1.0
        M[:, J] = sym6(Q @ E @ Q.T, False)
    return M


class Crystal:
    def __init__(self, b, n, C6, angles, rate_n, theta_0, tau_y, tau_v, voche_m=1.0, iD_v=0.0, miter=30,
                 atol=1e-5, atol1=1e-5, rtol=5e-5, rtol1=1e-5):
        g = kocks(angles)
        self.g = g
        bs, ns = b @ g, n @ g                                   # g^T applied to the crystal-frame vectors
        self.M0 = [0.5 * (np.outer(x, y) + np.outer(y, x)) for x, y in zip(bs, ns)]     # sym Schmid tensors
        self.W0 = [0.5 * (np.outer(x, y) - np.outer(y, x)) for x, y in zip(bs, ns)]     # skew parts
        R6 = rot6_stress(g.T)
        self.C = R6 @ C6 @ R6.T                                 # stiffness in the RVE frame
        self.n, self.theta_0, self.tau_y, self.tau_v, self.m, self.iD_v = rate_n, theta_0, tau_y, tau_v, voche_m, iD_v
        self.miter, self.atol, self.atol1, self.rtol, self.rtol1 = miter, atol, atol1, rtol, rtol1
        self.mts = None        # dict of the MTS parameters when `hardening mts`


def rvw(rt):
    """mm10_rt2rvw (mm10_a.f:1461-1479): rotates the (w23, w13, w12) skew vectors"""
    return np.array([
        [rt[1, 1] * rt[2, 2] - rt[1, 2] * rt[2, 1], rt[1, 0] * rt[2, 2] - rt[1, 2] * rt[2, 0], rt[1, 0] * rt[2, 1] - rt[1, 1] * rt[2, 0]],
        [rt[0, 1] * rt[2, 2] - rt[0, 2] * rt[2, 1], rt[0, 0] * rt[2, 2] - rt[0, 2] * rt[2, 0], rt[0, 0] * rt[2, 1] - rt[0, 1] * rt[2, 0]],
        [rt[0, 1] * rt[1, 2] - rt[0, 2] * rt[1, 1], rt[0, 0] * rt[1, 2] - rt[0, 2] * rt[1, 0], rt[0, 0] * rt[1, 1] - rt[0, 1] * rt[1, 0]]])


class Step:
    """mm10_setup (mm10_a.f:830-962): current Schmid vectors, dg, tinc.  The reference rotates the
    engineering-shear vectors ms0 with mm10_RT2RVE, which is the tensor-component (stress-type)
    operator (mm10_a.f:1400-1447): reproduced as such."""
    def __init__(self, cr, R, D, dt, Rpn):
        Q = Rpn.T
        RE, RW, RWC = rot6_stress(Q), rvw(Q), rvw(R @ Q)
        self.ms = [RE @ sym6(M, True) for M in cr.M0]
        self.qs = [RW @ skew3(W) for W in cr.W0]
        self.qc = [RWC @ skew3(W) for W in cr.W0]
        self.D, self.tinc = np.asarray(D, float), dt
        self.dg = np.sqrt(2.0 / 3.0 * (D[:3] @ D[:3] + 0.5 * (D[3:] @ D[3:])))

    def setup_mts(self, cr, temp, n):
        """mm10_setup_mts (mm10_a.f:2109-2175); n = dict(tt, u1, u2) of the n state (modified: flags < 0
        are replaced by the values of this step)"""
        m = cr.mts
        dgc = self.dg / self.tinc
        self.temp = temp
        self.mu = m["mu_0"] if temp == 0.0 else m["mu_0"] - m["D_0"] / (np.exp(m["T_0"] / temp) - 1.0)
        if dgc == 0.0:
            self.tau_v, self.tau_y = m["tau_hat_v"], m["tau_hat_y"]
        else:
            kT = m["boltz"] * temp / (self.mu * m["b"] ** 3)
            self.tau_v = m["tau_hat_v"] * (1.0 - (kT / m["G_0_v"] * np.log(m["eps_dot_0_v"] / dgc)) ** (1.0 / m["q_v"])) ** (1.0 / m["p_v"])
            self.tau_y = m["tau_hat_y"] * (1.0 - (kT / m["G_0_y"] * np.log(m["eps_dot_0_y"] / dgc)) ** (1.0 / m["q_y"])) ** (1.0 / m["p_y"])
        self.tau_y_n = self.tau_y if n["u1"] < 0.0 else n["u1"]
        self.mu_n = self.mu if n["u2"] < 0.0 else n["u2"]
        if n["tt"] < 0.0:
            n["tt"] = m["tau_a"] + (self.mu / m["mu_0"]) * self.tau_y + 0.1


def symsw(s, w):
    """mm10_symSW (mm10_b.f:1505-1524)"""
    return np.array([s[3] * w[2] - s[5] * w[1],
                     s[3] * w[2] - s[4] * w[0],
                     s[5] * w[1] + s[4] * w[0],
                     0.5 * (w[2] * (s[0] - s[1]) + w[0] * s[5] - w[1] * s[4]),
                     0.5 * (w[0] * (s[1] - s[2]) + w[1] * s[3] + w[2] * s[5]),
                     0.5 * (w[1] * (s[0] - s[2]) + w[0] * s[3] - w[2] * s[4])])


def slips(cr, st, sig, tt):
    rs = np.array([sig @ m for m in st.ms])
    return rs, st.dg / tt * np.abs(rs / tt) ** (cr.n - 1.0) * rs


def residual(cr, st, sn, ttn, x, both=True):
    sig, tt = x[:6], x[6]
    rs, gam = slips(cr, st, sig, tt)
    f = gam + rs * st.tinc * cr.iD_v
    dbarp = sum(fi * m for fi, m in zip(f, st.ms))
    wp = sum(fi * q for fi, q in zip(f, st.qc))
    R1 = sig - sn - cr.C @ (st.D - dbarp) + 2.0 * symsw(sig, wp)
    if not both:
        return np.concatenate([R1, [0.0]]), ttn
    if cr.mts:                                                  # mm10_h_mts (mm10_b.f:2080-2111), tau_l = 0
        m = cr.mts
        cta = (m["mu_0"] / st.mu) * tt - (m["mu_0"] / st.mu) * m["tau_a"] - st.tau_y
        ct = 1.0 - cta / st.tau_v
        h = (m["tau_a"] * (1.0 - st.mu / st.mu_n) + (st.mu / m["mu_0"]) * (st.tau_y - st.tau_y_n) + (st.mu / st.mu_n) * ttn
             + cr.theta_0 * (st.mu / m["mu_0"]) * np.sum(ct ** cr.m * np.abs(gam)))
        return np.concatenate([R1, [tt - h]]), h
    hterm = 1.0 - (tt - cr.tau_y) / cr.tau_v
    h = ttn + cr.theta_0 * np.sum(np.abs(hterm) ** cr.m * np.sign(hterm) * np.abs(gam))
    return np.concatenate([R1, [tt - h]]), h


def jacobian(cr, st, x):
    sig, tt = x[:6], x[6]
    rs, gam = slips(cr, st, sig, tt)
    dgdt = st.dg * cr.n / tt ** cr.n * np.abs(rs) ** (cr.n - 1.0) + st.tinc * cr.iD_v
    J = np.zeros((7, 7))
    wvec = [cr.C @ m + 2.0 * symsw(sig, q) for m, q in zip(st.ms, st.qc)]
    for w, m, d, g in zip(wvec, st.ms, dgdt, gam):
        J[:6, :6] += d * np.outer(w, m)
        J[:6, 6] += w * (-cr.n / tt * g)
    f = gam + rs * st.tinc * cr.iD_v
    wp = sum(fi * q for fi, q in zip(f, st.qc))
    IW = np.zeros((6, 6))                                       # d(2 symSW(sig, wp))/d sig  (mm10_iw, mm10_b.f:1605-1634)
    for j in range(6):
        e = np.zeros(6); e[j] = 1.0
        IW[:, j] = 2.0 * symsw(e, wp)
    J[:6, :6] += IW + np.eye(6)
    if cr.mts:                                                  # mm10_estress_mts / mm10_ehard_mts (mm10_b.f:2114-2186)
        m = cr.mts
        if not hasattr(st, "mu"):
            return J
        cta = (m["mu_0"] / st.mu) * tt - (m["mu_0"] / st.mu) * m["tau_a"] - st.tau_y
        ct = 1.0 - cta / st.tau_v
        ur = st.mu / m["mu_0"]
        et = sum(ct ** cr.m * np.abs(r) ** (cr.n - 2.0) * r * mm for r, mm in zip(rs, st.ms))
        J[6, :6] = -(cr.theta_0 * ur * et * cr.n * st.dg / tt ** cr.n)
        etau = np.sum((cr.m * (1.0 / st.tau_v) / ct + ur * cr.n / tt) * ct ** cr.m * np.abs(gam))
        J[6, 6] = 1.0 + cr.theta_0 * etau
        return J
    hterm = 1.0 - (tt - cr.tau_y) / cr.tau_v
    hp = np.abs(hterm) ** cr.m
    et = sum(hp * np.sign(hterm) * np.abs(r) ** (cr.n - 2.0) * r * m for r, m in zip(rs, st.ms))
    J[6, :6] = -(cr.theta_0 * st.dg * cr.n / tt ** cr.n * et)
    # reference quirk kept: sign(slipinc) in mm10_ehard_voche (mm10_b.f:1962-1975)
    etau = np.sum((cr.m * (-1.0 / cr.tau_v) * np.abs(gam) / np.abs(hterm)
                   - np.abs(gam) * cr.n / tt * np.sign(hterm) * np.where(gam >= 0, 1.0, -1.0)) * hp)
    J[6, 6] = 1.0 - cr.theta_0 * etau
    return J


def newton(cr, st, sn, ttn, x, nun, atol, rtol, mmin, inR_fallback=None):
    """one phase of mm10_solve: nun unknowns (6: stress predictor at fixed x[6]; 7: coupled update).
    Armijo halving line search c = 1e-4, <= 10 halvings.  Returns x, iterations, fail, J_last, h, inR."""
    both = (nun == 7)
    R, h = residual(cr, st, sn, ttn, x, both)
    nR = np.linalg.norm(R[:nun]); inR = nR
    if both and inR == 0.0:
        inR = inR_fallback
    it, J = 0, None
    while ((nR > atol) and (nR / inR > rtol)) or (it < mmin):
        J = jacobian(cr, st, x)
        if not both:
            J[:6, 6] = 0.0; J[6, :6] = 0.0; J[6, 6] = 1.0
        dx = np.zeros(7)
        dx[:nun] = np.linalg.solve(-J[:nun, :nun], R[:nun])
        ls1 = 0.5 * (R[:nun] @ R[:nun])
        ls2 = 1.0e-4 * (dx[:nun] @ (J[:nun, :nun].T @ R[:nun]))
        alpha, ls = 1.0, 0
        while True:
            xn = x + alpha * dx
            R, h = residual(cr, st, sn, ttn, xn, both)
            if 0.5 * (R[:nun] @ R[:nun]) <= ls1 + ls2 * alpha or ls > 10:
                x = xn
                break
            alpha *= 0.5; ls += 1
        nR = np.linalg.norm(R[:nun])
        it += 1
        if it > cr.miter or np.any(np.isnan(x)):
            return x, it, True, J, h, inR
    return x, it, False, J, h, inR


def update(cr, R, D, dt, sn, ttn, ttrate_n, Dn, Rpn, iter0, u1n=-1.0, u2n=-1.0):
    """mm10_solve_strup + mm10_tangent + mm10_update_rotation for one crystal.
    Returns dict(stress, tt, tangent, Rp, slip, iters=(predictor, update), fail[, u1, u2]).
    MTS: ttn, u1n, u2n are the raw history values (< 0 = flags of mm10_init_mts)."""
    D = np.asarray(D, float)
    full = Step(cr, R, D, dt, Rpn)
    nst = dict(tt=ttn, u1=u1n, u2=u2n)
    tt_raw = ttn
    if cr.mts:
        full.setup_mts(cr, 297.0, nst)                           # may replace the tau_tilde flag
        ttn = nst["tt"]
    no_load = (sn @ sn == 0.0) and (D @ D == 0.0)
    if iter0 or no_load:                                        # elastic predictor (mm10_a.f:2735-2751)
        # tt was copied before mm10_setup ran (mm10_a.f:2674-2677): the raw n value, flag included
        x = np.concatenate([sn, [tt_raw]])
        sig = sn.copy()
        if not no_load:
            sig = sn - residual(cr, full, sn, tt_raw, x, False)[0][:6]
        out = dict(stress=sig, tt=tt_raw, tangent=cr.C.copy(), Rp=None, slip=None, iters=(0, 0), fail=False)
        if cr.mts:
            out.update(u1=full.tau_y, u2=full.mu)
        return out

    def dev_dir(d):
        d = d.copy(); d[:3] -= d[:3].sum() / 3.0
        nrm = np.sqrt(d @ d)
        return d / nrm if nrm > 0 else 0.0 * d
    cos_ang = max(dev_dir(D) @ dev_dir(np.asarray(Dn, float)), 0.0)
    frac, stp, cuts = 0.0, 1.0, 0
    x = np.concatenate([sn, [ttn]]); ox = x.copy()
    itp = itu = 0
    fail, J, h, st = False, None, ttn, full
    while frac < 1.0:
        sc = stp + frac
        st = Step(cr, R, D * sc, dt * sc, Rpn)
        if cr.mts:
            st.setup_mts(cr, 297.0 * sc, nst)                    # curr%temp = 297 (step + frac): n%temp = 0 (mm10_a.f:2769)
        x[6] = ttn
        x0 = x.copy(); x0[6] = ttn + cos_ang * ttrate_n * (dt * stp)
        x1, i1, f1, _, _, inR1 = newton(cr, st, sn, ttn, x0, 6, cr.atol1, cr.rtol1, 0)
        itp += i1
        xs = x.copy() if f1 else x1                              # a failed predictor leaves x untouched; update still runs
        x2, i2, f2, J2, h, _ = newton(cr, st, sn, ttn, xs, 7, cr.atol, cr.rtol, 1, inR1)
        itu += i2
        J = J2
        if f1 or f2:
            x = ox.copy(); stp *= 0.5; cuts += 1
            if cuts > 4:
                fail = True
                break
        else:
            x = x2; ox = x.copy(); frac += stp
    if fail or np.any(np.isnan(x)):
        return dict(stress=sn.copy(), tt=ttn, tangent=cr.C.copy(), Rp=Rpn.copy(), slip=None, iters=(itp, itu), fail=True,
                    u1=u1n, u2=u2n)
    # tangent from the LAGGED Jacobian of the last sub-step (mm10_a.f:1137), Voce: JA = JB = 0
    sig, tt = x[:6], x[6]
    rs, gam = slips(cr, full, sig, tt)
    JR = cr.C.copy()
    if cr.mts:                                                  # mm10_dgdd_mts, mm10_ed_mts at the converged state, full step
        m = cr.mts
        d_mod = D.copy(); d_mod[3:] *= 0.5
        alpha = 2.0 / (3.0 * full.dg ** 2)
        # mm10_a.f:771-772 hands `symtqmat` (= its column 1), not `symtqmat(1,i)`, to mm10_a_mult_type_4: the sym(sigma W) of
        # the FIRST slip system enters every term.  Reproduced (established by executing the reference, test_reference_vectors.py).
        JA = sum(np.outer(cr.C @ ms + 2.0 * symsw(sig, full.qc[0]), alpha * g * d_mod) for ms, g in zip(full.ms, gam))
        dgc = full.dg / full.tinc
        kT = m["boltz"] * full.temp / (full.mu * m["b"] ** 3)
        lny, lnv = np.log(m["eps_dot_0_y"] / dgc), np.log(m["eps_dot_0_v"] / dgc)
        ty, tv = kT / m["G_0_y"] * lny, kT / m["G_0_v"] * lnv
        dydd = (2.0 * m["tau_hat_y"] / (3.0 * full.dg ** 2 * m["q_y"] * m["p_y"] * lny)
                * (1.0 - ty ** (1.0 / m["q_y"])) ** (1.0 / m["p_y"] - 1.0) * ty ** (1.0 / m["q_y"]) * d_mod)
        dvdd = (2.0 * m["tau_hat_v"] / (3.0 * full.dg ** 2 * m["q_v"] * m["p_v"] * lnv)
                * (1.0 - tv ** (1.0 / m["q_v"])) ** (1.0 / m["p_v"] - 1.0) * tv ** (1.0 / m["q_v"]) * d_mod)
        mnp0 = full.mu / m["mu_0"]
        sc_ = tt / mnp0 - m["tau_a"] / mnp0 - full.tau_y
        base = 1.0 - sc_ / full.tau_v
        ed = sum((cr.m / full.tau_v * base ** (cr.m - 1.0) * dydd + cr.m / full.tau_v ** 2 * sc_ * base ** (cr.m - 1.0) * dvdd
                  + 2.0 / (3.0 * full.dg ** 2) * base ** cr.m * d_mod) * abs(g) for g in gam)
        ed = cr.theta_0 * mnp0 * ed + mnp0 * dydd
        JB = np.outer(J[:6, 6], ed) / J[6, 6]
        JR = cr.C - JA - JB
    JJ = J[:6, :6] - np.outer(J[:6, 6], J[6, :6]) / J[6, 6]
    T = np.linalg.solve(JJ, JR)
    T = 0.5 * (T + T.T)
    f = gam + rs * dt * cr.iD_v
    wbar = sum(fi * q for fi, q in zip(f, full.qs))
    W = skewt(wbar)
    al = np.sqrt(W[1, 2] ** 2 + W[0, 2] ** 2 + W[0, 1] ** 2)
    ex = np.eye(3) if al < 1e-16 else np.eye(3) + (1.0 - np.cos(al)) / al ** 2 * (W @ W) + np.sin(al) / al * W
    out = dict(stress=sig, tt=tt, tangent=T, Rp=ex @ Rpn, slip=f, iters=(itp, itu), fail=False,
               tt_rate=(h - ttn) / st.tinc)
    if cr.mts:
        out.update(u1=full.tau_y, u2=full.mu)
    return out
