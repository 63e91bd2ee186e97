"""Independent numpy restatement of the bilinear Mises model with mixed hardening (mm01) and its
consistent tangent (cnst1), vectorised over points, written from the Fortran: mm01_init
(mm01.f:328-534), mm01_simple1 (:704-784), the elastic branch of mm01 (:186-195), mm01_sig_final
(:626-689), mm01_plastic_work (:546-614), cnst1 (:1222-1374); isothermal (dtemps = 0), lnelas off.
TEST INFRASTRUCTURE ONLY; shares no code with oracle/ or the CUDA sources.
State per point: cgn (9) = stress 6, energy, plastic work, accumulated plastic strain; history (11)."""
import numpy as np

ROOT2, ROOT3 = 1.414213562373095, 1.7320508075688          # the reference's literals
TWTHRD, ROOT23 = 0.666666666666667, 0.816496580927


def initial_history(n, yld, hprime):
    """mm01_set_history (mm01.f:240-312); slot 4 is the integer state word 3 packed in a double"""
    h = np.zeros((n, 11))
    h[:, 1] = yld / 1.73205080756888
    h[:, 3] = np.array([3], dtype=np.int64).view(np.float64)[0]
    h[:, 4] = hprime
    return h


def update(cgn, hist, deps, ym, nu, beta, hprime, yld):
    """one strain increment deps (n, 6) (engineering shear).  Returns cgn1 (n, 9), hist1 (n, 11),
    cep (n, 6, 6), yield flags."""
    n = len(deps)
    g = ym / 2.0 / (1.0 + nu)
    dvol = deps[:, :3].sum(axis=1)
    de = deps.copy(); de[:, :3] -= (dvol / 3.0)[:, None]
    een = np.empty((n, 6))                                      # elastic strain at n from the stress at n
    s = cgn[:, :6]
    een[:, 0] = (s[:, 0] - nu * (s[:, 1] + s[:, 2])) / ym
    een[:, 1] = (s[:, 1] - nu * (s[:, 0] + s[:, 2])) / ym
    een[:, 2] = (s[:, 2] - nu * (s[:, 0] + s[:, 1])) / ym
    een[:, 3:] = s[:, 3:6] / g
    eps_vol_n1 = een[:, :3].sum(axis=1) + dvol
    e = een.copy(); e[:, :3] -= (een[:, :3].sum(axis=1) / 3.0)[:, None]
    e += de
    dev_el = np.concatenate([2.0 * g * e[:, :3], g * e[:, 3:]], axis=1)      # trial deviator
    alpha_n = hist[:, 5:11]
    hbari, hbark = beta * hprime, (1.0 - beta) * hprime
    kbar = (yld + hbari * hist[:, 2]) / ROOT3
    rtse = dev_el - alpha_n                                                   # lk = 1 (isothermal)
    mrts = np.sqrt((rtse[:, :3] ** 2).sum(axis=1) + 2.0 * (rtse[:, 3:] ** 2).sum(axis=1))
    yf = mrts - ROOT2 * kbar
    yield_ = yf >= 0.0000001 * ROOT2 * kbar
    hist1 = np.zeros((n, 11))
    dev = np.empty((n, 6))
    # elastic points (mm01.f:186-195)
    el = ~yield_
    hist1[el, 0] = 0.0; hist1[el, 1] = kbar[el]; hist1[el, 2] = hist[el, 2]; hist1[el, 4] = hprime
    hist1[el, 5:11] = alpha_n[el]
    dev[el] = rtse[el] + alpha_n[el]
    # yielding points: radial return (mm01_simple1)
    y = yield_
    ldt = (mrts[y] - ROOT2 * kbar[y]) / (TWTHRD * (3.0 * g + hprime))
    k1 = kbar[y] + (ROOT2 / 3.0) * hbari * ldt
    hist1[y, 0] = ldt; hist1[y, 1] = k1; hist1[y, 2] = hist[y, 2] + ldt * ROOT23; hist1[y, 4] = hprime
    c1 = TWTHRD * hbark * ldt / mrts[y]
    c2 = ROOT2 * k1 / mrts[y]
    hist1[y, 5:11] = hist[y, 5:11] + c1[:, None] * rtse[y]
    dev[y] = hist1[y, 5:11] + c2[:, None] * rtse[y]
    state = np.where(yield_, 1, 3).astype(np.int64)
    hist1[:, 3] = state.view(np.float64)
    # mm01_sig_final
    cgn1 = np.zeros((n, 9))
    smean = eps_vol_n1 * (3.0 * ym * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)) + 2.0 * g) / 3.0
    cgn1[:, :6] = dev
    cgn1[:, :3] += smean[:, None]
    cgn1[:, 6] = cgn[:, 6] + 0.5 * (deps * (cgn1[:, :6] + cgn[:, :6])).sum(axis=1)
    # mm01_plastic_work
    cgn1[:, 7], cgn1[:, 8] = cgn[:, 7], cgn[:, 8]
    ds = cgn1[y, :6] - cgn[y, :6]
    dp = np.empty_like(ds)
    dp[:, 0] = deps[y, 0] - (ds[:, 0] - nu * (ds[:, 1] + ds[:, 2])) / ym
    dp[:, 1] = deps[y, 1] - (ds[:, 1] - nu * (ds[:, 0] + ds[:, 2])) / ym
    dp[:, 2] = deps[y, 2] - (ds[:, 2] - nu * (ds[:, 0] + ds[:, 1])) / ym
    dp[:, 3:] = deps[y, 3:] - ds[:, 3:] / g
    cgn1[y, 7] = cgn[y, 7] + 0.5 * (dp * (cgn1[y, :6] + cgn[y, :6])).sum(axis=1)
    f1 = (dp[:, 0] - dp[:, 1]) ** 2 + (dp[:, 1] - dp[:, 2]) ** 2 + (dp[:, 0] - dp[:, 2]) ** 2
    f2 = (dp[:, 3:] ** 2).sum(axis=1)
    cgn1[y, 8] = cgn[y, 8] + (ROOT2 / 3.0) * np.sqrt(f1 + 1.5 * f2)
    # cnst1
    cep = np.zeros((n, 6, 6))
    c_1 = ym / ((1.0 + nu) * (1.0 - 2.0 * nu))
    Ce = np.zeros((6, 6))
    Ce[:3, :3] = nu * c_1
    Ce[np.arange(3), np.arange(3)] = (1.0 - nu) * c_1
    Ce[np.arange(3, 6), np.arange(3, 6)] = (1.0 - 2.0 * nu) / 2.0 * c_1
    cep[el] = Ce
    lam = ym * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
    K = (3.0 * lam + 2.0 * g) / 3.0
    r = rtse[y]
    mrtsq = (r[:, :3] ** 2).sum(axis=1) + 2.0 * (r[:, 3:] ** 2).sum(axis=1)
    bb = (1.414213562 * k1 + (2.0 / 3.0) * (1.0 - beta) * hprime * ldt) / np.sqrt(mrtsq)    # cnst1's own root2 literal
    gamma = 1.0 / (1.0 + hprime / (3.0 * g))
    gambar = gamma - 1.0 + bb
    gbar = g * bb
    albar = K - 2.0 * gbar / 3.0
    thbar = 2.0 * g * gambar
    cy = -(thbar / mrtsq)[:, None, None] * (r[:, :, None] * r[:, None, :])
    cy[:, :3, :3] += albar[:, None, None]
    cy[:, np.arange(3), np.arange(3)] += 2.0 * gbar[:, None]
    cy[:, np.arange(3, 6), np.arange(3, 6)] += gbar[:, None]
    cep[y] = cy
    return cgn1, hist1, cep, yield_
