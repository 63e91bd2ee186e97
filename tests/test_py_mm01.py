"""oracle (C++) against the independent numpy restatement of mm01 + cnst1 (tests/py_mm01.py) over a
load path with elastic, first-yield, flow and unloading increments."""
import numpy as np
import pytest

import py_mm01
from helpers import deck


@pytest.fixture(scope="module")
def Oracle(oracle_built):
    from oracle import Oracle
    return Oracle


def test_oracle_matches_numpy_restatement(Oracle):
    p = deck("test_mm01.in")
    o = Oracle(p)
    N3 = p.N3
    mats = [p.materials[m - 1] for m in p.matlist]
    f32 = lambda name: np.array([np.float64(np.float32(getattr(m, name))) for m in mats])   # REAL*4 slots
    ym, nu, beta, tan_e, yld = f32("e"), f32("nu"), f32("beta"), f32("tan_e"), f32("yld_pt")
    hprime = tan_e * ym / (ym - tan_e)
    rng = np.random.default_rng(8)
    I = np.zeros((9, N3)); I[[0, 4, 8]] = 1.0
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = -0.3; bar[8] = -0.55; bar[1] = 0.2
    G = rng.standard_normal((9, N3))
    o.drive_eps_sig(1, 0)
    cgn = np.zeros((N3, 9))
    hist = None
    eps_n = np.zeros((N3, 6))
    nyield = 0
    for step, amp in enumerate((0.002, 0.012, 0.03, 0.022), start=1):       # load, yield, flow, unload
        for it in (0, 1):
            F1 = I + amp * (0.8 + 0.2 * it) * (bar + 0.2 * G)
            o.Fn1[:] = F1
            assert o.drive_eps_sig(step, it) == 0
            deps = o.eps_n1 - eps_n                                        # unrotated strain increment the oracle used
            res = []
            for v in range(N3):                                            # per point: material constants differ
                h0 = py_mm01.initial_history(1, yld[v], hprime[v]) if step == 1 else hist[v:v + 1]
                c0 = cgn[v:v + 1].copy()
                if step == 1:
                    c0[:, 7:9] = 0.0
                res.append(py_mm01.update(c0, h0, deps[v:v + 1], ym[v], nu[v], beta[v], hprime[v], yld[v]))
            cgn1 = np.concatenate([r[0] for r in res]); hist1 = np.concatenate([r[1] for r in res])
            cep = np.concatenate([r[2] for r in res]); yflag = np.concatenate([r[3] for r in res])
            scale = np.abs(o.urcs_n1[:, :6]).max()
            assert np.abs(cgn1[:, :6] - o.urcs_n1[:, :6]).max() <= 1e-12 * scale, (step, it)
            assert np.abs(cgn1[:, 6:] - o.urcs_n1[:, 6:]).max() <= 1e-11 * max(np.abs(o.urcs_n1[:, 6:]).max(), 1e-30)
            if it > 0:                                                      # history n+1 is scattered for iter > 0 only (rplstr.f:78)
                assert np.array_equal(hist1[:, 3].view(np.int64), o.hist_n1[:, 3].copy().view(np.int64))
                cols = [0, 1, 2, 4, 5, 6, 7, 8, 9, 10]
                assert np.abs(hist1[:, cols] - o.hist_n1[:, cols]).max() <= 1e-12 * max(np.abs(o.hist_n1[:, cols]).max(), 1.0)
                nyield += int(yflag.sum())
            # the consistent tangent [D] of cnst1, through the exact dP/dF it produces (cep2A probe)
            for v in range(0, N3, 37):
                A = Oracle.cep2A(o.Fn[:, v], F1[:, v], cgn1[v, :6], cep[v])
                assert np.abs(A - o.K4[:, v]).max() <= 1e-11 * np.abs(o.K4[:, v]).max(), (step, it, v)
            last = (cgn1, hist1, cep)
        cgn, hist = last[0], last[1]
        eps_n = o.eps_n1.copy()
        o.Fn[:] = F1
        o.update()
    assert 0 < nyield
