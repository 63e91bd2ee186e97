"""The CUDA material stage, through the C ABI, against outputs of the REFERENCE'S OWN SOURCE (tests/golden/reference_vectors.npz,
reference_global.npz: maranGit/CPFFT's Fortran executed by tools/fortran_subset.py; tests/test_reference_vectors.py and
tests/test_reference_global.py hold the oracle and the host build of the kernel source to the same data on the CPU).

Written after the round's GPU budget was spent: the bodies below were run on the CPU behind the Solver interface with the host
build of the kernel source (tests/test_gpu_test_bodies.py), not yet on a GPU.  The file sorts last so that a surprise here cannot
hide the established GPU tests under `pytest -x`.  Local Newton counts: device and host libm differ in the last bit, a point
sitting exactly on a convergence threshold may take an iteration more or less (tests/test_zz_gpu_new_features.py); the summed
counts of a case are therefore allowed to differ from the reference's by 2 (or 2 %), everything else is held to the per-voxel tolerances
measured on the host build with a margin for the polar-decomposition noise band."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COUNT_SLACK = 2


@pytest.fixture(scope="module")
def solver(oracle_built):
    from cpfft_b200 import Solver
    return Solver


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def slack(n):
    return max(COUNT_SLACK, int(0.02 * int(n)))


def counts_close(it, want, npts):
    got = (int(it[:npts, 0].sum()), int(it[:npts, 1].sum()))
    return abs(got[0] - int(want[0])) <= slack(want[0]) and abs(got[1] - int(want[1])) <= slack(want[1])


def voxel_case(Solver, k):
    """one voxel through do_nleps_block's sequence (tests/test_reference_vectors.py::test_kernel_source_voxel_against_reference)"""
    from cpfft_b200.polycrystal import polycrystal
    from cpfft_b200.problem import Crystal
    from test_reference_vectors import V, _hist_layout
    rate_n, theta_0, tau_y, tau_v, voche_m, iD_v, e, nu = V["voxel_params"][k]
    cr = Crystal(slip_type=int(V["voxel_slip_type"][k]), elastic_type=1, h_type=int(V["voxel_h_type"][k]), e=e, nu=nu, mu=e / 2.0 / (1.0 + nu),
                 harden_n=rate_n, theta_0=theta_0, tau_y=tau_y, tau_v=tau_v, voche_m=voche_m, iD_v=iD_v)
    if cr.h_type == 2:
        for name, val in zip(V["crystal_mts_names"], V["crystal_mts_params"]):
            if str(name) != "theta_0":
                setattr(cr, str(name), float(val))
    p = polycrystal(2, ngrains=1)
    p.crystals = [cr]
    p.angles = np.ascontiguousarray(np.tile(V["voxel_angles"][k], (p.N3, 1)))
    L = _hist_layout(12 if cr.slip_type == 1 else 48)
    ns = V["voxel_n_state"][k]
    Rp = ns[23:32].reshape(3, 3)
    s = Solver(p)
    s.drive_eps_sig(1, 0)
    hist = np.zeros((s.H, p.N3))
    for q in range(6):
        hist[L["c_stress"] + q] = ns[q]; hist[L["c_D"] + q] = ns[8 + q]; hist[L["c_eps"] + q] = ns[14 + q]
    hist[L["c_tt"]] = ns[6]; hist[L["c_ttrate"]] = ns[7]
    for q in range(3):
        hist[L["c_euler"] + q] = ns[20 + q]
    for i in range(3):
        for j in range(3):
            hist[L["c_Rp"] + 3 * j + i] = Rp[i, j]
            hist[63 + 3 * j + i] = 1.0 if i == j else 0.0
    if cr.h_type == 2:
        hist[L["c_u"]] = -1.0; hist[L["c_u"] + 1] = -1.0
    ones = np.ones((1, p.N3))
    for name, arr in (("HIST_N", hist), ("URCS_N", np.zeros((9, p.N3))), ("EPS_N", np.zeros((6, p.N3))),
                      ("FN", V["voxel_Fn"][k].reshape(9, 1) * ones), ("FN1", V["voxel_Fn1"][k].reshape(9, 1) * ones)):
        s.upload(name, arr)
    s.drive_eps_sig(2, 1)
    tol = 2e-9                                                # host build: 1e-13 .. 8e-11; the north-star per-voxel tolerance with the band's margin
    h1 = s.download("HIST_N1", 1)[0]
    assert rel(s.download("URCS_N1", 1)[0, :6], V["voxel_stress"][k]) <= tol
    assert rel(s.download("PN1")[:, 0], V["voxel_P"][k]) <= tol
    assert rel(s.download("K4")[:, 0], V["voxel_dPdF"][k]) <= 5.0 * tol
    assert abs(h1[L["c_tt"]] - V["voxel_tt"][k]) <= tol * V["voxel_tt"][k]
    Rp1 = np.array([[h1[L["c_Rp"] + 3 * j + i] for j in range(3)] for i in range(3)])
    assert np.abs(Rp1 - V["voxel_Rp"][k]).max() <= 1e-10
    assert np.abs(h1[L["c_eps"]:L["c_eps"] + 6] - V["voxel_eps"][k]).max() <= 1e-10
    assert counts_close(s.local_iters(), V["voxel_iters"][k] * 1, 1)


@pytest.mark.parametrize("k", range(5))
def test_voxel_through_the_c_abi(solver, k):
    """cpfft_drive_eps_sig on one voxel run end to end by the reference's source: fcc Voce at 1 %, 3 %, 0.2 % strain, bcc48, MTS"""
    voxel_case(solver, k)


def wrapper_case(Solver, name):
    """the reference's wrapper mm10 on Taylor points, MTS and the 48-system layout, two load steps with a commit
    (tests/test_reference_global.py::test_wrapper_cases)"""
    from cpfft_b200.polycrystal import polycrystal, taylor_polycrystal
    from cpfft_b200.problem import Crystal
    from helpers import compare_mm10_history
    from test_reference_global import V
    TOLW = 2e-8                                               # host build: worst group 3.4e-9 at these 0.15 % increments (polar band 1.3e-9 in R)
    W = lambda k: V[f"wrap_{name}_{k}"]
    ncry, slip_type, npts = int(W("ncry")), int(W("slip_type")), W("F1").shape[0]
    rate_n, theta_0, tau_y, tau_v, voche_m, iD_v, e, nu = V["params"]
    cr = Crystal(slip_type=slip_type, elastic_type=1, h_type=int(W("h_type")), e=e, nu=nu, mu=e / 2.0 / (1.0 + nu), harden_n=rate_n, theta_0=theta_0,
                 tau_y=tau_y, tau_v=tau_v, voche_m=voche_m, iD_v=iD_v)
    if cr.h_type == 2:
        for nm, val in zip(V["mts_names"], V["mts_params"]):
            setattr(cr, str(nm), float(val))
    p = taylor_polycrystal(2, ncrystals=ncry, ngrains=2) if ncry > 1 else polycrystal(2, ngrains=1)
    p.crystals = [cr]
    rep = p.N3 // npts
    ang = np.tile(W("angles"), (rep, 1, 1))
    p.angles = np.ascontiguousarray(ang if ncry > 1 else ang[:, 0, :])
    tile = lambda a: np.tile(a, (rep, 1))
    s = Solver(p)
    assert s.H == int(W("hist_size"))
    s.drive_eps_sig(1, 0)
    assert rel(s.download("K4").T, tile(W("K4_initial"))) <= 1e-12
    nslip = 12 if slip_type == 1 else 48
    for step, Fk in ((1, "F1"), (2, "F2")):
        F = tile(W(Fk)).T
        s.upload("FN1", F)
        s.drive_eps_sig(step, 1)
        assert rel(s.download("PN1").T, tile(W(f"P{step}"))) <= TOLW
        assert rel(s.download("K4").T, tile(W(f"K4_{step}"))) <= TOLW
        assert rel(s.download("URCS_N1", 1)[:, :6], tile(W(f"urcs{step}"))[:, :6]) <= TOLW
        assert compare_mm10_history(s.download("HIST_N1", 1)[:, :s.H], tile(W(f"hist{step}")), nslip, TOLW, ncry)
        assert counts_close(s.local_iters(), W(f"iters{step}"), npts)
        s.upload("FN", F)
        s.update()


@pytest.mark.parametrize("name", ["taylor", "mts", "bcc48"])
def test_wrapper_cases_through_the_c_abi(solver, name):
    wrapper_case(solver, name)


def job_sweep(Solver, job, k):
    """the closing sweep of load step k of a job the reference ran through FFT_nr3, on the reference's own state: 27 points in four
    2^3 grids (tests/test_reference_global.py::test_kernel_source_sweep_on_the_reference_state / ..._mm01_...)"""
    import test_reference_global as G
    V = G.V
    full = G.problem(job)
    n27 = full.N3
    F1, H1, U1, P1 = V[job + "step_Fn1"][k - 1], V[job + "step_hist"][k - 1], V[job + "step_urcs"][k - 1], V[job + "step_Pn1"][k - 1]
    Fn = np.tile(np.eye(3).reshape(9), (n27, 1)) if k == 1 else V[job + "step_Fn1"][k - 2]
    last = [r for r in V[job + "sweeps"] if r[0] == k][-1]
    H = int(V[job + "hist_size"])
    got_iters = np.zeros(2, dtype=np.int64)
    worst = {}
    for c0 in range(0, n27, 8):
        idx = np.array([i if i < n27 else c0 for i in range(c0, c0 + 8)])
        real = int(min(8, n27 - c0))
        p = G.problem(job)
        p.N = 2
        p.matlist = np.ascontiguousarray(full.matlist[idx])
        p.angles = np.ascontiguousarray(np.asarray(full.angles)[idx])
        s = Solver(p)
        assert s.H == H
        s.drive_eps_sig(1, 0)
        s.upload("FN", Fn[idx].T); s.upload("FN1", F1[idx].T)
        if k > 1:
            s.upload("HIST_N", V[job + "step_hist"][k - 2][idx].T)
            s.upload("URCS_N", V[job + "step_urcs"][k - 2][idx].T)
        s.drive_eps_sig(k, int(last[1]))
        hk = s.download("HIST_N1", 1)[:real, :H]
        ur = s.download("URCS_N1", 1)[:real]
        P = s.download("PN1").T[:real]
        sel = idx[:real]
        errs = dict(P=np.abs(P - P1[sel]).max() / np.abs(P1).max(), stress=np.abs(ur[:, :6] - U1[sel, :6]).max() / np.abs(U1[:, :6]).max())
        if job == "":
            cs = 75 + 12
            errs.update(D=np.abs(hk[:, :36] - H1[sel, :36]).max() / np.abs(H1[:, :36]).max(), Rp=np.abs(hk[:, cs + 9:cs + 18] - H1[sel, cs + 9:cs + 18]).max(),
                        tt=np.abs(hk[:, cs + 42] - H1[sel, cs + 42]).max() / np.abs(H1[:, cs + 42]).max())
            it = s.local_iters()
            got_iters += (int(it[:real, 0].sum()), int(it[:real, 1].sum()))
        else:
            mask = np.ones(11, dtype=bool); mask[3] = False
            errs.update(hist=np.abs(hk[:, mask] - H1[sel][:, mask]).max() / np.abs(H1[:, mask]).max())
            word = lambda a: np.ascontiguousarray(a[:, 3]).view(np.int64)
            assert np.array_equal(word(hk), word(H1[sel]))
        for key, v in errs.items():
            worst[key] = max(worst.get(key, 0.0), float(v))
    # job "": two principal stretches nearly equal, the closed-form polar decomposition's worst case (host build: P 5e-9, stress 7e-10);
    # the bound is the small-strain floor the GPU polycrystal tests use (tests/test_host_kernels.py header)
    tols = dict(P=5e-8, stress=2e-8, D=2e-8, Rp=1e-9, tt=5e-9, hist=1e-10) if job == "" else dict(P=5e-9, stress=5e-9, hist=1e-10)
    bad = {key: v for key, v in worst.items() if not v <= tols[key]}
    assert not bad, (bad, worst)
    if job == "":
        assert abs(int(got_iters[0]) - int(last[2])) <= slack(last[2]) and abs(int(got_iters[1]) - int(last[3])) <= slack(last[3]), (got_iters, last)


@pytest.mark.parametrize("job,k", [("", 1), ("", 2), ("", 3), ("m01_", 1), ("m01_", 2), ("m01_", 3)])
def test_job_sweeps_through_the_c_abi(solver, job, k):
    job_sweep(solver, job, k)


def deck_sweep(Solver, deck_name, job, k):
    """the closing sweep of load step k of a shipped deck as the reference ran it (all 343 points at once, the deck's own grid):
    examples/test_mm10.in exercises bcc48 with alter_mode on, examples/test_mm01.in the bilinear Mises update at 3 % strain per step"""
    from helpers import deck, mm10_layout
    import test_reference_global as G
    V = G.V
    p = deck(deck_name)
    n = p.N3
    F1, H1, U1, P1 = V[job + "step_Fn1"][k - 1], V[job + "step_hist"][k - 1], V[job + "step_urcs"][k - 1], V[job + "step_Pn1"][k - 1]
    Fn = np.tile(np.eye(3).reshape(9), (n, 1)) if k == 1 else V[job + "step_Fn1"][k - 2]
    last = [r for r in V[job + "sweeps"] if r[0] == k][-1]
    H = int(V[job + "hist_size"])
    s = Solver(p)
    assert s.H == H
    s.drive_eps_sig(1, 0)
    s.upload("FN", Fn.T); s.upload("FN1", F1.T)
    if k > 1:
        s.upload("HIST_N", V[job + "step_hist"][k - 2].T)
        s.upload("URCS_N", V[job + "step_urcs"][k - 2].T)
    s.drive_eps_sig(k, int(last[1]))
    hk = s.download("HIST_N1", 1)[:, :H]
    ur = s.download("URCS_N1", 1)
    assert rel(s.download("PN1").T, P1) <= 2e-8
    assert rel(ur[:, :6], U1[:, :6]) <= 2e-8
    if job in ("deckmts_", "decktaylor_"):                     # MTS hardening; two crystal types (bcc48 + fcc) per point
        from helpers import compare_mm10_history
        ncry = max(m.n_crystals for m in p.materials)
        nslip = 48 if any(c.slip_type == 8 for c in p.crystals) else 12
        assert compare_mm10_history(hk, H1, nslip, 5e-8, ncry)
        it = s.local_iters()
        got = (int(it[:, 0].sum()), int(it[:, 1].sum()))
        assert abs(got[0] - int(last[2])) <= slack(last[2]) and abs(got[1] - int(last[3])) <= slack(last[3]), (got, last)
    elif job == "deck_":
        L = mm10_layout(48)
        for key, tol in (("cep", 2e-8), ("stress", 2e-8), ("tau_tilde", 1e-9), ("slipinc", 5e-8)):
            a, b = hk[:, L[key][0]:L[key][1]], H1[:, L[key][0]:L[key][1]]
            assert rel(a, b) <= tol, (key, rel(a, b))
        a, b = hk[:, L["Rp"][0]:L["Rp"][1]], H1[:, L["Rp"][0]:L["Rp"][1]]
        assert np.abs(a - b).max() <= 1e-9
        it = s.local_iters()
        got = (int(it[:, 0].sum()), int(it[:, 1].sum()))
        assert abs(got[0] - int(last[2])) <= slack(last[2]) and abs(got[1] - int(last[3])) <= slack(last[3]), (got, last)
    else:
        mask = np.ones(11, dtype=bool); mask[3] = False
        assert rel(hk[:, mask], H1[:, mask]) <= 1e-10
        word = lambda a: np.ascontiguousarray(a[:, 3]).view(np.int64)
        assert np.array_equal(word(hk), word(H1))


@pytest.mark.parametrize("deck_name,job,k", [("test_mm10.in", "deck_", 1), ("test_mm10.in", "deck_", 3), ("test_mm10.in", "deck_", 10),
                                             ("test_mm01.in", "deck01_", 1), ("test_mm01.in", "deck01_", 10),
                                             ("mts_mm10.in", "deckmts_", 2), ("mts_mm10.in", "deckmts_", 4),
                                             ("taylor_mm10.in", "decktaylor_", 2), ("taylor_mm10.in", "decktaylor_", 5)])
def test_shipped_deck_sweeps_through_the_c_abi(solver, deck_name, job, k):
    deck_sweep(solver, deck_name, job, k)
