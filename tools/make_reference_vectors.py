#!/usr/bin/env python
"""Golden vectors produced by the reference's OWN source: leaf routines of /root/reference/src/*.f executed by the
Fortran-subset interpreter tools/fortran_subset.py on seeded inputs, saved as tests/golden/reference_vectors.npz.

    python tools/make_reference_vectors.py            # needs /root/reference (this container); writes the fixture

tests/test_reference_vectors.py (which needs neither /root/reference nor this script) holds the oracle, the numpy
restatements of tests/py_mm10.py and the spectral restatement to these outputs.  This is the pin the oracle has: the
reference cannot be compiled here (ifort + MKL), but these routines of it can be run, statement by statement, in the
order and precision the source states.  Routines (reference file:line of the subroutine statement):

  polar.f:18    rtcmp1  (+ irscp1 :58, ivcmp1 :133, evcmp1_new :226)   R of F = R U, closed form        -> K2
  polar.f:680   getrm1                                                   6x6 rotation operators, opt 1-3 -> K3
  cep2A.f:86    cep2A_a (+ multiply33, transpose33, ddot44, det33)       dP/dF from [D], sigma, F        -> K4
  mm10_a.f:1287 mm10_rotation_matrix                                     Kocks angles -> g               -> M2
  mm10_a.f:1400 mm10_rt2rve, :1461 mm10_rt2rvw                           slip-vector rotation operators  -> M4
  mm10_b.f      mm10_symSW                                               sym(S W) in Voigt form          -> M5
  mm10_a.f:830  mm10_setup (+ setup_voche, mult_type_*, ET2EV, WT2WV)   current ms, qs, qc, dg           -> M4
  mm10_b.f:1029 mm10_formR (+ formR1/R2, form_dbarp/wbarp/wp, slipinc, rs, h_voche, symSW)  residual   -> M5
  mm10_b.f:901  mm10_formJ (+ formJ11..J22, formarrs, dgdt/estress/ehard_voche, IW, symSWmat, DGER)   Jacobian -> M6
  mm10_a.f:1080 mm10_solve_crystal: mm10_solve_strup (:2628), mm10_solve (:2860, predictor / update Newton loops,
                line search, DGESV), mm10_tangent (:658), mm10_update_rotation (:3310), mm10_output (:3433)      -> M3, M7-M10
  FFT_init.f:272 formG                                                   Green operator table, odd N     -> G3
  G_K_dF.f:241  ddot42n                                                  K4 : x with its summation tree  -> G1
  do_nleps_block end to end for one mm10 voxel: (Fn, Fn1, n state) -> kinematics -> mm10_solve_crystal -> cs2p, cep2A_a -> P, dP/dF
  drive_eps_sig.f:1017 inv33, :1110 mul33, :1182 cs2p, qmply1.f:15 qmply1, in do_nleps_block's order (:203-300)   -> K1, K3
  mm01.f:28     mm01 (+ mm01_set_history, _init, _simple1, _sig_final, _plastic_work) and cnst1 (:1222)            -> M1
  G_K_dF.f:11   G_K_dF (+ fftfem3d :101, ifftfem3d :163, formfftshift FFT_init.f:355; DFTI by numpy)  the operator -> G2, G4, G5
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fortran_subset as F  # noqa: E402

REF = "/root/reference/src/"
FILES = ["param_def", "mod_crystals.f", "polar.f", "cep2A.f", "mm10_a.f", "mm10_b.f", "FFT_init.f", "G_K_dF.f", "mm01.f", "drive_eps_sig.f", "qmply1.f"]


def interpreter():
    it = F.Interpreter()
    it.add_constants(open(REF + "param_def").read())
    mc = open(REF + "mod_crystals.f").read()
    i0 = mc.index("      module mm10_constants")
    it.add_constants(mc[i0:mc.index("      end module", i0)])
    for f in FILES[2:8]:
        it.load(open(REF + f).read())
    return it


def rand_rotation(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    return q * np.sign(np.linalg.det(q))


def main():
    it = interpreter()
    mx = it.consts["mxvl"]
    rng = np.random.default_rng(20240607)
    out = {}

    # ---- rtcmp1: small / moderate / large stretches, with and without a superposed rotation
    Fs, amps = [], []
    for amp in (1e-4, 1e-3, 1e-2, 5e-2, 0.2, 0.5):
        for rot in (False, True):
            for _ in range(2):
                Fm = np.eye(3) + amp * rng.standard_normal((3, 3))
                Fs.append(rand_rotation(rng) @ Fm if rot else Fm)
                amps.append(amp)
    Fs = np.array(Fs)
    out["polar_amp"] = np.array(amps)
    n = len(Fs)
    fb = np.zeros((mx, 3, 3), order="F"); rb = np.zeros((mx, 3, 3), order="F")
    fb[:n] = Fs
    it.call("rtcmp1", n, fb, rb)
    out["polar_F"], out["polar_R"] = Fs, rb[:n].copy()

    # ---- getrm1: the three operator flavours for the rotations just computed
    qs = []
    for opt in (1, 2, 3):
        qb = np.zeros((mx, 6, 6), order="F")
        it.call("getrm1", n, qb, rb, opt)
        qs.append(qb[:n].copy())
    out["getrm1_q"] = np.array(qs)                       # (3 opts, n, 6, 6)

    # ---- cep2A_a
    cases = []
    out["cep2A_amp"] = np.repeat([1e-3, 2e-2, 0.1], 3)
    for amp in (1e-3, 2e-2, 0.1):
        for _ in range(3):
            Fn = np.eye(3) + amp * rng.standard_normal((3, 3))
            Fn1 = Fn + 0.3 * amp * rng.standard_normal((3, 3))
            t6 = 300.0 * rng.standard_normal(6)
            A = rng.standard_normal((6, 6)); C66 = 1e4 * (A @ A.T) + 8e4 * np.eye(6)
            Fnh = 0.5 * (Fn + Fn1)
            fb[:] = 0; rb[:] = 0
            fb[0], fb[1] = Fnh, Fn1
            it.call("rtcmp1", 2, fb, rb)
            Rh, R = rb[0].copy(), rb[1].copy()
            t33 = np.array([[t6[0], t6[3], t6[5]], [t6[3], t6[1], t6[4]], [t6[5], t6[4], t6[2]]])
            dPdF = np.zeros(81)
            it.call("cep2a_a", np.asfortranarray(Fn), np.asfortranarray(t33), np.asfortranarray(C66), np.asfortranarray(Rh),
                    float(np.linalg.det(Fnh)), np.asfortranarray(np.linalg.inv(Fnh)), np.asfortranarray(R), np.asfortranarray(Fn1),
                    np.asfortranarray(np.linalg.inv(Fn1)), float(np.linalg.det(Fn1)), dPdF)
            cases.append((Fn, Fn1, t6, C66, dPdF))
    for k, name in enumerate(("cep2A_Fn", "cep2A_Fn1", "cep2A_t6", "cep2A_C66", "cep2A_dPdF")):
        out[name] = np.array([c[k] for c in cases])

    # ---- mm10 rotation helpers
    rots = np.array([rand_rotation(rng) for _ in range(6)])
    rve, rvw = [], []
    for Q in rots:
        a, b = np.zeros((6, 6), order="F"), np.zeros((3, 3), order="F")
        it.call("mm10_rt2rve", np.asfortranarray(Q), a); it.call("mm10_rt2rvw", np.asfortranarray(Q), b)
        rve.append(a.copy()); rvw.append(b.copy())
    out["mm10_rt"], out["mm10_rt2rve"], out["mm10_rt2rvw"] = rots, np.array(rve), np.array(rvw)
    angs = np.array([[0, 0, 0], [30, 45, 60], [90, 90, 90], [12.5, 133.0, 271.25], [359, 1, 180], [45, 0, 45], [200, 70, 10], [77, 150, 300]], dtype=float)
    gs = []
    for a3 in angs:
        g = np.zeros((3, 3), order="F")
        it.call("mm10_rotation_matrix", a3.copy(), "kocks", "degrees", g, 6)
        gs.append(g.copy())
    out["kocks_angles"], out["kocks_g"] = angs, np.array(gs)
    S, W, SW = rng.standard_normal((6, 6)) * 100.0, rng.standard_normal((6, 3)) * 1e-3, []
    for s6, w3 in zip(S, W):
        sw = np.zeros(6)
        it.call("mm10_symsw", s6.copy(), w3.copy(), sw)
        SW.append(sw.copy())
    out["symsw_s"], out["symsw_w"], out["symsw_sw"] = S, W, np.array(SW)

    # ---- formG for odd grids (the sizes at which the reference's table is a projection)
    for N in (3, 5, 7):
        G = np.zeros((N ** 3, 81), order="F")
        it2 = F.Interpreter(it.consts); it2.units = it.units
        it2.module_vars = dict(n=N, nhalf=(N + 1) // 2, n3=N ** 3, ndim1=3, ndim2=9, ghat4=G)     # FFT_init.f:146
        it2.call("formg")
        out[f"formG_{N}"] = np.ascontiguousarray(G)

    # ---- ddot42n
    nv = 7
    A4 = np.asfortranarray(rng.standard_normal((nv, 81)) * 1e5); B2 = np.asfortranarray(rng.standard_normal((nv, 9)) * 1e-3)
    C2, tmp = np.zeros((nv, 9), order="F"), np.zeros((nv, 9), order="F")
    it.call("ddot42n", A4, B2, C2, tmp, nv)
    out["ddot42_A4"], out["ddot42_B2"], out["ddot42_C2"] = np.ascontiguousarray(A4), np.ascontiguousarray(B2), np.ascontiguousarray(C2)

    # ---- mm10: mm10_setup -> mm10_formR, mm10_formJ (Voce) at trial points, with a plastic rotation Rp_n and a
    #      polar rotation R.  props%ms / qs / ns / stiffness are built the way setup_mm10_rknstr does
    #      (drive_eps_sig.f:571-606, 975-986) from the reference's own mm10_rotation_matrix, mm10_RT2RVE, mm10_ET2EV,
    #      mm10_WT2WV; the crystal-frame slip vectors and elastic constants are inputs of the fixture.
    from types import SimpleNamespace as NS
    sys.path.insert(0, ROOT)
    from oracle import Oracle                                 # noqa: E402  only for the (b, n) tables extracted from mod_crystals.f

    def slip_vectors(slip_type):       # unit vectors k / sqrt(k.k) of mod_crystals.f:438-1205 (tools/extract_slip_tables.py)
        return Oracle.slip_table(slip_type)
    mu, ms_max = it.consts["max_uhard"], it.consts["max_slip_sys"]
    rec = {k: [] for k in ("slip_type", "angles", "D6", "x7", "n_stress", "n_tt", "Rp", "R", "R7", "J", "ms", "qs", "qc", "params")}
    for slip_type in (1, 8):                                  # fcc, bcc48
        b, nrm = slip_vectors(slip_type)
        nslip = len(b)
        for case in range(3):
            ang = rng.uniform(0.0, 360.0, 3)
            e_mod, nu = 200000.0, 0.3
            prm = dict(rate_n=20.0 if case < 2 else 7.5, theta_0=100.0, tau_y=100.0, tau_v=100.0, voche_m=1.0 if case != 1 else 1.7,
                       iD_v=0.0 if case != 2 else 1e-7, e=e_mod, nu=nu)
            Sf = np.zeros((6, 6)); Sf[:3, :3] = -nu / e_mod
            Sf[np.arange(3), np.arange(3)] = 1.0 / e_mod; Sf[np.arange(3, 6), np.arange(3, 6)] = 2.0 * (1.0 + nu) / e_mod
            Cc = np.linalg.inv(Sf); Cc = 0.5 * (Cc + Cc.T)
            g = np.zeros((3, 3), order="F")
            it.call("mm10_rotation_matrix", ang.copy(), "kocks", "degrees", g, 6)
            trot = np.asfortranarray(g.T)
            RE = np.zeros((6, 6), order="F")
            it.call("mm10_rt2rve", trot, RE)
            props = NS(nslip=nslip, num_hard=1, h_type=1, out=6, alter_mode=False, eps_dot_0_y=0.0, k_0=0.0, burgers=3.5e-7,
                       cp_031=0.0, rate_n=prm["rate_n"], theta_0=prm["theta_0"], tau_y=prm["tau_y"], tau_v=prm["tau_v"],
                       voche_m=prm["voche_m"], id_v=prm["iD_v"], stiffness=np.asfortranarray(RE @ Cc @ RE.T),
                       ms=np.zeros((6, ms_max), order="F"), qs=np.zeros((3, ms_max), order="F"), ns=np.zeros((3, ms_max), order="F"))
            for s_ in range(nslip):
                bs, ns_ = trot @ b[s_], trot @ nrm[s_]
                A = np.outer(bs, ns_)
                ev, wv = np.zeros(6), np.zeros(3)
                it.call("mm10_et2ev", np.asfortranarray(0.5 * (A + A.T)), ev)
                it.call("mm10_wt2wv", np.asfortranarray(0.5 * (A - A.T)), wv)
                props.ms[:, s_], props.qs[:, s_], props.ns[:, s_] = ev, wv, ns_
            w = rng.standard_normal(3) * 0.03
            Wm = np.array([[0, w[2], w[1]], [-w[2], 0, w[0]], [-w[1], -w[0], 0.0]])
            Rp = np.eye(3) + Wm + 0.5 * Wm @ Wm; Rp, _ = np.linalg.qr(Rp); Rp = Rp * np.sign(np.diag(Rp))[None, :]
            fb[:] = 0; rb[:] = 0; fb[0] = np.eye(3) + 0.02 * rng.standard_normal((3, 3))
            it.call("rtcmp1", 1, fb, rb)
            Rpol = rb[0].copy()
            D6 = rng.standard_normal(6) * 1e-3
            n_stress = rng.standard_normal(6) * 60.0
            n_tt = 100.0 + 20.0 * rng.random()
            x7 = np.concatenate([n_stress + rng.standard_normal(6) * 40.0, [n_tt + 3.0 * rng.random()]])
            nst = NS(rp=np.asfortranarray(Rp), r=np.asfortranarray(np.eye(3)), stress=n_stress.copy(), tau_tilde=np.full(mu, n_tt),
                     gradfeinv=np.zeros((3, 3, 3), order="F"))
            np1 = NS(d=D6.copy(), r=np.asfortranarray(Rpol), tinc=1.0, dg=0.0, mu_harden=0.0, ms=np.zeros((6, ms_max), order="F"),
                     qs=np.zeros((3, ms_max), order="F"), qc=np.zeros((3, ms_max), order="F"), tau_l=np.zeros(ms_max),
                     tt_rate=np.zeros(mu))
            it.call("mm10_setup", props, np1, nst)
            vec1, vec2 = np.zeros(mu), np.zeros(mu)
            arr1, arr2 = np.zeros((mu, mu), order="F"), np.zeros((mu, mu), order="F")
            Rv, Jm = np.zeros(7), np.zeros((7, 7), order="F")
            stress, tt = x7[:6].copy(), x7[6:7].copy()
            it.call("mm10_formr", props, np1, nst, vec1, vec2, stress, tt, Rv, 1)
            it.call("mm10_formj", props, np1, nst, vec1, vec2, arr1, arr2, stress, tt, Jm)
            pad = lambda a, k: np.concatenate([a[:, :nslip].T, np.zeros((48 - nslip, k))])
            for k, v in (("slip_type", slip_type), ("angles", ang), ("D6", D6), ("x7", x7), ("n_stress", n_stress), ("n_tt", n_tt), ("Rp", Rp),
                         ("R", Rpol), ("R7", Rv.copy()), ("J", np.ascontiguousarray(Jm)), ("ms", pad(np1.ms, 6)), ("qs", pad(np1.qs, 3)),
                         ("qc", pad(np1.qc, 3)), ("params", [prm[q] for q in ("rate_n", "theta_0", "tau_y", "tau_v", "voche_m", "iD_v", "e", "nu")])):
                rec[k].append(v)
    for k, v in rec.items():
        out["mm10_" + k] = np.array(v)

    # ---- mm10: the WHOLE update of one crystal, mm10_solve_crystal (mm10_a.f:1080-1157) = mm10_solve_strup (sub-stepping,
    #      mm10_setup_np1, mm10_solve with its predictor / update Newton loops, line search, DGESV) -> mm10_tangent ->
    #      mm10_a_make_symm_1 -> mm10_update_rotation -> mm10_output (DPOSV), from explicit n states.  Newton iteration counts
    #      = numbers of Jacobian formations (calls of mm10_formJ11 / mm10_formJ).
    it.module_vars["asymmetric_assembly"] = False

    def new_state():
        return NS(r=np.zeros((3, 3), order="F"), rp=np.zeros((3, 3), order="F"), stress=np.zeros(6), d=np.zeros(6), eps=np.zeros(6),
                  euler_angles=np.zeros(3), slip_incs=np.zeros(ms_max), tau_tilde=np.zeros(mu), tt_rate=np.zeros(mu), u=np.zeros(mu),
                  ep=np.zeros(6), ed=np.zeros(6), tangent=np.zeros((6, 6), order="F"), ms=np.zeros((6, ms_max), order="F"),
                  qs=np.zeros((3, ms_max), order="F"), qc=np.zeros((3, ms_max), order="F"), tau_l=np.zeros(ms_max),
                  gradfeinv=np.zeros((3, 3, 3), order="F"), dg=0.0, tinc=0.0, temp=0.0, mu_harden=0.0, work_inc=0.0, p_work_inc=0.0,
                  p_strain_inc=0.0, step=0, elem=0, iter=0, gp=0, tau_v=0.0, tau_y=0.0)
    it.derived_factories["crystal_state"] = new_state
    MTS = dict(theta_0=1500.0, tau_a=20.0, tau_hat_y=180.0, g_0_y=0.4, tau_hat_v=300.0, g_0_v=1.2, burgers=2.5e-7, mu_0=80000.0, D_0=3000.0,
               T_0=200.0, p_y=0.5, q_y=2.0, p_v=0.5, q_v=2.0, boltzman=1.3806e-20, eps_dot_0_y=1.0e10, eps_dot_0_v=1.0e10)   # tests/golden/decks/mts_mm10.in
    crec = {k: [] for k in ("h_type", "slip_type", "angles", "iter", "R", "D6", "n_state", "params", "stress", "tt", "tt_rate", "tangent", "Rp", "euler",
                            "eps", "slip_incs", "u", "ep", "ed", "iters", "fail")}
    for slip_type in (1, 8):
        b, nrm = slip_vectors(slip_type)
        nslip = len(b)
        for case in range(7):                                   # 5, 6: MTS hardening (h_type 2)
            mts = case >= 5
            ang = rng.uniform(0.0, 360.0, 3)
            e_mod, nu = 200000.0, 0.3
            prm = dict(rate_n=20.0 if case != 3 else 7.5, theta_0=100.0 if not mts else MTS["theta_0"], tau_y=100.0, tau_v=100.0, voche_m=1.0 if case != 3 else 1.7,
                       iD_v=0.0 if case != 2 else 1e-7, e=e_mod, nu=nu)
            Sf = np.zeros((6, 6)); Sf[:3, :3] = -nu / e_mod
            Sf[np.arange(3), np.arange(3)] = 1.0 / e_mod; Sf[np.arange(3, 6), np.arange(3, 6)] = 2.0 * (1.0 + nu) / e_mod
            Cc = np.linalg.inv(Sf); Cc = 0.5 * (Cc + Cc.T)
            g = np.zeros((3, 3), order="F")
            it.call("mm10_rotation_matrix", ang.copy(), "kocks", "degrees", g, 6)
            trot = np.asfortranarray(g.T)
            RE = np.zeros((6, 6), order="F")
            it.call("mm10_rt2rve", trot, RE)
            props = NS(nslip=nslip, num_hard=1, h_type=1, out=6, alter_mode=False, eps_dot_0_y=1.0e10, k_0=0.0, burgers=2.87e-7, cp_031=0.0,
                       rate_n=prm["rate_n"], theta_0=prm["theta_0"], tau_y=prm["tau_y"], tau_v=prm["tau_v"], voche_m=prm["voche_m"],
                       id_v=prm["iD_v"], stiffness=np.asfortranarray(RE @ Cc @ RE.T), ms=np.zeros((6, ms_max), order="F"),
                       qs=np.zeros((3, ms_max), order="F"), ns=np.zeros((3, ms_max), order="F"), debug=False, gpall=False, gpp=0, solver=True,
                       strategy=True, atol=1e-5, atol1=1e-5, rtol=5e-5, rtol1=1e-5, xtol=1e-4, xtol1=1e-4, miter=30, tang_calc=0,
                       g=np.asfortranarray(g), init_angles=ang.copy(), angle_type=1, angle_convention=1)
            if mts:
                props.h_type = 2
                props.burgers, props.id_v = MTS["burgers"], 0.0
                for k_, v_ in MTS.items():
                    if k_ not in ("theta_0", "burgers"):
                        setattr(props, k_.lower(), v_)
            for s_ in range(nslip):
                bs, ns_ = trot @ b[s_], trot @ nrm[s_]
                A = np.outer(bs, ns_)
                ev, wv = np.zeros(6), np.zeros(3)
                it.call("mm10_et2ev", np.asfortranarray(0.5 * (A + A.T)), ev)
                it.call("mm10_wt2wv", np.asfortranarray(0.5 * (A - A.T)), wv)
                props.ms[:, s_], props.qs[:, s_], props.ns[:, s_] = ev, wv, ns_
            # n state: case 0 virgin (step-1 values), others a loaded, rotated state
            n = new_state()
            n.r[...] = np.eye(3); n.rp[...] = np.eye(3); n.tau_tilde[0] = prm["tau_y"] + 1.0e-5; n.euler_angles[:] = ang
            if case > 0:
                w = rng.standard_normal(3) * 0.02
                Wm = np.array([[0, w[2], w[1]], [-w[2], 0, w[0]], [-w[1], -w[0], 0.0]])
                q_, r_ = np.linalg.qr(np.eye(3) + Wm + 0.5 * Wm @ Wm)
                n.rp[...] = q_ * np.sign(np.diag(r_))[None, :]
                n.stress[:] = rng.standard_normal(6) * 80.0
                n.tau_tilde[0] = 100.0 + 15.0 * rng.random()
                n.tt_rate[0] = 0.5 * rng.random()
                n.d[:] = rng.standard_normal(6) * 1e-3
                n.eps[:] = rng.standard_normal(6) * 1e-3
            if mts:
                n.u[0] = n.u[1] = -1.0                      # tau_y / mu_harden of the n state not set yet (mm10_init_mts)
                n.tau_tilde[0] = 150.0 + 20.0 * rng.random()
            np1 = new_state()
            np1.temp = 297.0                               # mm10.f: constant temperature of this code base; n%temp = 0
            fb[:] = 0; rb[:] = 0; fb[0] = np.eye(3) + 0.02 * rng.standard_normal((3, 3))
            it.call("rtcmp1", 1, fb, rb)
            np1.r[...] = rb[0] if case > 0 else np.eye(3)
            scale = {0: 2e-3, 1: 1.5e-3, 2: 1e-3, 3: 2e-3, 4: 2.5e-2, 5: 1.5e-3, 6: 4e-3}[case]        # case 4: a 2.5 % increment, sub-stepped
            np1.d[:] = rng.standard_normal(6) * scale
            np1.tinc, np1.step, np1.iter, np1.elem, np1.gp = 1.0, 2, (0 if (case == 1 and slip_type == 1) else 1), 1, 1
            n_state = np.concatenate([n.stress, [n.tau_tilde[0], n.tt_rate[0]], n.d, n.eps, n.euler_angles, np.asarray(n.rp).ravel(), np.asarray(n.r).ravel()])
            it.calls.clear()
            res = it.call("mm10_solve_crystal", props, np1, n, False, 6, False, 1, np.zeros(6), np1.iter == 0)
            nj, nj11 = it.calls.get("mm10_formj", 0), it.calls.get("mm10_formj11", 0)
            for k, v in (("h_type", props.h_type), ("slip_type", slip_type), ("angles", ang), ("iter", np1.iter), ("R", np.asarray(np1.r).copy()), ("D6", np1.d.copy()),
                         ("n_state", n_state), ("params", [prm[q] for q in ("rate_n", "theta_0", "tau_y", "tau_v", "voche_m", "iD_v", "e", "nu")]),
                         ("stress", np1.stress.copy()), ("tt", np1.tau_tilde[0]), ("tt_rate", np1.tt_rate[0]), ("tangent", np.ascontiguousarray(np1.tangent)),
                         ("Rp", np.ascontiguousarray(np1.rp)), ("euler", np1.euler_angles.copy()), ("eps", np1.eps.copy()),
                         ("slip_incs", np1.slip_incs[:48].copy()), ("u", np1.u[:15].copy()), ("ep", np1.ep.copy()), ("ed", np1.ed.copy()),
                         ("iters", [nj11 - nj, nj]), ("fail", bool(res.get("cut", False)))):
                crec[k].append(v)
    for k, v in crec.items():
        out["crystal_" + k] = np.array(v)
    out["crystal_mts_params"] = np.array([MTS[k] for k in sorted(MTS)])
    out["crystal_mts_names"] = np.array(sorted(MTS))

    # ---- one voxel end to end, the way do_nleps_block runs it (drive_eps_sig.f:203-300): (Fn, Fn1, n state) -> kinematics ->
    #      mm10_solve_crystal -> rotation of the stress, cs2p -> P, cep2A_a -> dP/dF.  Every routine is the reference's; the few
    #      assignments between them (Fnh, dFn, t33 from the stress vector) are do_nleps_block's / cep2A's.
    it.load(open(REF + "drive_eps_sig.f").read()); it.load(open(REF + "qmply1.f").read())
    vrec = {k: [] for k in ("h_type", "slip_type", "angles", "params", "Fn", "Fn1", "n_state", "R", "uddt", "stress", "tt", "tt_rate", "tangent", "Rp",
                            "euler", "eps", "slip_incs", "P", "dPdF", "iters")}
    rng_main, rng = rng, np.random.default_rng(20240608)        # own stream: the sections below keep their inputs
    for (slip_type, mts_) in ((1, False), (1, False), (1, False), (8, False), (1, True)):
        b, nrm = slip_vectors(slip_type)
        nslip = len(b)
        case = len(vrec["Fn"])
        ang = rng.uniform(0.0, 360.0, 3)
        e_mod, nu = 200000.0, 0.3
        prm = dict(rate_n=20.0, theta_0=100.0 if not mts_ else MTS["theta_0"], tau_y=100.0, tau_v=100.0, voche_m=1.0, iD_v=0.0, e=e_mod, nu=nu)
        Sf = np.zeros((6, 6)); Sf[:3, :3] = -nu / e_mod
        Sf[np.arange(3), np.arange(3)] = 1.0 / e_mod; Sf[np.arange(3, 6), np.arange(3, 6)] = 2.0 * (1.0 + nu) / e_mod
        Cc = np.linalg.inv(Sf); Cc = 0.5 * (Cc + Cc.T)
        g = np.zeros((3, 3), order="F")
        it.call("mm10_rotation_matrix", ang.copy(), "kocks", "degrees", g, 6)
        trot = np.asfortranarray(g.T)
        RE = np.zeros((6, 6), order="F")
        it.call("mm10_rt2rve", trot, RE)
        props = NS(nslip=nslip, num_hard=1, h_type=2 if mts_ else 1, out=6, alter_mode=False, eps_dot_0_y=1.0e10, k_0=0.0, burgers=2.87e-7, cp_031=0.0,
                   rate_n=prm["rate_n"], theta_0=prm["theta_0"], tau_y=prm["tau_y"], tau_v=prm["tau_v"], voche_m=prm["voche_m"],
                   id_v=prm["iD_v"], stiffness=np.asfortranarray(RE @ Cc @ RE.T), ms=np.zeros((6, ms_max), order="F"),
                   qs=np.zeros((3, ms_max), order="F"), ns=np.zeros((3, ms_max), order="F"), debug=False, gpall=False, gpp=0, solver=True,
                   strategy=True, atol=1e-5, atol1=1e-5, rtol=5e-5, rtol1=1e-5, xtol=1e-4, xtol1=1e-4, miter=30, tang_calc=0,
                   g=np.asfortranarray(g), init_angles=ang.copy(), angle_type=1, angle_convention=1)
        if mts_:
            props.burgers = MTS["burgers"]
            for k_, v_ in MTS.items():
                if k_ not in ("theta_0", "burgers"):
                    setattr(props, k_.lower(), v_)
        for s_ in range(nslip):
            bs, ns_ = trot @ b[s_], trot @ nrm[s_]
            A = np.outer(bs, ns_)
            ev, wv = np.zeros(6), np.zeros(3)
            it.call("mm10_et2ev", np.asfortranarray(0.5 * (A + A.T)), ev)
            it.call("mm10_wt2wv", np.asfortranarray(0.5 * (A - A.T)), wv)
            props.ms[:, s_], props.qs[:, s_], props.ns[:, s_] = ev, wv, ns_
        # a loaded n state and a deformation step
        n = new_state()
        n.r[...] = np.eye(3); n.euler_angles[:] = ang
        w = rng.standard_normal(3) * 0.02
        Wm = np.array([[0, w[2], w[1]], [-w[2], 0, w[0]], [-w[1], -w[0], 0.0]])
        q_, r_ = np.linalg.qr(np.eye(3) + Wm + 0.5 * Wm @ Wm)
        n.rp[...] = q_ * np.sign(np.diag(r_))[None, :]
        n.stress[:] = rng.standard_normal(6) * 80.0
        n.tau_tilde[0] = (150.0 if mts_ else 100.0) + 15.0 * rng.random()
        n.tt_rate[0] = 0.5 * rng.random()
        n.d[:] = rng.standard_normal(6) * 1e-3
        n.eps[:] = rng.standard_normal(6) * 1e-3
        if mts_:
            n.u[0] = n.u[1] = -1.0
        amp = (0.01, 0.03, 0.002, 0.01, 0.01)[case]
        Fn = np.eye(3) + amp * rng.standard_normal((3, 3))
        Fn1 = Fn + 0.25 * amp * rng.standard_normal((3, 3))
        B3 = lambda: np.zeros((mx, 3, 3), order="F")
        fnb, fn1b, fnh, dfn, rnh, Rb, fnhinv, fn1inv = (B3() for _ in range(8))
        fnb[0], fn1b[0] = Fn, Fn1
        fnh[0] = 0.5 * (fnb[0] + fn1b[0]); dfn[0] = fn1b[0] - fnb[0]
        it.call("rtcmp1", 1, fnh, rnh); it.call("rtcmp1", 1, fn1b, Rb)
        detFh, detF = np.zeros(mx), np.zeros(mx)
        it.call("inv33", 1, 1, fnh, fnhinv, detFh)
        ddt, uddt, cs, urb = (np.zeros((mx, 6), order="F") for _ in range(4))
        it.call("mul33", 1, 1, dfn, fnhinv, ddt, 6)
        q1, q2 = np.zeros((mx, 6, 6), order="F"), np.zeros((mx, 6, 6), order="F")
        it.call("getrm1", 1, q1, rnh, 1)
        it.call("qmply1", 1, mx, 6, q1, ddt, uddt)
        np1 = new_state()
        np1.temp = 297.0
        np1.r[...] = Rb[0]; np1.d[:] = uddt[0]
        np1.tinc, np1.step, np1.iter, np1.elem, np1.gp = 1.0, 2, 1, 1, 1
        n_state = np.concatenate([n.stress, [n.tau_tilde[0], n.tt_rate[0]], n.d, n.eps, n.euler_angles, np.asarray(n.rp).ravel(), np.asarray(n.r).ravel()])
        it.calls.clear()
        res = it.call("mm10_solve_crystal", props, np1, n, False, 6, False, 1, np.zeros(6), False)
        assert not res.get("cut", False)
        nj, nj11 = it.calls.get("mm10_formj", 0), it.calls.get("mm10_formj11", 0)
        urb[0] = np1.stress
        it.call("getrm1", 1, q2, Rb, 2)
        it.call("qmply1", 1, mx, 6, q2, urb, cs)
        it.call("inv33", 1, 1, fn1b, fn1inv, detF)
        Pb = np.zeros((mx, 9), order="F")
        it.call("cs2p", 1, 1, cs, fn1inv, detF, Pb)
        t6 = np1.stress
        t33 = np.array([[t6[0], t6[3], t6[5]], [t6[3], t6[1], t6[4]], [t6[5], t6[4], t6[2]]])
        dPdF = np.zeros(81)
        it.call("cep2a_a", np.asfortranarray(Fn), np.asfortranarray(t33), np.asfortranarray(np1.tangent.copy()), np.asfortranarray(rnh[0].copy()),
                float(detFh[0]), np.asfortranarray(fnhinv[0].copy()), np.asfortranarray(Rb[0].copy()), np.asfortranarray(Fn1),
                np.asfortranarray(fn1inv[0].copy()), float(detF[0]), dPdF)
        for k, v in (("h_type", props.h_type), ("slip_type", slip_type), ("angles", ang), ("params", [prm[q] for q in ("rate_n", "theta_0", "tau_y", "tau_v", "voche_m", "iD_v", "e", "nu")]),
                     ("Fn", Fn), ("Fn1", Fn1), ("n_state", n_state), ("R", Rb[0].copy()), ("uddt", uddt[0].copy()), ("stress", np1.stress.copy()),
                     ("tt", np1.tau_tilde[0]), ("tt_rate", np1.tt_rate[0]), ("tangent", np.ascontiguousarray(np1.tangent)), ("Rp", np.ascontiguousarray(np1.rp)),
                     ("euler", np1.euler_angles.copy()), ("eps", np1.eps.copy()), ("slip_incs", np1.slip_incs[:48].copy()), ("P", Pb[0].copy()),
                     ("dPdF", dPdF.copy()), ("iters", [nj11 - nj, nj])):
            vrec[k].append(v)
    for k, v in vrec.items():
        out["voxel_" + k] = np.array(v)
    rng = rng_main

    # ---- the spectral operator itself: G_K_dF (G_K_dF.f:11-87) = ddot42n with K4 -> fftfem3d (phase ramp of formfftshift,
    #      MKL DFTI forward, split real / imaginary storage) -> two ddot42n with Ghat4 -> ifftfem3d, for odd N.  DFTI is
    #      computed by numpy's FFT (tools/fortran_subset.py), everything else is the reference's own statements.
    it.derived_factories["dfti_descriptor"] = NS
    for N in (3, 5):
        N3 = N ** 3
        G = np.zeros((N3, 81), order="F"); c1 = np.zeros((N, N, N), order="F"); c2 = np.zeros((N, N, N), order="F")
        K4 = np.asfortranarray(rng.standard_normal((N3, 81)) * 1e4 + 1e5 * np.eye(9).reshape(81)[None, :])
        it3 = F.Interpreter(it.consts); it3.units = it.units; it3.derived_factories = it.derived_factories
        it3.module_vars = dict(n=N, nhalf=(N + 1) // 2, n3=N3, ndim1=3, ndim2=9, ghat4=G, k4=K4, dims=np.array([N, N, N]), coeffs1=c1, coeffs2=c2,
                               real1=np.zeros((N3, 9), order="F"), real2=np.zeros((N3, 9), order="F"), real3=np.zeros((N3, 9), order="F"))
        it3.call("formg"); it3.call("formfftshift", c1, c2)
        Fm = np.asfortranarray(rng.standard_normal((N3, 9)) * 1e-3)
        res = []
        for flg in (True, False):
            GKF = np.zeros((N3, 9), order="F")
            it3.call("g_k_df", Fm, GKF, flg)
            res.append(np.ascontiguousarray(GKF))
        out[f"GKdF_{N}_K4"], out[f"GKdF_{N}_F"] = np.ascontiguousarray(K4), np.ascontiguousarray(Fm)
        out[f"GKdF_{N}_with_K4"], out[f"GKdF_{N}_without_K4"] = res
        out[f"fftshift_{N}"] = np.array([np.ascontiguousarray(c1), np.ascontiguousarray(c2)])

    # ---- the kinematics of do_nleps_block around the material call (drive_eps_sig.f:203-300): the call sequence is the
    #      block driver's, every routine is the reference's: rtcmp1 (Fnh, Fn1), inv33, mul33, getrm1 opt 1, qmply1 -> uddt;
    #      getrm1 opt 2, qmply1, inv33, cs2p -> P
    it.load(open(REF + "drive_eps_sig.f").read()); it.load(open(REF + "qmply1.f").read())
    kin = {k: [] for k in ("amp", "Fn", "Fn1", "ur6", "R", "uddt", "P", "detF")}
    for amp in (1e-3, 1e-2, 0.1, 0.3):
        for _ in range(3):
            Fn = rand_rotation(rng) @ (np.eye(3) + amp * rng.standard_normal((3, 3))) if amp >= 0.1 else np.eye(3) + amp * rng.standard_normal((3, 3))
            Fn1 = Fn + 0.3 * amp * rng.standard_normal((3, 3))
            ur6 = 200.0 * rng.standard_normal(6)
            B3 = lambda: np.zeros((mx, 3, 3), order="F")
            fnb, fn1b, fnh, dfn, rnh, Rb, fnhinv, fn1inv = (B3() for _ in range(8))
            fnb[0], fn1b[0] = Fn, Fn1
            fnh[0] = 0.5 * (fnb[0] + fn1b[0]); dfn[0] = fn1b[0] - fnb[0]
            it.call("rtcmp1", 1, fnh, rnh); it.call("rtcmp1", 1, fn1b, Rb)
            detFh, detF = np.zeros(mx), np.zeros(mx)
            it.call("inv33", 1, 1, fnh, fnhinv, detFh)
            ddt, uddt, cs, urb = (np.zeros((mx, 6), order="F") for _ in range(4))
            it.call("mul33", 1, 1, dfn, fnhinv, ddt, 6)
            q1, q2 = np.zeros((mx, 6, 6), order="F"), np.zeros((mx, 6, 6), order="F")
            it.call("getrm1", 1, q1, rnh, 1)
            it.call("qmply1", 1, mx, 6, q1, ddt, uddt)
            urb[0] = ur6
            it.call("getrm1", 1, q2, Rb, 2)
            it.call("qmply1", 1, mx, 6, q2, urb, cs)
            it.call("inv33", 1, 1, fn1b, fn1inv, detF)
            Pb = np.zeros((mx, 9), order="F")
            it.call("cs2p", 1, 1, cs, fn1inv, detF, Pb)
            for k, v in (("amp", amp), ("Fn", Fn), ("Fn1", Fn1), ("ur6", ur6), ("R", Rb[0].copy()), ("uddt", uddt[0].copy()), ("P", Pb[0].copy()), ("detF", detF[0])):
                kin[k].append(v)
    for k, v in kin.items():
        out["kin_" + k] = np.array(v)

    # ---- mm01 (bilinear Mises plasticity, mixed hardening) + cnst1: a four-increment path (elastic, plastic, plastic in
    #      another direction, unloading) on 6 points with beta = 0, 0.5, 1; the history carries the packed state word
    #      (equivalence of a double and two integers, mm01.f:253-255)
    it.load(open(REF + "mm01.f").read())
    npt = 6
    ym = np.zeros(mx); nuv = np.zeros(mx); beta = np.zeros(mx); hp = np.zeros(mx); yld = np.zeros(mx)
    ym[:npt], nuv[:npt], yld[:npt] = 69000.0, 0.33, 100.0
    beta[:npt] = [0.0, 0.5, 1.0, 0.0, 0.5, 1.0]
    tan_e = 1000.0
    hp[:npt] = tan_e * 69000.0 / (69000.0 - tan_e)
    lnelas = np.zeros(mx)
    cgn = np.zeros((mx, 9), order="F"); hist = np.zeros((npt, 11), order="F")
    dirs = rng.standard_normal((4, npt, 6))
    amps = [4e-4, 3e-3, 2e-3, -1.5e-3]
    m01 = {k: [] for k in ("deps", "cgn", "hist", "cgn1", "hist1", "cep")}
    for step in range(1, 5):
        deps = np.zeros((mx, 6), order="F"); deps[:npt] = amps[step - 1] * dirs[step - 1] * np.array([1, 1, 1, 2, 2, 2.0])
        cgn1 = np.zeros((mx, 9), order="F"); hist1 = np.zeros((npt, 11), order="F"); rtse = np.zeros((mx, 6), order="F")
        it.call("mm01", npt, 1, 1, step, 1, ym, nuv, beta, hp, lnelas, yld, cgn, cgn1, deps, hist, hist1, rtse, np.zeros(mx), ym, nuv, 6)
        cep = np.zeros((mx, 6, 6), order="F")
        it.call("cnst1", npt, cep, rtse, nuv, ym, np.asfortranarray(hist1[:, 1].copy()), np.asfortranarray(hist1[:, 4].copy()), beta,
                np.asfortranarray(hist1[:, 0].copy()), np.asfortranarray(hist1[:, 3].copy()), 1, 6)
        for k, v in (("deps", deps[:npt]), ("cgn", cgn[:npt]), ("hist", hist), ("cgn1", cgn1[:npt]), ("hist1", hist1), ("cep", cep[:npt])):
            m01[k].append(np.ascontiguousarray(v).copy())
        cgn, hist = cgn1, hist1
    for k, v in m01.items():
        out["mm01_" + k] = np.array(v)
    out["mm01_props"] = np.array([69000.0, 0.33, 100.0, hp[0]])
    out["mm01_beta"] = beta[:npt].copy()

    prov = "; ".join(f"{f} sha256 {hashlib.sha256(open(REF + f, 'rb').read()).hexdigest()[:16]}" for f in FILES)
    out["provenance"] = np.array("maranGit/CPFFT src: " + prov + "; executed by tools/fortran_subset.py (tools/make_reference_vectors.py)")
    dst = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
