#!/usr/bin/env python
"""Whole jobs run by the reference's OWN source: the global Newton loop FFT_nr3 with its stress-controlled outer loop, fftPcg,
tangent_homo, NBC_update, the operator G_K_dF and -- inside the Python stand-in for the block driver -- the material routines
(the crystal-plasticity wrapper mm10 with everything below it, or mm01 + cnst1), executed statement by statement by the
Fortran-subset interpreter tools/fortran_subset.py.  Output: tests/golden/reference_global.npz.

    python tools/make_reference_global.py     # needs /root/reference (this container); about twenty-five minutes
    GLOBAL_ONLY=deck,deckmts python ...       # regenerate single jobs, keep the rest of the fixture; GLOBAL_DECK_STEPS=0: no deck jobs

Jobs (key prefix in the fixture):
  ""            3^3 fcc polycrystal, Voce hardening, uniaxial tension under mixed boundary conditions (F_xx, P_yy = P_zz = 0), 3 steps
  m01_          3^3, mm01 + cnst1 (mm01.f) in drive_01_update's sequence (rstgp1.f:330-450), strain-controlled, 3 steps
  deck_         the shipped deck examples/test_mm10.in as it stands (7^3, bcc48, alter_mode on, three blocks), all ten load steps
  deck01_       the shipped deck examples/test_mm01.in as it stands (7^3, two bilinear materials), all ten load steps
  deck01nbc_, deck10nbc_   the shipped decks with F_xx driven and P_yy = P_zz = 0, 3 / 2 steps
  deckmts_      tests/golden/decks/mts_mm10.in (5^3, MTS hardening), 4 steps
  decktaylor_   tests/golden/decks/taylor_mm10.in (5^3, bcc48 + fcc crystal per point from a flat file, Taylor average), 5 steps
  wrap_*        the block-driver sequence and mm10 on polycrystalline points (three crystals per point), MTS hardening and the
                48-system layout, two load steps each, with mm10_set_history_locs' layout tables

tests/test_reference_global.py (which needs neither /root/reference nor this script) holds the oracle's solver and the
kernel source to it.  What is executed from the reference (file:line of the subroutine statement):

  FFT_nr3.f:14    FFT_nr3      load steps, boundary-condition split, the Newton loop on the fluctuation, the outer loop on
                                the prescribed mean stress, convergence tests, the commit of a step
  FFT_nr3.f:214   fftPcg       right-hand side norm, tolerance checks, the RCI loop and its stopping test
  FFT_nr3.f:375   NBC_update   the mixed 9 x 9 system for the mean deformation gradient (DGESV)
  tangent_homo.f:11 tangent_homo  nine CG solves for d(dF)/d(Fbar), C_homo = <K4 : dF/dFbar>
  G_K_dF.f:11     G_K_dF, ddot42n, fftfem3d, ifftfem3d; FFT_init.f:272 formG, :355 formfftshift
  mm10_a.f:28     mm10         the per-point wrapper: history initialisation (mm10_init_general_hist, _uout_hist, _slip_hist,
                                mm10_init_cc_hist0), mm10_init_cc_props, mm10_copy_cc_hist, mm10_setup_np1,
                                mm10_solve_crystal (everything tools/make_reference_vectors.py lists), the crystal averages,
                                mm10_store_cryhist and the common history block
  polar.f:18 rtcmp1, :680 getrm1; drive_eps_sig.f:1017 inv33, :1110 mul33, :1182 cs2p; qmply1.f:15 qmply1; cep2A.f:14 cep2A

What is NOT the reference's text, and why:
  * MKL's reverse-communication CG (dcg_init / dcg_check / dcg / dcg_get) is a closed library: restated here from its
    documented algorithm (plain CG on tmp(:,1..3) = direction, A x direction, residual; user stopping test, ipar(10) = 1).
  * do_nleps_block / rstgp1 / dupstr_blocked / rplstr (drive_eps_sig.f:16-330, rstgp1.f) -- the gather / scatter between the
    global arrays and the 128-point block work space -- are replaced by `drive_eps_sig` below, which calls the reference
    routines listed above in do_nleps_block's order on blocks of at most 128 points.  drive_10_cnst's copy of history(1:36) into the
    block tangent (gptns1.f:519-567) and update.f's n+1 -> n copies are one assignment each, done here.  rknstr_finish_cp (the
    lattice-curvature fit) is not run: its output only enters through k_0, which is zero.
  * mm10_set_cons returns at once for Voce / MTS hardening (mm10_a.f:383-388) and is skipped.  The history layout is computed by
    the reference's mm10_set_history_locs (mm10_d.f:25-331) from hand-filled material tables (one CP material; one crystal type, or
    the crystal numbers per point of `crystal_input file`).
  * MKL DFTI is numpy's FFT (tools/fortran_subset.py).
"""
import hashlib
import os
import sys
import time
from types import SimpleNamespace as NS

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import fortran_subset as F  # noqa: E402
import make_reference_vectors as G  # noqa: E402

REF = G.REF
FILES = G.FILES + ["FFT_nr3.f", "tangent_homo.f", "mm10_d.f"]


class Defaulting(NS):
    """crystal_properties as the deck reader leaves it: every parameter the deck does not set is zero"""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return 0.0


def mkl_rci_cg():
    """MKL RCI CG (dcg_init, dcg_check, dcg, dcg_get), restated from the library's documented scheme: x_0 given, r = b - A x_0,
    p = r, then alpha = (r.r)/(p.Ap), x += alpha p, r -= alpha Ap, beta = (r.r)_new/(r.r)_old, p = r + beta p; RCI_request 1 asks
    for tmp(:,2) = A tmp(:,1), RCI_request 2 for the user's stopping test on tmp(:,3) = r (ipar(10) = 1, ipar(9) = 0), ipar(4) counts
    the iterations."""
    st = {}

    def dcg_init(n, x, b, rci, ipar, dpar, tmp):
        st.clear(); st.update(stage=0, it=0, rho=0.0)
        return 0

    def dcg_check(n, x, b, rci, ipar, dpar, tmp):
        return 0

    def dcg(n, x, b, rci, ipar, dpar, tmp):
        n = int(n)
        p, q, r = tmp[0:n], tmp[n:2 * n], tmp[2 * n:3 * n]
        if st["stage"] == 0:                      # residual of the initial guess
            p[:] = x[:n]; st["stage"] = 1
            return 1
        if st["stage"] == 1:
            r[:] = b[:n] - q; st["stage"] = 2
            return 2
        if st["stage"] == 2:                      # the user's test did not stop the iteration: next direction
            rho = float(np.dot(r, r))
            if st["it"] == 0:
                p[:] = r
            else:
                p[:] = r + (rho / st["rho"]) * p
            st["rho"] = rho; st["stage"] = 3
            return 1
        alpha = st["rho"] / float(np.dot(p, q))
        x[:n] = x[:n] + alpha * p
        r[:] = r - alpha * q
        st["it"] += 1; ipar[3] = st["it"]; st["stage"] = 2
        if ipar[7] != 0 and st["it"] > ipar[4]:
            return -1
        return 2

    def dcg_get(n, x, b, rci, ipar, dpar, tmp, itercount):
        return st["it"]
    return st, dcg_init, dcg_check, dcg, dcg_get


MTS = dict(theta_0=1500.0, tau_a=20.0, tau_hat_y=180.0, g_0_y=0.4, tau_hat_v=300.0, g_0_v=1.2, burgers=2.5e-7, mu_0=80000.0, D_0=3000.0,
           T_0=200.0, p_y=0.5, q_y=2.0, p_v=0.5, q_v=2.0, boltzman=1.3806e-20, eps_dot_0_y=1.0e10, eps_dot_0_v=1.0e10)   # tests/golden/decks/mts_mm10.in
MTS_FIELD = dict(theta_0="theta_o", tau_hat_y="tauhat_y", g_0_y="go_y", tau_hat_v="tauhat_v", g_0_v="go_v", mu_0="mu_o", D_0="d_o", T_0="t_o",
                 eps_dot_0_y="eps_dot_o_y", eps_dot_0_v="eps_dot_o_v")         # crystal_properties spells them differently (mm10_a.f:1822-1845)
PRM = dict(rate_n=20.0, theta_0=100.0, tau_y=100.0, tau_v=100.0, voche_m=1.0, iD_v=0.0, e=200000.0, nu=0.3)
Z = lambda *s: np.zeros(s, order="F")


def _die(*a):
    raise RuntimeError("die_abort")


class Harness:
    """The material stage of npts points, one block-driver call at a time: the reference's routines in do_nleps_block's order
    (drive_eps_sig.f:203-300), with its own interpreter because the history layout is module data.  slip_type 1 (fcc, 12
    systems) or 8 (bcc48, which selects the maximum-size layout, mm10_d.f:136-141)."""

    def __init__(self, npts, angles, slip_type=1, mts=False, extra_module_vars=None, files=FILES, mm01=None, prm=None, dt=1.0, types=None, ids=None):
        """types / ids: several crystal types (list of dict(slip_type=, prm=)) and the 1-based type of every crystal of every point
        (`crystal_input file`); otherwise one type (slip_type, prm, mts) for all"""
        self.npts, self.slip_type, self.mts, self.mm01 = npts, slip_type, mts, mm01
        PRM = {**globals()["PRM"], "alter_mode": False, "eps_dot_0_y": 1.0e10, **(prm or {})}
        if angles is None:
            angles = np.zeros((npts, 1, 3))
        it = self.it = F.Interpreter()
        it.add_constants(open(REF + "param_def").read())
        mc_src = open(REF + "mod_crystals.f").read()
        i0 = mc_src.index("      module mm10_constants")
        it.add_constants(mc_src[i0:mc_src.index("      end module", i0)])
        it.consts.setdefault("out", 6)
        mx, ms_max, mu_ = it.consts["mxvl"], it.consts["max_slip_sys"], it.consts["max_uhard"]
        self.mx = mx
        angles = np.asarray(angles, dtype=np.float64).reshape(npts, -1, 3)
        self.ncry = ncry = angles.shape[1]
        from oracle import Oracle
        if types is None:
            types, ids = [dict(slip_type=slip_type, prm=prm)], np.ones((npts, ncry), dtype=np.int64)
        tables = [Oracle.slip_table(t["slip_type"]) for t in types]
        nslip = self.nslip = max(len(t[0]) for t in tables)

        # module mm10_defs: the history layout, computed by the reference's own mm10_set_history_locs (mm10_d.f:25-331) from the
        # material / crystal tables a deck with one crystal-plasticity material of `ncry` crystals per point leaves behind
        mcr = it.consts["max_crystals"]
        mm10_defs = dict(indexes_common=np.zeros((5, 2), dtype=np.int64, order="F"), index_crys_hist=np.zeros((mcr, 11, 2), dtype=np.int64, order="F"),
                         length_comm_hist=np.zeros(5, dtype=np.int64), length_crys_hist=np.zeros(11, dtype=np.int64), num_common_indexes=0,
                         num_crystal_terms=0, one_crystal_hist_size=0, common_hist_size=0, asymmetric_assembly=False)
        matprp = np.zeros((300, 500), order="F"); matprp[8, 0] = 10                      # matprp(9, 1): material model 10
        imatprp = np.zeros((300, 500), dtype=np.int64, order="F")
        imatprp[100, 0], imatprp[103, 0], imatprp[104, 0] = ncry, (1 if len(types) == 1 else 2), 1   # imatprp(101 / 104 / 105, 1): crystals per point,
        c_array = np.empty(mcr, dtype=object)                                                        # crystal type single / from file, its number
        for k_, tb in enumerate(tables):
            c_array[k_] = NS(nslip=len(tb[0]), num_hard=1)
        tables_mod = dict(matprp=matprp, imatprp=imatprp, matlist=np.ones(npts, dtype=np.int64), c_array=c_array,
                          crystal_input=np.asfortranarray(np.asarray(ids, dtype=np.int64)), data_offset=np.arange(1, npts + 1, dtype=np.int64))
        it.consts.update(nummat=1, noelem=npts)
        it.module_vars.update(mm10_defs); it.module_vars.update(tables_mod)
        it.load(open(REF + "mm10_d.f").read())
        it.call("mm10_set_history_locs")
        for k_ in tables_mod:
            del it.module_vars[k_]
        ic, lcr = it.module_vars["indexes_common"], it.module_vars["length_crys_hist"]
        self.hist_sz = hist_sz = int(ic[4, 1] + ncry * it.module_vars["one_crystal_hist_size"])
        it.module_members["mm10_defs"] = set(mm10_defs)
        if extra_module_vars:
            it.module_vars.update(extra_module_vars); it.module_members["fft"] = set(extra_module_vars)
        for f in files[2:]:
            if f != "mm10_d.f":
                it.load(open(REF + f).read())
        del it.units["mm10_set_cons"]                            # returns at once for Voce / MTS (mm10_a.f:383-388); its allocate(..., stat=) is not interpretable
        F.BUILTIN_SUBS.setdefault("mm10_set_cons", lambda *a: None); F.BUILTIN_ARRAY_ARGS.setdefault("mm10_set_cons", ())
        F.BUILTIN_SUBS.setdefault("die_abort", _die); F.BUILTIN_ARRAY_ARGS.setdefault("die_abort", ())
        F.BUILTIN_SUBS.setdefault("die_gracefully", _die); F.BUILTIN_ARRAY_ARGS.setdefault("die_gracefully", ())
        for name in ("drive_eps_sig", "update"):
            it.units.pop(name, None)

        def new_state():
            return NS(r=Z(3, 3), rp=Z(3, 3), stress=np.zeros(6), d=np.zeros(6), eps=np.zeros(6), euler_angles=np.zeros(3), slip_incs=np.zeros(ms_max),
                      tau_tilde=np.zeros(mu_), tt_rate=np.zeros(mu_), u=np.zeros(mu_), ep=np.zeros(6), ed=np.zeros(6), tangent=Z(6, 6), ms=Z(6, ms_max),
                      qs=Z(3, ms_max), qc=Z(3, ms_max), tau_l=np.zeros(ms_max), gradfeinv=Z(3, 3, 3), dg=0.0, tinc=0.0, temp=0.0, mu_harden=0.0,
                      work_inc=0.0, p_work_inc=0.0, p_strain_inc=0.0, step=0, elem=0, iter=0, gp=0, tau_v=0.0, tau_y=0.0)
        it.derived_factories["crystal_state"] = new_state
        it.derived_factories["crystal_props"] = lambda: NS(g=Z(3, 3), ms=Z(6, ms_max), qs=Z(3, ms_max), ns=Z(3, ms_max), stiffness=Z(6, 6),
                                                           st_it=np.zeros(3, dtype=np.int64), init_angles=np.zeros(3), out=6)
        it.derived_factories["dfti_descriptor"] = NS

        # crystal_properties the way setup_mm10_rknstr fills it (drive_eps_sig.f:571-606, 975-986) with the reference's own
        # mm10_rotation_matrix, mm10_RT2RVE, mm10_ET2EV, mm10_WT2WV; every parameter the deck does not set is zero
        base_prm = PRM
        self.c_props = c_props = np.empty((npts, ncry), dtype=object)
        for e in range(npts):
            for c in range(ncry):
                ty = types[int(ids[e][c]) - 1]
                PRM = {**base_prm, **(ty.get("prm") or {})}
                slip_type_c = ty["slip_type"]
                bvec, nvec = tables[int(ids[e][c]) - 1]
                nslip = len(bvec)
                e_mod, nu = PRM["e"], PRM["nu"]
                Sf = np.zeros((6, 6)); Sf[:3, :3] = -nu / e_mod
                Sf[np.arange(3), np.arange(3)] = 1.0 / e_mod; Sf[np.arange(3, 6), np.arange(3, 6)] = 2.0 * (1.0 + nu) / e_mod
                Cc = np.linalg.inv(Sf); Cc = 0.5 * (Cc + Cc.T)
                g = Z(3, 3)
                it.call("mm10_rotation_matrix", angles[e, c].copy(), "kocks", "degrees", g, 6)
                trot = np.asfortranarray(g.T)
                RE = Z(6, 6)
                it.call("mm10_rt2rve", trot, RE)
                cp = Defaulting(raten=PRM["rate_n"], theta_o=PRM["theta_0"], tau_y=PRM["tau_y"], tau_v=PRM["tau_v"], voche_m=PRM["voche_m"], id_v=PRM["iD_v"],
                                burgers=2.87e-7, eps_dot_o_y=PRM["eps_dot_0_y"], solver=True, strategy=True, gpall=False, gpp=0, method=0, miter=30, atol=1e-5,
                                atol1=1e-5, rtol=5e-5, rtol1=1e-5, xtol=1e-4, xtol1=1e-4, alter_mode=bool(PRM["alter_mode"]), nslip=nslip, h_type=2 if mts else 1, num_hard=1,
                                tang_calc=0, s_type=slip_type_c, cnum=int(ids[e][c]), st_it=np.zeros(3, dtype=np.int64), rotation_g=np.asfortranarray(g), ms=Z(6, ms_max),
                                qs=Z(3, ms_max), ns=Z(3, ms_max), init_elast_stiff=np.asfortranarray(RE @ Cc @ RE.T), init_angles=angles[e, c].copy())
                if mts:
                    for k_, v_ in (mts if isinstance(mts, dict) else MTS).items():
                        setattr(cp, MTS_FIELD.get(k_, k_), v_)
                for s_ in range(nslip):
                    bs, ns_ = trot @ bvec[s_], trot @ nvec[s_]
                    A = np.outer(bs, ns_)
                    ev, wv = np.zeros(6), np.zeros(3)
                    it.call("mm10_et2ev", np.asfortranarray(0.5 * (A + A.T)), ev)
                    it.call("mm10_wt2wv", np.asfortranarray(0.5 * (A - A.T)), wv)
                    cp.ms[:, s_], cp.qs[:, s_], cp.ns[:, s_] = ev, wv, ns_
                c_props[e, c] = cp

        # the block work space and the global state the block driver gathers from / scatters to
        mk = lambda span: NS(dt=float(dt), blk=1, span=span, felem=1, gpn=1, step=1, iter=0, iout=6, mat_type=10, material_cut_step=False,
                             debug_flag=np.zeros(mx, dtype=bool), c_props=np.empty((mx, ncry), dtype=object), angle_type=np.ones(mx, dtype=np.int64),
                             angle_convention=np.ones(mx, dtype=np.int64), fn=Z(mx, 3, 3), fn1=Z(mx, 3, 3), urcs_blk_n=Z(mx, 9, 1),
                             urcs_blk_n1=Z(mx, 9, 1), rot_blk_n1=Z(mx, 9, 1))
        self.lw, self.lw1 = mk(min(npts, mx)), mk(1)
        self.urcs_n, self.urcs_n1 = np.zeros((npts, 9)), np.zeros((npts, 9))
        if mm01 is not None:                      # mm01: 11 history words per point (mm01.f:219-262), properties per point
            self.hist_sz = hist_sz = 11
            pad = lambda a: np.asarray(a, dtype=np.float64)
            self.m1 = NS(e=pad(mm01["e"]), nu=pad(mm01["nu"]), beta=pad(mm01["beta"]), yld=pad(mm01["yld"]),
                         h=pad(np.asarray(mm01["tan_e"]) * np.asarray(mm01["e"]) / (np.asarray(mm01["e"]) - np.asarray(mm01["tan_e"]))))
            self.eps_n, self.eps_n1 = Z(npts, 6), Z(npts, 6)
        self.hist_n, self.hist_n1 = Z(npts, hist_sz), Z(npts, hist_sz)
        self.h_n, self.h_n1, self.u1 = Z(1, hist_sz), Z(1, hist_sz), Z(mx, 6)
        self.ncrystals = np.full(mx, ncry, dtype=np.int64)
        self.sweeps = []

    def sweep(self, step, iter_, Fn, Fn1):
        """(npts, 9) row-major F_n, F_n+1 -> P (npts, 9), dP/dF (npts, 81); the n+1 history and stresses are kept in the harness.
        Blocks of at most mxvl = 128 points, as the reference's automatic blocking makes them (FFT_init.f:449-475)."""
        Fn, Fn1 = np.asarray(Fn), np.asarray(Fn1)
        P, K = np.zeros((self.npts, 9)), np.zeros((self.npts, 81))
        self.hist_n1[...] = 0.0; self.urcs_n1[...] = 0.0
        nj0, nj110 = self.it.calls.get("mm10_formj", 0), self.it.calls.get("mm10_formj11", 0)
        for e0 in range(0, self.npts, self.mx):
            span = min(self.mx, self.npts - e0)
            P[e0:e0 + span], K[e0:e0 + span] = self._sweep_block(step, iter_, Fn[e0:e0 + span], Fn1[e0:e0 + span], e0, span)
        nj, nj11 = self.it.calls.get("mm10_formj", 0) - nj0, self.it.calls.get("mm10_formj11", 0) - nj110
        self.sweeps.append((int(step), int(iter_), nj11 - nj, nj))
        return P, K

    def _sweep_block(self, step, iter_, Fn, Fn1, e0, span):
        it, lw, lw1, mx = self.it, self.lw, self.lw1, self.mx
        lw.step, lw.iter, lw.material_cut_step, lw.span, lw.felem = int(step), int(iter_), False, span, e0 + 1
        lw.fn[...] = 0.0; lw.fn1[...] = 0.0
        lw.fn[:span] = Fn.reshape(span, 3, 3)                            # Fn(e, 1..9) = F11, F12, F13, F21, ... (drive_eps_sig.f:190-214)
        lw.fn1[:span] = Fn1.reshape(span, 3, 3)
        lw.urcs_blk_n[...] = 0.0; lw.urcs_blk_n[:span, :, 0] = self.urcs_n[e0:e0 + span]
        fnh, dfn, rnh, fnhinv, fn1inv = (Z(mx, 3, 3) for _ in range(5))
        fnh[:span] = 0.5 * (lw.fn[:span] + lw.fn1[:span]); dfn[...] = lw.fn1 - lw.fn
        it.call("rtcmp1", span, fnh, rnh); it.call("rtcmp1", span, lw.fn1, lw.rot_blk_n1)
        detFh, detF = np.zeros(mx), np.zeros(mx)
        it.call("inv33", span, 1, fnh, fnhinv, detFh)
        ddt, uddt, cs = Z(mx, 6), Z(mx, 6), Z(mx, 6)
        it.call("mul33", span, 1, dfn, fnhinv, ddt, 6)
        qnhalf, qtn1 = Z(mx, 6, 6), Z(mx, 6, 6)
        it.call("getrm1", span, qnhalf, rnh, 1)
        it.call("qmply1", span, mx, 6, qnhalf, ddt, uddt)
        lw.urcs_blk_n1[...] = 0.0
        if self.mm01 is not None:          # drive_01_update (rstgp1.f:330-450): total strain, mm01, cnst1, [D] kept as its upper triangle
            m1 = self.m1
            blk = lambda a: np.concatenate([a[e0:e0 + span], np.zeros(mx - span)])
            e_, nu_, beta_, h_, yld_ = blk(m1.e), blk(m1.nu), blk(m1.beta), blk(m1.h), blk(m1.yld)
            self.eps_n1[e0:e0 + span] = self.eps_n[e0:e0 + span] + uddt[:span]        # rstgp1_update_strains
            cgn1, rtse, cep = Z(mx, 9), Z(mx, 6), Z(mx, 6, 6)
            hn, h1 = np.asfortranarray(self.hist_n[e0:e0 + span]), Z(span, 11)
            it.call("mm01", span, e0 + 1, 1, int(step), int(iter_), e_, nu_, beta_, h_, np.zeros(mx), yld_, lw.urcs_blk_n, cgn1, uddt, hn, h1, rtse,
                    np.zeros(mx), e_, nu_, 6)
            lw.urcs_blk_n1[:, :, 0] = cgn1
            self.hist_n1[e0:e0 + span] = h1
            self.hist_n[e0:e0 + span] = hn                               # step 1 initialises the n history in place (mm01_set_history)
            it.call("cnst1", span, cep, rtse, nu_, e_, np.asfortranarray(h1[:, 1].copy()), np.asfortranarray(h1[:, 4].copy()), beta_,
                    np.asfortranarray(h1[:, 0].copy()), np.asfortranarray(h1[:, 3].copy()), e0 + 1, 6)
            for i in range(span):                                         # rstgp1_store_cep keeps 21 terms, drive_01_cnst mirrors them
                cep[i] = np.triu(cep[i]) + np.triu(cep[i], 1).T
            self.cep_mm01 = cep
        for e in range(span if self.mm01 is None else 0):              # one-point blocks: mm10 addresses the history through history(iloop, 1)
            g = e0 + e                                                   # with an assumed-size dummy, which for span > 1 runs past whole columns
            lw1.step, lw1.iter, lw1.felem, lw1.material_cut_step = lw.step, lw.iter, g + 1, False
            lw1.c_props[0, :] = self.c_props[g, :]
            lw1.rot_blk_n1[0] = lw.rot_blk_n1[e]; lw1.urcs_blk_n[0] = lw.urcs_blk_n[e]; lw1.urcs_blk_n1[...] = 0.0
            self.u1[0] = uddt[e]; self.h_n[0] = self.hist_n[g]; self.h_n1[...] = 0.0
            it.call("mm10", 1, 1, self.ncrystals, self.hist_sz, self.h_n, self.h_n1, lw1, self.u1, np.full(mx, 297.0), np.zeros(mx), 6, False, False,
                    Z(mx, 1), 1, int(iter_) == 0)                        # rstgp1.f:862-880: iteration 0 is always the linear-elastic estimate
            lw.material_cut_step = lw.material_cut_step or lw1.material_cut_step
            self.hist_n1[g] = self.h_n1[0]; lw.urcs_blk_n1[e] = lw1.urcs_blk_n1[0]
            if step == 1:
                self.hist_n[g] = self.h_n[0]                             # step 1 initialises the n history in place (mm10_a.f:73-78, 237-244)
        if lw.material_cut_step:
            raise RuntimeError("material_cut_step")
        self.urcs_n1[e0:e0 + span] = lw.urcs_blk_n1[:span, :, 0]
        it.call("getrm1", span, qtn1, lw.rot_blk_n1, 2)
        it.call("qmply1", span, mx, 6, qtn1, lw.urcs_blk_n1, cs)
        it.call("inv33", span, 1, lw.fn1, fn1inv, detF)
        P_blk, A_blk, cep = Z(mx, 9), Z(mx, 81), Z(mx, 6, 6)
        it.call("cs2p", span, 1, cs, fn1inv, detF, P_blk)
        for i in range(span):                                             # drive_10_cnst, gptns1.f:562-567
            cep[i] = self.hist_n1[e0 + i, 0:36].reshape(6, 6, order="F") if self.mm01 is None else self.cep_mm01[i]
        it.call("cep2a", lw, cep, rnh, detF, detFh, fnhinv, fn1inv, A_blk)
        return P_blk[:span].copy(), A_blk[:span].copy()

    def update(self):
        self.hist_n[...] = self.hist_n1; self.urcs_n[...] = self.urcs_n1            # update.f:85-93
        if self.mm01 is not None:
            self.eps_n[...] = self.eps_n1


def wrapper_cases(out):
    """the per-point wrapper mm10 on what the 3^3 job does not reach: polycrystalline points (three crystals, Taylor average,
    mm10_a.f:112-197), MTS hardening with its history (mm10_init_mts, tau_y / mu_harden carried in u(1:2)), and the 48-system
    layout.  Two load steps each: step 1 from the virgin state (iteration 0 = elastic estimate, then a plastic sweep), commit,
    step 2."""
    rng = np.random.default_rng(20240610)
    for name, slip_type, mts, ncry in (("taylor", 1, False, 3), ("mts", 1, True, 1), ("bcc48", 8, False, 1)):
        npts = 4
        angles = rng.uniform(0.0, 360.0, (npts, ncry, 3))
        H = Harness(npts, angles, slip_type=slip_type, mts=mts)
        eye = np.tile(np.eye(3).reshape(9), (npts, 1))
        F1 = eye + 0.0015 * rng.standard_normal((npts, 9))
        F2 = F1 + 0.0012 * rng.standard_normal((npts, 9))
        P0, K0 = H.sweep(1, 0, eye, eye)
        mv = H.it.module_vars
        rec = dict(angles=angles, F1=F1, F2=F2, K4_initial=K0, hist_size=H.hist_sz, slip_type=slip_type, h_type=2 if mts else 1, ncry=ncry,
                   layout_common=mv["indexes_common"].copy(), layout_crystal=mv["index_crys_hist"][:ncry].copy(),      # mm10_set_history_locs' tables
                   layout_sizes=np.array([mv["common_hist_size"], mv["one_crystal_hist_size"]]))
        for step, (Fa, Fb) in ((1, (eye, F1)), (2, (F1, F2))):
            P, K = H.sweep(step, 1, Fa, Fb)
            rec[f"P{step}"], rec[f"K4_{step}"] = P, K
            rec[f"hist{step}"] = np.ascontiguousarray(H.hist_n1).copy()
            rec[f"urcs{step}"] = H.urcs_n1.copy()
            rec[f"iters{step}"] = np.array(H.sweeps[-1][2:])
            H.update()
        for k, v in rec.items():
            out[f"wrap_{name}_{k}"] = np.asarray(v)
        print("wrapper case", name, "hist", H.hist_sz, "iters", rec["iters1"], rec["iters2"], flush=True)
    out["mts_names"] = np.array(sorted(MTS)); out["mts_params"] = np.array([MTS[k] for k in sorted(MTS)])


def run_job(out, prefix, N, nstep, FP_max, isNBC, make_harness, t_start, mults=None, maxiter=10):
    """FFT_init's state (FFT_init.f:141-172), the initial sweep and FFT_nr3 (FFT_finite_3d.f:145-146) for one job; results
    under `prefix` in `out`"""
    N3 = N ** 3
    fft = dict(n=N, nhalf=(N + 1) // 2, n3=N3, ndim1=3, ndim2=9, veclen=9 * N3, dims=np.array([N, N, N]), ghat4=Z(N3, 81), k4=Z(N3, 81),
               coeffs1=Z(N, N, N), coeffs2=Z(N, N, N), real1=Z(N3, 9), real2=Z(N3, 9), real3=Z(N3, 9), b=Z(N3, 9), fn=Z(N3, 9), fn1=Z(N3, 9),
               pn=Z(N3, 9), pn1=Z(N3, 9), dfm=Z(N3, 9), tmppcg=Z(9 * N3, 4), isnbc=np.asarray(isNBC, dtype=bool), bc_all=Z(9, nstep),
               straininc=0.0, tolpcg=1.0e-10, tolnr=1.0e-5, maxiter=maxiter, nstep=nstep, out_step=np.zeros(nstep, dtype=bool))
    mults = np.ones(nstep) if mults is None else np.asarray(mults, dtype=np.float64)[:nstep]
    bc = np.cumsum(np.outer(mults, FP_max), axis=0)           # inlod.f:57-63
    for d in (0, 4, 8):
        if not fft["isnbc"][d]:
            bc[:, d] += 1.0
    fft["bc_all"][...] = bc.T
    fft["fn"][:, [0, 4, 8]] = 1.0; fft["fn1"][...] = fft["fn"]
    H = make_harness(fft)
    it = H.it
    log = dict(cg=[], steps=[])

    def drive_eps_sig(step, iter_):
        P, K = H.sweep(step, iter_, fft["fn"], fft["fn1"])
        fft["pn1"][...] = P; fft["k4"][...] = K
        print(f"  {prefix}sweep step {step} iter {iter_}: {time.time() - t_start:.0f} s", flush=True)

    def update():
        H.update()
        log["steps"].append(dict(Fn1=np.ascontiguousarray(fft["fn1"]).copy(), Pn1=np.ascontiguousarray(fft["pn1"]).copy(),
                                 hist=np.ascontiguousarray(H.hist_n1).copy(), urcs=H.urcs_n1.copy(),
                                 n_sweeps=len(H.sweeps), n_cg=len(log["cg"]), n_tangent_homo=it.calls.get("tangent_homo", 0)))
    st, dcg_init, dcg_check, dcg, dcg_get = mkl_rci_cg()

    def thyme(a, b):
        if int(b) == 1:
            log["cg"].append([it.calls.get("tangent_homo", 0), 0])
        else:
            log["cg"][-1][1] = st.get("it", 0)
    F.BUILTIN_SUBS.update(drive_eps_sig=drive_eps_sig, update=update, thyme=thyme, mkl_free_buffers=lambda *a: None,
                          dcg_init=dcg_init, dcg_check=dcg_check, dcg=dcg, dcg_get=dcg_get, ouresult=lambda *a: None)
    F.BUILTIN_ARRAY_ARGS.update(drive_eps_sig=(), update=(), thyme=(), mkl_free_buffers=(), ouresult=(),
                                dcg_init=(1, 2, 4, 5, 6), dcg_check=(1, 2, 4, 5, 6), dcg=(1, 2, 4, 5, 6), dcg_get=(1, 2, 4, 5, 6))
    F.BUILTIN_INFO_ARG.update(dcg_init=3, dcg_check=3, dcg=3, dcg_get=7)

    it.call("formg"); it.call("formfftshift", fft["coeffs1"], fft["coeffs2"])
    drive_eps_sig(1, 0)                                                  # FFT_finite_3d.f:145
    K4_initial = np.ascontiguousarray(fft["k4"]).copy()
    it.call("fft_nr3")                                                   # FFT_finite_3d.f:146

    res = dict(N=N, nstep=nstep, FP_max=FP_max, isNBC=fft["isnbc"].astype(np.int32), mults=mults, tolNR=fft["tolnr"], tolPCG=fft["tolpcg"],
               maxIter=fft["maxiter"], K4_initial=K4_initial, hist_size=H.hist_sz)
    for k in ("Fn1", "Pn1", "hist", "urcs", "n_sweeps", "n_cg", "n_tangent_homo"):
        res["step_" + k] = np.array([s[k] for s in log["steps"]])
    res["sweeps"] = np.array(H.sweeps)                            # (step, global iteration, predictor Jacobians, update Jacobians) summed over the block
    res["cg"] = np.array(log["cg"])                                    # (tangent_homo calls so far, CG iterations) per fftPcg call, in call order
    for k, v in res.items():
        out[prefix + k] = v
    return H


def main():
    """GLOBAL_ONLY=<comma list of wrap, job1, m01, deck, deck01>: regenerate only those parts and keep the rest of the existing fixture"""
    t_start = time.time()
    N, nstep = 3, int(os.environ.get("GLOBAL_NSTEP", "3"))
    N3 = N ** 3
    path = os.path.join(ROOT, "tests", "golden", "reference_global.npz")
    only = [x for x in os.environ.get("GLOBAL_ONLY", "").split(",") if x]
    want = lambda name: not only or name in only
    out = {}
    if only:
        old = np.load(path)
        out.update({k: old[k] for k in old.files})
    if want("wrap"):
        wrapper_cases(out)

    # ---- job 1: one fcc crystal per voxel, Voce hardening, its own orientation; uniaxial tension along x under mixed boundary
    #      conditions: F_xx prescribed, P_yy = P_zz = 0, no mean shear
    if want("job1"):
        rng = np.random.default_rng(20240609)
        angles = rng.uniform(0.0, 360.0, (N3, 3))
        FP_max = np.zeros(9); FP_max[0] = 0.002
        isNBC = np.zeros(9, dtype=bool); isNBC[[4, 8]] = True
        run_job(out, "", N, nstep, FP_max, isNBC, lambda fft: Harness(N3, angles, extra_module_vars=fft), t_start)
        out.update(angles=angles, params=np.array([PRM[q] for q in ("rate_n", "theta_0", "tau_y", "tau_v", "voche_m", "iD_v", "e", "nu")]))

    # ---- job 2: mm01 (bilinear Mises plasticity), the two materials of the shipped deck examples/test_mm01.in scattered over the
    #      grid with kinematic / mixed / isotropic hardening per voxel, strain-controlled (all nine mean components prescribed):
    #      the deck's 3 % tension with lateral contraction per step plus a shear component
    single = lambda d: {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in d.items()}      # matprp is single precision (mod_fft.f:20)
    if want("m01"):
        rng = np.random.default_rng(20240611)
        incl = rng.random(N3) < 0.4
        m01 = dict(e=np.where(incl, 24000.0, 12000.0), nu=np.full(N3, 0.3), yld=np.where(incl, 200.0, 100.0), tan_e=np.full(N3, 1000.0),
                   beta=rng.choice([0.0, 0.5, 1.0], N3))
        FP_max = np.array([0.03, 0.005, 0.0, 0.0, -0.01, 0.0, 0.0, 0.0, -0.01])
        run_job(out, "m01_", N, nstep, FP_max, np.zeros(9, dtype=bool), lambda fft: Harness(N3, None, extra_module_vars=fft, mm01=single(m01)), t_start)
        for k, v in m01.items():
            out["m01_prop_" + k] = np.asarray(v, dtype=np.float64)

    # ---- jobs 3, 4: the reference's shipped decks examples/test_mm10.in (7^3, bcc48, Voce with alter_mode on, orientations from
    #      angle_bc.in, F_xx 0.03 / F_yy = F_zz -0.01 in ten steps, time step 10) and examples/test_mm01.in (7^3, two bilinear materials,
    #      F_xx 0.3 / F_yy = F_zz -0.1 in ten steps) as they stand, all load steps.  The decks are read by this repository's reader
    #      (cpfft_b200/deck.py); everything from the material properties on is the reference's text.
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import deck
    nd = int(os.environ.get("GLOBAL_DECK_STEPS", "10"))
    if want("deck") and nd > 0:
        pd = deck("test_mm10.in")
        cd = pd.crystals[0]
        prm = dict(rate_n=cd.harden_n, theta_0=cd.theta_0, tau_y=cd.tau_y, tau_v=cd.tau_v, voche_m=cd.voche_m, iD_v=cd.iD_v, e=cd.e, nu=cd.nu,
                   alter_mode=bool(cd.alter_mode), eps_dot_0_y=cd.eps_dot_0_y)
        run_job(out, "deck_", pd.N, nd, np.asarray(pd.FP_max, dtype=np.float64), np.asarray(pd.isNBC, dtype=bool),
                lambda fft: Harness(pd.N3, np.asarray(pd.angles), slip_type=cd.slip_type, extra_module_vars=fft, prm=prm, dt=pd.tstep), t_start,
                mults=pd.mults, maxiter=pd.maxIter)
    if want("deck01") and nd > 0:
        pd1 = deck("test_mm01.in")
        mat = [pd1.materials[m - 1] for m in pd1.matlist]
        d01 = dict(e=[m.e for m in mat], nu=[m.nu for m in mat], yld=[m.yld_pt for m in mat], tan_e=[m.tan_e for m in mat], beta=[m.beta for m in mat])
        run_job(out, "deck01_", pd1.N, nd, np.asarray(pd1.FP_max, dtype=np.float64), np.asarray(pd1.isNBC, dtype=bool),
                lambda fft: Harness(pd1.N3, None, extra_module_vars=fft, mm01=single(d01), dt=pd1.tstep), t_start, mults=pd1.mults, maxiter=pd1.maxIter)

    # ---- jobs 5, 6: the derived mixed decks of SURVEY.md 8d -- the shipped decks with F_xx driven and P_yy = P_zz = 0 (tests/helpers.py
    #      stress_bc_variant): the outer loop on the mean stress with tangent_homo / NBC_update on the shipped materials
    from helpers import stress_bc_variant
    if want("deck01nbc") and nd > 0:
        pn = stress_bc_variant(deck("test_mm01.in"))
        mat = [pn.materials[m - 1] for m in pn.matlist]
        d01 = dict(e=[m.e for m in mat], nu=[m.nu for m in mat], yld=[m.yld_pt for m in mat], tan_e=[m.tan_e for m in mat], beta=[m.beta for m in mat])
        run_job(out, "deck01nbc_", pn.N, min(nd, 3), np.asarray(pn.FP_max, dtype=np.float64), np.asarray(pn.isNBC, dtype=bool),
                lambda fft: Harness(pn.N3, None, extra_module_vars=fft, mm01=single(d01), dt=pn.tstep), t_start, mults=pn.mults, maxiter=pn.maxIter)
    if want("deck10nbc") and nd > 0:
        pn = stress_bc_variant(deck("test_mm10.in"))
        cd = pn.crystals[0]
        prm = dict(rate_n=cd.harden_n, theta_0=cd.theta_0, tau_y=cd.tau_y, tau_v=cd.tau_v, voche_m=cd.voche_m, iD_v=cd.iD_v, e=cd.e, nu=cd.nu,
                   alter_mode=bool(cd.alter_mode), eps_dot_0_y=cd.eps_dot_0_y)
        run_job(out, "deck10nbc_", pn.N, min(nd, 2), np.asarray(pn.FP_max, dtype=np.float64), np.asarray(pn.isNBC, dtype=bool),
                lambda fft: Harness(pn.N3, np.asarray(pn.angles), slip_type=cd.slip_type, extra_module_vars=fft, prm=prm, dt=pn.tstep), t_start,
                mults=pn.mults, maxiter=pn.maxIter)

    # ---- job 7: the derived MTS deck tests/golden/decks/mts_mm10.in (5^3 fcc polycrystal, MTS hardening, 0.1 % strain per step,
    #      four steps): the mechanical-threshold-stress law through the whole loop
    if want("deckmts") and nd > 0:
        pm = deck("mts_mm10.in")
        cm = pm.crystals[0]
        prm = dict(rate_n=cm.harden_n, theta_0=cm.theta_0, tau_y=cm.tau_y, tau_v=cm.tau_v, voche_m=cm.voche_m, iD_v=cm.iD_v, e=cm.e, nu=cm.nu,
                   alter_mode=bool(cm.alter_mode), eps_dot_0_y=cm.eps_dot_0_y)
        mts = {k: getattr(cm, k) for k in MTS}
        run_job(out, "deckmts_", pm.N, min(nd, pm.nstep), np.asarray(pm.FP_max, dtype=np.float64), np.asarray(pm.isNBC, dtype=bool),
                lambda fft: Harness(pm.N3, np.asarray(pm.angles), slip_type=cm.slip_type, mts=mts, extra_module_vars=fft, prm=prm, dt=pm.tstep), t_start,
                mults=pm.mults, maxiter=pm.maxIter)

    # ---- job 8: the derived Taylor deck tests/golden/decks/taylor_mm10.in (5^3, two crystals per material point -- bcc48 and fcc, crystal
    #      numbers and orientations from a flat file --, Taylor average, five steps)
    if want("decktaylor") and nd > 0:
        pt = deck("taylor_mm10.in")
        types = [dict(slip_type=c.slip_type, prm=dict(rate_n=c.harden_n, theta_0=c.theta_0, tau_y=c.tau_y, tau_v=c.tau_v, voche_m=c.voche_m, iD_v=c.iD_v,
                                                      e=c.e, nu=c.nu, alter_mode=bool(c.alter_mode), eps_dot_0_y=c.eps_dot_0_y)) for c in pt.crystals]
        run_job(out, "decktaylor_", pt.N, min(nd, pt.nstep), np.asarray(pt.FP_max, dtype=np.float64), np.asarray(pt.isNBC, dtype=bool),
                lambda fft: Harness(pt.N3, np.asarray(pt.angles), extra_module_vars=fft, dt=pt.tstep, types=types, ids=np.asarray(pt.crystal_ids)), t_start,
                mults=pt.mults, maxiter=pt.maxIter)

    h = hashlib.sha256()
    for f in FILES:
        h.update(open(REF + f, "rb").read())
    out["provenance"] = ("maranGit/CPFFT src/{" + ", ".join(FILES) + "} executed by tools/fortran_subset.py (tools/make_reference_global.py); sha256 of the sources "
                         + h.hexdigest() + "; interpreter sha256 " + hashlib.sha256(open(os.path.join(ROOT, "tools", "fortran_subset.py"), "rb").read()).hexdigest())
    np.savez_compressed(path, **out)
    print("wrote", path, f"{time.time() - t_start:.0f} s", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
