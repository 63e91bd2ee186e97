// cpfft_b200: the per-voxel stress update (drive_eps_sig.f:57-339), one voxel per call.
//
//   upd_mm01_voxel / upd_mm10_voxel : kinematics -> material model -> scatter of the n+1 state
//                                     (rstgp1.f dispatch, rplstr.f:62-85 scatter rules)
//   upd_pk1_voxel                   : P = J sigma F^-T and A = dP/dF (cs2p + gptns1 + cep2A)
//
// The kernels of material.cu call these with one thread per voxel; the host build
// (tests/native/material_host.cpp) calls the same code voxel by voxel.  All state is
// structure-of-arrays field[comp * n3 + voxel].
#pragma once
#include "compat.cuh"
#include "material_types.h"
#include "kin.cuh"
#include "mm01.cuh"
#include "mm10.cuh"

struct UpdArgs {
  const double* Fn; const double* Fn1;
  const double* urcs_n; double* urcs_n1;
  const double* eps_n; double* eps_n1;
  double* rot_n1;
  const double* hist_n; double* hist_n1;
  double* cep;
  const int32_t* matidx; const int32_t* grain;   // grain: (ncmax, n3), crystal ci of voxel e at [ci * n3 + e]
  const int32_t* grain_cry;                      // per grain-table entry: 0-based crystal library index
  const CpfMatDev* mats; const CpfCryDev* crys; const double* grains;
  int32_t* fail; int32_t* liters; int* failcnt;
  int64_t n3; int step, iter; double dt;
  CpfHistLayout L;
  // The model's crystal constants when every grain refers to the same crystal-library entry
  // (uni_cry = 1; the usual polycrystal: one crystal definition, many orientations).  The UNI
  // kernels read them from the kernel-parameter constant bank -- instruction operands, no loads
  // and no registers held across the Newton loop -- and the exponent test of the slip-rate power
  // becomes warp-uniform.
  CpfCryDev cr0; int uni_cry;
};


CPF_DI void upd_mm01_voxel(const UpdArgs& a, const int64_t e) {
  const CpfMatDev mp = a.mats[a.matidx[e]];
  if (mp.type != 1) return;
  const int64_t n3 = a.n3;
  double fn[9], fn1[9], Rh[9], R[9], fhinv[9], detFh, de[6];
#pragma unroll
  for (int k = 0; k < 9; ++k) { fn[k] = a.Fn[k * n3 + e]; fn1[k] = a.Fn1[k * n3 + e]; }
  voxel_kinematics(fn, fn1, Rh, R, fhinv, &detFh, de);
  double hn[11], sn[9], s1[9], h1[11], cep[36];
#pragma unroll
  for (int k = 0; k < 11; ++k) hn[k] = a.hist_n[k * n3 + e];
#pragma unroll
  for (int k = 0; k < 9; ++k) sn[k] = a.urcs_n[k * n3 + e];
  mm01_update(a.step, mp, hn, sn, de, s1, h1, cep);
#pragma unroll
  for (int k = 0; k < 9; ++k) a.urcs_n1[k * n3 + e] = s1[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) a.eps_n1[k * n3 + e] = a.eps_n[k * n3 + e] + de[k];
  if (a.iter > 0) {  // rplstr.f:74-85
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) a.rot_n1[(3 * j + i) * n3 + e] = R[3 * i + j];
#pragma unroll
    for (int k = 0; k < 11; ++k) a.hist_n1[k * n3 + e] = h1[k];
  } else {
    // iter 0 leaves history n+1 untouched in the reference, i.e. equal to history n since the last
    // commit; cpfft_update exchanges the two buffers instead of copying, so carry the n values over
#pragma unroll
    for (int k = 0; k < 11; ++k) a.hist_n1[k * n3 + e] = a.hist_n[k * n3 + e];
  }
#pragma unroll
  for (int k = 0; k < 36; ++k) a.cep[k * n3 + e] = cep[k];
}

// sm: this thread's slice of the kernel's shared memory (element k at sm[k * MM10_THREADS]).
// MULTI = false: one crystal per material point (every shipped deck, the benchmark).
// MULTI = true : polycrystalline material points, n_crystals > 1 (mm10_a.f:112-197): the
//   crystals of the point are integrated one after the other under the same R and D, each
//   with its own history block (common block + ci * one_crystal_hist_size, mm10_a.f:640-641),
//   and stress, tangent, slip and work increments are Taylor-averaged
//   (mm10_a_crystal_avgs, mm10_a.f:139-164).  The sums of the 36 tangent entries and of the
//   slip increments are kept in the voxel's own n+1 slots (a.cep, hist_n1 slip sums).
// HARD: hardening law of the material's crystals, MM10_VOCE or MM10_MTS (`hardening mts`:
//   thresholds tau_y, tau_v of every (sub)step from the strain rate, mm10_setup_mts
//   mm10_a.f:2109-2175; h / estress / ehard mm10_b.f:2080-2186; tangent terms JA, JB from
//   dgamma/dD and ed, mm10_a.f:740-810, mm10_b.f:2189-2345 -- both proportional to the strain
//   increment, so C - JA - JB is a rank-one correction of C).
template <bool MULTI, int HARD, bool LF = false, bool UNI = false>
CPF_DI void upd_mm10_voxel(const UpdArgs& a, const int64_t e, double* sm) {
  const CpfMatDev mp = a.mats[a.matidx[e]];
  if (mp.type != 10) return;
  if ((mp.ncry > 1) != MULTI) return;
  if (mp.hard != HARD) return;
  const int64_t n3 = a.n3;
  const CpfHistLayout& L = a.L;
  double R[9], de[6];
  {
    double fn[9], fn1[9], Rh[9], fhinv[9], detFh;
#pragma unroll
    for (int k = 0; k < 9; ++k) { fn[k] = a.Fn[k * n3 + e]; fn1[k] = a.Fn1[k * n3 + e]; }
    voxel_kinematics(fn, fn1, Rh, R, fhinv, &detFh, de);
  }
  // Taylor sums (MULTI only)
  double sig_sum[6] = {0, 0, 0, 0, 0, 0}, winc_sum[3] = {0, 0, 0};
  int fail_any = 0, itp_sum = 0, itu_sum = 0;
  if (MULTI) {
#pragma unroll 1
    for (int k = 0; k < 36; ++k) a.cep[k * n3 + e] = 0.0;
#pragma unroll 1
    for (int s = 0; s < L.len_slip; ++s) a.hist_n1[(L.slipsum + s) * n3 + e] = 0.0;
  }
  const int ncry = MULTI ? mp.ncry : 1;
#pragma unroll 1
  for (int ci = 0; ci < ncry; ++ci) {
    const int co = MULTI ? ci * (L.total - L.c_stress) : 0;            // offset of this crystal's history block
    const int gi = a.grain[(MULTI ? (int64_t)ci * n3 : (int64_t)0) + e];
    CpfCryDev cr_l;
    if (!UNI) cr_l = a.crys[a.grain_cry[gi]];           // the grain-table entry knows its crystal (crystal_input single or file)
    const CpfCryDev& cr = UNI ? a.cr0 : cr_l;
    const double* gt = a.grains + (int64_t)gi * CPF_GRAIN_STRIDE;
    const int nslip = cr.nslip;
    Mm10Ctx c;
    c.ms0 = gt + CPF_GRAIN_B; c.C = gt + CPF_GRAIN_C;
    c.nslip = nslip; c.rate_int = cr.rate_int; c.miter = cr.miter;
    c.rate_n = cr.rate_n; c.theta_0 = cr.theta_0; c.tau_y = cr.tau_y; c.tau_v = cr.tau_v;
    c.voche_m = cr.voche_m; c.iD_v = cr.iD_v;
    c.atol = cr.atol; c.atol1 = cr.atol1; c.rtol = cr.rtol; c.rtol1 = cr.rtol1;
    // ---- state at n (mm10_copy_cc_hist), or the step-1 initial state (mm10_a.f:92-97,220-227)
    double Rpn[9], Dn[6], ttrate_n, work_n[3];
    const bool first = (a.step == 1);
  #pragma unroll
    for (int k = 0; k < 6; ++k) {
      c.sn[k] = first ? a.urcs_n[k * n3 + e] : a.hist_n[((L.c_stress + co) + k) * n3 + e];
      Dn[k] = first ? 0.0 : a.hist_n[((L.c_D + co) + k) * n3 + e];
    }
  #pragma unroll
    for (int j = 0; j < 3; ++j)
  #pragma unroll
      for (int i = 0; i < 3; ++i)
        Rpn[3 * i + j] = first ? ((i == j) ? 1.0 : 0.0) : a.hist_n[((L.c_Rp + co) + 3 * j + i) * n3 + e];
    c.ttn = first ? ((HARD == MM10_MTS) ? -1.0 : (cr.tau_y + 1.0e-5)) : a.hist_n[(L.c_tt + co) * n3 + e];
    // MTS: tau_y and mu_harden of the n state live in u(1:2); < 0 = not set yet (mm10_init_mts)
    double u1n = -1.0, u2n = -1.0, tau_y_full = 0.0, tau_v_full = 0.0;
    Mm10Mts mts;
    if (HARD == MM10_MTS) {
      if (!first) { u1n = a.hist_n[((L.c_u + co) + 0) * n3 + e]; u2n = a.hist_n[((L.c_u + co) + 1) * n3 + e]; }
      mts.tau_hat_y = cr.tau_hat_y; mts.tau_hat_v = cr.tau_hat_v; mts.ky = 0.0; mts.kv = 0.0;
      mts.iq_y = cr.iq_y; mts.ip_y = cr.ip_y; mts.iq_v = cr.iq_v; mts.ip_v = cr.ip_v;
      mts.eps_dot_0_y = cr.eps_dot_0_y; mts.eps_dot_0_v = cr.eps_dot_0_v;
      c.ur = 1.0; c.tau_a = cr.tau_a; c.iD_v = 0.0; c.h0 = 0.0;
    }
    ttrate_n = first ? 0.0 : a.hist_n[(L.c_ttrate + co) * n3 + e];
  #pragma unroll
    for (int k = 0; k < 3; ++k) work_n[k] = first ? 0.0 : a.hist_n[(L.work + k) * n3 + e];
    // ---- mm10_setup: Q = Rp_n^T, RW(R), dg, tau_l ----
  #pragma unroll
    for (int i = 0; i < 3; ++i)
  #pragma unroll
      for (int j = 0; j < 3; ++j) c.Q[3 * i + j] = Rpn[3 * j + i];
    c.J.p = sm + MM10_SM_J * MM10_THREADS;
    c.RWQ.p = sm + MM10_SM_RWQ * MM10_THREADS;
    c.RWR.p = sm + MM10_SM_RWR * MM10_THREADS;
    c.acc.p = sm + MM10_SM_ACC * MM10_THREADS;
    cpf_rvw(c.Q, c.RWQ);
    cpf_rvw(R, c.RWR);
    const double dt = a.dt;
    {
      const double mu_h = CPF_LDG(c.C + 35);
      const double alpha = 1.0 / 3.0;
      const double cst = cr.k_0 * cr.burgers * alpha * alpha * mu_h * mu_h / 2.0 / cr.theta_0;
      c.taul = cst * 0.0;
    }
    const double t1 = de[0] * de[0] + de[1] * de[1] + de[2] * de[2];
    const double t2 = de[3] * de[3] + de[4] * de[4] + de[5] * de[5];
    const bool alter = (HARD == MM10_VOCE) && cr.alter_mode;       // mm10_setup_voche only (mm10_a.f:2073)
    const double dg_full = alter ? cr.eps_dot_0_y * dt : sqrt((2.0 / 3.0) * (t1 + 0.5 * t2));
    double sn2 = 0.0;
  #pragma unroll
    for (int k = 0; k < 6; ++k) sn2 += c.sn[k] * c.sn[k];
    const bool no_load = (sn2 == 0.0) && ((t1 + t2) == 0.0);
    const bool elastic = (a.iter == 0) || no_load;  // iter_0_extrapolate_off (rstgp1.f:870-877)

    // MTS, full step (np1) at 297 K: mu, thresholds, and the n-state hardening value when the
    // history still holds the flag (mm10_a.f:2170-2173).  The elastic path stores the RAW n value
    // (mm10_solve_strup copies tt before mm10_setup runs, mm10_a.f:2674-2677).
    const double ttn_raw = c.ttn;
    double mu_full = 0.0;
    if (HARD == MM10_MTS) {
      mts_at_temperature(cr, 297.0, &mu_full, &mts);
      mts_thresholds(mts, dg_full / dt, &tau_y_full, &tau_v_full);
      c.ur = mu_full / cr.mu_0; c.tau_y = tau_y_full; c.tau_v = tau_v_full;
      if (c.ttn < 0.0) c.ttn = cr.tau_a + c.ur * tau_y_full + 0.1;
    }
    double x[7];
  #pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = c.sn[k];
    x[6] = (HARD == MM10_MTS && elastic) ? ttn_raw : c.ttn;
    double tang[36];   // row-major
    double tt_rate = 0.0;
    int itp = 0, itu = 0;
    bool fail = false;
  #pragma unroll
    for (int k = 0; k < 6; ++k) c.D[k] = de[k];
    c.dg = dg_full; c.tinc = dt;
    if (elastic) {
  #pragma unroll
      for (int k = 0; k < 36; ++k) tang[k] = CPF_LDG(c.C + k);
      if (!no_load) {
        double R1[7];
        mm10_resid<HARD>(c, x, x[6], R1, false);
  #pragma unroll
        for (int k = 0; k < 6; ++k) x[k] = x[k] - R1[k];
      }
    } else {
      // cosine of the angle between the deviatoric strain increments (mm10_a.f:2993-3021)
      double cos_ang;
      {
        double d1[6], d2[6];
        double tr = (de[0] + de[1] + de[2]) / 3.0;
  #pragma unroll
        for (int k = 0; k < 6; ++k) d1[k] = de[k] - ((k < 3) ? tr : 0.0);
        double a1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], a2 = d1[3] * d1[3] + d1[4] * d1[4] + d1[5] * d1[5];
        double s1 = (a1 + a2 == 0.0) ? 0.0 : 1.0 / sqrt(a1 + a2);
        tr = (Dn[0] + Dn[1] + Dn[2]) / 3.0;
  #pragma unroll
        for (int k = 0; k < 6; ++k) d2[k] = Dn[k] - ((k < 3) ? tr : 0.0);
        a1 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]; a2 = d2[3] * d2[3] + d2[4] * d2[4] + d2[5] * d2[5];
        double s2 = (a1 + a2 == 0.0) ? 0.0 : 1.0 / sqrt(a1 + a2);
        // the reference divides each vector by its norm, then takes the dot product; a
        // uniform scaling of the sub-step strain does not change the direction
        double p1 = 0.0, p2 = 0.0;
  #pragma unroll
        for (int k = 0; k < 3; ++k) p1 += (d1[k] * s1) * (d2[k] * s2);
  #pragma unroll
        for (int k = 3; k < 6; ++k) p2 += (d1[k] * s1) * (d2[k] * s2);
        cos_ang = fmax(p1 + p2, 0.0);
      }
      double frac = 0.0, stp = 1.0, ox[7], h_last = c.ttn;
      int cuts = 0, lu_piv = 0;
  #pragma unroll
      for (int k = 0; k < 7; ++k) ox[k] = x[k];
      while (frac < 1.0) {  // mm10_solve_strup_iterate (mm10_a.f:2759-2843)
        const double sc = stp + frac;
  #pragma unroll
        for (int k = 0; k < 6; ++k) c.D[k] = de[k] * sc;
        c.tinc = dt * sc;
        c.dg = alter ? cr.eps_dot_0_y * c.tinc : sqrt((2.0 / 3.0) * ((t1 * sc * sc) + 0.5 * (t2 * sc * sc)));
        if (!alter && sc == 1.0) c.dg = dg_full;
        if (HARD == MM10_MTS) {
          // mm10_setup_mts for the sub-step state `curr`; its temperature is 297 (step + frac) because
          // n%temp = 0 in this code base (mm10_a.f:2486, 2769)
          double mu_s, ty, tv;
          mts_at_temperature(cr, 297.0 * sc, &mu_s, &mts);
          mts_thresholds(mts, c.dg / c.tinc, &ty, &tv);
          const double ty_n = (u1n < 0.0) ? ty : u1n, mu_n = (u2n < 0.0) ? mu_s : u2n;
          c.ur = mu_s / cr.mu_0; c.tau_y = ty; c.tau_v = tv;
          c.h0 = cr.tau_a * (1.0 - mu_s / mu_n) + c.ur * (ty - ty_n) + (mu_s / mu_n) * c.ttn;
        }
        x[6] = c.ttn;
        fail = mm10_solve<HARD, LF>(c, x, cos_ang * ttrate_n * (dt * stp), &itp, &itu, &h_last, &lu_piv);
        if (fail) {
  #pragma unroll
          for (int k = 0; k < 7; ++k) x[k] = ox[k];
          stp = stp * 0.5; cuts = cuts + 1;
          if (cuts > 4) break;
          fail = false;
        } else {
  #pragma unroll
          for (int k = 0; k < 7; ++k) ox[k] = x[k];
          frac = frac + stp;
        }
      }
      bool nan = false;
  #pragma unroll
      for (int k = 0; k < 7; ++k) nan = nan || isnan(x[k]);
      fail = fail || nan;
      tt_rate = (h_last - c.ttn) / c.tinc;
      if (!fail) {
        // restore the full-step context for tangent / rotation / output (np1, not curr)
  #pragma unroll
        for (int k = 0; k < 6; ++k) c.D[k] = de[k];
        c.dg = dg_full; c.tinc = dt;
        // MTS: C - JA - JB = C - w (x) d_mod, w = alpha va + (ce / J22) J12 with the lagged J12, J22,
        //   va = sum_s slip_s (C ms_s + 2 symSW(sigma, qc_s)) and ed = ce d_mod at the converged state
        //   (mm10_dgdd_mts, mm10_ed_mts with tau_l = 0; mm10_a.f:760-805)
        double wv[7] = {0, 0, 0, 0, 0, 0, 0}, dmod[6] = {0, 0, 0, 0, 0, 0};
        if (HARD == MM10_MTS) {
          mts_at_temperature(cr, 297.0, &mu_full, &mts);
          c.ur = mu_full / cr.mu_0; c.tau_y = tau_y_full; c.tau_v = tau_v_full;
          // va: the reference hands `symtqmat` (its column 1), not `symtqmat(1,i)`, to mm10_a_mult_type_4 (mm10_a.f:771-772),
          // so sym(sigma W) of the FIRST slip system enters every term: 2 symSW(sigma, qc_1) sum_s slip_s.  Reproduced
          // (established by executing the reference's mm10_tangent, tests/test_reference_vectors.py).
          double dps[6] = {0, 0, 0, 0, 0, 0}, wqs[3] = {0, 0, 0}, sabs = 0.0, ssum = 0.0;
          const double tt = x[6], itt = 1.0 / tt, dgtt = c.dg / tt;
  #pragma unroll 1
          for (int s = 0; s < nslip; ++s) {
            double ms[6], qs[3];
            mm10_slip_geom(c, s, ms, qs);
            const double rs = x[0] * ms[0] + x[1] * ms[1] + x[2] * ms[2] + x[3] * ms[3] + x[4] * ms[4] + x[5] * ms[5];
            const double slip = dgtt * cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0) * rs;
  #pragma unroll
            for (int k = 0; k < 6; ++k) dps[k] += slip * ms[k];
            if (s == 0) {
  #pragma unroll
              for (int k = 0; k < 3; ++k) wqs[k] = qs[k];
            }
            sabs += fabs(slip); ssum += slip;
          }
  #pragma unroll
          for (int k = 0; k < 3; ++k) wqs[k] *= ssum;
          double wc[3], sw[6];
          cpf_mv3(c.RWR, wqs, wc);
          cpf_symsw(x, wc, sw);
          const double alpha = 2.0 / (3.0 * c.dg * c.dg);
          const double dgc = c.dg / c.tinc;
          const double lny = log(mts.eps_dot_0_y / dgc), lnv = log(mts.eps_dot_0_v / dgc);
          const double ty = mts.ky * lny, tv = mts.kv * lnv;
          const double cy = 2.0 * cr.tau_hat_y / (3.0 * c.dg * c.dg * cr.q_y * cr.p_y * lny) *
                            cpf_pow(1.0 - cpf_pow(ty, mts.iq_y), mts.ip_y - 1.0) * cpf_pow(ty, mts.iq_y);
          const double cv = 2.0 * cr.tau_hat_v / (3.0 * c.dg * c.dg * cr.q_v * cr.p_v * lnv) *
                            cpf_pow(1.0 - cpf_pow(tv, mts.iq_v), mts.ip_v - 1.0) * cpf_pow(tv, mts.iq_v);
          const double scc = tt / c.ur - c.tau_a / c.ur - c.tau_y;
          const double base = 1.0 - scc / c.tau_v;
          const double bm1 = cpf_pow(base, c.voche_m - 1.0);
          const double ce = c.theta_0 * c.ur * ((c.voche_m / c.tau_v * bm1) * cy + (c.voche_m / (c.tau_v * c.tau_v) * scc * bm1) * cv +
                                               alpha * (bm1 * base)) * sabs + c.ur * cy;
          const double j22 = c.acc[MM10_SM_STASH + 6];     // raw J12, J22 of the lagged Jacobian (c.J holds its LU factors)
  #pragma unroll
          for (int i = 0; i < 6; ++i) {
            double va = 2.0 * sw[i];
  #pragma unroll
            for (int k = 0; k < 6; ++k) va += CPF_LDG(c.C + 6 * i + k) * dps[k];
            wv[i] = alpha * va + (ce / j22) * c.acc[MM10_SM_STASH + i];
            dmod[i] = (i < 3) ? de[i] : 0.5 * de[i];
          }
        }
        // mm10_tangent (Voce: ed = 0, dgammadd = 0): T = (J11 - J12 J21 / J22)^-1 C.  The Schur
        // complement is never formed: (J11 - J12 J21 / J22)^-1 B is the leading 6 rows of J^-1 [B; 0], and
        // c.J already holds the LU factors of the lagged Jacobian (the last Newton step's), so a column
        // of the tangent is one pair of triangular solves.
        if (HARD == MM10_MTS) mm10_lu7_solve(c.J, lu_piv, wv);      // JJ^-1 w (w was formed from the stashed J12, J22 above)
  #pragma unroll 1
        for (int col = 0; col < 6; ++col) {
          double b7[7];
  #pragma unroll
          for (int k = 0; k < 6; ++k) b7[k] = CPF_LDG(c.C + 6 * k + col);
          b7[6] = 0.0;
          mm10_lu7_solve(c.J, lu_piv, b7);
  #pragma unroll
          for (int k = 0; k < 6; ++k) c.acc[6 * k + col] = b7[k];
        }
  #pragma unroll
        for (int k = 0; k < 36; ++k) tang[k] = c.acc[k];
        if (HARD == MM10_MTS) {       // T = JJ^-1 C - (JJ^-1 w) (x) d_mod
  #pragma unroll
          for (int i = 0; i < 6; ++i)
  #pragma unroll
            for (int j = 0; j < 6; ++j) tang[6 * i + j] -= wv[i] * dmod[j];
        }
  #pragma unroll
        for (int i = 0; i < 6; ++i)   // mm10_a_make_symm_1
  #pragma unroll
          for (int j = i + 1; j < 6; ++j) {
            const double v = (tang[6 * i + j] + tang[6 * j + i]) * 0.5;
            tang[6 * i + j] = v; tang[6 * j + i] = v;
          }
      }
    }
    // ---- outputs: update_rotation + mm10_output (skipped on the elastic path, where the
    //      reference stores the zero-initialised np1 fields) ----
    double Rp1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, euler[3] = {0, 0, 0}, eps6[6] = {0, 0, 0, 0, 0, 0};
    double ep6[6] = {0, 0, 0, 0, 0, 0}, ed6[6] = {0, 0, 0, 0, 0, 0};
    if (fail) {
      // material_cut_step.  The reference prints a warning, resets stress / tau_tilde to the n
      // state (mm10_a.f:2838-2841) and leaves the rest of the block un-updated (:125-127), i.e.
      // undefined data.  Defined behaviour here (identical in the oracle): the point keeps its n
      // state (stress, tau_tilde, Rp, Euler angles, lattice strain), no slip, elastic tangent;
      // the sweep goes on and the failure is counted (cpfft_material_failures).
      if (MULTI) fail_any = 1;
      else { a.fail[e] = 1; CPF_ATOMIC_INC(a.failcnt); CPF_ATOMIC_INC(a.failcnt + 1); }
  #pragma unroll
      for (int k = 0; k < 6; ++k) x[k] = c.sn[k];
      x[6] = c.ttn;
      tt_rate = 0.0;
  #pragma unroll
      for (int k = 0; k < 36; ++k) tang[k] = CPF_LDG(c.C + k);
  #pragma unroll
      for (int k = 0; k < 9; ++k) Rp1[k] = Rpn[k];
  #pragma unroll
      for (int k = 0; k < 3; ++k)
        euler[k] = first ? CPF_LDG(gt + CPF_GRAIN_ANG + k) : a.hist_n[((L.c_euler + co) + k) * n3 + e];
  #pragma unroll
      for (int k = 0; k < 6; ++k) eps6[k] = first ? 0.0 : a.hist_n[((L.c_eps + co) + k) * n3 + e];
    } else if (!MULTI) a.fail[e] = 0;
    if (MULTI) { itp_sum += itp; itu_sum += itu; }
    else { a.liters[2 * e] = itp; a.liters[2 * e + 1] = itu; }
    double u6 = 0, u7 = 0, u8 = 0, u11 = 0, u12 = 0, u13 = 0, u14 = 0, u15 = 0;
    double work_inc = 0, p_work_inc = 0, p_strain_inc = 0;
    const bool full = !elastic && !fail;
    if (full) {
      double dbarp[6] = {0, 0, 0, 0, 0, 0}, wq[3] = {0, 0, 0}, edv[6] = {0, 0, 0, 0, 0, 0}, Nv[6] = {0, 0, 0, 0, 0, 0};
      const double tt = x[6], itt = 1.0 / tt, dgtt = c.dg / tt, dif = dt * c.iD_v, dgn = c.dg * c.rate_n / tt;
      double maxslip = 0.0; int sysID = 0;
      for (int s = 0; s < nslip; ++s) {
        double ms[6], qs[3];
        mm10_slip_geom(c, s, ms, qs);
        const double rs = x[0] * ms[0] + x[1] * ms[1] + x[2] * ms[2] + x[3] * ms[3] + x[4] * ms[4] + x[5] * ms[5];
        const double p = cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0);
        const double slip = dgtt * p * rs, dslp = rs * dif, dgdt = dgn * p + dif;
  #pragma unroll
        for (int k = 0; k < 6; ++k) { dbarp[k] += slip * ms[k]; edv[k] += dslp * ms[k]; Nv[k] += (rs * dgdt) * ms[k]; }
  #pragma unroll
        for (int k = 0; k < 3; ++k) wq[k] += (slip + dslp) * qs[k];
        const double tot = slip + dslp;
        a.hist_n1[((L.c_slipinc + co) + s) * n3 + e] = tot;
        if (MULTI) a.hist_n1[(L.slipsum + s) * n3 + e] += tot;
        else a.hist_n1[(L.slipsum + s) * n3 + e] = (first ? 0.0 : a.hist_n[(L.slipsum + s) * n3 + e]) + tot;
        if (fabs(tot) > maxslip) { maxslip = fabs(tot); sysID = s + 1; }
      }
      int numAct = 0;
      for (int s = 0; s < nslip; ++s)
        if (fabs(a.hist_n1[((L.c_slipinc + co) + s) * n3 + e]) >= 0.1 * maxslip) numAct++;
      u6 = maxslip / dt; u7 = (double)sysID; u8 = (double)numAct;
      // plastic rotation update: Rp = exp(Wbar_p) Rp_n (mm10_a.f:3310-3414)
      {
        double W[9] = {0, wq[2], wq[1], -wq[2], 0, wq[0], -wq[1], -wq[0], 0}, ex[9], W2[9];
        const double al = sqrt(W[5] * W[5] + W[2] * W[2] + W[1] * W[1]);
        if (al < 1.0e-16) {
  #pragma unroll
          for (int k = 0; k < 9; ++k) ex[k] = 0.0;
        } else {
          m3_mul(W, W, W2);
          const double ca = (1.0 - cos(al)) / (al * al), cb = sin(al) / al;
  #pragma unroll
          for (int k = 0; k < 9; ++k) ex[k] = ca * W2[k] + cb * W[k];
        }
        ex[0] += 1.0; ex[4] += 1.0; ex[8] += 1.0;
        m3_mul(ex, Rpn, Rp1);
      }
      // Euler angles (mm10_a.f:1171-1233)
      {
        double w1[9], fr[9], g[9];
  #pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = CPF_LDG(gt + CPF_GRAIN_G + k);
        m3_mul_nt(Rp1, R, w1);
        m3_mul(g, w1, fr);
        const double PI = 3.141592653589793;
        double psi = cpf_atan2(fr[7], fr[6]); if (psi < 0.0) psi += 2.0 * PI;
        double phi = cpf_atan2(fr[5], fr[2]); if (phi < 0.0) phi += 2.0 * PI;
        double f33 = fr[8]; if (f33 > 1.0) f33 = 1.0;
        const double th = acos(f33);
        euler[0] = 180.0 / PI * psi; euler[1] = 180.0 / PI * th; euler[2] = 180.0 / PI * phi;
      }
      // diffusion strain
  #pragma unroll
      for (int k = 0; k < 6; ++k) ed6[k] = edv[k] / dt;
      u15 = sqrt(2.0 / 3.0 * ((edv[0] * edv[0] + edv[1] * edv[1] + edv[2] * edv[2]) +
                             0.5 * (edv[3] * edv[3] + edv[4] * edv[4] + edv[5] * edv[5]))) / dt;
      work_inc = x[0] * de[0] + x[1] * de[1] + x[2] * de[2] + x[3] * de[3] + x[4] * de[4] + x[5] * de[5];
      // lattice strain: ee = RE(R) (C^-1 sigma)
      {
        double eu[7];   // C^-1 sigma through the same LU site (C padded to 7x7 in shared memory)
  #pragma unroll
        for (int i = 0; i < 6; ++i) {
  #pragma unroll
          for (int j = 0; j < 6; ++j) c.J[7 * i + j] = CPF_LDG(c.C + 6 * i + j);
          c.J[7 * i + 6] = 0.0; c.J[42 + i] = 0.0;
        }
        c.J[48] = 1.0;
  #pragma unroll
        for (int k = 0; k < 6; ++k) eu[k] = x[k];
        eu[6] = 0.0;
        mm10_lu7_solve(c.J, mm10_lu7_factor(c.J), eu);
        // ee = RT2RVE(R) eeunrot: the stress-type operator (mm10_a.f:3538-3539), i.e. R E~ R^T
        double E[9], T[9], S2[9];
        v6_to_m3(eu, E);
        m3_mul(R, E, T);
        m3_mul_nt(T, R, S2);
        eps6[0] = S2[0]; eps6[1] = S2[4]; eps6[2] = S2[8];
        eps6[3] = S2[1]; eps6[4] = S2[5]; eps6[5] = S2[2];
      }
      double wp[3], ew[6], ep[6];
      cpf_mv3(c.RWR, wq, wp);
      cpf_symsw(eps6, wp, ew);
  #pragma unroll
      for (int k = 0; k < 6; ++k) { ep[k] = dbarp[k] + ew[k]; ep6[k] = ep[k] / dt; }
      u11 = sqrt(2.0 / 3.0 * ((ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2]) +
                             0.5 * (ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5]))) / dt;
  #pragma unroll
      for (int k = 0; k < 6; ++k) ep[k] = ep[k] + edv[k];
      p_strain_inc = sqrt(2.0 / 3.0 * ((ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2]) +
                                      0.5 * (ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5])));
      p_work_inc = x[0] * ep[0] + x[1] * ep[1] + x[2] * ep[2] + x[3] * ep[3] + x[4] * ep[4] + x[5] * ep[5];
      const double ec_dot = p_strain_inc / dt;
      double n_eff;
      if (ec_dot > 0.0) {
        const double a1 = Nv[0] * ep[0] + Nv[1] * ep[1] + Nv[2] * ep[2];
        const double a2 = Nv[3] * ep[3] + Nv[4] * ep[4] + Nv[5] * ep[5];
        n_eff = (2.0 / 3.0) * ((a1 + 0.5 * a2) / dt) / ec_dot / ec_dot / dt;
      } else n_eff = 1.0e10;
      u12 = n_eff;
      {
        const double st = (x[0] + x[1] + x[2]) / 3.0;
        const double s0 = x[0] - st, s1 = x[1] - st, s2 = x[2] - st;
        u13 = sqrt(1.5 * ((s0 * s0 + s1 * s1 + s2 * s2) + 2.0 * (x[3] * x[3] + x[4] * x[4] + x[5] * x[5])));
      }
      if (ec_dot < 1.e-100) u14 = 0.0;
      else if (n_eff > 100.0) u14 = -1.0;
      else u14 = ec_dot / cpf_pow(u13, n_eff);
    } else {
      for (int s = 0; s < nslip; ++s) {
        a.hist_n1[((L.c_slipinc + co) + s) * n3 + e] = 0.0;
        if (!MULTI) a.hist_n1[(L.slipsum + s) * n3 + e] = first ? 0.0 : a.hist_n[(L.slipsum + s) * n3 + e];
      }
    }
    // ---- scatter of the crystal's history block (mm10_store_cryhist; rplstr: mat 10 always saves hist1)
  #pragma unroll
    for (int k = 0; k < 6; ++k) {
      a.hist_n1[((L.c_stress + co) + k) * n3 + e] = x[k];
      a.hist_n1[((L.c_D + co) + k) * n3 + e] = de[k];
      a.hist_n1[((L.c_eps + co) + k) * n3 + e] = eps6[k];
      a.hist_n1[((L.c_ep + co) + k) * n3 + e] = ep6[k];
      a.hist_n1[((L.c_ed + co) + k) * n3 + e] = ed6[k];
    }
  #pragma unroll
    for (int k = 0; k < 3; ++k) a.hist_n1[((L.c_euler + co) + k) * n3 + e] = euler[k];
  #pragma unroll
    for (int j = 0; j < 3; ++j)
  #pragma unroll
      for (int i = 0; i < 3; ++i) a.hist_n1[((L.c_Rp + co) + 3 * j + i) * n3 + e] = Rp1[3 * i + j];
    a.hist_n1[(L.c_tt + co) * n3 + e] = x[6];
    a.hist_n1[(L.c_ttrate + co) * n3 + e] = tt_rate;
    if (HARD == MM10_MTS) {   // np1%u(1:2) = tau_y, mu_harden of the full step; a failed point keeps the n values
      a.hist_n1[((L.c_u + co) + 0) * n3 + e] = fail ? u1n : tau_y_full;
      a.hist_n1[((L.c_u + co) + 1) * n3 + e] = fail ? u2n : mu_full;
    }
    a.hist_n1[((L.c_u + co) + 5) * n3 + e] = u6;
    a.hist_n1[((L.c_u + co) + 6) * n3 + e] = u7;
    a.hist_n1[((L.c_u + co) + 7) * n3 + e] = u8;
    a.hist_n1[((L.c_u + co) + 10) * n3 + e] = u11;
    a.hist_n1[((L.c_u + co) + 11) * n3 + e] = u12;
    a.hist_n1[((L.c_u + co) + 12) * n3 + e] = u13;
    a.hist_n1[((L.c_u + co) + 13) * n3 + e] = u14;
    a.hist_n1[((L.c_u + co) + 14) * n3 + e] = u15;
    if (MULTI) {
      // sums for the Taylor average (mm10_a.f:228-238)
  #pragma unroll
      for (int k = 0; k < 6; ++k) sig_sum[k] = sig_sum[k] + x[k];
      winc_sum[0] = winc_sum[0] + work_inc; winc_sum[1] = winc_sum[1] + p_work_inc; winc_sum[2] = winc_sum[2] + p_strain_inc;
  #pragma unroll
      for (int k = 0; k < 36; ++k) a.cep[k * n3 + e] += tang[k];
    } else {
      // ---- point-level store for the single crystal (mm10_a_store_crystal) ----
  #pragma unroll
      for (int k = 0; k < 6; ++k) {
        a.urcs_n1[k * n3 + e] = x[k];
        a.eps_n1[k * n3 + e] = a.eps_n[k * n3 + e] + de[k];
      }
      a.urcs_n1[6 * n3 + e] = a.urcs_n[6 * n3 + e] + work_inc;
      a.urcs_n1[7 * n3 + e] = a.urcs_n[7 * n3 + e] + p_work_inc;
      a.urcs_n1[8 * n3 + e] = a.urcs_n[8 * n3 + e] + p_strain_inc;
      a.hist_n1[(L.work + 0) * n3 + e] = work_n[0] + work_inc;
      a.hist_n1[(L.work + 1) * n3 + e] = work_n[1] + p_work_inc;
      a.hist_n1[(L.work + 2) * n3 + e] = work_n[2] + p_strain_inc;
  #pragma unroll
      for (int j = 0; j < 3; ++j)
  #pragma unroll
        for (int i = 0; i < 3; ++i) {
          a.hist_n1[(L.R + 3 * j + i) * n3 + e] = R[3 * i + j];
          if (a.iter > 0) a.rot_n1[(3 * j + i) * n3 + e] = R[3 * i + j];
        }
  #pragma unroll
      for (int i = 0; i < 6; ++i)
  #pragma unroll
        for (int j = 0; j < 6; ++j) {
          a.hist_n1[(L.cep + 6 * j + i) * n3 + e] = tang[6 * i + j];  // column-major in history
          a.cep[(6 * i + j) * n3 + e] = tang[6 * i + j];
        }
    }
  }   // crystals of the point
  if (MULTI) {
    // ---- mm10_a_crystal_avgs + mm10_a_store_crystal (mm10_a.f:139-164, 285-318) ----
    const bool first = (a.step == 1);
    const double rncry = (double)ncry;
    a.fail[e] = fail_any;
    if (fail_any) { CPF_ATOMIC_INC(a.failcnt); CPF_ATOMIC_INC(a.failcnt + 1); }
    a.liters[2 * e] = itp_sum; a.liters[2 * e + 1] = itu_sum;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      a.urcs_n1[k * n3 + e] = sig_sum[k] / rncry;
      a.eps_n1[k * n3 + e] = a.eps_n[k * n3 + e] + de[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double inc = winc_sum[k] / rncry;
      a.urcs_n1[(6 + k) * n3 + e] = a.urcs_n[(6 + k) * n3 + e] + inc;
      a.hist_n1[(L.work + k) * n3 + e] = (first ? 0.0 : a.hist_n[(L.work + k) * n3 + e]) + inc;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        a.hist_n1[(L.R + 3 * j + i) * n3 + e] = R[3 * i + j];
        if (a.iter > 0) a.rot_n1[(3 * j + i) * n3 + e] = R[3 * i + j];
      }
#pragma unroll 1
    for (int i = 0; i < 6; ++i)
#pragma unroll 1
      for (int j = 0; j < 6; ++j) {
        const double t = a.cep[(6 * i + j) * n3 + e] / rncry;
        a.cep[(6 * i + j) * n3 + e] = t;
        a.hist_n1[(L.cep + 6 * j + i) * n3 + e] = t;               // column-major in history
      }
#pragma unroll 1
    for (int s = 0; s < L.len_slip; ++s)
      a.hist_n1[(L.slipsum + s) * n3 + e] =
          (first ? 0.0 : a.hist_n[(L.slipsum + s) * n3 + e]) + a.hist_n1[(L.slipsum + s) * n3 + e] / rncry;
  }
}

// P and K4 of one voxel from (Fn, Fn1, unrotated stress, [D]).
CPF_DI void upd_pk1_voxel(const double* Fn, const double* Fn1, const double* urcs_n1, const double* cep,
                          double* Pn1, double* K4, const int64_t n3, const int64_t e, const Pk1Scratch S) {
  double fn[9], fn1[9], t6[6], P[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { fn[k] = Fn[k * n3 + e]; fn1[k] = Fn1[k * n3 + e]; }
#pragma unroll
  for (int k = 0; k < 6; ++k) t6[k] = urcs_n1[k * n3 + e];
  pk1_and_tangent(fn, fn1, t6, cep + e, n3, P, nullptr, K4 + e, n3, S);
#pragma unroll
  for (int k = 0; k < 9; ++k) Pn1[k * n3 + e] = P[k];
}
