// cpfft_b200: crystal plasticity (mm10) with Voce hardening, one thread per voxel.
//
// Device-side replacement of the per-point work of
//   setup_mm10_rknstr   (drive_eps_sig.f:537-1002)   -> per-grain table, built once
//   mm10 / mm10_solve_crystal / mm10_solve_strup / mm10_solve (mm10_a.f:29-355,1080-1157,
//                                                       2628-3295)
//   mm10_setup(+_voche), mm10_tangent, mm10_update_rotation, mm10_output
//                                                      (mm10_a.f:830-962,658-815,3310-3682)
//   mm10_formR1/R2/J11/J12/J21/J22 and the Voce law   (mm10_b.f:63-481,1065-1381,1805-2071)
//
// Design differences (same mathematics, same iteration logic, different arithmetic layout):
//  * the current Schmid vectors of a slip system are regenerated from the per-grain
//    reference vectors (ms0, qs0) and Rp_n^T on every use -- 54 FMA -- instead of being
//    stored per point (12 or 48 systems x 12 doubles would not fit in registers);
//  * skew parts are accumulated in the lattice frame and rotated once by RW(R)
//    (RW(R Rp^T) = RW(R) RW(Rp^T)), so qc is never formed per system;
//  * the Jacobian blocks are assembled from slip-system sums (S = sum dgdt m (x) m etc.) and
//    multiplied by the stiffness once, instead of a DGER per system;
//  * |rs/tt|^(n-1) is evaluated ONCE per system per evaluation (the reference calls
//    mm10_slipinc up to three times) and by repeated squaring when n-1 is a small integer.
// Control flow that decides iteration counts (tolerances, Armijo test, at least one update
// iteration, lagged Jacobian for the tangent, sub-stepping) follows the reference exactly.
#pragma once
#include "kin.cuh"
#include "common.cuh"

CPF_DI double cpf_sgn(double x) { return x >= 0.0 ? 1.0 : -1.0; }  // Fortran sign(one,x)

CPF_DI double cpf_pow_abs(double x, int ie, double fe) {  // x >= 0
  if (ie < 0) return pow(x, fe);
  double r = 1.0, b = x;
  int e = ie;
  while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
  return r;
}

// symmetric part of S*W in Voigt form (mm10_b.f:1505-1524)
CPF_DI void cpf_symsw(const double* s, const double* w, double* sw) {
  sw[0] = s[3] * w[2] - s[5] * w[1];
  sw[1] = s[3] * w[2] - s[4] * w[0];
  sw[2] = s[5] * w[1] + s[4] * w[0];
  sw[3] = 0.5 * (w[2] * (s[0] - s[1]) + w[0] * s[5] - w[1] * s[4]);
  sw[4] = 0.5 * (w[0] * (s[1] - s[2]) + w[1] * s[3] + w[2] * s[5]);
  sw[5] = 0.5 * (w[1] * (s[0] - s[2]) + w[0] * s[3] - w[2] * s[4]);
}
// Per-thread array kept in shared memory, element k of thread t at p[k * blockDim + t]:
// conflict-free, and it takes the lagged Jacobian and the two skew-rotation operators out of
// the register budget of the Newton loops.
#ifndef MM10_UNROLL
#define MM10_UNROLL 1
#endif
#define MM10_PRAGMA_(x) _Pragma(#x)
#define MM10_PRAGMA(x) MM10_PRAGMA_(x)
#ifndef MM10_THREADS
#define MM10_THREADS 128
#endif
struct SArr {
  double* p;
  CPF_DI double& operator[](int k) const { return p[k * MM10_THREADS]; }
};

// mm10_rt2rvw (mm10_a.f:1461-1479)
template <class Out>
CPF_DI void cpf_rvw(const double* rt, Out rv) {
  rv[0] = rt[4] * rt[8] - rt[5] * rt[7]; rv[1] = rt[3] * rt[8] - rt[5] * rt[6]; rv[2] = rt[3] * rt[7] - rt[4] * rt[6];
  rv[3] = rt[1] * rt[8] - rt[2] * rt[7]; rv[4] = rt[0] * rt[8] - rt[2] * rt[6]; rv[5] = rt[0] * rt[7] - rt[1] * rt[6];
  rv[6] = rt[1] * rt[5] - rt[2] * rt[4]; rv[7] = rt[0] * rt[5] - rt[2] * rt[3]; rv[8] = rt[0] * rt[4] - rt[1] * rt[3];
}
template <class Mat>
CPF_DI void cpf_mv3(const Mat& M, const double* v, double* o) {
  o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}

// Dense solve with partial pivoting (stand-in for DGESV), A is N x N row-major, B is N x NR.
template <int N, int NR>
CPF_DI void cpf_lu_solve(double* A, double* B) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(A[k * N + k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      double v = fabs(A[i * N + k]);
      if (v > best) { best = v; piv = i; }
    }
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      if (i == piv) {
#pragma unroll
        for (int j = 0; j < N; ++j) { double t = A[k * N + j]; A[k * N + j] = A[i * N + j]; A[i * N + j] = t; }
#pragma unroll
        for (int j = 0; j < NR; ++j) { double t = B[k * NR + j]; B[k * NR + j] = B[i * NR + j]; B[i * NR + j] = t; }
      }
    }
    const double inv = 1.0 / A[k * N + k];
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double l = A[i * N + k] * inv;
#pragma unroll
      for (int j = k + 1; j < N; ++j) A[i * N + j] -= l * A[k * N + j];
#pragma unroll
      for (int j = 0; j < NR; ++j) B[i * NR + j] -= l * B[k * NR + j];
    }
  }
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    const double inv = 1.0 / A[k * N + k];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      double s = B[k * NR + j];
#pragma unroll
      for (int c = k + 1; c < N; ++c) s -= A[k * N + c] * B[c * NR + j];
      B[k * NR + j] = s * inv;
    }
  }
}

struct Mm10Ctx {
  const double* __restrict__ ms0;    // grain table: per system ms0[6], qs0[3] (drive_eps_sig.f:975-986)
  const double* __restrict__ C;      // rotated stiffness, 36 row-major
  int nslip, rate_int, miter;
  double rate_n, theta_0, tau_y, tau_v, voche_m, iD_v;
  double atol, atol1, rtol, rtol1;
  double Q[9];     // Rp_n^T
  SArr RWQ;        // RW(Rp_n^T), 9 entries (shared memory)
  SArr RWR;        // RW(R), 9 entries (shared memory)
  double sn[6];    // stress at n
  double ttn;      // tau_tilde at n
  double D[6];     // strain increment of the (sub)step
  double dg, tinc, taul;
};

// Current Schmid vectors of system s.  The reference forms ms = RT2RVE(Rp_n^T) ms0 and
// qs = RT2RVW(Rp_n^T) qs0 (mm10_a.f:867-876).  RT2RVE is the stress-type 6x6 operator, so in
// tensor form ms = V6( Q M~ Q^T ) with M~ the symmetric tensor whose Voigt vector (no shear
// doubling) is ms0, Q = Rp_n^T; that is what is evaluated here (45 FMA, no 6x6 operator).
CPF_DI void mm10_slip_geom(const Mm10Ctx& c, int s, double* ms, double* qs) {
  const double* t = c.ms0 + 9 * s;
  const double m0 = __ldg(t), m1 = __ldg(t + 1), m2 = __ldg(t + 2), m3 = __ldg(t + 3), m4 = __ldg(t + 4), m5 = __ldg(t + 5);
  const double w0 = __ldg(t + 6), w1 = __ldg(t + 7), w2 = __ldg(t + 8);
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a = c.Q[3 * i], b = c.Q[3 * i + 1], d = c.Q[3 * i + 2];
    T[3 * i + 0] = a * m0 + b * m3 + d * m5;
    T[3 * i + 1] = a * m3 + b * m1 + d * m4;
    T[3 * i + 2] = a * m5 + b * m4 + d * m2;
  }
  ms[0] = T[0] * c.Q[0] + T[1] * c.Q[1] + T[2] * c.Q[2];
  ms[1] = T[3] * c.Q[3] + T[4] * c.Q[4] + T[5] * c.Q[5];
  ms[2] = T[6] * c.Q[6] + T[7] * c.Q[7] + T[8] * c.Q[8];
  ms[3] = T[0] * c.Q[3] + T[1] * c.Q[4] + T[2] * c.Q[5];
  ms[4] = T[3] * c.Q[6] + T[4] * c.Q[7] + T[5] * c.Q[8];
  ms[5] = T[0] * c.Q[6] + T[1] * c.Q[7] + T[2] * c.Q[8];
  qs[0] = c.RWQ[0] * w0 + c.RWQ[1] * w1 + c.RWQ[2] * w2;
  qs[1] = c.RWQ[3] * w0 + c.RWQ[4] * w1 + c.RWQ[5] * w2;
  qs[2] = c.RWQ[6] * w0 + c.RWQ[7] * w1 + c.RWQ[8] * w2;
}

CPF_DI double mm10_hfac(const Mm10Ctx& c, double tt, double* hterm_out) {
  const double hterm = 1.0 - (tt - c.tau_y) / c.tau_v + c.taul / (tt - c.tau_y);
  *hterm_out = hterm;
  const double ah = fabs(hterm);
  const double pw = (c.voche_m == 1.0) ? ah : pow(ah, c.voche_m);
  return pw * cpf_sgn(hterm);
}

// Residual (mm10_formR / formR1 / formR2).  R[0..5] = R1, R[6] = R2 (if want2); returns the
// hardening target h (np1%tt_rate = (h - tt_n)/tinc).  wq_out: lattice-frame sum of
// (slip + diffusion) * qs, i.e. wbarp.
CPF_DI double mm10_resid(const Mm10Ctx& c, const double* sig, double tt, double* R, bool want2, double* wq_out) {
  double dbarp[6] = {0, 0, 0, 0, 0, 0}, wq[3] = {0, 0, 0}, sabs = 0.0;
  const double itt = 1.0 / tt, dgtt = c.dg / tt, dif = c.tinc * c.iD_v;
  MM10_PRAGMA(unroll MM10_UNROLL)
  for (int s = 0; s < c.nslip; ++s) {
    double ms[6], qs[3];
    mm10_slip_geom(c, s, ms, qs);
    const double rs = sig[0] * ms[0] + sig[1] * ms[1] + sig[2] * ms[2] + sig[3] * ms[3] + sig[4] * ms[4] + sig[5] * ms[5];
    const double p = cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0);
    const double slip = dgtt * p * rs;
    const double f = rs * dif + slip;
#pragma unroll
    for (int k = 0; k < 6; ++k) dbarp[k] += f * ms[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) wq[k] += f * qs[k];
    sabs += fabs(slip);
  }
  double wp[3], sw[6], w1[6];
  cpf_mv3(c.RWR, wq, wp);
  cpf_symsw(sig, wp, sw);
#pragma unroll
  for (int k = 0; k < 6; ++k) w1[k] = c.D[k] - dbarp[k];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += __ldg(c.C + 6 * i + j) * w1[j];
    R[i] = sig[i] - c.sn[i] - s + 2.0 * sw[i];
  }
  if (wq_out) { wq_out[0] = wq[0]; wq_out[1] = wq[1]; wq_out[2] = wq[2]; }
  double h = 0.0;
  if (want2) {
    double ht;
    const double hf = mm10_hfac(c, tt, &ht);
    h = c.ttn + c.theta_0 * (hf * sabs);
    R[6] = tt - h;
  }
  return h;
}

// Jacobian (mm10_formJ): J is NJ x NJ row-major, NJ = 6 (J11 only, predictor) or 7.
template <int NJ, class JT>
CPF_DI void mm10_jacobian(const Mm10Ctx& c, const double* sig, double tt, JT J) {
  double S[21], T[18], dps[6], wqs[3], wqf[3], es[6], sabs = 0.0, ssum = 0.0;
#pragma unroll
  for (int k = 0; k < 21; ++k) S[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 18; ++k) T[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) { dps[k] = 0.0; es[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < 3; ++k) { wqs[k] = 0.0; wqf[k] = 0.0; }
  const double itt = 1.0 / tt, dgtt = c.dg / tt, dif = c.tinc * c.iD_v, dgn = c.dg * c.rate_n / tt;
  MM10_PRAGMA(unroll MM10_UNROLL)
  for (int s = 0; s < c.nslip; ++s) {
    double ms[6], qs[3];
    mm10_slip_geom(c, s, ms, qs);
    const double rs = sig[0] * ms[0] + sig[1] * ms[1] + sig[2] * ms[2] + sig[3] * ms[3] + sig[4] * ms[4] + sig[5] * ms[5];
    const double p = cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0);
    const double slip = dgtt * p * rs;
    const double dgdt = dgn * p + dif;
    const double f = rs * dif + slip;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const double da = dgdt * ms[a];
#pragma unroll
      for (int b = a; b < 6; ++b) S[q++] += da * ms[b];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dk = dgdt * qs[k];
#pragma unroll
      for (int b = 0; b < 6; ++b) T[6 * k + b] += dk * ms[b];
      wqf[k] += f * qs[k];
      if (NJ == 7) wqs[k] += slip * qs[k];
    }
    if (NJ == 7) {
      const double sp = cpf_sgn(rs) * p;
#pragma unroll
      for (int k = 0; k < 6; ++k) { dps[k] += slip * ms[k]; es[k] += sp * ms[k]; }
      sabs += fabs(slip); ssum += slip;
    }
  }
  // J11 = C S + 2 Lsig (RWR T) + IW(wp) + I
  double Sf[36];
  { int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) { Sf[6 * a + b] = S[q]; Sf[6 * b + a] = S[q]; ++q; } }
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    double tcol[3] = {T[b], T[6 + b], T[12 + b]}, tc[3], sw[6];
    cpf_mv3(c.RWR, tcol, tc);
    cpf_symsw(sig, tc, sw);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double s = 2.0 * sw[a];
#pragma unroll
      for (int k = 0; k < 6; ++k) s += __ldg(c.C + 6 * a + k) * Sf[6 * k + b];
      J[NJ * a + b] = s;
    }
  }
  double w[3];
  cpf_mv3(c.RWR, wqf, w);
  J[NJ * 0 + 3] += 2.0 * w[2]; J[NJ * 0 + 5] += -2.0 * w[1];
  J[NJ * 1 + 3] += 2.0 * w[2]; J[NJ * 1 + 4] += -2.0 * w[0];
  J[NJ * 2 + 4] += 2.0 * w[0]; J[NJ * 2 + 5] += 2.0 * w[1];
  J[NJ * 3 + 0] += w[2]; J[NJ * 3 + 1] += -w[2]; J[NJ * 3 + 4] += -w[1]; J[NJ * 3 + 5] += w[0];
  J[NJ * 4 + 1] += w[0]; J[NJ * 4 + 2] += -w[0]; J[NJ * 4 + 3] += w[1]; J[NJ * 4 + 5] += w[2];
  J[NJ * 5 + 0] += w[1]; J[NJ * 5 + 2] += -w[1]; J[NJ * 5 + 3] += w[0]; J[NJ * 5 + 4] += -w[2];
#pragma unroll
  for (int a = 0; a < 6; ++a) J[NJ * a + a] += 1.0;
  if (NJ == 7) {
    // J12 = -(n/tt) [C dps + 2 symSW(sig, RWR wqs)]
    double wc[3], sw[6];
    cpf_mv3(c.RWR, wqs, wc);
    cpf_symsw(sig, wc, sw);
    const double nt = -c.rate_n / tt;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double s = 2.0 * sw[a];
#pragma unroll
      for (int k = 0; k < 6; ++k) s += __ldg(c.C + 6 * a + k) * dps[k];
      J[NJ * a + 6] = nt * s;
    }
    // J21 = -theta0 dg n / tt * hfac * sum sgn(rs) |rs/tt|^(n-1) ms
    double ht;
    const double hf = mm10_hfac(c, tt, &ht);
    const double fac = c.theta_0 * dgn * hf;
#pragma unroll
    for (int k = 0; k < 6; ++k) J[NJ * 6 + k] = -(fac * es[k]);
    // J22 (mm10_ehard_voche)
    const double ah = fabs(ht);
    const double pw = (c.voche_m == 1.0) ? ah : pow(ah, c.voche_m);
    const double A = -1.0 / c.tau_v - c.taul / ((tt - c.tau_y) * (tt - c.tau_y));
    const double etau = (c.voche_m * A * sabs / ah - ssum * c.rate_n / tt * cpf_sgn(ht)) * pw;
    J[NJ * 6 + 6] = 1.0 - c.theta_0 * etau;
  }
}

// mm10_solve (mm10_a.f:2860-3295): predictor on the stress with extrapolated hardening, then
// the coupled update.  x[7] in/out.  J7 receives the last Jacobian formed (lagged).
// Returns true on failure.
CPF_DI bool mm10_solve(const Mm10Ctx& c, double* x, double cos_ang_ttrate_dt, SArr J7, int* it_pred,
                       int* it_upd, double* h_last) {
  const double cc = 1.0e-4, red = 0.5;
  const int mls = 10, mmin = 1;
  bool fail = false;
  double inR1;
  {  // ---- predictor ----
    double x1[6], R1[7];
#pragma unroll
    for (int k = 0; k < 6; ++k) x1[k] = x[k];
    const double x2 = x[6] + cos_ang_ttrate_dt;
    mm10_resid(c, x1, x2, R1, false, nullptr);
    double nR1 = sqrt(R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] + R1[5] * R1[5]);
    inR1 = nR1;
    int iter = 0;
    while ((nR1 > c.atol1) && (nR1 / inR1 > c.rtol1)) {
      double J[36], mJ[36], dx[6], wv[6];
      mm10_jacobian<6, double*>(c, x1, x2, J);
#pragma unroll
      for (int k = 0; k < 36; ++k) mJ[k] = -J[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) dx[k] = R1[k];
      double dot = R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] + R1[5] * R1[5];
      const double ls1 = 0.5 * dot;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        wv[j] = J[j] * R1[0] + J[6 + j] * R1[1] + J[12 + j] * R1[2] + J[18 + j] * R1[3] + J[24 + j] * R1[4] + J[30 + j] * R1[5];
      cpf_lu_solve<6, 1>(mJ, dx);
      const double ls2 = cc * (dx[0] * wv[0] + dx[1] * wv[1] + dx[2] * wv[2] + dx[3] * wv[3] + dx[4] * wv[4] + dx[5] * wv[5]);
      double alpha = 1.0;
      int ls = 0;
      for (;;) {
        const double nlsx = ls1 + ls2 * alpha;
        double xn[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) xn[k] = x1[k] + alpha * dx[k];
        mm10_resid(c, xn, x2, R1, false, nullptr);
        dot = R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] + R1[5] * R1[5];
        nR1 = sqrt(dot);
        if ((0.5 * dot <= nlsx) || (ls > mls)) {
#pragma unroll
          for (int k = 0; k < 6; ++k) x1[k] = xn[k];
          break;
        }
        alpha = red * alpha; ls = ls + 1;
      }
      iter = iter + 1;
      bool nan = false;
#pragma unroll
      for (int k = 0; k < 6; ++k) nan = nan || isnan(x1[k]);
      if ((iter > c.miter) || nan) { fail = true; break; }
    }
    *it_pred += iter;
    if (!fail) {
#pragma unroll
      for (int k = 0; k < 6; ++k) x[k] = x1[k];
      x[6] = x2;
    }
  }
  if (fail) return true;  // reference still runs the update but discards its result (fail stays set)
  {  // ---- coupled update ----
    double R[7];
    double h = mm10_resid(c, x, x[6], R, true, nullptr);
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) dot += R[k] * R[k];
    double nR = sqrt(dot), inR = nR;
    if (inR == 0.0) inR = inR1;
    int iter = 0;
    while (((nR > c.atol) && (nR / inR > c.rtol)) || (iter < mmin)) {
      double mJ[49], dx[7], wv[7];
      mm10_jacobian<7, SArr>(c, x, x[6], J7);
#pragma unroll
      for (int k = 0; k < 49; ++k) mJ[k] = -J7[k];
      dot = 0.0;
#pragma unroll
      for (int k = 0; k < 7; ++k) { dx[k] = R[k]; dot += R[k] * R[k]; }
      const double ls1 = 0.5 * dot;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 7; ++i) s += J7[7 * i + j] * R[i];
        wv[j] = s;
      }
      cpf_lu_solve<7, 1>(mJ, dx);
      double d = 0.0;
#pragma unroll
      for (int k = 0; k < 7; ++k) d += dx[k] * wv[k];
      const double ls2 = cc * d;
      double alpha = 1.0;
      int ls = 0;
      for (;;) {
        const double nlsx = ls1 + ls2 * alpha;
        double xn[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) xn[k] = x[k] + alpha * dx[k];
        h = mm10_resid(c, xn, xn[6], R, true, nullptr);
        dot = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) dot += R[k] * R[k];
        nR = sqrt(dot);
        if ((0.5 * dot <= nlsx) || (ls > mls)) {
#pragma unroll
          for (int k = 0; k < 7; ++k) x[k] = xn[k];
          break;
        }
        alpha = red * alpha; ls = ls + 1;
      }
      iter = iter + 1;
      bool nan = false;
#pragma unroll
      for (int k = 0; k < 7; ++k) nan = nan || isnan(x[k]);
      if ((iter > c.miter) || nan) { fail = true; break; }
    }
    *it_upd += iter;
    *h_last = h;
  }
  return fail;
}
