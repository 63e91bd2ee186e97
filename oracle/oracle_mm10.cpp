// CPFFT ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Crystal plasticity model mm10, Voce hardening, Newton-Raphson local solver, one crystal
// per material point, restated from mm10_a.f / mm10_b.f / mm10_d.f / mod_crystals.f and
// setup_mm10_rknstr (drive_eps_sig.f:537-1002).
#include "oracle_internal.hpp"
#include <vector>
#include "slip_tables.inc"

namespace orc {

static const double PI = 3.141592653589793;  // mod_crystals.f:121

// ----------------------------------------------------------------------------
// small dense helpers (stand-ins for DGESV / DPOSV / DSYTRI)
static int lu_solve(int n, double* A /*n x n row-major, destroyed*/, double* b, int nrhs /*b is n x nrhs row-major*/) {
  for (int k = 0; k < n; ++k) {
    int piv = k; double best = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      double v = std::fabs(A[i * n + k]);
      if (v > best) { best = v; piv = i; }
    }
    if (best == 0.0) return k + 1;
    if (piv != k) {
      for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
      for (int j = 0; j < nrhs; ++j) { double t = b[k * nrhs + j]; b[k * nrhs + j] = b[piv * nrhs + j]; b[piv * nrhs + j] = t; }
    }
    double inv = 1.0 / A[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      double l = A[i * n + k] * inv;
      A[i * n + k] = l;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
      for (int j = 0; j < nrhs; ++j) b[i * nrhs + j] -= l * b[k * nrhs + j];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    for (int j = 0; j < nrhs; ++j) {
      double s = b[k * nrhs + j];
      for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * b[c * nrhs + j];
      b[k * nrhs + j] = s / A[k * n + k];
    }
  }
  return 0;
}

// ----------------------------------------------------------------------------
// mm10 rotation / Voigt helpers (mm10_a.f:1287-1534)
static void rotation_matrix_kocks_deg(const double* ang, M33 r) {  // mm10_a.f:1287-1345
  double psi = ang[0] * PI / 180.0, theta = ang[1] * PI / 180.0, phi = ang[2] * PI / 180.0;
  r[0][0] = -std::sin(psi) * std::sin(phi) - std::cos(psi) * std::cos(phi) * std::cos(theta);
  r[0][1] = std::cos(psi) * std::sin(phi) - std::sin(psi) * std::cos(phi) * std::cos(theta);
  r[0][2] = std::cos(phi) * std::sin(theta);
  r[1][0] = std::sin(psi) * std::cos(phi) - std::cos(psi) * std::sin(phi) * std::cos(theta);
  r[1][1] = -std::cos(psi) * std::cos(phi) - std::sin(psi) * std::sin(phi) * std::cos(theta);
  r[1][2] = std::sin(phi) * std::sin(theta);
  r[2][0] = std::cos(psi) * std::sin(theta);
  r[2][1] = std::sin(psi) * std::sin(theta);
  r[2][2] = std::cos(theta);
}
static void rt2rve(const M33 rt, M66 rv) {  // mm10_a.f:1400-1447
  const double two = 2.0;
  rv[0][0] = rt[0][0] * rt[0][0]; rv[0][1] = rt[0][1] * rt[0][1]; rv[0][2] = rt[0][2] * rt[0][2];
  rv[0][3] = two * rt[0][0] * rt[0][1]; rv[0][4] = two * rt[0][2] * rt[0][1]; rv[0][5] = two * rt[0][0] * rt[0][2];
  rv[1][0] = rt[1][0] * rt[1][0]; rv[1][1] = rt[1][1] * rt[1][1]; rv[1][2] = rt[1][2] * rt[1][2];
  rv[1][3] = two * rt[1][0] * rt[1][1]; rv[1][4] = two * rt[1][2] * rt[1][1]; rv[1][5] = two * rt[1][0] * rt[1][2];
  rv[2][0] = rt[2][0] * rt[2][0]; rv[2][1] = rt[2][1] * rt[2][1]; rv[2][2] = rt[2][2] * rt[2][2];
  rv[2][3] = two * rt[2][0] * rt[2][1]; rv[2][4] = two * rt[2][2] * rt[2][1]; rv[2][5] = two * rt[2][0] * rt[2][2];
  rv[3][0] = rt[0][0] * rt[1][0]; rv[3][1] = rt[0][1] * rt[1][1]; rv[3][2] = rt[0][2] * rt[1][2];
  rv[3][3] = rt[0][0] * rt[1][1] + rt[1][0] * rt[0][1];
  rv[3][4] = rt[0][1] * rt[1][2] + rt[0][2] * rt[1][1];
  rv[3][5] = rt[0][0] * rt[1][2] + rt[0][2] * rt[1][0];
  rv[4][0] = rt[1][0] * rt[2][0]; rv[4][1] = rt[2][1] * rt[1][1]; rv[4][2] = rt[1][2] * rt[2][2];
  rv[4][3] = rt[1][0] * rt[2][1] + rt[1][1] * rt[2][0];
  rv[4][4] = rt[1][1] * rt[2][2] + rt[2][1] * rt[1][2];
  rv[4][5] = rt[1][0] * rt[2][2] + rt[1][2] * rt[2][0];
  rv[5][0] = rt[0][0] * rt[2][0]; rv[5][1] = rt[0][1] * rt[2][1]; rv[5][2] = rt[0][2] * rt[2][2];
  rv[5][3] = rt[0][0] * rt[2][1] + rt[0][1] * rt[2][0];
  rv[5][4] = rt[0][1] * rt[2][2] + rt[0][2] * rt[2][1];
  rv[5][5] = rt[0][0] * rt[2][2] + rt[2][0] * rt[0][2];
}
static void rt2rvw(const M33 rt, M33 rv) {  // mm10_a.f:1461-1479
  rv[0][0] = rt[1][1] * rt[2][2] - rt[1][2] * rt[2][1];
  rv[0][1] = rt[1][0] * rt[2][2] - rt[1][2] * rt[2][0];
  rv[0][2] = rt[1][0] * rt[2][1] - rt[1][1] * rt[2][0];
  rv[1][0] = rt[0][1] * rt[2][2] - rt[0][2] * rt[2][1];
  rv[1][1] = rt[0][0] * rt[2][2] - rt[0][2] * rt[2][0];
  rv[1][2] = rt[0][0] * rt[2][1] - rt[0][1] * rt[2][0];
  rv[2][0] = rt[0][1] * rt[1][2] - rt[0][2] * rt[1][1];
  rv[2][1] = rt[0][0] * rt[1][2] - rt[0][2] * rt[1][0];
  rv[2][2] = rt[0][0] * rt[1][1] - rt[0][1] * rt[1][0];
}
static inline void mat33(const M33 b, const M33 c, M33 a) {  // mm10_a_mult_type_1: a = b*c
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) a[i][j] = b[i][0] * c[0][j] + b[i][1] * c[1][j] + b[i][2] * c[2][j];
}
static inline void matvec6(const M66 b, const double* c, double* a) {  // mm10_a_mult_type_2
  for (int i = 0; i < 6; ++i) a[i] = b[i][0] * c[0];
  for (int j = 1; j < 6; ++j)
    for (int i = 0; i < 6; ++i) a[i] = a[i] + b[i][j] * c[j];
}
static inline void matvec3(const M33 b, const double* c, double* a) {  // mm10_a_mult_type_3
  for (int i = 0; i < 3; ++i) a[i] = b[i][0] * c[0] + b[i][1] * c[1] + b[i][2] * c[2];
}

// ----------------------------------------------------------------------------
void finalize_crystal(CrystalLib& c) {  // mod_crystals.f:414-1931 (fcc, bcc48; isotropic/cubic)
  const int (*tab)[6]; int n;
  if (c.in.slip_type == 1) { tab = ORC_SLIP_FCC; n = 12; }
  else if (c.in.slip_type == 2) { tab = ORC_SLIP_BCC; n = 12; }
  else if (c.in.slip_type == 3) { tab = ORC_SLIP_SINGLE; n = 1; }
  else if (c.in.slip_type == 6) { tab = ORC_SLIP_ROTERS; n = 12; }
  else if (c.in.slip_type == 7) { tab = ORC_SLIP_BCC12; n = 12; }
  else if (c.in.slip_type == 8) { tab = ORC_SLIP_BCC48; n = 48; }
  else { std::fprintf(stderr, "oracle: unsupported slip_type %d\n", c.in.slip_type); n = 0; tab = ORC_SLIP_FCC; }
  c.nslip = n;
  for (int s = 0; s < n; ++s) {
    double sb = 0, sn = 0;
    for (int k = 0; k < 3; ++k) { sb += tab[s][k] * tab[s][k]; sn += tab[s][3 + k] * tab[s][3 + k]; }
    for (int k = 0; k < 3; ++k) {
      c.bi[s][k] = (double)tab[s][k] / std::sqrt(sb);
      c.ni[s][k] = (double)tab[s][3 + k] / std::sqrt(sn);
    }
  }
  double e = c.in.e, v = c.in.nu, u = c.in.mu;
  double flex[36] = {0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) flex[i * 6 + j] = (i == j) ? 1 / e : -v / e;
  double sh = (c.in.elastic_type == 1) ? 2 * (1 + v) / e : 1 / u;
  flex[3 * 6 + 3] = flex[4 * 6 + 4] = flex[5 * 6 + 5] = sh;
  double inv[36] = {0};
  for (int i = 0; i < 6; ++i) inv[i * 6 + i] = 1.0;
  lu_solve(6, flex, inv, 6);  // stand-in for mm10_invsym (DSYTRF/DSYTRI), mod_crystals.f:1878-1879
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) c.elast_stiff[i][j] = 0.5 * (inv[i * 6 + j] + inv[j * 6 + i]);
}

HistLayout mm10_history_layout(int nslip, int num_hard) {  // mm10_d.f:137-331
  HistLayout L;
  L.use_max = (num_hard == 48 || nslip == 48) ? 1 : 0;
  L.nslip = nslip; L.num_hard = num_hard;
  int lc5 = L.use_max ? 48 : nslip;
  L.cep = 0; L.gradfe = 36; L.R = 63; L.work = 72; L.slipsum = 75;
  int common = 75 + lc5;
  int l6 = L.use_max ? 48 : nslip, l7 = L.use_max ? 48 : num_hard, l8 = L.use_max ? 48 : 15,
      l9 = L.use_max ? 48 : num_hard;
  L.len_slip = l6; L.len_u = l8;
  L.c_stress = common; L.c_euler = L.c_stress + 6; L.c_Rp = L.c_euler + 3; L.c_D = L.c_Rp + 9;
  L.c_eps = L.c_D + 6; L.c_slipinc = L.c_eps + 6; L.c_tt = L.c_slipinc + l6; L.c_u = L.c_tt + l7;
  L.c_ttrate = L.c_u + l8; L.c_ep = L.c_ttrate + l9; L.c_ed = L.c_ep + 6;
  L.total = L.c_ed + 6;
  return L;
}

// ----------------------------------------------------------------------------
struct Props {   // crystal_props subset actually reached by the Voce / MTS NR path
  int nslip, alter_mode, miter, h_type;
  double rate_n, theta_0, tau_y, tau_v, voche_m, iD_v, eps_dot_0_y, k_0, burgers;
  // MTS (mm10_a.f:2109-2175, mm10_b.f:2080-2345)
  double tau_a, tau_hat_y, G_0_y, tau_hat_v, G_0_v, p_y, q_y, p_v, q_v, boltzman, eps_dot_0_v, mu_0, D_0, T_0;
  double atol, atol1, rtol, rtol1;
  M33 g;
  double ms[ORC_MAX_SLIP][6], qs[ORC_MAX_SLIP][3], ns[ORC_MAX_SLIP][3];
  M66 stiffness;
};
struct State {   // crystal_state subset
  M33 R, Rp;
  double stress[6], D[6], eps[6], euler[3], slip_incs[ORC_MAX_SLIP];
  double tau_tilde, tt_rate, u[16], ep[6], ed[6];
  M66 tangent;
  double ms[ORC_MAX_SLIP][6], qs[ORC_MAX_SLIP][3], qc[ORC_MAX_SLIP][3], tau_l[ORC_MAX_SLIP];
  double dg, tinc, temp, mu_harden, work_inc, p_work_inc, p_strain_inc;
  double tau_y, tau_v;   // MTS threshold contributions of the (sub)step
};

// setup_mm10_rknstr (drive_eps_sig.f:537-1002): per point, per call
static void setup_props(const CrystalLib& cry, const double* angles, Props& p) {
  const orc_crystal& in = cry.in;
  p.nslip = cry.nslip; p.alter_mode = in.alter_mode; p.miter = in.miter;
  p.rate_n = in.harden_n; p.theta_0 = in.theta_0; p.tau_y = in.tau_y; p.tau_v = in.tau_v;
  p.voche_m = in.voche_m; p.iD_v = in.iD_v; p.eps_dot_0_y = in.eps_dot_0_y; p.k_0 = in.k_0;
  p.burgers = in.burgers; p.atol = in.atol; p.atol1 = in.atol1; p.rtol = in.rtol; p.rtol1 = in.rtol1;
  p.h_type = in.h_type;
  p.tau_a = in.tau_a; p.tau_hat_y = in.tau_hat_y; p.G_0_y = in.g_0_y; p.tau_hat_v = in.tau_hat_v; p.G_0_v = in.g_0_v;
  p.p_y = in.p_y; p.q_y = in.q_y; p.p_v = in.p_v; p.q_v = in.q_v; p.boltzman = in.boltzman;
  p.eps_dot_0_v = in.eps_dot_0_v; p.mu_0 = in.mu_0; p.D_0 = in.D_0; p.T_0 = in.T_0;
  // the diffusion slip rs dt iD_v exists in the Voce branches only (mm10_b.f:1180-1200, 1253-1270,
  // 1325-1345, 2005-2030): with iD_v = 0 the shared code below adds exact zeros for MTS
  if (p.h_type == 2) p.iD_v = 0.0;
  rotation_matrix_kocks_deg(angles, p.g);
  M33 trot;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) trot[i][j] = p.g[j][i];
  for (int s = 0; s < p.nslip; ++s) {
    double bs[3], ns[3];
    matvec3(trot, cry.bi[s], bs);
    matvec3(trot, cry.ni[s], ns);
    for (int k = 0; k < 3; ++k) p.ns[s][k] = ns[k];
    M33 A, sy, as;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i][j] = bs[i] * ns[j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { sy[i][j] = 0.5 * (A[i][j] + A[j][i]); as[i][j] = 0.5 * (A[i][j] - A[j][i]); }
    p.ms[s][0] = sy[0][0]; p.ms[s][1] = sy[1][1]; p.ms[s][2] = sy[2][2];
    p.ms[s][3] = 2.0 * sy[0][1]; p.ms[s][4] = 2.0 * sy[1][2]; p.ms[s][5] = 2.0 * sy[0][2];
    p.qs[s][0] = as[1][2]; p.qs[s][1] = as[0][2]; p.qs[s][2] = as[0][1];
  }
  M66 Rs, tmp;
  rt2rve(trot, Rs);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += cry.elast_stiff[i][k] * Rs[j][k];  // C * Rstiff^T
      tmp[i][j] = s;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += Rs[i][k] * tmp[k][j];
      p.stiffness[i][j] = s;
    }
}

// mm10_setup_np1 (mm10_a.f:1659-1732)
static void setup_np1(const M33 R, const double* D, double dt, State& s) {
  std::memset(&s, 0, sizeof(State));
  s.temp = 297.0; s.tinc = dt;
  for (int i = 0; i < 6; ++i) s.D[i] = D[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) s.R[i][j] = R[i][j];
}

// mm10_setup_mts (mm10_a.f:2109-2175).  Also fills the n-state values that are kept as flags
// (< 0) in the history until the first plastic update: n.tau_y, n.mu_harden, n.tau_tilde.
static void mm10_setup_mts(const Props& p, State& np1, State& n) {
  const double init_hard = 0.1;
  const double dgc = np1.dg / np1.tinc;
  if (np1.temp == 0.0) np1.mu_harden = p.mu_0;
  else np1.mu_harden = p.mu_0 - p.D_0 / (std::exp(p.T_0 / np1.temp) - 1.0);
  if (dgc == 0.0) {
    np1.tau_v = p.tau_hat_v;
    np1.tau_y = p.tau_hat_y;
  } else {
    np1.tau_v = p.tau_hat_v * std::pow(1.0 - std::pow(p.boltzman * np1.temp / (np1.mu_harden * std::pow(p.burgers, 3.0) * p.G_0_v) *
                                                         std::log(p.eps_dot_0_v / dgc), 1.0 / p.q_v), 1.0 / p.p_v);
    np1.tau_y = p.tau_hat_y * std::pow(1.0 - std::pow(p.boltzman * np1.temp / (np1.mu_harden * std::pow(p.burgers, 3.0) * p.G_0_y) *
                                                         std::log(p.eps_dot_0_y / dgc), 1.0 / p.q_y), 1.0 / p.p_y);
  }
  np1.u[0] = np1.tau_y;
  if (n.u[0] < 0.0) n.tau_y = np1.tau_y; else n.tau_y = n.u[0];
  np1.u[1] = np1.mu_harden;
  if (n.u[1] < 0.0) n.mu_harden = np1.mu_harden; else n.mu_harden = n.u[1];
  if (n.tau_tilde < 0.0) n.tau_tilde = p.tau_a + (np1.mu_harden / p.mu_0) * np1.tau_y + init_hard;
}

// mm10_setup + mm10_setup_voche (mm10_a.f:830-962, 2057-2075)
static void mm10_setup(const Props& p, State& np1, State& n) {
  double t1 = np1.D[0] * np1.D[0] + np1.D[1] * np1.D[1] + np1.D[2] * np1.D[2];
  double t2 = np1.D[3] * np1.D[3] + np1.D[4] * np1.D[4] + np1.D[5] * np1.D[5];
  const double twothirds = 2.0 / 3.0;
  np1.dg = std::sqrt(twothirds * (t1 + 0.5 * t2));
  M33 tRp, work, RW, RWC; M66 RE;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tRp[i][j] = n.Rp[j][i];
  rt2rve(tRp, RE);
  rt2rvw(tRp, RW);
  mat33(np1.R, tRp, work);
  rt2rvw(work, RWC);
  for (int i = 0; i < p.nslip; ++i) {
    matvec6(RE, p.ms[i], np1.ms[i]);
    matvec3(RW, p.qs[i], np1.qs[i]);
    matvec3(RWC, p.qs[i], np1.qc[i]);
  }
  if (p.h_type == 2) mm10_setup_mts(p, np1, n);
  else {
    np1.mu_harden = p.stiffness[5][5];
    if (p.alter_mode) np1.dg = p.eps_dot_0_y * np1.tinc;
  }
  // geometric hardening: curvature is identically zero (gradFeinv never filled)
  const double alpha = 1.0 / 3.0;
  double cst = p.k_0 * p.burgers * alpha * alpha * np1.mu_harden * np1.mu_harden / 2.0 / p.theta_0;
  for (int t = 0; t < p.nslip; ++t) np1.tau_l[t] = cst * std::sqrt(0.0);
}

static inline double mm10_rs(const State& np1, const double* stress, int i) {  // mm10_b.f:1837-1858
  return stress[0] * np1.ms[i][0] + stress[1] * np1.ms[i][1] + stress[2] * np1.ms[i][2] +
         stress[3] * np1.ms[i][3] + stress[4] * np1.ms[i][4] + stress[5] * np1.ms[i][5];
}
static inline double mm10_slipinc(const Props& p, const State& np1, const double* stress, double tt, int i) {
  double rs = mm10_rs(np1, stress, i);                                   // mm10_b.f:1805-1824
  return np1.dg / tt * std::pow(std::fabs(rs / tt), p.rate_n - 1.0) * rs;
}
static void symsw(const double* s, const double* w, double* sw) {  // mm10_b.f:1505-1524
  sw[0] = s[3] * w[2] - s[5] * w[1];
  sw[1] = s[3] * w[2] - s[4] * w[0];
  sw[2] = s[5] * w[1] + s[4] * w[0];
  sw[3] = 0.5 * (w[2] * (s[0] - s[1]) + w[0] * s[5] - w[1] * s[4]);
  sw[4] = 0.5 * (w[0] * (s[1] - s[2]) + w[1] * s[3] + w[2] * s[5]);
  sw[5] = 0.5 * (w[1] * (s[0] - s[2]) + w[0] * s[3] - w[2] * s[4]);
}
static void form_dbarp(const Props& p, const State& np1, const double* stress, double tt, double* dbar) {
  for (int k = 0; k < 6; ++k) dbar[k] = 0.0;                              // mm10_b.f:1161-1219
  for (int i = 0; i < p.nslip; ++i) {
    double slipinc = mm10_slipinc(p, np1, stress, tt, i);
    double rs = mm10_rs(np1, stress, i);
    double f = (rs * np1.tinc * p.iD_v + slipinc);
    for (int k = 0; k < 6; ++k) dbar[k] = dbar[k] + f * np1.ms[i][k];
  }
}
static void form_wp(const Props& p, const State& np1, const double* stress, double tt, double* w) {
  w[0] = w[1] = w[2] = 0.0;                                               // mm10_b.f:1304-1381
  for (int i = 0; i < p.nslip; ++i) {
    double slipinc = mm10_slipinc(p, np1, stress, tt, i);
    for (int k = 0; k < 3; ++k) w[k] = w[k] + slipinc * np1.qc[i][k];
    double rs = mm10_rs(np1, stress, i);
    for (int k = 0; k < 3; ++k) w[k] = w[k] + rs * np1.tinc * p.iD_v * np1.qc[i][k];
  }
}
static void form_wbarp(const Props& p, const State& np1, const double* stress, double tt, double* w) {
  w[0] = w[1] = w[2] = 0.0;                                               // mm10_b.f:1231-1290
  for (int i = 0; i < p.nslip; ++i) {
    double slipinc = mm10_slipinc(p, np1, stress, tt, i);
    for (int k = 0; k < 3; ++k) w[k] = w[k] + slipinc * np1.qs[i][k];
    double rs = mm10_rs(np1, stress, i);
    for (int k = 0; k < 3; ++k) w[k] = w[k] + rs * np1.tinc * p.iD_v * np1.qs[i][k];
  }
}
static void formR1(const Props& p, const State& np1, const State& n, const double* stress, double tt, double* R1) {
  double dbarp[6], wp[3], symTW[6], w1[6], w2[6];                          // mm10_b.f:1065-1101
  form_dbarp(p, np1, stress, tt, dbarp);
  form_wp(p, np1, stress, tt, wp);
  symsw(stress, wp, symTW);
  for (int k = 0; k < 6; ++k) w1[k] = np1.D[k] - dbarp[k];
  matvec6(p.stiffness, w1, w2);
  for (int k = 0; k < 6; ++k) R1[k] = stress[k] - n.stress[k] - w2[k] + 2.0 * symTW[k];
}
static inline double sgn(double x) { return x >= 0.0 ? 1.0 : -1.0; }       // Fortran sign(one,x)
static double h_voche(const Props& p, const State& np1, const State& n, const double* stress, double tt) {
  double h = 0.0;                                                           // mm10_b.f:1888-1915
  for (int i = 0; i < p.nslip; ++i) {
    double slipinc = mm10_slipinc(p, np1, stress, tt, i);
    double h_term = 1.0 - (tt - p.tau_y) / p.tau_v + np1.tau_l[i] / (tt - p.tau_y);
    h = h + std::pow(std::fabs(h_term), p.voche_m) * sgn(h_term) * std::fabs(slipinc);
  }
  return n.tau_tilde + p.theta_0 * h;
}
static double h_mts(const Props& p, const State& np1, const State& n, const double* stress, double tt) {
  const double cta = (p.mu_0 / np1.mu_harden) * tt - (p.mu_0 / np1.mu_harden) * p.tau_a - np1.tau_y;   // mm10_b.f:2080-2111
  const double ct = 1.0 - cta / np1.tau_v;
  double h = 0.0;
  for (int i = 0; i < p.nslip; ++i) {
    double slipinc = mm10_slipinc(p, np1, stress, tt, i);
    h = h + std::pow(ct + np1.tau_l[i] / cta, p.voche_m) * std::fabs(slipinc);
  }
  return p.tau_a * (1.0 - np1.mu_harden / n.mu_harden) + (np1.mu_harden / p.mu_0) * (np1.tau_y - n.tau_y) +
         (np1.mu_harden / n.mu_harden) * n.tau_tilde + p.theta_0 * (np1.mu_harden / p.mu_0) * h;
}
static void formR2(const Props& p, State& np1, const State& n, const double* stress, double tt, double* R2) {
  double h = (p.h_type == 2) ? h_mts(p, np1, n, stress, tt) : h_voche(p, np1, n, stress, tt);   // mm10_b.f:63-113
  *R2 = tt - h;
  np1.tt_rate = (h - n.tau_tilde) / np1.tinc;
}
static void formR(const Props& p, State& np1, const State& n, const double* x, double* R) {
  formR1(p, np1, n, x, x[6], R);                                           // mm10_b.f:1029-1051
  formR2(p, np1, n, x, x[6], R + 6);
}
static void symswmat_col(const double* s, const double* w, double* sw) { symsw(s, w, sw); }

static void dgdt_voche(const Props& p, const State& np1, const double* stress, double tt, double* dg) {
  for (int s = 0; s < p.nslip; ++s) {                                       // mm10_b.f:2005-2030
    double rs = mm10_rs(np1, stress, s);
    double d = std::pow(std::fabs(rs), p.rate_n - 1.0);
    d = np1.dg * p.rate_n / std::pow(tt, p.rate_n) * d;
    dg[s] = d + np1.tinc * p.iD_v;
  }
}
static void formJ11(const Props& p, const State& np1, const double* stress, double tt, M66 J11) {
  double dgdt[ORC_MAX_SLIP];                                                // mm10_b.f:177-257
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) J11[i][j] = 0.0;
  dgdt_voche(p, np1, stress, tt, dgdt);
  for (int i = 0; i < p.nslip; ++i) {
    double symtq[6], wv[6];
    symswmat_col(stress, np1.qc[i], symtq);
    for (int k = 0; k < 6; ++k) wv[k] = 2.0 * symtq[k];                     // mm10_b_mult_type_4
    for (int j = 0; j < 6; ++j)
      for (int k = 0; k < 6; ++k) wv[k] = wv[k] + p.stiffness[k][j] * np1.ms[i][j];
    for (int a = 0; a < 6; ++a)                                             // DGER
      for (int b = 0; b < 6; ++b) J11[a][b] += dgdt[i] * wv[a] * np1.ms[i][b];
  }
  double w[3];
  form_wp(p, np1, stress, tt, w);
  // mm10_iw (mm10_b.f:1605-1634)
  J11[0][3] += 2.0 * w[2]; J11[0][5] += -2.0 * w[1];
  J11[1][3] += 2.0 * w[2]; J11[1][4] += -2.0 * w[0];
  J11[2][4] += 2.0 * w[0]; J11[2][5] += 2.0 * w[1];
  J11[3][0] += w[2]; J11[3][1] += -w[2]; J11[3][4] += -w[1]; J11[3][5] += w[0];
  J11[4][1] += w[0]; J11[4][2] += -w[0]; J11[4][3] += w[1]; J11[4][5] += w[2];
  J11[5][0] += w[1]; J11[5][2] += -w[1]; J11[5][3] += w[0]; J11[5][4] += -w[2];
  for (int i = 0; i < 6; ++i) J11[i][i] = J11[i][i] + 1.0;
}
static void formJ(const Props& p, const State& np1, const double* x, double J[7][7]) {
  const double* stress = x; double tt = x[6];                               // mm10_b.f:901-954
  M66 J11;
  formJ11(p, np1, stress, tt, J11);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) J[i][j] = J11[i][j];
  // J12 (mm10_b.f:271-353) with dgdh_voche (mm10_b.f:2033-2053)
  for (int k = 0; k < 6; ++k) J[k][6] = 0.0;
  for (int i = 0; i < p.nslip; ++i) {
    double symtq[6], tm[6];
    symswmat_col(stress, np1.qc[i], symtq);
    for (int k = 0; k < 6; ++k) tm[k] = 2.0 * symtq[k];
    for (int j = 0; j < 6; ++j)
      for (int k = 0; k < 6; ++k) tm[k] = tm[k] + p.stiffness[k][j] * np1.ms[i][j];
    double dgam = mm10_slipinc(p, np1, stress, tt, i);
    double dgdtt = -p.rate_n / tt * dgam;
    for (int k = 0; k < 6; ++k) J[k][6] += tm[k] * dgdtt;
  }
  if (p.h_type == 2) {
    // J21 = -estress_mts (mm10_b.f:2114-2147), J22 = ehard_mts (mm10_b.f:2150-2186)
    const double cta = (p.mu_0 / np1.mu_harden) * tt - (p.mu_0 / np1.mu_harden) * p.tau_a - np1.tau_y;
    const double ct = 1.0 - cta / np1.tau_v;
    const double ur = np1.mu_harden / p.mu_0;
    double et[6] = {0, 0, 0, 0, 0, 0}, etau = 0.0;
    for (int i = 0; i < p.nslip; ++i) {
      double rs = mm10_rs(np1, stress, i);
      double f = std::pow(ct + np1.tau_l[i] / cta, p.voche_m) * std::pow(std::fabs(rs), p.rate_n - 2.0) * rs;
      for (int k = 0; k < 6; ++k) et[k] = et[k] + f * np1.ms[i][k];
      double slipinc = mm10_slipinc(p, np1, stress, tt, i);
      etau = etau + (p.voche_m * (1.0 / np1.tau_v + np1.tau_l[i] / (cta * cta)) * std::pow(ct + np1.tau_l[i] / cta, -1.0) +
                     ur * p.rate_n / tt) * std::pow(ct + np1.tau_l[i] / cta, p.voche_m) * std::fabs(slipinc);
    }
    const double fac = p.theta_0 * (np1.mu_harden / p.mu_0);
    for (int k = 0; k < 6; ++k) J[6][k] = -(fac * et[k] * p.rate_n * np1.dg / std::pow(tt, p.rate_n));
    etau = -p.theta_0 * etau;
    J[6][6] = 1.0 - etau;
    return;
  }
  // J21 = -estress (mm10_b.f:369-417, 1918-1948)
  double et[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < p.nslip; ++i) {
    double rs = mm10_rs(np1, stress, i);
    double h_term = 1.0 - (tt - p.tau_y) / p.tau_v + np1.tau_l[i] / (tt - p.tau_y);
    double f = std::pow(std::fabs(h_term), p.voche_m) * sgn(h_term) * std::pow(std::fabs(rs), p.rate_n - 2.0) * rs;
    for (int k = 0; k < 6; ++k) et[k] = et[k] + f * np1.ms[i][k];
  }
  double fac = p.theta_0 * np1.dg * p.rate_n / std::pow(tt, p.rate_n);
  for (int k = 0; k < 6; ++k) J[6][k] = -(fac * et[k]);
  // J22 = ehard (mm10_b.f:431-481, 1951-1982)
  double etau = 0.0;
  for (int i = 0; i < p.nslip; ++i) {
    double slipinc = mm10_slipinc(p, np1, stress, tt, i);
    double h_term = 1.0 - (tt - p.tau_y) / p.tau_v + np1.tau_l[i] / (tt - p.tau_y);
    etau = etau + (p.voche_m * (-1.0 / p.tau_v - np1.tau_l[i] / ((tt - p.tau_y) * (tt - p.tau_y))) *
                       std::fabs(slipinc) / std::fabs(h_term) -
                   std::fabs(slipinc) * p.rate_n / tt * sgn(h_term) * sgn(slipinc)) *
                      (std::pow(std::fabs(h_term), p.voche_m));
  }
  etau = p.theta_0 * etau;
  J[6][6] = 1.0 - etau;
}

// mm10_solve (mm10_a.f:2860-3295).  Returns fail flag.  J is the last Jacobian formed.
static bool mm10_solve(const Props& p, State& np1, const State& n, double* stress, double* tt,
                       double dtinc, double Jout[7][7], int* iters) {
  const double c = 1.0e-4, red = 0.5; const int mls = 10, mmin = 1;
  double x[7]; for (int k = 0; k < 6; ++k) x[k] = stress[k]; x[6] = *tt;
  bool fail = false;
  // ---- mm10_solve_predict (mm10_a.f:2975-3156) ----
  double inR1;
  {
    int iter = 0;
    double x1[6]; for (int k = 0; k < 6; ++k) x1[k] = x[k];
    double d1[6], d2[6];
    for (int k = 0; k < 6; ++k) d1[k] = np1.D[k];
    double dtrace = (d1[0] + d1[1] + d1[2]) / 3.0;
    d1[0] -= dtrace; d1[1] -= dtrace; d1[2] -= dtrace;
    double t1 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
    double t2 = d1[3] * d1[3] + d1[4] * d1[4] + d1[5] * d1[5];
    if (t1 + t2 == 0.0) { for (int k = 0; k < 6; ++k) d1[k] = 0.0; }
    else { double s = std::sqrt(t1 + t2); for (int k = 0; k < 6; ++k) d1[k] = d1[k] / s; }
    for (int k = 0; k < 6; ++k) d2[k] = n.D[k];
    dtrace = (d2[0] + d2[1] + d2[2]) / 3.0;
    d2[0] -= dtrace; d2[1] -= dtrace; d2[2] -= dtrace;
    t1 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
    t2 = d2[3] * d2[3] + d2[4] * d2[4] + d2[5] * d2[5];
    if (t1 + t2 == 0.0) { for (int k = 0; k < 6; ++k) d2[k] = 0.0; }
    else { double s = std::sqrt(t1 + t2); for (int k = 0; k < 6; ++k) d2[k] = d2[k] / s; }
    t1 = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
    t2 = d1[3] * d2[3] + d1[4] * d2[4] + d1[5] * d2[5];
    double cos_ang = std::fmax(t1 + t2, 0.0);
    double x2 = x[6] + cos_ang * n.tt_rate * dtinc;
    double R1[6];
    formR1(p, np1, n, x1, x2, R1);
    double nR1 = std::sqrt(R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] + R1[5] * R1[5]);
    inR1 = nR1;
    while ((nR1 > p.atol1) && (nR1 / inR1 > p.rtol1)) {
      M66 J11;
      formJ11(p, np1, x1, x2, J11);
      double dx1[6], mJ[36];
      for (int k = 0; k < 6; ++k) dx1[k] = R1[k];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) mJ[i * 6 + j] = -J11[i][j];
      lu_solve(6, mJ, dx1, 1);
      double alpha = 1.0;
      double dotR1 = (R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] + R1[5] * R1[5]);
      double ls1 = 0.5 * dotR1;
      double wv[6];
      for (int j = 0; j < 6; ++j)  // trans(J11) x R1, mm10_a_mult_type_2t
        wv[j] = J11[0][j] * R1[0] + J11[1][j] * R1[1] + J11[2][j] * R1[2] + J11[3][j] * R1[3] +
                J11[4][j] * R1[4] + J11[5][j] * R1[5];
      double ls2 = c * (dx1[0] * wv[0] + dx1[1] * wv[1] + dx1[2] * wv[2] + dx1[3] * wv[3] + dx1[4] * wv[4] + dx1[5] * wv[5]);
      int ls = 0;
      for (;;) {
        double nlsx = ls1 + ls2 * alpha;
        double xnew1[6];
        for (int k = 0; k < 6; ++k) xnew1[k] = x1[k] + alpha * dx1[k];
        formR1(p, np1, n, xnew1, x2, R1);
        dotR1 = (R1[0] * R1[0] + R1[1] * R1[1] + R1[2] * R1[2] + R1[3] * R1[3] + R1[4] * R1[4] + R1[5] * R1[5]);
        nR1 = std::sqrt(dotR1);
        double nRs = 0.5 * dotR1;
        if ((nRs <= nlsx) || (ls > mls)) { for (int k = 0; k < 6; ++k) x1[k] = xnew1[k]; break; }
        alpha = red * alpha; ls = ls + 1;
      }
      iter = iter + 1;
      bool anynan = false;
      for (int k = 0; k < 6; ++k) anynan = anynan || std::isnan(x1[k]);
      if ((iter > p.miter) || anynan) { fail = true; break; }
    }
    iters[0] += iter;
    if (!fail) { for (int k = 0; k < 6; ++k) x[k] = x1[k]; x[6] = x2; }
  }
  // ---- mm10_solve_update (mm10_a.f:3162-3293): runs even after a failed predictor; fail stays set
  {
    int iter = 0;
    double R[7];
    formR(p, np1, n, x, R);
    double dot = 0; for (int k = 0; k < 7; ++k) dot += R[k] * R[k];
    double nR = std::sqrt(dot), inR = nR;
    if (inR == 0.0) inR = inR1;
    while (((nR > p.atol) && (nR / inR > p.rtol)) || (iter < mmin)) {
      double J[7][7];
      formJ(p, np1, x, J);
      for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) Jout[i][j] = J[i][j];
      double dx[7], mJ[49];
      for (int k = 0; k < 7; ++k) dx[k] = R[k];
      for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) mJ[i * 7 + j] = -J[i][j];
      lu_solve(7, mJ, dx, 1);
      double alpha = 1.0;
      double dotR = 0; for (int k = 0; k < 7; ++k) dotR += R[k] * R[k];
      double ls1 = 0.5 * dotR;
      double wv[7];
      for (int j = 0; j < 7; ++j) { double s = 0; for (int i = 0; i < 7; ++i) s += J[i][j] * R[i]; wv[j] = s; }
      double d = 0; for (int k = 0; k < 7; ++k) d += dx[k] * wv[k];
      double ls2 = c * d;
      int ls = 0;
      for (;;) {
        double nlsx = ls1 + ls2 * alpha;
        double xnew[7];
        for (int k = 0; k < 7; ++k) xnew[k] = x[k] + alpha * dx[k];
        formR(p, np1, n, xnew, R);
        dotR = 0; for (int k = 0; k < 7; ++k) dotR += R[k] * R[k];
        nR = std::sqrt(dotR);
        double nRs = 0.5 * dotR;
        if ((nRs <= nlsx) || (ls > mls)) { for (int k = 0; k < 7; ++k) x[k] = xnew[k]; break; }
        alpha = red * alpha; ls = ls + 1;
      }
      iter = iter + 1;
      bool anynan = false;
      for (int k = 0; k < 7; ++k) anynan = anynan || std::isnan(x[k]);
      if ((iter > p.miter) || anynan) { fail = true; break; }
    }
    iters[1] += iter;
  }
  for (int k = 0; k < 6; ++k) stress[k] = x[k];
  *tt = x[6];
  return fail;
}

// mm10_ed_mts (mm10_b.f:2189-2264): derivative of the hardening function wrt the strain increment
static void ed_mts(const Props& p, const State& np1, const double* stress, double tt, double* ed) {
  double d_mod[6], dydd[6], dvdd[6];
  for (int k = 0; k < 6; ++k) d_mod[k] = (k < 3) ? np1.D[k] : 0.5 * np1.D[k];
  const double dgc = np1.dg / np1.tinc;
  const double b3 = std::pow(p.burgers, 3.0);
  const double lny = std::log(p.eps_dot_0_y / dgc);
  const double ty = p.boltzman * np1.temp / (np1.mu_harden * b3 * p.G_0_y) * lny;
  const double cy = 2.0 * p.tau_hat_y / (3.0 * np1.dg * np1.dg * p.q_y * p.p_y * lny) *
                    std::pow(1.0 - std::pow(ty, 1.0 / p.q_y), 1.0 / p.p_y - 1.0) * std::pow(ty, 1.0 / p.q_y);
  const double lnv = std::log(p.eps_dot_0_v / dgc);
  const double tv = p.boltzman * np1.temp / (np1.mu_harden * b3 * p.G_0_v) * lnv;
  const double cv = 2.0 * p.tau_hat_v / (3.0 * np1.dg * np1.dg * p.q_v * p.p_v * lnv) *
                    std::pow(1.0 - std::pow(tv, 1.0 / p.q_v), 1.0 / p.p_v - 1.0) * std::pow(tv, 1.0 / p.q_v);
  for (int k = 0; k < 6; ++k) { dydd[k] = cy * d_mod[k]; dvdd[k] = cv * d_mod[k]; }
  const double mnp0 = np1.mu_harden / p.mu_0;
  const double sc = tt / mnp0 - p.tau_a / mnp0 - np1.tau_y;
  for (int k = 0; k < 6; ++k) ed[k] = 0.0;
  for (int s = 0; s < p.nslip; ++s) {
    const double slipinc = mm10_slipinc(p, np1, stress, tt, s);
    const double base = 1.0 - sc / np1.tau_v + np1.tau_l[s] / sc;
    const double a = p.voche_m * (1.0 / np1.tau_v + np1.tau_l[s] / (sc * sc)) * std::pow(base, p.voche_m - 1.0);
    const double b = p.voche_m / (np1.tau_v * np1.tau_v) * sc * std::pow(base, p.voche_m - 1.0);
    const double c = 2.0 / (3.0 * np1.dg * np1.dg) * std::pow(base, p.voche_m);
    for (int k = 0; k < 6; ++k) ed[k] = ed[k] + (a * dydd[k] + b * dvdd[k] + c * d_mod[k]) * std::fabs(slipinc);
  }
  for (int k = 0; k < 6; ++k) ed[k] = p.theta_0 * mnp0 * ed[k] + mnp0 * dydd[k];
}

// mm10_tangent (mm10_a.f:658-815).  Voce: ed = 0, dgammadd = 0 => JA = JB = 0.  MTS: JA from
// dgammadd (mm10_b.f:2316-2345), JB = J12 J22^-1 ed.
static void mm10_tangent(const Props& p, State& np1, const double J[7][7]) {
  double JJ[36], JR[36];
  double JA[6][6], JB[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { JA[i][j] = 0.0; JB[i][j] = 0.0; }
  if (p.h_type == 2) {
    double ed[6], d_mod[6];
    ed_mts(p, np1, np1.stress, np1.tau_tilde, ed);
    for (int k = 0; k < 6; ++k) d_mod[k] = (k < 3) ? np1.D[k] : 0.5 * np1.D[k];
    const double alpha = 2.0 / (3.0 * np1.dg * np1.dg);
    for (int i = 0; i < p.nslip; ++i) {
      double symtq[6], wv[6], dgdd[6];
      // the reference passes `symtqmat`, not `symtqmat(1,i)`, to mm10_a_mult_type_4 (mm10_a.f:771-772): every slip system
      // gets column 1, i.e. sym(sigma W) of the FIRST system.  Reproduced (found by executing the reference's
      // mm10_tangent, tests/test_reference_vectors.py: with the per-system column the MTS tangent is 1e-4 off).
      symswmat_col(np1.stress, np1.qc[0], symtq);
      for (int k = 0; k < 6; ++k) wv[k] = 2.0 * symtq[k];                   // mm10_a_mult_type_4
      for (int j = 0; j < 6; ++j)
        for (int k = 0; k < 6; ++k) wv[k] = wv[k] + p.stiffness[k][j] * np1.ms[i][j];
      const double dgam = mm10_slipinc(p, np1, np1.stress, np1.tau_tilde, i);
      for (int k = 0; k < 6; ++k) dgdd[k] = alpha * dgam * d_mod[k];
      for (int a = 0; a < 6; ++a)                                           // mm10_a_mult_type_5
        for (int b = 0; b < 6; ++b) JA[a][b] += wv[a] * dgdd[b];
    }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) JB[a][b] = J[a][6] * (ed[b] / J[6][6]);   // J12 * (J22^-1 ed), len = 1
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      // beta = J22^-1 J21 (DGESV len=1), JJ = J11 - J12 * beta
      double beta = J[6][j] / J[6][6];
      JJ[i * 6 + j] = J[i][j] + (-1.0) * J[i][6] * beta;
      JR[i * 6 + j] = p.stiffness[i][j] - JA[i][j] - JB[i][j];
    }
  lu_solve(6, JJ, JR, 6);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) np1.tangent[i][j] = JR[i * 6 + j];
}

static double atan2_0_2pi(double a, double b) {  // mm10_a.f:1224-1233
  double v = std::atan2(a, b);
  if (v < 0.0) v = v + 2.0 * PI;
  return v;
}

// mm10_output (mm10_a.f:3429-3682), Voce branch, halite add-on off (cp_031 = 0)
static void mm10_output(const Props& p, State& np1, const State& /*n*/) {
  const int nslip = p.nslip;
  for (int i = 0; i < nslip; ++i) np1.slip_incs[i] = mm10_slipinc(p, np1, np1.stress, np1.tau_tilde, i);
  // mm10_update_euler_angles (mm10_a.f:1171-1220)
  {
    M33 work1, full;
    for (int i = 0; i < 3; ++i)  // mm10_a_mult_type_3t: a = b * c^T
      for (int j = 0; j < 3; ++j)
        work1[i][j] = np1.Rp[i][0] * np1.R[j][0] + np1.Rp[i][1] * np1.R[j][1] + np1.Rp[i][2] * np1.R[j][2];
    mat33(p.g, work1, full);
    double psiK = atan2_0_2pi(full[2][1], full[2][0]);
    double phiK = atan2_0_2pi(full[1][2], full[0][2]);
    if (full[2][2] > 1.0) full[2][2] = 1.0;
    double thetaK = std::acos(full[2][2]);
    np1.euler[0] = 180.0 / PI * psiK;
    np1.euler[1] = 180.0 / PI * thetaK;
    np1.euler[2] = 180.0 / PI * phiK;
  }
  double dif_slp[ORC_MAX_SLIP];
  for (int i = 0; i < nslip; ++i) {
    double rs = mm10_rs(np1, np1.stress, i);
    dif_slp[i] = rs * np1.tinc * p.iD_v;
  }
  double ed[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nslip; ++i)
    for (int k = 0; k < 6; ++k) ed[k] = ed[k] + dif_slp[i] * np1.ms[i][k];
  for (int k = 0; k < 6; ++k) np1.ed[k] = ed[k] / np1.tinc;
  double t1 = ed[0] * ed[0] + ed[1] * ed[1] + ed[2] * ed[2];
  double t2 = ed[3] * ed[3] + ed[4] * ed[4] + ed[5] * ed[5];
  double ed_dot = std::sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
  ed_dot = ed_dot / np1.tinc;
  np1.u[14] = ed_dot;
  np1.work_inc = np1.stress[0] * np1.D[0] + np1.stress[1] * np1.D[1] + np1.stress[2] * np1.D[2] +
                 np1.stress[3] * np1.D[3] + np1.stress[4] * np1.D[4] + np1.stress[5] * np1.D[5];
  // lattice strain: S^-1 stress (DPOSV), rotated by RT2RVE(R)
  double S[36], eeun[6], ee[6];
  for (int i = 0; i < 6; ++i) { eeun[i] = np1.stress[i]; for (int j = 0; j < 6; ++j) S[i * 6 + j] = p.stiffness[i][j]; }
  lu_solve(6, S, eeun, 1);
  M66 erot;
  rt2rve(np1.R, erot);
  matvec6(erot, eeun, ee);
  double dbarp[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nslip; ++i)
    for (int k = 0; k < 6; ++k) dbarp[k] = dbarp[k] + np1.slip_incs[i] * np1.ms[i][k];
  double wp[3], ewwe[6], ep[6];
  form_wp(p, np1, np1.stress, np1.tau_tilde, wp);
  symsw(ee, wp, ewwe);
  for (int k = 0; k < 6; ++k) { ep[k] = dbarp[k] + ewwe[k]; np1.ep[k] = ep[k] / np1.tinc; }
  t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
  t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
  double ep_dot = std::sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
  ep_dot = ep_dot / np1.tinc;
  np1.u[10] = ep_dot;
  for (int k = 0; k < 6; ++k) ep[k] = ep[k] + ed[k];
  t1 = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
  t2 = ep[3] * ep[3] + ep[4] * ep[4] + ep[5] * ep[5];
  np1.p_strain_inc = std::sqrt(2.0 / 3.0 * (t1 + 0.5 * t2));
  np1.p_work_inc = np1.stress[0] * ep[0] + np1.stress[1] * ep[1] + np1.stress[2] * ep[2] +
                   np1.stress[3] * ep[3] + np1.stress[4] * ep[4] + np1.stress[5] * ep[5];
  for (int k = 0; k < 6; ++k) np1.eps[k] = ee[k];
  double ec_dot = np1.p_strain_inc / np1.tinc, n_eff;
  if (ec_dot > 0.0) {
    for (int k = 0; k < 6; ++k) ep[k] = ep[k] / np1.tinc;
    double dgdt[ORC_MAX_SLIP];
    dgdt_voche(p, np1, np1.stress, np1.tau_tilde, dgdt);
    n_eff = 0.0;
    for (int i = 0; i < nslip; ++i) {
      double rs = mm10_rs(np1, np1.stress, i);
      double a1 = np1.ms[i][0] * ep[0] + np1.ms[i][1] * ep[1] + np1.ms[i][2] * ep[2];
      double a2 = np1.ms[i][3] * ep[3] + np1.ms[i][4] * ep[4] + np1.ms[i][5] * ep[5];
      double ec_slip = a1 + 0.5 * a2;
      double b1 = rs * dgdt[i] / ec_dot;
      double b2 = ec_slip / ec_dot / np1.tinc;
      n_eff = n_eff + (2.0 / 3.0) * b1 * b2;
    }
  } else n_eff = 1.0e10;
  np1.u[11] = n_eff;
  double s_trace = (np1.stress[0] + np1.stress[1] + np1.stress[2]) / 3.0;
  double se[6];
  for (int k = 0; k < 6; ++k) se[k] = np1.stress[k];
  se[0] -= s_trace; se[1] -= s_trace; se[2] -= s_trace;
  t1 = se[0] * se[0] + se[1] * se[1] + se[2] * se[2];
  t2 = se[3] * se[3] + se[4] * se[4] + se[5] * se[5];
  s_trace = std::sqrt(1.5 * (t1 + 2.0 * t2));
  np1.u[12] = s_trace;
  double B_eff;
  if (ec_dot < 1.e-100) B_eff = 0.0;
  else if (n_eff > 100.0) B_eff = -1.0;
  else B_eff = ec_dot / std::pow(s_trace, n_eff);
  np1.u[13] = B_eff;
  for (int i = 0; i < nslip; ++i) np1.slip_incs[i] = np1.slip_incs[i] + dif_slp[i];
  double maxslip = 0.0; int sysID = 0;
  for (int i = 0; i < nslip; ++i) {
    double cur = std::fabs(np1.slip_incs[i]);
    if (cur > maxslip) { maxslip = cur; sysID = i + 1; }
  }
  np1.u[5] = maxslip / np1.tinc;
  np1.u[6] = (double)sysID;
  int numAct = 0;
  maxslip = 0.1 * maxslip;
  for (int i = 0; i < nslip; ++i) if (std::fabs(np1.slip_incs[i]) >= maxslip) numAct++;
  np1.u[7] = (double)numAct;
}

// ----------------------------------------------------------------------------
// One crystal of one point: mm10_a_do_crystal (mm10_a.f:166-243) incl. mm10_solve_crystal
// (:1080-1157), mm10_solve_strup (:2628-2845) and mm10_store_cryhist (:976-1055).
// L carries the history offsets of THIS crystal (per-crystal terms already shifted).  The
// crystal's contributions to the point averages come back in `out`.  Returns 1 when the local
// solve failed (material_cut_step).
struct CrystalOut { double stress[6], tangent[6][6], slip_incs[ORC_MAX_SLIP], work_inc, p_work_inc, p_strain_inc; };
static int mm10_crystal(int step, int iter, const CrystalLib& cry, const double* angles, const HistLayout& L,
                        double dt, const double* rot9, const double* uddt, double* hn, double* h1,
                        const double* urcs_n, int* local_iters, CrystalOut& out) {
  const bool iter_0_extrapolate_off = (iter == 0);  // rstgp1.f:870-877
  static thread_local Props p;
  static thread_local State n, np1, curr;
  setup_props(cry, angles, p);
  const int nslip = p.nslip;
  std::memset(&out, 0, sizeof(out));
  if (step == 1) {  // mm10_init_cc_hist0 (mm10_a.f:220-227)
    for (int k = 0; k < 6; ++k) hn[L.c_stress + k] = urcs_n[k];
    for (int k = 0; k < 3; ++k) hn[L.c_euler + k] = angles[k];
    for (int k = 0; k < 9; ++k) hn[L.c_Rp + k] = (k % 4 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < 6; ++k) { hn[L.c_D + k] = 0.0; hn[L.c_eps + k] = 0.0; }
    for (int k = 0; k < L.len_slip; ++k) hn[L.c_slipinc + k] = 0.0;
    hn[L.c_tt] = p.tau_y + 1.0e-5;  // mm10_init_voche (mm10_a.f:2042-2055); other slots undefined there
    if (p.h_type == 2) { hn[L.c_tt] = -1.0; hn[L.c_u] = -1.0; hn[L.c_u + 1] = -1.0; }   // mm10_init_mts: flags (mm10_a.f:2090-2106)
  }
  // mm10_copy_cc_hist (mm10_a.f:2464-2561)
  std::memset(&n, 0, sizeof(State));
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) { n.R[i][j] = hn[L.R + 3 * j + i]; n.Rp[i][j] = hn[L.c_Rp + 3 * j + i]; }
  for (int k = 0; k < 6; ++k) { n.stress[k] = hn[L.c_stress + k]; n.D[k] = hn[L.c_D + k]; n.eps[k] = hn[L.c_eps + k]; }
  for (int k = 0; k < 3; ++k) n.euler[k] = hn[L.c_euler + k];
  n.tau_tilde = hn[L.c_tt];
  n.tt_rate = hn[L.c_ttrate];
  for (int k = 0; k < 14; ++k) n.u[k] = hn[L.c_u + k];      // n%u(1:len2-1) (mm10_a.f:2544)
  // mm10_setup_np1
  M33 R;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) R[i][j] = rot9[3 * j + i];
  setup_np1(R, uddt, dt, np1);

  // ---- mm10_solve_strup ----
  double stress[6], ostress[6], tt, ott;
  for (int k = 0; k < 6; ++k) { stress[k] = n.stress[k]; ostress[k] = stress[k]; }
  tt = n.tau_tilde; ott = tt;
  mm10_setup(p, np1, n);
  bool fail = false;
  double temp1 = 0, temp2 = 0;
  for (int k = 0; k < 6; ++k) { temp1 += np1.D[k] * np1.D[k]; temp2 += stress[k] * stress[k]; }
  bool no_load = (temp2 == 0.0) && (temp1 == 0.0);
  double Jmat[7][7];
  double curr_tt_rate = 0.0;
  if (iter_0_extrapolate_off || no_load) {
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) np1.tangent[i][j] = p.stiffness[i][j];
    if (!no_load) {
      double R1[6];
      formR1(p, np1, n, stress, tt, R1);
      for (int k = 0; k < 6; ++k) stress[k] = stress[k] - R1[k];
    }
    curr_tt_rate = 0.0;
  } else {
    double frac = 0.0, stp = 1.0; int cuts = 0; const double mult = 0.5; const int mcuts = 4;
    while (frac < 1.0) {
      double Dw[6];
      for (int k = 0; k < 6; ++k) Dw[k] = np1.D[k] * (stp + frac);
      double tinc_work = np1.tinc * (stp + frac);
      setup_np1(np1.R, Dw, tinc_work, curr);
      curr.temp = (np1.temp - 0.0) * (stp + frac) + 0.0;  // n%temp = 0 (mm10_a.f:2486, :2769); unused by Voce
      mm10_setup(p, curr, n);
      tt = n.tau_tilde;
      fail = mm10_solve(p, curr, n, stress, &tt, np1.tinc * stp, Jmat, local_iters);
      if (fail) {
        for (int k = 0; k < 6; ++k) stress[k] = ostress[k];
        tt = ott;
        stp = stp * mult; cuts = cuts + 1;
        if (cuts > mcuts) break;
        fail = false;
      } else {
        for (int k = 0; k < 6; ++k) ostress[k] = stress[k];
        ott = tt;
        frac = frac + stp;
      }
    }
    bool anynan = std::isnan(tt);
    for (int k = 0; k < 6; ++k) anynan = anynan || std::isnan(stress[k]);
    if (fail || anynan) {
      // material_cut_step.  The reference prints a warning, sets np1 stress / tau_tilde back to
      // the n state (mm10_a.f:2838-2841), returns from mm10 and leaves the REST OF THE BLOCK
      // un-updated (mm10_a.f:125-127) -- undefined data.  Defined behaviour of this project
      // (oracle and GPU alike): the crystal keeps its n state (stress, tau_tilde, Rp, Euler
      // angles, lattice strain), no slip, elastic tangent; the other crystals of the point and
      // the sweep go on and the failure is counted.
      for (int k = 0; k < 6; ++k) { h1[L.c_stress + k] = n.stress[k]; out.stress[k] = n.stress[k]; }
      for (int k = 0; k < 3; ++k) h1[L.c_euler + k] = n.euler[k];
      for (int k = 0; k < 9; ++k) h1[L.c_Rp + k] = hn[L.c_Rp + k];
      for (int k = 0; k < 6; ++k) { h1[L.c_D + k] = np1.D[k]; h1[L.c_eps + k] = n.eps[k]; h1[L.c_ep + k] = 0.0; h1[L.c_ed + k] = 0.0; }
      for (int k = 0; k < L.len_slip; ++k) h1[L.c_slipinc + k] = 0.0;
      h1[L.c_tt] = n.tau_tilde; h1[L.c_ttrate] = 0.0;
      for (int k = 0; k < 15; ++k) h1[L.c_u + k] = 0.0;
      if (p.h_type == 2) { h1[L.c_u] = n.u[0]; h1[L.c_u + 1] = n.u[1]; }    // MTS: tau_y / mu_harden of the n state (or their flags)
      for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) out.tangent[i][j] = p.stiffness[i][j];
      return 1;
    }
    curr_tt_rate = curr.tt_rate;
  }
  for (int k = 0; k < 6; ++k) np1.stress[k] = stress[k];
  np1.tau_tilde = tt;
  np1.tt_rate = curr_tt_rate;

  // ---- rest of mm10_solve_crystal ----
  if (!(iter_0_extrapolate_off || no_load)) {
    mm10_tangent(p, np1, Jmat);
    M66 w;  // mm10_a_make_symm_1
    for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) w[i][j] = (np1.tangent[i][j] + np1.tangent[j][i]) * 0.5;
    for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) np1.tangent[i][j] = w[i][j];
    // mm10_update_rotation (mm10_a.f:3310-3414)
    double wbarp[3]; M33 W, expw;
    form_wbarp(p, np1, np1.stress, np1.tau_tilde, wbarp);
    W[0][0] = 0; W[0][1] = wbarp[2]; W[0][2] = wbarp[1];
    W[1][0] = -wbarp[2]; W[1][1] = 0; W[1][2] = wbarp[0];
    W[2][0] = -wbarp[1]; W[2][1] = -wbarp[0]; W[2][2] = 0;
    double alpha = std::sqrt(W[1][2] * W[1][2] + W[0][2] * W[0][2] + W[0][1] * W[0][1]);
    if (alpha < 1.0e-16) {
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) expw[i][j] = 0.0;
    } else {
      M33 W2; mat33(W, W, W2);
      double ca = (1.0 - std::cos(alpha)) / (alpha * alpha), cb = std::sin(alpha) / alpha;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) expw[i][j] = ca * W2[i][j] + cb * W[i][j];
    }
    expw[0][0] += 1.0; expw[1][1] += 1.0; expw[2][2] += 1.0;
    mat33(expw, n.Rp, np1.Rp);
    mm10_output(p, np1, n);
  }
  // ---- mm10_store_cryhist (mm10_a.f:976-1055) ----
  for (int k = 0; k < 6; ++k) h1[L.c_stress + k] = np1.stress[k];
  for (int k = 0; k < 3; ++k) h1[L.c_euler + k] = np1.euler[k];
  for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) h1[L.c_Rp + 3 * j + i] = np1.Rp[i][j];
  for (int k = 0; k < 6; ++k) { h1[L.c_D + k] = np1.D[k]; h1[L.c_eps + k] = np1.eps[k]; }
  for (int k = 0; k < L.len_slip; ++k) h1[L.c_slipinc + k] = (k < ORC_MAX_SLIP) ? np1.slip_incs[k] : 0.0;
  h1[L.c_tt] = np1.tau_tilde;
  for (int k = 0; k < 15; ++k) h1[L.c_u + k] = np1.u[k];
  h1[L.c_ttrate] = np1.tt_rate;
  for (int k = 0; k < 6; ++k) { h1[L.c_ep + k] = np1.ep[k]; h1[L.c_ed + k] = np1.ed[k]; }
  // contributions to the point averages (mm10_a.f:228-238)
  for (int k = 0; k < 6; ++k) out.stress[k] = np1.stress[k];
  for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) out.tangent[i][j] = np1.tangent[i][j];
  for (int k = 0; k < nslip; ++k) out.slip_incs[k] = np1.slip_incs[k];
  out.work_inc = np1.work_inc; out.p_work_inc = np1.p_work_inc; out.p_strain_inc = np1.p_strain_inc;
  return 0;
}

// mm10 driver for one point (mm10_a.f:29-355): loop over the crystals of the point, Taylor
// average of stress, tangent, slip and work increments (mm10_a_crystal_avgs :139-164), store
// (mm10_a_store_crystal :285-318).  crystals[c] / angles[3 c ..] describe crystal c.
int mm10_point(int step, int iter, int ncrystals, const CrystalLib* const* crystals, const double* angles,
               const HistLayout& L, double dt, const double* rot9, const double* uddt, double* hn, double* h1,
               const double* urcs_n, double* urcs_n1, int* local_iters) {
  local_iters[0] = local_iters[1] = 0;
  if (step == 1) {  // mm10_init_general_hist / uout / slip (mm10_a.f:92-97)
    for (int k = 0; k < 36; ++k) hn[L.cep + k] = 0.0;
    for (int k = 0; k < 27; ++k) hn[L.gradfe + k] = 0.0;
    for (int k = 0; k < 9; ++k) hn[L.R + k] = (k % 4 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < 3; ++k) hn[L.work + k] = 0.0;
    for (int k = 0; k < L.len_slip; ++k) hn[L.slipsum + k] = 0.0;
  }
  double sig_avg[6] = {0, 0, 0, 0, 0, 0}, tang_avg[6][6], slip_avg[ORC_MAX_SLIP];
  double t_work_inc = 0.0, p_work_inc = 0.0, p_strain_inc = 0.0;
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) tang_avg[i][j] = 0.0;
  for (int k = 0; k < ORC_MAX_SLIP; ++k) slip_avg[k] = 0.0;
  const int per = L.total - L.c_stress;    // one_crystal_hist_size
  int failed = 0;
  static thread_local CrystalOut out;
  for (int c = 0; c < ncrystals; ++c) {
    HistLayout Lc = L;
    const int co = c * per;
    Lc.c_stress += co; Lc.c_euler += co; Lc.c_Rp += co; Lc.c_D += co; Lc.c_eps += co; Lc.c_slipinc += co;
    Lc.c_tt += co; Lc.c_u += co; Lc.c_ttrate += co; Lc.c_ep += co; Lc.c_ed += co;
    int li[2] = {0, 0};
    failed += mm10_crystal(step, iter, *crystals[c], angles + 3 * c, Lc, dt, rot9, uddt, hn, h1, urcs_n, li, out);
    local_iters[0] += li[0]; local_iters[1] += li[1];
    for (int k = 0; k < 6; ++k) sig_avg[k] = sig_avg[k] + out.stress[k];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) tang_avg[i][j] = tang_avg[i][j] + out.tangent[i][j];
    for (int k = 0; k < L.len_slip && k < ORC_MAX_SLIP; ++k) slip_avg[k] = slip_avg[k] + out.slip_incs[k];
    t_work_inc = t_work_inc + out.work_inc;
    p_work_inc = p_work_inc + out.p_work_inc;
    p_strain_inc = p_strain_inc + out.p_strain_inc;
  }
  const double rncry = (double)ncrystals;   // mm10_a_crystal_avgs
  for (int k = 0; k < 6; ++k) sig_avg[k] = sig_avg[k] / rncry;
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) tang_avg[i][j] = tang_avg[i][j] / rncry;
  for (int k = 0; k < ORC_MAX_SLIP; ++k) slip_avg[k] = slip_avg[k] / rncry;
  t_work_inc = t_work_inc / rncry; p_work_inc = p_work_inc / rncry; p_strain_inc = p_strain_inc / rncry;
  // mm10_a_store_crystal
  for (int k = 0; k < 6; ++k) urcs_n1[k] = sig_avg[k];
  for (int k = 0; k < 9; ++k) h1[L.R + k] = rot9[k];
  for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) h1[L.cep + 6 * j + i] = tang_avg[i][j];
  for (int k = 0; k < L.len_slip; ++k) h1[L.slipsum + k] = hn[L.slipsum + k] + ((k < ORC_MAX_SLIP) ? slip_avg[k] : 0.0);
  urcs_n1[6] = urcs_n[6] + t_work_inc;
  urcs_n1[7] = urcs_n[7] + p_work_inc;
  urcs_n1[8] = urcs_n[8] + p_strain_inc;
  h1[L.work + 0] = hn[L.work + 0] + t_work_inc;
  h1[L.work + 1] = hn[L.work + 1] + p_work_inc;
  h1[L.work + 2] = hn[L.work + 2] + p_strain_inc;
  return failed ? 1 : 0;
}

} // namespace orc

// unit probe for the pinning tests: residual R(x) (7) and Jacobian J(x) (7x7 row-major) of the
// local Newton system of one crystal (mm10_formR / mm10_formJ, mm10_b.f:1029-1051, 901-954) for a
// strain increment D6 over dt with R = I, Rp_n = I, n state (stress, tau_tilde); the MTS n-state
// values tau_y, mu_harden are taken from the step itself (history flags < 0)
extern "C" void orc_mm10_residual_jacobian(const orc_crystal* c, const double* angles, const double* D6, double dt,
                                           const double* x7, const double* n_stress6, double n_tt, double* R7, double* J49) {
  using namespace orc;
  CrystalLib L; L.in = *c; finalize_crystal(L);
  static thread_local Props p; static thread_local State n, np1;
  setup_props(L, angles, p);
  std::memset(&n, 0, sizeof(State));
  M33 I = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { n.R[i][j] = I[i][j]; n.Rp[i][j] = I[i][j]; }
  for (int k = 0; k < 6; ++k) n.stress[k] = n_stress6[k];
  n.tau_tilde = n_tt; n.u[0] = -1.0; n.u[1] = -1.0;
  setup_np1(I, D6, dt, np1);
  mm10_setup(p, np1, n);
  formR(p, np1, n, x7, R7);
  double J[7][7];
  formJ(p, np1, x7, J);
  for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) J49[7 * i + j] = J[i][j];
}

// the same probe with a plastic rotation Rp_n and a polar rotation R of the step (row-major 3x3): exercises the
// rotation operators of mm10_setup (mm10_a.f:830-962); also returns the current Schmid vectors (tests/test_reference_vectors.py)
extern "C" void orc_mm10_residual_jacobian_rot(const orc_crystal* c, const double* angles, const double* D6, double dt,
                                               const double* x7, const double* n_stress6, double n_tt, const double* Rp9,
                                               const double* R9, double* R7, double* J49, double* ms_out, double* qs_out, double* qc_out) {
  using namespace orc;
  CrystalLib L; L.in = *c; finalize_crystal(L);
  static thread_local Props p; static thread_local State n, np1;
  setup_props(L, angles, p);
  std::memset(&n, 0, sizeof(State));
  M33 I = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, Rm;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { n.R[i][j] = I[i][j]; n.Rp[i][j] = Rp9[3 * i + j]; Rm[i][j] = R9[3 * i + j]; }
  for (int k = 0; k < 6; ++k) n.stress[k] = n_stress6[k];
  n.tau_tilde = n_tt; n.u[0] = -1.0; n.u[1] = -1.0;
  setup_np1(Rm, D6, dt, np1);
  mm10_setup(p, np1, n);
  formR(p, np1, n, x7, R7);
  double J[7][7];
  formJ(p, np1, x7, J);
  for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) J49[7 * i + j] = J[i][j];
  for (int s = 0; s < p.nslip; ++s) {
    for (int k = 0; k < 6; ++k) ms_out[6 * s + k] = np1.ms[s][k];
    for (int k = 0; k < 3; ++k) { qs_out[3 * s + k] = np1.qs[s][k]; qc_out[3 * s + k] = np1.qc[s][k]; }
  }
}

// the whole update of one crystal (mm10_solve_crystal + history store, mm10_a.f:1080-1157, 976-1055) from an explicit n
// state, for tests/test_reference_vectors.py.  3x3 matrices row-major.  n_state: stress[6], tau_tilde, tt_rate, D[6],
// eps[6], euler[3], Rp[9], R[9] (41 doubles).  out: stress[6], tau_tilde, tt_rate, tangent[36] (row-major), Rp[9],
// euler[3], eps[6], slip_incs[48], u[15], ep[6], ed[6] (137 doubles); iters[2] = predictor / update Newton iterations;
// returns the failure flag.
extern "C" int orc_mm10_crystal_probe(const orc_crystal* c, const double* angles, double dt, const double* R9, const double* D6,
                                      int iter, const double* n_state, double* out137, int* iters) {
  using namespace orc;
  CrystalLib L; L.in = *c; finalize_crystal(L);
  const HistLayout H = mm10_history_layout(L.nslip, 1);
  std::vector<double> hn(H.total + 64, 0.0), h1(H.total + 64, 0.0);
  const double* ns = n_state;
  for (int k = 0; k < 6; ++k) { hn[H.c_stress + k] = ns[k]; hn[H.c_D + k] = ns[8 + k]; hn[H.c_eps + k] = ns[14 + k]; }
  hn[H.c_tt] = ns[6]; hn[H.c_ttrate] = ns[7];
  for (int k = 0; k < 3; ++k) hn[H.c_euler + k] = ns[20 + k];
  for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) { hn[H.c_Rp + 3 * j + i] = ns[23 + 3 * i + j]; hn[H.R + 3 * j + i] = ns[32 + 3 * i + j]; }
  hn[H.c_u] = -1.0; hn[H.c_u + 1] = -1.0;
  double rot9[9], urcs_n[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) rot9[3 * j + i] = R9[3 * i + j];
  static thread_local CrystalOut co;
  iters[0] = iters[1] = 0;
  const int fail = mm10_crystal(2, iter, L, angles, H, dt, rot9, D6, hn.data(), h1.data(), urcs_n, iters, co);
  double* o = out137;
  for (int k = 0; k < 6; ++k) o[k] = h1[H.c_stress + k];
  o[6] = h1[H.c_tt]; o[7] = h1[H.c_ttrate];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) o[8 + 6 * i + j] = co.tangent[i][j];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[44 + 3 * i + j] = h1[H.c_Rp + 3 * j + i];
  for (int k = 0; k < 3; ++k) o[53 + k] = h1[H.c_euler + k];
  for (int k = 0; k < 6; ++k) o[56 + k] = h1[H.c_eps + k];
  for (int k = 0; k < 48; ++k) o[62 + k] = (k < H.len_slip) ? h1[H.c_slipinc + k] : 0.0;
  for (int k = 0; k < 15; ++k) o[110 + k] = h1[H.c_u + k];
  for (int k = 0; k < 6; ++k) { o[125 + k] = h1[H.c_ep + k]; o[131 + k] = h1[H.c_ed + k]; }
  return fail;
}

extern "C" void orc_crystal_stiffness(const orc_crystal* c, double* C36) {
  orc::CrystalLib L; L.in = *c; orc::finalize_crystal(L);
  for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) C36[6 * j + i] = L.elast_stiff[i][j];
}
extern "C" void orc_slip_table(int slip_type, int* nslip, double* b, double* n) {
  orc::CrystalLib L; std::memset(&L, 0, sizeof(L)); L.in.slip_type = slip_type; L.in.e = 1; L.in.nu = 0.3; L.in.mu = 1; L.in.elastic_type = 1;
  orc::finalize_crystal(L);
  *nslip = L.nslip;
  for (int s = 0; s < L.nslip; ++s) for (int k = 0; k < 3; ++k) { b[3 * s + k] = L.bi[s][k]; n[3 * s + k] = L.ni[s][k]; }
}
