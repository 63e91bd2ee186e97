// cpfft_b200: host-side construction of the material tables the update kernels read:
// per-material and per-crystal constants, voxel -> material / grain indices and the per-grain
// table (setup_mm10_rknstr, drive_eps_sig.f:571-606, 975-986; mm10_rotation_matrix
// mm10_a.f:1287-1345; mm10_RT2RVE mm10_a.f:1400-1447; crystal stiffness finalize_new_crystal
// mod_crystals.f:1793-1931).  Pure C++ (no CUDA): material.cu uploads the result, the host
// build of the per-voxel code (tests/native/material_host.cpp) uses it directly.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/cpfft_b200.h"
#include "material_types.h"
#include "slip_tables.cuh"

struct CpfMatTables {
  std::vector<CpfMatDev> md;
  std::vector<CpfCryDev> cd;
  std::vector<int32_t> midx;         // per voxel: 0-based material
  std::vector<int32_t> gidx;         // (ncmax, n3): grain-table entry of crystal ci of voxel e at [ci * n3 + e]
  std::vector<int32_t> gcry;         // per grain-table entry: 0-based crystal library index
  int ncmax = 1;                     // crystals per material point the tables are sized for
  bool has_taylor = false;           // some cp material has n_crystals > 1
  bool kern[2][3] = {{false, false, false}, {false, false, false}};   // [n_crystals > 1][hardening law]: kernels to launch
  std::vector<double> gtab;          // ngrains x CPF_GRAIN_STRIDE
  int ngrains = 0, nslip_max = 0;
  bool has_mm01 = false, has_mm10 = false;
  CpfHistLayout L;
  int H = 11;                        // history components per voxel
};

static void host_rt2rve(const double rt[3][3], double rv[6][6]) {
  const int a[6] = {0, 1, 2, 0, 1, 0}, b[6] = {0, 1, 2, 1, 2, 2};
  // strain-type (engineering shear) rotation operator of the tensor map E -> rt E rt^T
  for (int I = 0; I < 6; ++I)
    for (int Jc = 0; Jc < 6; ++Jc) {
      int i = a[I], j = b[I], k = a[Jc], l = b[Jc];
      double v;
      if (Jc < 3) v = rt[i][k] * rt[j][k];
      else v = rt[i][k] * rt[j][l] + rt[i][l] * rt[j][k];
      if (I < 3 && Jc >= 3) v = 2.0 * rt[i][k] * rt[i][l];
      rv[I][Jc] = v;
    }
}
static void host_inv6(const double in[6][6], double out[6][6]) {
  double A[6][12];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { A[i][j] = in[i][j]; A[i][6 + j] = (i == j); }
  for (int k = 0; k < 6; ++k) {
    int p = k;
    for (int i = k + 1; i < 6; ++i) if (std::fabs(A[i][k]) > std::fabs(A[p][k])) p = i;
    if (p != k) for (int j = 0; j < 12; ++j) std::swap(A[k][j], A[p][j]);
    double inv = 1.0 / A[k][k];
    for (int j = 0; j < 12; ++j) A[k][j] *= inv;
    for (int i = 0; i < 6; ++i) if (i != k) { double l = A[i][k]; for (int j = 0; j < 12; ++j) A[i][j] -= l * A[k][j]; }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[i][j] = A[i][6 + j];
}

// returns 0, or CPFFT_ERR_USAGE with the reason in `err`
// angles: (n3, ncmax, 3) Kocks degrees; crystal_ids: (n3, ncmax) 1-based crystal numbers or
// nullptr = the material's own crystal (crystal_input single); a voxel uses the first
// n_crystals entries of its material (setup_mm10_rknstr case 2, drive_eps_sig.f:744-777).
static int cpf_build_material_tables(const std::vector<cpfft_material>& mats, const std::vector<cpfft_crystal>& crys,
                                     const int32_t* matlist, int ncmax, const double* angles,
                                     const int32_t* crystal_ids, int64_t n3,
                                     CpfMatTables& T, std::string& err) {
    const int nmat = (int)mats.size(), ncry = (int)crys.size();
  std::vector<CpfMatDev>& md = T.md; md.assign(nmat, CpfMatDev());
  T.has_mm01 = T.has_mm10 = false;
  int nslip_max = 0;
  for (int i = 0; i < nmat; ++i) {
    const cpfft_material& m = mats[i];
    md[i].type = m.type; md[i].crystal = m.crystal - 1;
    md[i].ncry = (m.type == 10 && m.n_crystals > 1) ? m.n_crystals : 1; md[i].hard = 0;   // hard: from the crystals in use, below
    if (md[i].ncry > ncmax) { err = "n_crystals of a material exceeds the crystals per voxel of cpfft_set_voxels_taylor"; return CPFFT_ERR_USAGE; }
    if (md[i].ncry > 1) T.has_taylor = true;
    md[i].ym = (double)m.e; md[i].nu = (double)m.nu; md[i].beta = (double)m.beta;
    md[i].tan_e = (double)m.tan_e; md[i].yld = (double)m.yld_pt;
    md[i].hprime = (m.type == 1) ? md[i].tan_e * md[i].ym / (md[i].ym - md[i].tan_e) : 0.0;
    if (m.type == 1) T.has_mm01 = true;
    else if (m.type == 10) {
      T.has_mm10 = true;
      if (!crystal_ids && (m.crystal < 1 || m.crystal > ncry)) { err = "material refers to an undefined crystal"; return CPFFT_ERR_USAGE; }
    } else { err = "unsupported material type (1 = bilinear, 10 = cp)"; return CPFFT_ERR_USAGE; }
  }
  std::vector<CpfCryDev>& cd = T.cd; cd.assign(std::max(1, ncry), CpfCryDev());
  std::vector<std::array<double, 36>> stiff(std::max(1, ncry));
  std::vector<std::vector<double>> bi(std::max(1, ncry)), ni(std::max(1, ncry));
  for (int i = 0; i < ncry; ++i) {
    const cpfft_crystal& c = crys[i];
    CpfCryDev& d = cd[i];
    if (c.h_type != 1 && c.h_type != 2) { err = "unsupported hardening law (h_type 1 = voce, 2 = mts)"; return CPFFT_ERR_USAGE; }
    const signed char (*tb)[3]; const signed char (*tn)[3];
    if (c.slip_type == 1) { d.nslip = 12; tb = CPF_FCC_B; tn = CPF_FCC_N; }
    else if (c.slip_type == 2) { d.nslip = 12; tb = CPF_BCC_B; tn = CPF_BCC_N; }
    else if (c.slip_type == 3) { d.nslip = 1; tb = CPF_SINGLE_B; tn = CPF_SINGLE_N; }
    else if (c.slip_type == 6) { d.nslip = 12; tb = CPF_ROTERS_B; tn = CPF_ROTERS_N; }
    else if (c.slip_type == 7) { d.nslip = 12; tb = CPF_BCC12_B; tn = CPF_BCC12_N; }
    else if (c.slip_type == 8) { d.nslip = 48; tb = CPF_BCC48_B; tn = CPF_BCC48_N; }
    else { err = "unsupported slip_type (1 fcc, 2 bcc, 3 single, 6 roters, 7 bcc12, 8 bcc48; hcp needs the ti6242 elasticity)"; return CPFFT_ERR_USAGE; }
    bi[i].resize(3 * d.nslip); ni[i].resize(3 * d.nslip);
    for (int s = 0; s < d.nslip; ++s) {
      double sb = 0, sn = 0;
      for (int k = 0; k < 3; ++k) { sb += tb[s][k] * tb[s][k]; sn += tn[s][k] * tn[s][k]; }
      for (int k = 0; k < 3; ++k) { bi[i][3 * s + k] = (double)tb[s][k] / std::sqrt(sb); ni[i][3 * s + k] = (double)tn[s][k] / std::sqrt(sn); }
    }
    d.alter_mode = c.alter_mode; d.miter = c.miter;
    d.rate_n = c.harden_n; d.theta_0 = c.theta_0; d.tau_y = c.tau_y; d.tau_v = c.tau_v; d.voche_m = c.voche_m;
    d.iD_v = c.iD_v; d.eps_dot_0_y = c.eps_dot_0_y; d.k_0 = c.k_0; d.burgers = c.burgers;
    d.atol = c.atol; d.atol1 = c.atol1; d.rtol = c.rtol; d.rtol1 = c.rtol1;
    d.hard = c.h_type; d.pad_ = 0;
    {  // MTS constants (mm10_setup_mts, mm10_a.f:2109-2175)
      d.tau_a = c.tau_a; d.tau_hat_y = c.tau_hat_y; d.tau_hat_v = c.tau_hat_v; d.eps_dot_0_v = c.eps_dot_0_v;
      d.mu_0 = c.mu_0; d.D_0 = c.D_0; d.T_0 = c.T_0;
      const double b3 = std::pow(c.burgers, 3.0);
      d.kby = c.boltzman / (b3 * c.g_0_y);
      d.kbv = c.boltzman / (b3 * c.g_0_v);
      d.p_y = c.p_y; d.q_y = c.q_y; d.p_v = c.p_v; d.q_v = c.q_v;
      d.iq_y = 1.0 / c.q_y; d.ip_y = 1.0 / c.p_y; d.iq_v = 1.0 / c.q_v; d.ip_v = 1.0 / c.p_v;
      if (c.h_type == 2) d.iD_v = 0.0;      // the diffusion slip exists in the Voce branches only
    }
    const double em1 = c.harden_n - 1.0;
    d.rate_int = (em1 >= 0.0 && em1 <= 64.0 && em1 == std::floor(em1)) ? (int)em1 : -1;
    double flex[6][6] = {{0}}, st[6][6];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) flex[a][b] = (a == b) ? 1 / c.e : -c.nu / c.e;
    const double sh = (c.elastic_type == 1) ? 2 * (1 + c.nu) / c.e : 1 / c.mu;
    flex[3][3] = flex[4][4] = flex[5][5] = sh;
    host_inv6(flex, st);
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) stiff[i][6 * a + b] = 0.5 * (st[a][b] + st[b][a]);
  }
  // voxel -> material index, grain dedup
  std::vector<int32_t>& midx = T.midx; std::vector<int32_t>& gidx = T.gidx;
  midx.assign(n3, 0); gidx.assign((size_t)n3 * ncmax, 0);
  T.ncmax = ncmax; T.gcry.clear();
  std::map<std::array<double, 4>, int> gmap;
  std::vector<double>& gtab = T.gtab; gtab.clear();
  const double PI = 3.141592653589793;
  for (int64_t e = 0; e < n3; ++e) {
    const int m = matlist[e] - 1;
    if (m < 0 || m >= nmat) { err = "matlist entry out of range"; return CPFFT_ERR_USAGE; }
    midx[e] = m;
    if (mats[m].type != 10) continue;
    for (int cc = 0; cc < md[m].ncry; ++cc) {
    const int ci = (crystal_ids ? crystal_ids[(size_t)e * ncmax + cc] : mats[m].crystal) - 1;
    if (ci < 0 || ci >= ncry) { err = "voxel refers to an undefined crystal"; return CPFFT_ERR_USAGE; }
    nslip_max = std::max(nslip_max, cd[ci].nslip);       // over the crystals in use (mm10_d.f:85-110)
    if (md[m].hard == 0) md[m].hard = cd[ci].hard;
    else if (md[m].hard != cd[ci].hard) { err = "the crystals of one cp material must share one hardening law"; return CPFFT_ERR_USAGE; }
    const double* av = angles + ((size_t)e * ncmax + cc) * 3;
    std::array<double, 4> key = {(double)ci, av[0], av[1], av[2]};
    auto it = gmap.find(key);
    if (it != gmap.end()) { gidx[(size_t)cc * n3 + e] = it->second; continue; }
    const int g = (int)gmap.size();
    gmap[key] = g; gidx[(size_t)cc * n3 + e] = g;
    T.gcry.push_back(ci);
    gtab.resize((size_t)(g + 1) * CPF_GRAIN_STRIDE, 0.0);
    double* t = &gtab[(size_t)g * CPF_GRAIN_STRIDE];
    const double psi = key[1] * PI / 180.0, th = key[2] * PI / 180.0, phi = key[3] * PI / 180.0;
    double r[3][3];
    r[0][0] = -std::sin(psi) * std::sin(phi) - std::cos(psi) * std::cos(phi) * std::cos(th);
    r[0][1] = std::cos(psi) * std::sin(phi) - std::sin(psi) * std::cos(phi) * std::cos(th);
    r[0][2] = std::cos(phi) * std::sin(th);
    r[1][0] = std::sin(psi) * std::cos(phi) - std::cos(psi) * std::sin(phi) * std::cos(th);
    r[1][1] = -std::cos(psi) * std::cos(phi) - std::sin(psi) * std::sin(phi) * std::cos(th);
    r[1][2] = std::sin(phi) * std::sin(th);
    r[2][0] = std::cos(psi) * std::sin(th);
    r[2][1] = std::sin(psi) * std::sin(th);
    r[2][2] = std::cos(th);
    double tr[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { t[CPF_GRAIN_G + 3 * i + j] = r[i][j]; tr[i][j] = r[j][i]; }
    for (int k = 0; k < 3; ++k) t[CPF_GRAIN_ANG + k] = key[1 + k];
    double Rs[6][6], tmp[6][6];
    host_rt2rve(tr, Rs);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += stiff[ci][6 * i + k] * Rs[j][k]; tmp[i][j] = s; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += Rs[i][k] * tmp[k][j]; t[CPF_GRAIN_C + 6 * i + j] = s; }
    const int ns = cd[ci].nslip;
    for (int s = 0; s < ns; ++s) {
      double vb[3], vn[3], A[3][3];
      for (int i = 0; i < 3; ++i) {
        vb[i] = tr[i][0] * bi[ci][3 * s] + tr[i][1] * bi[ci][3 * s + 1] + tr[i][2] * bi[ci][3 * s + 2];
        vn[i] = tr[i][0] * ni[ci][3 * s] + tr[i][1] * ni[ci][3 * s + 1] + tr[i][2] * ni[ci][3 * s + 2];
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = vb[i] * vn[j];
      double* o = t + CPF_GRAIN_B + CPF_SLIP_STRIDE * s;
      o[0] = 0.5 * (A[0][0] + A[0][0]); o[1] = 0.5 * (A[1][1] + A[1][1]); o[2] = 0.5 * (A[2][2] + A[2][2]);
      o[3] = 2.0 * (0.5 * (A[0][1] + A[1][0])); o[4] = 2.0 * (0.5 * (A[1][2] + A[2][1])); o[5] = 2.0 * (0.5 * (A[0][2] + A[2][0]));
      o[6] = 0.5 * (A[1][2] - A[2][1]); o[7] = 0.5 * (A[0][2] - A[2][0]); o[8] = 0.5 * (A[0][1] - A[1][0]);
    }
    }   // crystals of the voxel
  }
  T.ngrains = (int)gmap.size();
  if (gtab.empty()) { gtab.assign(CPF_GRAIN_STRIDE, 0.0); T.gcry.assign(1, 0); }
  for (int i = 0; i < nmat; ++i)
    if (md[i].type == 10 && md[i].hard >= 1 && md[i].hard <= 2) T.kern[md[i].ncry > 1 ? 1 : 0][md[i].hard] = true;
  T.nslip_max = nslip_max;
  T.H = 11;
  if (T.has_mm10) {
    T.L = cpf_hist_layout(nslip_max, 1);
    int need = 1;
    for (int i = 0; i < nmat; ++i) if (md[i].type == 10) need = std::max(need, md[i].ncry);
    // common block + n_crystals per-crystal blocks (mm10_set_sizes_special, mm10_a.f:640-641)
    T.H = std::max(T.H, T.L.c_stress + need * (T.L.total - T.L.c_stress));
  }
  else std::memset(&T.L, 0, sizeof(T.L));
  return 0;
}
