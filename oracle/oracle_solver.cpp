// CPFFT ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Model container, drive_eps_sig sweep, spectral operator G_K_dF, conjugate gradients and the
// FFT_nr3 step loop, restated from drive_eps_sig.f, G_K_dF.f, FFT_init.f, FFT_nr3.f,
// tangent_homo.f, rplstr.f and update.f.
#include "oracle_internal.hpp"
#include <complex>
#include <chrono>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;
typedef std::complex<double> cplx;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ----------------------------------------------------------------------------
// 1-D complex DFT of arbitrary length, stand-in for MKL DFTI (G_K_dF.f:130-154): Stockham
// autosort, one pass per prime factor (4 taken as one radix), explicit butterflies for radix
// 2, 3, 4, 5 and an O(R^2) butterfly for any other prime; twiddles precomputed per stage.
// exponent sign = dir (-1 forward, +1 backward), unscaled.
struct Fft1d {
  int n;
  std::vector<int> radices;
  std::vector<std::vector<cplx>> tw;     // stage s: tw[s][k * R + r] = exp(-2 pi i k r / (Ns R)), k < Ns
  std::vector<std::vector<cplx>> twc;    // conjugates (backward transform)
  std::vector<std::vector<cplx>> roots;  // generic radix R: exp(-2 pi i q / R), q < R
  void init(int n_) {
    n = n_; radices.clear(); tw.clear(); twc.clear(); roots.clear();
    int m = n;
    while (m % 4 == 0) { radices.push_back(4); m /= 4; }
    for (int p = 2; p * p <= m;) { if (m % p == 0) { radices.push_back(p); m /= p; } else ++p; }
    if (m > 1) radices.push_back(m);
    int Ns = 1;
    for (int R : radices) {
      std::vector<cplx> t((size_t)Ns * R), rt(R);
      for (int k = 0; k < Ns; ++k)
        for (int r = 0; r < R; ++r) {
          double a = -2.0 * M_PI * (double)k * (double)r / ((double)Ns * (double)R);
          t[(size_t)k * R + r] = cplx(std::cos(a), std::sin(a));
        }
      for (int q = 0; q < R; ++q) { double a = -2.0 * M_PI * (double)q / (double)R; rt[q] = cplx(std::cos(a), std::sin(a)); }
      std::vector<cplx> tc(t.size());
      for (size_t i = 0; i < t.size(); ++i) tc[i] = std::conj(t[i]);
      tw.push_back(t); twc.push_back(tc); roots.push_back(rt);
      Ns *= R;
    }
  }
  static inline cplx mul_i(cplx a, int dir) { return dir < 0 ? cplx(a.imag(), -a.real()) : cplx(-a.imag(), a.real()); }  // a * (dir * i)
  // one Stockham pass with a compile-time radix: values in registers, loops unrolled
  template <int R>
  void pass(const cplx* in, cplx* out, const cplx* t, int Ns, int dir) const {
    const int M = n / R;
    for (int j0 = 0; j0 < M; j0 += Ns)
      for (int k = 0; k < Ns; ++k) {
        const int j = j0 + k;
        cplx v[R];
        v[0] = in[j];
        for (int r = 1; r < R; ++r) v[r] = in[j + (size_t)r * M] * t[(size_t)k * R + r];
        cplx* o = out + (size_t)j0 * R + k;
        if (R == 2) {
          o[0] = v[0] + v[1]; o[Ns] = v[0] - v[1];
        } else if (R == 4) {
          const cplx a0 = v[0] + v[2], a1 = v[0] - v[2], a2 = v[1] + v[3], a3 = mul_i(v[1] - v[3], dir);
          o[0] = a0 + a2; o[Ns] = a1 + a3; o[2 * Ns] = a0 - a2; o[3 * Ns] = a1 - a3;
        } else if (R == 3) {
          const double c = -0.5, sn = dir * 0.86602540378443864676;
          const cplx t1 = v[1] + v[2], t2 = v[0] + c * t1, t3 = mul_i(v[1] - v[2], 1) * sn;
          o[0] = v[0] + t1; o[Ns] = t2 + t3; o[2 * Ns] = t2 - t3;
        } else {   // R == 5
          const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
          const double s1 = dir * 0.95105651629515357212, s2 = dir * 0.58778525229247312917;
          const cplx a1 = v[1] + v[4 % R], a2 = v[2] + v[3], b1 = v[1] - v[4 % R], b2 = v[2] - v[3];
          const cplx m1 = v[0] + c1 * a1 + c2 * a2, m2 = v[0] + c2 * a1 + c1 * a2;
          const cplx n1 = mul_i(s1 * b1 + s2 * b2, 1), n2 = mul_i(s2 * b1 - s1 * b2, 1);
          o[0] = v[0] + a1 + a2; o[Ns] = m1 + n1; o[4 * Ns] = m1 - n1; o[2 * Ns] = m2 + n2; o[3 * Ns] = m2 - n2;
        }
      }
  }
  // any other prime radix: O(R^2) butterfly, scratch behind the ping-pong buffer
  void pass_generic(const cplx* in, cplx* out, const cplx* t, const cplx* rt, cplx* vv, int R, int Ns, int dir) const {
    const int M = n / R;
    for (int j0 = 0; j0 < M; j0 += Ns)
      for (int k = 0; k < Ns; ++k) {
        const int j = j0 + k;
        vv[0] = in[j];
        for (int r = 1; r < R; ++r) vv[r] = in[j + (size_t)r * M] * t[(size_t)k * R + r];
        cplx* o = out + (size_t)j0 * R + k;
        for (int q = 0; q < R; ++q) {
          cplx acc(0, 0);
          for (int r = 0; r < R; ++r) {
            cplx w = rt[(int)(((long long)q * r) % R)];
            if (dir > 0) w = std::conj(w);
            acc += vv[r] * w;
          }
          o[(size_t)q * Ns] = acc;
        }
      }
  }
  // data (n) is transformed; tmp (2 n) is scratch; the result ends in `data`
  void run(cplx* data, cplx* tmp, int dir) const {
    cplx* in = data; cplx* out = tmp;
    int Ns = 1;
    for (size_t s = 0; s < radices.size(); ++s) {
      const int R = radices[s];
      const cplx* t = (dir < 0 ? tw[s] : twc[s]).data();
      switch (R) {
        case 2: pass<2>(in, out, t, Ns, dir); break;
        case 3: pass<3>(in, out, t, Ns, dir); break;
        case 4: pass<4>(in, out, t, Ns, dir); break;
        case 5: pass<5>(in, out, t, Ns, dir); break;
        default: pass_generic(in, out, t, roots[s].data(), tmp + n, R, Ns, dir);
      }
      std::swap(in, out);
      Ns *= R;
    }
    if (in != data) for (int i = 0; i < n; ++i) data[i] = in[i];
  }
};

struct orc_model {
  int N, N3, Nhalf, shift;
  bool even_fix;
  std::vector<orc_material> mats;
  std::vector<CrystalLib> crys;
  std::vector<int32_t> matlist;
  std::vector<double> angles;
  int ncmax;                          // crystals per voxel the tables below are sized for
  std::vector<int32_t> cry_ids;       // (N3, ncmax) 1-based crystal numbers, empty = material's own
  HistLayout L; int H;
  double tolNR, tolPCG, tstep; int maxIter;
  std::vector<double> Fn, Fn1, Pn, Pn1, K4, dFm, b, tmp1, tmp2, tmp3;
  std::vector<double> hist_n, hist_n1, urcs_n, urcs_n1, eps_n, eps_n1, rot_n1;
  std::vector<int32_t> fail, liters;
  std::vector<double> qtab;      // frequency table per axis index
  std::vector<double> c1, c2;    // 1-D phase-ramp tables (phase separable in i+j+k)
  Fft1d fft;
  std::vector<cplx> work;        // 9 * N3 complex
  double t_pcg, t_sig; int64_t n_apply, n_sweep, n_cg, n_fail, n_fail_final;
  // FFT_nr3 locals that persist across load steps (FFT_nr3.f:23-34)
  double barF[9], barF_t[9], P_bar[9], C_homo[81]; bool have_chomo;
};

static int g_threads = 0;
extern "C" void orc_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

extern "C" orc_model* orc_create(int N, int nmat, const orc_material* mats, int ncry, const orc_crystal* crys,
                                 const int32_t* matlist, const double* angles) {
  orc_model* m = new orc_model();
  m->N = N; m->N3 = N * N * N;
  // FFT_init.f:146-147
  m->Nhalf = (N % 2 == 1) ? (N + 1) / 2 : N / 2 + 1;
  // Even N: the reference's shift (Nhalf = N/2+1) is inconsistent with its frequency table
  // (SURVEY.md fact 4).  Convention used by this project for even N: shift = N/2,
  // q = index - N/2, Green operator zero on the Nyquist planes.
  m->even_fix = (N % 2 == 0);
  m->shift = m->even_fix ? N / 2 : m->Nhalf;
  m->mats.assign(mats, mats + nmat);
  m->crys.resize(ncry);
  int nslip_max = 0, nhard_max = 0; bool has_cp = false;
  for (int c = 0; c < ncry; ++c) { m->crys[c].in = crys[c]; finalize_crystal(m->crys[c]); }
  m->matlist.assign(matlist, matlist + m->N3);
  m->angles.assign(angles, angles + 3 * (size_t)m->N3);
  m->ncmax = 1;
  for (int i = 0; i < nmat; ++i)
    if (mats[i].type == 10) {
      has_cp = true;
      // crystal_input file: the material names no crystal; orc_set_taylor sizes the history then
      if (mats[i].crystal >= 1 && mats[i].crystal <= ncry) nslip_max = std::max(nslip_max, m->crys[mats[i].crystal - 1].nslip);
      else nslip_max = std::max(nslip_max, 12);
      nhard_max = std::max(nhard_max, 1);
    }
  m->H = 11;
  if (has_cp) { m->L = mm10_history_layout(nslip_max, nhard_max); m->H = std::max(m->H, m->L.total); }
  size_t n3 = m->N3;
  m->Fn.assign(9 * n3, 0.0);
  for (int c : {0, 4, 8}) for (size_t e = 0; e < n3; ++e) m->Fn[c * n3 + e] = 1.0;  // FFT_init.f:157-158
  m->Fn1 = m->Fn;
  m->Pn.assign(9 * n3, 0.0); m->Pn1.assign(9 * n3, 0.0); m->K4.assign(81 * n3, 0.0);
  m->dFm.assign(9 * n3, 0.0); m->b.assign(9 * n3, 0.0);
  m->tmp1.assign(9 * n3, 0.0); m->tmp2.assign(9 * n3, 0.0); m->tmp3.assign(9 * n3, 0.0);
  m->hist_n.assign((size_t)m->H * n3, 0.0); m->hist_n1.assign((size_t)m->H * n3, 0.0);
  m->urcs_n.assign(9 * n3, 0.0); m->urcs_n1.assign(9 * n3, 0.0);
  m->eps_n.assign(6 * n3, 0.0); m->eps_n1.assign(6 * n3, 0.0); m->rot_n1.assign(9 * n3, 0.0);
  m->fail.assign(n3, 0); m->liters.assign(2 * n3, 0);
  m->tolNR = 1e-5; m->tolPCG = 1e-10; m->maxIter = 10; m->tstep = 1.0;
  // formG frequency table (FFT_init.f:311-318) and formfftshift (FFT_init.f:354-385)
  m->qtab.resize(N);
  for (int i = 0; i < N; ++i) m->qtab[i] = m->even_fix ? (double)(i - N / 2) : (double)(i + 1 - m->Nhalf);
  m->c1.resize(3 * N); m->c2.resize(3 * N);
  double NhN = (double)m->shift * 2.0 * (4.0 * std::atan(1.0)) / (double)N;
  for (int s = 0; s < 3 * N; ++s) { double t = NhN * (double)s; m->c1[s] = std::cos(t); m->c2[s] = -std::sin(t); }
  m->fft.init(N);
  m->work.resize(9 * n3);
  m->t_pcg = m->t_sig = 0; m->n_apply = m->n_sweep = m->n_cg = 0; m->n_fail = m->n_fail_final = 0;
  for (int i = 0; i < 9; ++i) { m->barF[i] = m->barF_t[i] = (i % 4 == 0) ? 1.0 : 0.0; m->P_bar[i] = 0.0; }
  m->have_chomo = false;
  return m;
}
extern "C" int orc_set_taylor(orc_model* m, int ncmax, const double* angles, const int32_t* crystal_ids) {
  if (ncmax < 1 || ncmax > ORC_MAX_CRYSTALS_PER_POINT) return 1;
  const size_t n3 = m->N3;
  int nslip_max = 0, need = 1;
  for (size_t e = 0; e < n3; ++e) {
    const orc_material& mat = m->mats[m->matlist[e] - 1];
    if (mat.type != 10) continue;
    const int nc = std::max(1, (int)mat.n_crystals);
    if (nc > ncmax) return 2;
    need = std::max(need, nc);
    for (int c = 0; c < nc; ++c) {
      const int id = crystal_ids ? crystal_ids[e * ncmax + c] : mat.crystal;
      if (id < 1 || id > (int)m->crys.size()) return 3;
      nslip_max = std::max(nslip_max, m->crys[id - 1].nslip);
    }
  }
  m->ncmax = ncmax;
  m->angles.assign(angles, angles + 3 * n3 * (size_t)ncmax);
  if (crystal_ids) m->cry_ids.assign(crystal_ids, crystal_ids + n3 * (size_t)ncmax); else m->cry_ids.clear();
  if (nslip_max > 0) {
    m->L = mm10_history_layout(nslip_max, 1);
    const int per = m->L.total - m->L.c_stress;
    m->H = std::max(11, m->L.c_stress + need * per);     // mm10_set_sizes_special (mm10_a.f:640-641)
    m->hist_n.assign((size_t)m->H * n3, 0.0); m->hist_n1.assign((size_t)m->H * n3, 0.0);
  }
  return 0;
}
extern "C" void orc_destroy(orc_model* m) { delete m; }
extern "C" void orc_set_params(orc_model* m, double tolNR, double tolPCG, int maxIter, double tstep) {
  m->tolNR = tolNR; m->tolPCG = tolPCG; m->maxIter = maxIter; m->tstep = tstep;
}
extern "C" int orc_hist_size(const orc_model* m) { return m->H; }
extern "C" double* orc_Fn(orc_model* m) { return m->Fn.data(); }
extern "C" double* orc_Fn1(orc_model* m) { return m->Fn1.data(); }
extern "C" double* orc_Pn(orc_model* m) { return m->Pn.data(); }
extern "C" double* orc_Pn1(orc_model* m) { return m->Pn1.data(); }
extern "C" double* orc_K4(orc_model* m) { return m->K4.data(); }
extern "C" double* orc_dFm(orc_model* m) { return m->dFm.data(); }
extern "C" double* orc_b(orc_model* m) { return m->b.data(); }
extern "C" double* orc_hist_n(orc_model* m) { return m->hist_n.data(); }
extern "C" double* orc_hist_n1(orc_model* m) { return m->hist_n1.data(); }
extern "C" double* orc_urcs_n(orc_model* m) { return m->urcs_n.data(); }
extern "C" double* orc_urcs_n1(orc_model* m) { return m->urcs_n1.data(); }
extern "C" double* orc_eps_n(orc_model* m) { return m->eps_n.data(); }
extern "C" double* orc_eps_n1(orc_model* m) { return m->eps_n1.data(); }
extern "C" double* orc_rot_n1(orc_model* m) { return m->rot_n1.data(); }
extern "C" int32_t* orc_fail_flags(orc_model* m) { return m->fail.data(); }
extern "C" int32_t* orc_local_iters(orc_model* m) { return m->liters.data(); }

// ----------------------------------------------------------------------------
// do_nleps_block for one voxel (drive_eps_sig.f:57-339) + rplstr.f:62-85
static int update_point(orc_model* m, size_t e, int step, int iter, const double* Fn9, const double* Fn19,
                        double* P9, double* A81, bool scatter) {
  const int matno = m->matlist[e];
  const orc_material& mat = m->mats[matno - 1];
  const int H = m->H;
  M33 fn, fn1, fnh, dfn, rnh, R, fnhinv, fn1inv;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { fn[i][j] = Fn9[3 * i + j]; fn1[i][j] = Fn19[3 * i + j]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { fnh[i][j] = 0.5 * (fn[i][j] + fn1[i][j]); dfn[i][j] = fn1[i][j] - fn[i][j]; }
  rtcmp1(fnh, rnh);
  rtcmp1(fn1, R);
  double detFh, detF, ddt[6], uddt[6];
  inv33(fnh, fnhinv, &detFh);
  mul33(dfn, fnhinv, ddt);
  M66 qnhalf, qtn1;
  getrm1(qnhalf, rnh, 1);
  qmply1(qnhalf, ddt, uddt);
  // gather n state (dupstr.f)
  std::vector<double> hbuf(2 * (size_t)H);
  double* hn = hbuf.data(); double* h1 = hn + H;
  for (int k = 0; k < H; ++k) { hn[k] = m->hist_n[e * H + k]; h1[k] = 0.0; }
  double urn[9], ur1[9] = {0}, eps1[6], rot9[9];
  for (int k = 0; k < 9; ++k) urn[k] = m->urcs_n[e * 9 + k];
  for (int k = 0; k < 6; ++k) eps1[k] = m->eps_n[e * 6 + k] + uddt[k];  // rstgp1_update_strains
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) rot9[3 * j + i] = R[i][j];
  M66 cep;
  int rc = 0;
  if (mat.type == 1) {
    Mm01Props pr;  // setup_mm01_rknstr (drive_eps_sig.f:486-521): REAL*4 -> REAL*8
    pr.ym = (double)mat.e; pr.nu = (double)mat.nu; pr.beta = (double)mat.beta;
    pr.tan_e = (double)mat.tan_e; pr.yld = (double)mat.yld_pt;
    pr.hprime = pr.tan_e * pr.ym / (pr.ym - pr.tan_e);
    for (int k = 0; k < 11; ++k) h1[k] = m->hist_n1[e * H + k];  // not rewritten slots keep block values
    mm01_point(step, pr, hn, urn, uddt, ur1, h1, cep);
  } else {
    const int nc = std::min(m->ncmax, std::max(1, (int)mat.n_crystals));   // > 1 needs orc_set_taylor
    const CrystalLib* cl[ORC_MAX_CRYSTALS_PER_POINT];
    for (int c = 0; c < nc; ++c)
      cl[c] = &m->crys[(m->cry_ids.empty() ? mat.crystal : m->cry_ids[e * m->ncmax + c]) - 1];
    for (int k = 0; k < H; ++k) h1[k] = m->hist_n1[e * H + k];
    int li[2];
    rc = mm10_point(step, iter, nc, cl, &m->angles[3 * e * (size_t)m->ncmax], m->L, m->tstep, rot9, uddt, hn, h1, urn, ur1, li);
    if (scatter) { m->liters[2 * e] = li[0]; m->liters[2 * e + 1] = li[1]; }
    for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) cep[i][j] = h1[m->L.cep + 6 * j + i];
  }
  if (scatter) {
    m->fail[e] = rc;
    for (int k = 0; k < 9; ++k) m->urcs_n1[e * 9 + k] = ur1[k];
    for (int k = 0; k < 6; ++k) m->eps_n1[e * 6 + k] = eps1[k];
    if (iter > 0) for (int k = 0; k < 9; ++k) m->rot_n1[e * 9 + k] = rot9[k];
    if (iter > 0 || mat.type == 10) for (int k = 0; k < H; ++k) m->hist_n1[e * H + k] = h1[k];
  }
  double cs[6];
  getrm1(qtn1, R, 2);
  qmply1(qtn1, ur1, cs);
  inv33(fn1, fn1inv, &detF);
  cs2p(cs, fn1inv, detF, P9);
  M33 t;
  t[0][0] = ur1[0]; t[1][0] = ur1[3]; t[2][0] = ur1[5];
  t[0][1] = ur1[3]; t[1][1] = ur1[1]; t[2][1] = ur1[4];
  t[0][2] = ur1[5]; t[1][2] = ur1[4]; t[2][2] = ur1[2];
  cep2A_a(fn, t, cep, rnh, detFh, fnhinv, R, fn1, fn1inv, detF, A81);
  return rc;
}

extern "C" int orc_drive_eps_sig(orc_model* m, int step, int iter) {
  double t0 = now_s();
  const size_t n3 = m->N3;
  int nfail = 0;
#pragma omp parallel for schedule(dynamic, 128) reduction(+ : nfail)
  for (long long ee = 0; ee < (long long)n3; ++ee) {
    size_t e = (size_t)ee;
    double f0[9], f1[9], P[9], A[81];
    for (int c = 0; c < 9; ++c) { f0[c] = m->Fn[c * n3 + e]; f1[c] = m->Fn1[c * n3 + e]; }
    int rc = update_point(m, e, step, iter, f0, f1, P, A, true);
    nfail += rc;
    for (int c = 0; c < 9; ++c) m->Pn1[c * n3 + e] = P[c];
    for (int c = 0; c < 81; ++c) m->K4[c * n3 + e] = A[c];
  }
  m->t_sig += now_s() - t0; m->n_sweep++;
  return nfail;
}

extern "C" void orc_point_update(orc_model* m, int voxel, int step, int iter, const double* Fn, const double* Fn1,
                                 double* P9, double* A81) {
  update_point(m, (size_t)voxel, step, iter, Fn, Fn1, P9, A81, false);
}

extern "C" void orc_update(orc_model* m) {  // update.f:75-106 (history, eps, urcs; not rot)
  m->hist_n = m->hist_n1; m->eps_n = m->eps_n1; m->urcs_n = m->urcs_n1;
}

// ----------------------------------------------------------------------------
// ddot42n (G_K_dF.f:241-268) with its fixed summation tree, A4 given by a functor
template <class AF>
static inline void ddot42_point(AF a4, const double* B, double* C) {
  for (int i = 0; i < 9; ++i) {
    double t[9];
    for (int j = 0; j < 9; ++j) t[j] = a4(9 * i + j) * B[j];
    C[i] = t[0] + (((t[1] + t[5]) + (t[3] + t[7])) + ((t[2] + t[6]) + (t[4] + t[8])));
  }
}

// formG (FFT_init.f:272-340) evaluated on the fly for frequency indices (ii,jj,kk), 0-based
static inline void green_entry(const orc_model* m, int ii, int jj, int kk, double* G81) {
  double q[3] = {m->qtab[ii], m->qtab[jj], m->qtab[kk]};
  double qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  bool zero = std::fabs(qq) <= 1e-10;
  if (m->even_fix) {
    const int h = m->N / 2;
    if (ii - h == -h || jj - h == -h || kk - h == -h) zero = true;  // Nyquist planes
  }
  for (int t = 0; t < 81; ++t) {
    int i = t / 27, j = (t / 9) % 3, k = (t / 3) % 3, l = t % 3;
    double v = 0.0;
    if (i == k) v = q[j] * q[l];
    G81[t] = zero ? 0.0 : v / qq;
  }
}
extern "C" void orc_formG_entry(int N, int ii, int jj, int kk, double* G81) {
  orc_model m; m.N = N; m.Nhalf = (N % 2 == 1) ? (N + 1) / 2 : N / 2 + 1; m.even_fix = (N % 2 == 0);
  m.qtab.resize(N);
  for (int i = 0; i < N; ++i) m.qtab[i] = m.even_fix ? (double)(i - N / 2) : (double)(i + 1 - m.Nhalf);
  green_entry(&m, ii, jj, kk, G81);
}

// 3-D complex transform of one component, strides (N^2, N, 1), by three sweeps of 1-D DFTs
// one axis pass of the 3-D transform of all 9 components (the reference runs the 9 components
// in parallel with OpenMP and lets MKL thread inside, G_K_dF.f:50-57; here (component, plane)
// pairs are the parallel tasks)
static void fft3d_all(const orc_model* m, cplx* W, int dir) {
  const int N = m->N; const size_t n3 = m->N3;
  for (int axis = 0; axis < 3; ++axis) {
    const size_t stride = axis == 0 ? 1 : (axis == 1 ? (size_t)N : (size_t)N * N);
#pragma omp parallel
    {
      constexpr int NB = 8;                                  // lines gathered together on the strided axes
      std::vector<cplx> lines((size_t)NB * N), tmp(2 * (size_t)N);
#pragma omp for collapse(2) schedule(static)
      for (int c = 0; c < 9; ++c)
        for (int u = 0; u < N; ++u) {
          cplx* a = W + (size_t)c * n3;
          if (axis == 0) {
            for (int v = 0; v < N; ++v) m->fft.run(a + ((size_t)u * N + v) * N, tmp.data(), dir);
          } else {
            // axis 1: lines (x = u, z = v), stride N; axis 2: lines (y = u, z = v), stride N^2; consecutive v are adjacent
            const size_t base0 = axis == 1 ? (size_t)u * N * N : (size_t)u * N;
            for (int v0 = 0; v0 < N; v0 += NB) {
              const int nb = std::min(NB, N - v0);
              for (int k = 0; k < N; ++k)
                for (int b = 0; b < nb; ++b) lines[(size_t)b * N + k] = a[base0 + v0 + b + k * stride];
              for (int b = 0; b < nb; ++b) m->fft.run(lines.data() + (size_t)b * N, tmp.data(), dir);
              for (int k = 0; k < N; ++k)
                for (int b = 0; b < nb; ++b) a[base0 + v0 + b + k * stride] = lines[(size_t)b * N + k];
            }
          }
        }
    }
  }
}

// G_K_dF (G_K_dF.f:11-87)
extern "C" void orc_G_K_dF(orc_model* m, const double* F, double* GKF, int flgK) {
  const size_t n3 = m->N3; const int N = m->N;
  double* real1 = m->tmp1.data();
  if (flgK) {
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)n3; ++e) {
      double B[9], C[9];
      for (int c = 0; c < 9; ++c) B[c] = F[c * n3 + e];
      const double* K4 = m->K4.data();
      ddot42_point([&](int col) { return K4[(size_t)col * n3 + e]; }, B, C);
      for (int c = 0; c < 9; ++c) real1[c * n3 + e] = C[c];
    }
  } else {
    std::memcpy(real1, F, sizeof(double) * 9 * n3);
  }
  cplx* W = m->work.data();
  // fftfem3d (G_K_dF.f:101-157): phase ramp then forward C2C, unscaled
#pragma omp parallel for schedule(static)
  for (int c = 0; c < 9; ++c) {
    cplx* a = W + (size_t)c * n3;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) for (int k = 0; k < N; ++k) {
      size_t e = ((size_t)i * N + j) * N + k;
      double x = real1[c * n3 + e];
      a[e] = cplx(x * m->c1[i + j + k], x * m->c2[i + j + k]);
    }
  }
  fft3d_all(m, W, -1);
  // Ghat4 contraction of real and imaginary parts separately (G_K_dF.f:64-65)
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < (long long)n3; ++e) {
    int k = (int)(e % N), j = (int)((e / N) % N), i = (int)(e / ((size_t)N * N));
    double G[81], Br[9], Bi[9], Cr[9], Ci[9];
    green_entry(m, i, j, k, G);
    for (int c = 0; c < 9; ++c) { Br[c] = W[(size_t)c * n3 + e].real(); Bi[c] = W[(size_t)c * n3 + e].imag(); }
    ddot42_point([&](int col) { return G[col]; }, Br, Cr);
    ddot42_point([&](int col) { return G[col]; }, Bi, Ci);
    for (int c = 0; c < 9; ++c) W[(size_t)c * n3 + e] = cplx(Cr[c], Ci[c]);
  }
  // ifftfem3d (G_K_dF.f:171-227): backward scaled 1/N3, out = re*c1 + im*c2
  const double scale = 1.0 / (double)n3;
  fft3d_all(m, W, +1);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < 9; ++c) {
    cplx* a = W + (size_t)c * n3;
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) for (int k = 0; k < N; ++k) {
      size_t e = ((size_t)i * N + j) * N + k;
      double re = a[e].real() * scale, im = a[e].imag() * scale;
      double v = re * m->c1[i + j + k];
      double w2 = im * m->c2[i + j + k];
      GKF[c * n3 + e] = v + w2;
    }
  }
  m->n_apply++;
}

static double nrm2(const double* x, size_t n) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long long i = 0; i < (long long)n; ++i) s += x[i] * x[i];
  return std::sqrt(s);
}
static double dot(const double* x, const double* y, size_t n) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long long i = 0; i < (long long)n; ++i) s += x[i] * y[i];
  return s;
}

// fftPcg (FFT_nr3.f:214-360): unpreconditioned CG (stand-in for MKL RCI dcg) with the user
// stopping test ||r|| <= tol*||b|| or ||r|| <= tol, evaluated before every iteration.
// cap > 0: leave the loop after `cap` iterations without an error (bench.py's bounded CPU sample)
static int fftPcg_impl(orc_model* m, const double* b, double* x, double tol, int* iters, double* relres, int cap) {
  double t0 = now_s();
  const size_t n = 9 * (size_t)m->N3;
  const int maxIter = 1000;
  const double eps = 2.220446049250313e-16;
  for (size_t i = 0; i < n; ++i) x[i] = 0.0;
  *iters = 0; if (relres) *relres = 0.0;
  double n2b = nrm2(b, n), tolb = tol * n2b;
  if (tol <= eps || tol >= 1.0) return 3;   // improper tolerance -> die_abort
  if (n2b <= eps) { m->t_pcg += now_s() - t0; return 0; }  // zero rhs, zero solution
  std::vector<double> p(n), Ap(n), r(b, b + n);
  double rr = dot(r.data(), r.data(), n), rr_old = 0.0, resnorm = 0.0;
  int it = 0, rc = 0;
  for (;;) {
    resnorm = nrm2(r.data(), n);
    if (resnorm <= tolb || resnorm <= tol) break;
    if (it >= maxIter) { rc = 2; break; }  // FFT_nr3.f:335
    if (cap > 0 && it >= cap) break;
    if (it == 0) p = r;
    else {
      double beta = rr / rr_old;
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < (long long)n; ++i) p[i] = r[i] + beta * p[i];
    }
    orc_G_K_dF(m, p.data(), Ap.data(), 1);
    double alpha = rr / dot(p.data(), Ap.data(), n);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    rr_old = rr; rr = dot(r.data(), r.data(), n);
    ++it;
  }
  *iters = it; m->n_cg += it;
  if (relres) *relres = resnorm / n2b;
  m->t_pcg += now_s() - t0;
  return rc;
}

extern "C" int orc_fftPcg(orc_model* m, const double* b, double* x, double tol, int* iters, double* relres) {
  return fftPcg_impl(m, b, x, tol, iters, relres, 0);
}
extern "C" int orc_fftPcg_capped(orc_model* m, const double* b, double* x, double tol, int cap, int* iters, double* relres) {
  return fftPcg_impl(m, b, x, tol, iters, relres, cap);
}
extern "C" void orc_counters(const orc_model* m, int64_t* c3, double* t2) {
  c3[0] = m->n_apply; c3[1] = m->n_sweep; c3[2] = m->n_cg;
  t2[0] = m->t_pcg; t2[1] = m->t_sig;
}

extern "C" void orc_mean_P(orc_model* m, double* Pbar) {
  const size_t n3 = m->N3;
  for (int c = 0; c < 9; ++c) {
    double s = 0;
    for (size_t e = 0; e < n3; ++e) s = s + m->Pn1[c * n3 + e];
    Pbar[c] = s / (double)n3;
  }
}

// tangent_homo (tangent_homo.f:11-73).  C_homo flat index k = 9 (j-1) + i (1-based).
extern "C" int orc_tangent_homo(orc_model* m, double* C_homo) {
  const size_t n3 = m->N3, n = 9 * n3;
  std::vector<double> Cij(n);
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) std::memcpy(&Cij[j * n3], &m->K4[(size_t)(9 * j + i) * n3], sizeof(double) * n3);
    orc_G_K_dF(m, Cij.data(), m->b.data(), 0);
    for (size_t k = 0; k < n; ++k) m->b[k] = -m->b[k];
    int it; int rc = orc_fftPcg(m, m->b.data(), Cij.data(), m->tolPCG, &it, nullptr);
    if (rc) return rc;
    for (size_t e = 0; e < n3; ++e) Cij[i * n3 + e] += 1.0;
    for (int j = 0; j < 9; ++j)
      C_homo[9 * j + i] = dot(&m->K4[(size_t)(9 * j) * n3], Cij.data(), n) / (double)n3;
  }
  return 0;
}

// NBC_update (FFT_nr3.f:375-433): 9x9 mixed system by LU with partial pivoting (DGESV)
static int nbc_update(const double* C_homo, double* DbarF, const double* P_bar, const double* PBC, const int32_t* isNBC) {
  double A[81], bb[9];
  for (int i = 0; i < 9; ++i) {
    if (isNBC[i]) {
      for (int mcol = 0; mcol < 9; ++mcol) A[i * 9 + mcol] = C_homo[mcol + 9 * i];  // AAA(i,:) = C_homo(:,i)
      bb[i] = PBC[i] - P_bar[i];
    } else {
      for (int mcol = 0; mcol < 9; ++mcol) A[i * 9 + mcol] = 0.0;
      A[i * 9 + i] = 1.0; bb[i] = DbarF[i];
    }
  }
  // Gaussian elimination with partial pivoting
  for (int k = 0; k < 9; ++k) {
    int piv = k; double best = std::fabs(A[k * 9 + k]);
    for (int r = k + 1; r < 9; ++r) if (std::fabs(A[r * 9 + k]) > best) { best = std::fabs(A[r * 9 + k]); piv = r; }
    if (best == 0.0) return 1;
    if (piv != k) { for (int c = 0; c < 9; ++c) std::swap(A[k * 9 + c], A[piv * 9 + c]); std::swap(bb[k], bb[piv]); }
    for (int r = k + 1; r < 9; ++r) {
      double l = A[r * 9 + k] / A[k * 9 + k];
      for (int c = k; c < 9; ++c) A[r * 9 + c] -= l * A[k * 9 + c];
      bb[r] -= l * bb[k];
    }
  }
  for (int k = 8; k >= 0; --k) { double s = bb[k]; for (int c = k + 1; c < 9; ++c) s -= A[k * 9 + c] * bb[c]; bb[k] = s / A[k * 9 + k]; }
  for (int i = 0; i < 9; ++i) DbarF[i] = bb[i];
  return 0;
}

// FFT_nr3 (FFT_nr3.f:14-200).  Error codes: 1 Newton not converged (:116), 2 CG not converged
// (:335), 3 bad tolerance (:248), 4 stress BC not reached (:153), 5 P_bar update failed (:418),
// 6 material model failure (mm10_a.f:2811).
extern "C" int orc_FFT_nr3_from(orc_model* m, int first_step, int nstep, const double* BC_all, const int32_t* isNBC,
                                int32_t* nr_iters, int32_t* cg_iters, int cg_cap, double* Pbar_out, double* buckets,
                                int64_t* counters) {
  const size_t n3 = m->N3, n = 9 * n3;
  double* barF = m->barF; double* barF_t = m->barF_t; double* P_bar = m->P_bar; double* C_homo = m->C_homo;
  double DbarF[9] = {0}, FBC[9], PBC[9];
  bool existNBC = false;
  for (int i = 0; i < 9; ++i) if (isNBC[i]) existNBC = true;
  m->t_pcg = m->t_sig = 0; m->n_apply = m->n_sweep = m->n_cg = 0; m->n_fail = m->n_fail_final = 0;
  double t_start = now_s();
  int rc = 0;
  if (!m->have_chomo) { rc = orc_tangent_homo(m, C_homo); if (rc) return rc; m->have_chomo = true; }
  for (int sidx = 0; sidx < nstep; ++sidx) {
    const int step = first_step + sidx;
    int ncg = 0;
    auto push_cg = [&](int it) { if (ncg < cg_cap - 1) cg_iters[(size_t)sidx * cg_cap + ncg++] = it; };
    for (int i = 0; i < 9; ++i) {
      PBC[i] = 0; FBC[i] = 0; DbarF[i] = 0;
      if (isNBC[i]) PBC[i] = BC_all[(size_t)sidx * 9 + i];
      else { FBC[i] = BC_all[(size_t)sidx * 9 + i]; DbarF[i] = FBC[i] - barF_t[i]; }
    }
    if (existNBC) { if (nbc_update(C_homo, DbarF, P_bar, PBC, isNBC)) return 5; }
    for (int i = 0; i < 9; ++i) barF[i] = barF_t[i] + DbarF[i];
    int iiter_NBC = 0, total_nr = 0;
    for (;;) {
      for (int c = 0; c < 9; ++c) for (size_t e = 0; e < n3; ++e) m->dFm[c * n3 + e] = DbarF[c];
      double Fnorm = nrm2(m->Fn1.data(), n);
      for (size_t k = 0; k < n; ++k) m->Fn1[k] += m->dFm[k];
      orc_G_K_dF(m, m->dFm.data(), m->b.data(), 1);
      for (size_t k = 0; k < n; ++k) m->b[k] = -m->b[k];
      int it; rc = orc_fftPcg(m, m->b.data(), m->dFm.data(), m->tolPCG, &it, nullptr);
      if (rc) return rc;
      push_cg(it);
      for (size_t k = 0; k < n; ++k) m->Fn1[k] += m->dFm[k];
      double resfft = 1.0;
      int iiter_EBC = 0;
      while (resfft > m->tolNR) {
        m->n_fail += orc_drive_eps_sig(m, step, iiter_EBC);  // local failures are counted, not fatal
        orc_G_K_dF(m, m->Pn1.data(), m->b.data(), 0);
        for (size_t k = 0; k < n; ++k) m->b[k] = -m->b[k];
        rc = orc_fftPcg(m, m->b.data(), m->dFm.data(), m->tolPCG, &it, nullptr);
        if (rc) return rc;
        push_cg(it);
        for (size_t k = 0; k < n; ++k) m->Fn1[k] += m->dFm[k];
        resfft = nrm2(m->dFm.data(), n) / Fnorm;
        if (iiter_EBC == m->maxIter) return 1;
        iiter_EBC++;
      }
      total_nr += iiter_EBC;
      m->n_fail_final = orc_drive_eps_sig(m, step, iiter_EBC); m->n_fail += m->n_fail_final;
      double resP1 = 0, resP2 = 0, resP3;
      orc_mean_P(m, P_bar);
      for (int i = 0; i < 9; ++i) {
        resP2 += P_bar[i] * P_bar[i];
        if (!isNBC[i]) continue;
        resP1 += (P_bar[i] - PBC[i]) * (P_bar[i] - PBC[i]);
      }
      if (resP2 < 1.0e-8) resP3 = std::sqrt(resP1); else resP3 = std::sqrt(resP1 / resP2);
      if (resP3 <= m->tolNR) break;
      if (iiter_NBC > m->maxIter) return 4;
      rc = orc_tangent_homo(m, C_homo);
      if (rc) return rc;
      for (int i = 0; i < 9; ++i) DbarF[i] = 0;
      if (nbc_update(C_homo, DbarF, P_bar, PBC, isNBC)) return 5;
      for (int i = 0; i < 9; ++i) barF[i] += DbarF[i];
      iiter_NBC++;
    }
    for (int i = 0; i < 9; ++i) barF_t[i] = barF[i];
    m->Fn = m->Fn1; m->Pn = m->Pn1;
    orc_update(m);
    nr_iters[sidx] = total_nr;
    cg_iters[(size_t)sidx * cg_cap + ncg] = -1;
    for (int i = 0; i < 9; ++i) Pbar_out[(size_t)sidx * 9 + i] = P_bar[i];
  }
  if (buckets) { buckets[0] = m->t_pcg; buckets[1] = m->t_sig; buckets[2] = now_s() - t_start; }
  if (counters) { counters[0] = m->n_apply; counters[1] = m->n_sweep; counters[2] = m->n_cg; counters[3] = m->n_fail; counters[4] = m->n_fail_final; }
  return 0;
}

extern "C" int orc_FFT_nr3(orc_model* m, int nstep, const double* BC_all, const int32_t* isNBC, int32_t* nr_iters,
                           int32_t* cg_iters, int cg_cap, double* Pbar_out, double* buckets, int64_t* counters) {
  for (int i = 0; i < 9; ++i) { m->barF[i] = m->barF_t[i] = (i % 4 == 0) ? 1.0 : 0.0; m->P_bar[i] = 0.0; }
  m->have_chomo = false;
  return orc_FFT_nr3_from(m, 1, nstep, BC_all, isNBC, nr_iters, cg_iters, cg_cap, Pbar_out, buckets, counters);
}

// arithmetic of the polar decomposition (oracle_kin.cpp): 0 = double, the literal restatement with the
// reference's small-strain noise; 1 = the same formulas in __float128 (default)
extern "C" void orc_set_polar_precision(int quad) { set_polar_precision(quad); }
extern "C" int orc_get_polar_precision() { return get_polar_precision(); }

// ---- unit probes ----
extern "C" void orc_rtcmp1(const double* F9, double* R9) {
  M33 f, r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) f[i][j] = F9[3 * i + j];
  rtcmp1(f, r);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[3 * i + j] = r[i][j];
}
// unit probes for tests/test_reference_vectors.py: getrm1 (polar.f:680-802) and the summation tree of ddot42n (G_K_dF.f:241-268)
extern "C" void orc_getrm1(const double* R9, int opt, double* q36) {
  M33 r; M66 q;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = R9[3 * i + j];
  getrm1(q, r, opt);
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) q36[6 * i + j] = q[i][j];
}
extern "C" void orc_ddot42_point(const double* A81, const double* B9, double* C9) {
  ddot42_point([&](int col) { return A81[col]; }, B9, C9);
}
// the kinematics of do_nleps_block around the material call (drive_eps_sig.f:203-300), for tests/test_reference_vectors.py:
// (Fn, Fn1) -> R of Fn1, unrotated strain increment uddt (rtcmp1, inv33, mul33, getrm1 opt 1, qmply1), and for an
// unrotated stress ur6 the first Piola-Kirchhoff stress P (getrm1 opt 2, qmply1, inv33, cs2p); 3x3 row-major
extern "C" void orc_kinematics_probe(const double* Fn9, const double* Fn19, const double* ur6, double* R9, double* uddt6, double* P9) {
  M33 fn, fn1, fnh, dfn, rnh, R, fnhinv, fn1inv;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { fn[i][j] = Fn9[3 * i + j]; fn1[i][j] = Fn19[3 * i + j]; }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { fnh[i][j] = 0.5 * (fn[i][j] + fn1[i][j]); dfn[i][j] = fn1[i][j] - fn[i][j]; }
  rtcmp1(fnh, rnh); rtcmp1(fn1, R);
  double detFh, detF, ddt[6], cs[6];
  inv33(fnh, fnhinv, &detFh);
  mul33(dfn, fnhinv, ddt);
  M66 qnhalf, qtn1;
  getrm1(qnhalf, rnh, 1);
  qmply1(qnhalf, ddt, uddt6);
  getrm1(qtn1, R, 2);
  qmply1(qtn1, ur6, cs);
  inv33(fn1, fn1inv, &detF);
  cs2p(cs, fn1inv, detF, P9);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[3 * i + j] = R[i][j];
}
extern "C" void orc_cep2A(const double* Fn9, const double* Fn19, const double* t6, const double* cep36, double* A81) {
  M33 fn, fn1, fnh, rnh, R, fnhinv, fn1inv, t; M66 C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { fn[i][j] = Fn9[3 * i + j]; fn1[i][j] = Fn19[3 * i + j]; fnh[i][j] = 0.5 * (fn[i][j] + fn1[i][j]); }
  for (int j = 0; j < 6; ++j) for (int i = 0; i < 6; ++i) C[i][j] = cep36[6 * j + i];
  rtcmp1(fnh, rnh); rtcmp1(fn1, R);
  double dh, d1; inv33(fnh, fnhinv, &dh); inv33(fn1, fn1inv, &d1);
  t[0][0] = t6[0]; t[1][1] = t6[1]; t[2][2] = t6[2];
  t[0][1] = t[1][0] = t6[3]; t[1][2] = t[2][1] = t6[4]; t[0][2] = t[2][0] = t6[5];
  cep2A_a(fn, t, C, rnh, dh, fnhinv, R, fn1, fn1inv, d1, A81);
}
