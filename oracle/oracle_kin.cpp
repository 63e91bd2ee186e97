// CPFFT ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Kinematics, rotation operators, exact dP/dF and the bilinear Mises model, restated
// from polar.f, qmply1.f, drive_eps_sig.f:1017-1224, cep2A.f and mm01.f.
#include "oracle_internal.hpp"
#include <quadmath.h>

namespace orc {

// ---------------------------------------------------------------------------
// Polar decomposition of the reference: evcmp1_new (polar.f:224-307), closed-form (Cardano)
// eigenvalues of the metric tensor, c in upper-triangular order (11,12,22,13,23,33), and
// rtcmp1 / irscp1 / ivcmp1 (polar.f:18-211), R = F U^-1 from the invariants of U.
//
// Written once for a real type T.  T = double is the literal restatement: like the reference
// binary it loses the angle phi to round-off when the principal stretches differ by less than a
// few 1e-3 (the discriminant cancels), which puts a noise of order strain^3 <= ~3e-8 on R, hence
// on every stress, whatever the compiler makes of the expressions.  T = __float128 evaluates the
// SAME formulas with 113-bit arithmetic and rounds R to double at the end: the value the
// reference's algorithm defines, free of that noise.  orc_set_polar_precision() selects which
// one the model uses (default: __float128, the value the GPU results are held to at 1e-9; the
// double version measures the reference's own noise band, tests/test_oracle_material.py).
static inline double r_sqrt(double x) { return std::sqrt(x); }
static inline double r_fabs(double x) { return std::fabs(x); }
static inline double r_atan2(double y, double x) { return std::atan2(y, x); }
static inline double r_cos(double x) { return std::cos(x); }
static inline double r_sin(double x) { return std::sin(x); }
static inline __float128 r_sqrt(__float128 x) { return sqrtq(x); }
static inline __float128 r_fabs(__float128 x) { return fabsq(x); }
static inline __float128 r_atan2(__float128 y, __float128 x) { return atan2q(y, x); }
static inline __float128 r_cos(__float128 x) { return cosq(x); }
static inline __float128 r_sin(__float128 x) { return sinq(x); }

template <class T>
static void evcmp1_new_t(const T c[6], T lam[3]) {
  const T third = T(1) / T(3), oneroot3 = T(1) / r_sqrt(T(3));
  T m11 = c[0], m12 = c[1], m13 = c[3], m22 = c[2], m23 = c[4], m33 = c[5];
  T de = m12 * m23, dd = m12 * m12, ee = m23 * m23, ff = m13 * m13;
  T m = m11 + m22 + m33;
  T c1 = (m11 * m22 + m11 * m33 + m22 * m33) - (dd + ee + ff);
  T c0 = m33 * dd + m11 * ee + m22 * ff - m11 * m22 * m33 - T(2) * m13 * de;
  T p = m * m - T(3) * c1;
  T q = m * (p - T(1.5) * c1) - T(13.5) * c0;
  T sqrtp = r_sqrt(r_fabs(p));
  T phi = T(27) * (T(0.25) * c1 * c1 * (p - c1) + c0 * (q + T(6.75) * c0));
  phi = third * r_atan2(r_sqrt(r_fabs(phi)), q);
  T cphi = sqrtp * r_cos(phi);
  T sphi = oneroot3 * sqrtp * r_sin(phi);
  T e2 = third * (m - cphi);
  T e3 = e2 + sphi;
  T e1 = e2 + cphi;
  e2 = e2 - sphi;
  if (e2 < e1) { T s = e1; e1 = e2; e2 = s; }
  if (e3 < e1) { T s = e1; e1 = e3; e3 = s; }
  if (e3 < e2) { T s = e2; e2 = e3; e3 = s; }
  lam[0] = e1; lam[1] = e2; lam[2] = e3;
}

template <class T>
static void rtcmp1_t(const M33 fd, M33 r) {
  T f[3][3], c[6], cc[6], ev[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) f[i][j] = T(fd[i][j]);
  c[0] = f[0][0] * f[0][0] + f[1][0] * f[1][0] + f[2][0] * f[2][0];
  c[1] = f[0][0] * f[0][1] + f[1][0] * f[1][1] + f[2][0] * f[2][1];
  c[2] = f[0][1] * f[0][1] + f[1][1] * f[1][1] + f[2][1] * f[2][1];
  c[3] = f[0][0] * f[0][2] + f[1][0] * f[1][2] + f[2][0] * f[2][2];
  c[4] = f[0][1] * f[0][2] + f[1][1] * f[1][2] + f[2][1] * f[2][2];
  c[5] = f[0][2] * f[0][2] + f[1][2] * f[1][2] + f[2][2] * f[2][2];
  cc[0] = c[0] * c[0] + c[1] * c[1] + c[3] * c[3];
  cc[1] = c[0] * c[1] + c[1] * c[2] + c[3] * c[4];
  cc[2] = c[1] * c[1] + c[2] * c[2] + c[4] * c[4];
  cc[3] = c[0] * c[3] + c[1] * c[4] + c[3] * c[5];
  cc[4] = c[1] * c[3] + c[2] * c[4] + c[4] * c[5];
  cc[5] = c[3] * c[3] + c[4] * c[4] + c[5] * c[5];
  evcmp1_new_t<T>(c, ev);
  ev[0] = r_sqrt(ev[0]); ev[1] = r_sqrt(ev[1]); ev[2] = r_sqrt(ev[2]);
  T iu = ev[0] + ev[1] + ev[2];
  T iiu = ev[0] * ev[1] + ev[1] * ev[2] + ev[0] * ev[2];
  T iiiu = ev[0] * ev[1] * ev[2];
  T a2 = T(1) / (iiiu * (iu * iiu - iiiu));
  T b2 = iu * iiu * iiu - iiiu * (iu * iu + iiu);
  T c2 = -iiiu - iu * (iu * iu - T(2) * iiu);
  T d2 = iu;
  T ui[6];
  ui[0] = a2 * (b2 + c2 * c[0] + d2 * cc[0]);
  ui[1] = a2 * (c2 * c[1] + d2 * cc[1]);
  ui[2] = a2 * (b2 + c2 * c[2] + d2 * cc[2]);
  ui[3] = a2 * (c2 * c[3] + d2 * cc[3]);
  ui[4] = a2 * (c2 * c[4] + d2 * cc[4]);
  ui[5] = a2 * (b2 + c2 * c[5] + d2 * cc[5]);
  for (int i = 0; i < 3; ++i) {
    r[i][0] = (double)(f[i][0] * ui[0] + f[i][1] * ui[1] + f[i][2] * ui[3]);
    r[i][1] = (double)(f[i][0] * ui[1] + f[i][1] * ui[2] + f[i][2] * ui[4]);
    r[i][2] = (double)(f[i][0] * ui[3] + f[i][1] * ui[4] + f[i][2] * ui[5]);
  }
}

static int g_polar_quad = 1;
void set_polar_precision(int quad) { g_polar_quad = quad ? 1 : 0; }
int get_polar_precision() { return g_polar_quad; }
void rtcmp1(const M33 f, M33 r) {
  if (g_polar_quad) rtcmp1_t<__float128>(f, r);
  else rtcmp1_t<double>(f, r);
}

// getrm1 (polar.f:680-802). opt 1: {d} = q{D}, d = R^T D R (engineering shear);
// opt 2: {T} = q{t}, T = R t R^T.
void getrm1(M66 q, const M33 r, int opt) {
  const double two = 2.0;
  if (opt == 1) {
    q[0][0] = r[0][0] * r[0][0]; q[0][1] = r[1][0] * r[1][0]; q[0][2] = r[2][0] * r[2][0];
    q[0][3] = r[0][0] * r[1][0]; q[0][4] = r[2][0] * r[1][0]; q[0][5] = r[0][0] * r[2][0];
    q[1][0] = r[0][1] * r[0][1]; q[1][1] = r[1][1] * r[1][1]; q[1][2] = r[2][1] * r[2][1];
    q[1][3] = r[0][1] * r[1][1]; q[1][4] = r[2][1] * r[1][1]; q[1][5] = r[0][1] * r[2][1];
    q[2][0] = r[0][2] * r[0][2]; q[2][1] = r[1][2] * r[1][2]; q[2][2] = r[2][2] * r[2][2];
    q[2][3] = r[0][2] * r[1][2]; q[2][4] = r[2][2] * r[1][2]; q[2][5] = r[0][2] * r[2][2];
    q[3][0] = two * r[0][0] * r[0][1]; q[3][1] = two * r[1][0] * r[1][1]; q[3][2] = two * r[2][0] * r[2][1];
    q[3][3] = r[0][0] * r[1][1] + r[0][1] * r[1][0];
    q[3][4] = r[1][0] * r[2][1] + r[2][0] * r[1][1];
    q[3][5] = r[0][0] * r[2][1] + r[2][0] * r[0][1];
    q[4][0] = two * r[0][1] * r[0][2]; q[4][1] = two * r[1][2] * r[1][1]; q[4][2] = two * r[2][1] * r[2][2];
    q[4][3] = r[0][1] * r[1][2] + r[1][1] * r[0][2];
    q[4][4] = r[1][1] * r[2][2] + r[1][2] * r[2][1];
    q[4][5] = r[0][1] * r[2][2] + r[2][1] * r[0][2];
    q[5][0] = two * r[0][0] * r[0][2]; q[5][1] = two * r[1][0] * r[1][2]; q[5][2] = two * r[2][0] * r[2][2];
    q[5][3] = r[0][0] * r[1][2] + r[1][0] * r[0][2];
    q[5][4] = r[1][0] * r[2][2] + r[2][0] * r[1][2];
    q[5][5] = r[0][0] * r[2][2] + r[0][2] * r[2][0];
  } else {
    q[0][0] = r[0][0] * r[0][0]; q[0][1] = r[0][1] * r[0][1]; q[0][2] = r[0][2] * r[0][2];
    q[0][3] = two * r[0][0] * r[0][1]; q[0][4] = two * r[0][2] * r[0][1]; q[0][5] = two * r[0][0] * r[0][2];
    q[1][0] = r[1][0] * r[1][0]; q[1][1] = r[1][1] * r[1][1]; q[1][2] = r[1][2] * r[1][2];
    q[1][3] = two * r[1][0] * r[1][1]; q[1][4] = two * r[1][2] * r[1][1]; q[1][5] = two * r[1][0] * r[1][2];
    q[2][0] = r[2][0] * r[2][0]; q[2][1] = r[2][1] * r[2][1]; q[2][2] = r[2][2] * r[2][2];
    q[2][3] = two * r[2][0] * r[2][1]; q[2][4] = two * r[2][2] * r[2][1]; q[2][5] = two * r[2][0] * r[2][2];
    q[3][0] = r[0][0] * r[1][0]; q[3][1] = r[0][1] * r[1][1]; q[3][2] = r[0][2] * r[1][2];
    q[3][3] = r[0][0] * r[1][1] + r[1][0] * r[0][1];
    q[3][4] = r[0][1] * r[1][2] + r[0][2] * r[1][1];
    q[3][5] = r[0][0] * r[1][2] + r[0][2] * r[1][0];
    q[4][0] = r[1][0] * r[2][0]; q[4][1] = r[2][1] * r[1][1]; q[4][2] = r[1][2] * r[2][2];
    q[4][3] = r[1][0] * r[2][1] + r[1][1] * r[2][0];
    q[4][4] = r[1][1] * r[2][2] + r[2][1] * r[1][2];
    q[4][5] = r[1][0] * r[2][2] + r[1][2] * r[2][0];
    q[5][0] = r[0][0] * r[2][0]; q[5][1] = r[0][1] * r[2][1]; q[5][2] = r[0][2] * r[2][2];
    q[5][3] = r[0][0] * r[2][1] + r[0][1] * r[2][0];
    q[5][4] = r[0][1] * r[2][2] + r[0][2] * r[2][1];
    q[5][5] = r[0][0] * r[2][2] + r[2][0] * r[0][2];
  }
}

// qmply1.f:15-36: m2 = q m1, summed left to right
void qmply1(const M66 q, const double* m1, double* m2) {
  for (int i = 0; i < 6; ++i)
    m2[i] = q[i][0] * m1[0] + q[i][1] * m1[1] + q[i][2] * m1[2] + q[i][3] * m1[3] +
            q[i][4] * m1[4] + q[i][5] * m1[5];
}

void inv33(const M33 jac, M33 gama, double* dj) {
  double j1 = jac[1][1] * jac[2][2] - jac[1][2] * jac[2][1];
  double j2 = jac[1][0] * jac[2][2] - jac[1][2] * jac[2][0];
  double j3 = jac[1][0] * jac[2][1] - jac[1][1] * jac[2][0];
  double d = jac[0][0] * j1 - jac[0][1] * j2 + jac[0][2] * j3;
  *dj = d;
  gama[0][0] = j1 / d;
  gama[1][0] = -j2 / d;
  gama[2][0] = j3 / d;
  gama[0][1] = (jac[2][1] * jac[0][2] - jac[0][1] * jac[2][2]) / d;
  gama[1][1] = (jac[0][0] * jac[2][2] - jac[2][0] * jac[0][2]) / d;
  gama[2][1] = (jac[0][1] * jac[2][0] - jac[0][0] * jac[2][1]) / d;
  gama[0][2] = (jac[0][1] * jac[1][2] - jac[0][2] * jac[1][1]) / d;
  gama[1][2] = (jac[0][2] * jac[1][0] - jac[0][0] * jac[1][2]) / d;
  gama[2][2] = (jac[0][0] * jac[1][1] - jac[0][1] * jac[1][0]) / d;
}

// mul33: Voigt (xx,yy,zz,xy,yz,xz) of A*B with summed (engineering) shear terms
void mul33(const M33 A, const M33 B, double* C) {
  C[0] = A[0][0] * B[0][0] + A[0][1] * B[1][0] + A[0][2] * B[2][0];
  C[3] = A[1][0] * B[0][0] + A[1][1] * B[1][0] + A[1][2] * B[2][0] + A[0][0] * B[0][1] +
         A[0][1] * B[1][1] + A[0][2] * B[2][1];
  C[5] = A[2][0] * B[0][0] + A[2][1] * B[1][0] + A[2][2] * B[2][0] + A[0][0] * B[0][2] +
         A[0][1] * B[1][2] + A[0][2] * B[2][2];
  C[1] = A[1][0] * B[0][1] + A[1][1] * B[1][1] + A[1][2] * B[2][1];
  C[4] = A[2][0] * B[0][1] + A[2][1] * B[1][1] + A[2][2] * B[2][1] + A[1][0] * B[0][2] +
         A[1][1] * B[1][2] + A[1][2] * B[2][2];
  C[2] = A[2][0] * B[0][2] + A[2][1] * B[1][2] + A[2][2] * B[2][2];
}

void cs2p(const double* cs, const M33 finv, double detF, double* P) {
  // Voigt: 0-11, 1-22, 2-33, 3-12, 4-23, 5-13
  const int row[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      P[3 * i + j] = detF * (cs[row[i][0]] * finv[j][0] + cs[row[i][1]] * finv[j][1] +
                             cs[row[i][2]] * finv[j][2]);
}

// ---------------------------------------------------------------------------
// cep2A_a (cep2A.f:86-284).  Index m (0-based) = 27 i + 9 j + 3 k + l.
static inline void mult33(const M33 A, const M33 B, M33 C) {
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
}
static inline void trans33(const M33 A, M33 B) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[j][i] = A[i][j];
}
static inline double det33(const M33 A) {
  double d = A[0][0] * A[1][1] * A[2][2];
  d = d + A[0][1] * A[1][2] * A[2][0];
  d = d + A[0][2] * A[1][0] * A[2][1];
  d = d - A[0][0] * A[1][2] * A[2][1];
  d = d - A[0][1] * A[1][0] * A[2][2];
  d = d - A[0][2] * A[1][1] * A[2][0];
  return d;
}

void cep2A_a(const M33 Fn, const M33 t, const M66 C, const M33 Rh, double /*detFnh*/,
             const M33 fnhinv, const M33 R, const M33 Fn1, const M33 finv, double detFn1,
             double* dPdF) {
  M33 transR, U, Y, temp, sigma, Uinv, t_Uinv, R_t, RY, RYR, RtRF, dJdF, dFn, Fnh;
  M33 transRh, Uh, Yh, RYRh, RYh, FpFinv, L33, D2;
  trans33(R, transR);
  mult33(transR, Fn1, U);
  double traceU = U[0][0] + U[1][1] + U[2][2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Y[i][j] = (i == j ? traceU : 0.0) - U[i][j];
  mult33(R, t, temp);
  mult33(temp, transR, sigma);
  mult33(finv, R, Uinv);
  mult33(t, Uinv, t_Uinv);
  mult33(R, t, R_t);
  mult33(R, Y, RY);
  mult33(RY, transR, RYR);
  double detY_1 = 1.0 / det33(Y);
  mult33(R_t, Uinv, RtRF);
  trans33(finv, dJdF);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      dJdF[i][j] *= detFn1;
      dFn[i][j] = Fn1[i][j] - Fn[i][j];
      Fnh[i][j] = 0.5 * (Fn1[i][j] + Fn[i][j]);
    }
  trans33(Rh, transRh);
  mult33(transRh, Fnh, Uh);
  double traceUh = Uh[0][0] + Uh[1][1] + Uh[2][2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Yh[i][j] = (i == j ? traceUh : 0.0) - Uh[i][j];
  mult33(Rh, Yh, temp);
  mult33(temp, transRh, RYRh);
  double detYh_1 = 1.0 / det33(Yh);
  mult33(Rh, Yh, RYh);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) FpFinv[i][j] = 0.5 * fnhinv[i][j];
  mult33(dFn, FpFinv, L33);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) D2[i][j] = L33[i][j] + L33[j][i];

  // cep99 = C(index_voigt, index_voigt), index over (row,col) of a 3x3 listed 11,12,13,21,...
  const int iv[9] = {0, 3, 5, 3, 1, 4, 5, 4, 2};
  double cep99[9][9];
  for (int a = 0; a < 9; ++a)
    for (int b = 0; b < 9; ++b) cep99[a][b] = C[iv[a]][iv[b]];

  double dRdF[81], dRhdF[81], dLdF[81], dtdF1[81], dtdF2[81], dtdF3[81], dtdF[81];
  for (int m = 0; m < 81; ++m) {
    int i = m / 27, j = (m / 9) % 3, k = (m / 3) % 3, l = m % 3;
    dRdF[m] = detY_1 * (RYR[i][k] * Y[l][j] - RY[i][l] * RY[k][j]);
    dRhdF[m] = detYh_1 * (RYRh[i][k] * Yh[l][j] - RYh[i][l] * RYh[k][j]);
  }
  for (int m = 0; m < 81; ++m) dRhdF[m] *= 0.5;
  for (int m = 0; m < 81; ++m) {
    int i = m / 27, j = (m / 9) % 3, k = (m / 3) % 3, l = m % 3;
    dLdF[m] = -L33[i][k] * FpFinv[l][j];
    if (i == k) dLdF[m] = dLdF[m] + FpFinv[l][j];
  }
  for (int a = 0; a < 81; ++a) {
    int i = a / 27, j = (a / 9) % 3, p = (a / 3) % 3, q = a % 3;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int m = 0; m < 3; ++m)
      for (int n = 0; n < 3; ++n) {
        int tmp = m * 27 + i * 9 + p * 3 + q;
        s1 = s1 + dRhdF[tmp] * D2[m][n] * Rh[n][j];
        tmp = m * 27 + n * 9 + p * 3 + q;
        s2 = s2 + Rh[m][i] * dLdF[tmp] * Rh[n][j];
        tmp = n * 27 + m * 9 + p * 3 + q;
        s2 = s2 + Rh[m][i] * dLdF[tmp] * Rh[n][j];
        tmp = n * 27 + j * 9 + p * 3 + q;
        s3 = s3 + Rh[m][i] * D2[m][n] * dRhdF[tmp];
      }
    dtdF1[a] = s1; dtdF2[a] = s2; dtdF3[a] = s3;
  }
  double dd[81];
  for (int a = 0; a < 81; ++a) dd[a] = dtdF1[a] + dtdF2[a] + dtdF3[a];
  // cep2A_ddot44 (cep2A.f:334-348): C4(ii,jj) = sum_r cep99(r,ii) * B4(tmp2(r)+jj),
  // r runs over (m,n) column-major: r = 3 n + m
  for (int ii = 0; ii < 9; ++ii)
    for (int jj = 0; jj < 9; ++jj) {
      double s = 0.0;
      for (int r = 0; r < 9; ++r) {
        int m = r % 3, n = r / 3;
        s += cep99[r][ii] * dd[m * 27 + n * 9 + jj];
      }
      dtdF[ii * 9 + jj] = s;
    }
  for (int p = 0; p < 81; ++p) {
    int i = p / 27, j = (p / 9) % 3, k = (p / 3) % 3, l = p % 3;
    double d0 = RtRF[i][j] * dJdF[k][l];
    double d1 = 0.0, d2 = 0.0, d3 = 0.0, d4 = 0.0;
    for (int m = 0; m < 3; ++m) {
      d1 = d1 - detFn1 * sigma[i][m] * finv[j][k] * finv[l][m];
      int tmp = i * 27 + m * 9 + k * 3 + l;
      d2 = d2 + detFn1 * dRdF[tmp] * t_Uinv[m][j];
      for (int n = 0; n < 3; ++n) {
        tmp = m * 27 + n * 9 + k * 3 + l;
        d3 = d3 + detFn1 * R[i][m] * dtdF[tmp] * Uinv[n][j];
        tmp = n * 27 + m * 9 + k * 3 + l;
        d4 = d4 + detFn1 * R_t[i][m] * dRdF[tmp] * finv[j][n];
      }
    }
    dPdF[p] = d0 + d1 + d2 + d3 + d4;
  }
}

// ---------------------------------------------------------------------------
// mm01 (mm01.f:28-784) + cnst1 (mm01.f:1222-1374), one material point, isothermal.
static inline double state_word(int istate) {  // integer packed in a double via equivalence
  int64_t w = (int64_t)(uint32_t)istate;
  double d; std::memcpy(&d, &w, 8); return d;
}
static inline int state_of(double d) { int64_t w; std::memcpy(&w, &d, 8); return (int)(w & 0xffffffff); }

void mm01_point(int step, const Mm01Props& pr, double* history, double* cgn, const double* deps,
                double* cgn1, double* history1, M66 cep) {
  const double root3_hist = 1.73205080756888;     // mm01.f:262
  const double root3_init = 1.7320508075688;      // mm01.f:362 (sic)
  const double root2 = 1.414213562373095, yld_tol = 0.0000001, htol = 0.000001;
  const double twthrd = 0.666666666666667, root23 = 0.816496580927;
  const double ym = pr.ym, nu = pr.nu, beta = pr.beta, hprime = pr.hprime, yld = pr.yld;

  if (step == 1) {  // mm01_set_history (mm01.f:240-313): every step-1 call
    double kn = yld / root3_hist;
    double h0[11] = {0, kn, 0, state_word(3), hprime, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 11; ++k) { history[k] = h0[k]; history1[k] = h0[k]; }
    cgn[7] = 0.0; cgn[8] = 0.0;
  }
  // ---- mm01_init (mm01.f:328-534); ym_n = ym, nu_n = nu (setup_mm01_rknstr) ----
  double deps_vol = deps[0] + deps[1] + deps[2];
  double eps_mean = deps_vol / 3.0;
  double de[6] = {deps[0] - eps_mean, deps[1] - eps_mean, deps[2] - eps_mean, deps[3], deps[4], deps[5]};
  double e_n = ym, nu_n = nu;
  double g_n = e_n / 2.0 / (1.0 + nu_n);
  double een1 = (cgn[0] - nu_n * (cgn[1] + cgn[2])) / e_n;
  double een2 = (cgn[1] - nu_n * (cgn[0] + cgn[2])) / e_n;
  double een3 = (cgn[2] - nu_n * (cgn[0] + cgn[1])) / e_n;
  double een4 = cgn[3] / g_n, een5 = cgn[4] / g_n, een6 = cgn[5] / g_n;
  double eps_vol_n1 = een1 + een2 + een3 + deps_vol;
  double eps_mean_n = (een1 + een2 + een3) / 3.0;
  double e[6] = {(een1 - eps_mean_n) + de[0], (een2 - eps_mean_n) + de[1], (een3 - eps_mean_n) + de[2],
                 een4 + de[3], een5 + de[4], een6 + de[5]};
  double shear_mod = ym / (2.0 * (1.0 + nu));
  double dse[6] = {2.0 * shear_mod * e[0], 2.0 * shear_mod * e[1], 2.0 * shear_mod * e[2],
                   shear_mod * e[3], shear_mod * e[4], shear_mod * e[5]};
  double alpha_n[6];
  for (int k = 0; k < 6; ++k) alpha_n[k] = history[5 + k];
  double hbari_np1 = beta * hprime;
  double hbark_np1 = (1.0 - beta) * hprime;
  double hbark_n = (1.0 - beta) * history[4];
  double kbar = (yld + hbari_np1 * history[2]) / root3_init;
  double lk = 1.0;
  if (std::fabs(hbark_n) > htol) lk = hbark_np1 / hbark_n;
  double rtse[6];
  for (int k = 0; k < 6; ++k) rtse[k] = dse[k] - alpha_n[k] * lk;
  double mrts = std::sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                          2.0 * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
  double yf = mrts - root2 * kbar;
  for (int k = 0; k < 6; ++k) rtse[k] = dse[k] - alpha_n[k];
  mrts = std::sqrt(rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                   2.0 * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]));
  int instat = 3; bool yield = false;
  if (yf >= yld_tol * root2 * kbar) { yield = true; instat = 1; }

  double devstr[6];
  if (yield) {  // mm01_simple1 (mm01.f:704-784)
    double lambda_deltat = (mrts - root2 * kbar) / (twthrd * (3.0 * shear_mod + hprime));
    double k_np1 = kbar + (root2 / 3.0) * hbari_np1 * lambda_deltat;
    history1[0] = lambda_deltat;
    history1[1] = k_np1;
    history1[2] = history[2] + lambda_deltat * root23;
    history1[4] = hprime;
    double const1 = twthrd * hbark_np1 * lambda_deltat / mrts;
    double const2 = root2 * k_np1 / mrts;
    for (int k = 0; k < 6; ++k) history1[5 + k] = history[5 + k] + const1 * rtse[k];
    for (int k = 0; k < 6; ++k) devstr[k] = history1[5 + k] + const2 * rtse[k];
  } else {      // mm01.f:190-199
    history1[0] = 0.0;
    history1[1] = kbar;
    history1[2] = history[2];
    history1[4] = hprime;
    for (int k = 0; k < 6; ++k) history1[5 + k] = alpha_n[k] * lk;
    for (int k = 0; k < 6; ++k) devstr[k] = rtse[k] + alpha_n[k];
  }
  // mm01_sig_final (mm01.f:626-689)
  double sig_mean = eps_vol_n1 * (3.0 * ym * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)) + 2.0 * shear_mod) / 3.0;
  cgn1[0] = devstr[0] + sig_mean; cgn1[1] = devstr[1] + sig_mean; cgn1[2] = devstr[2] + sig_mean;
  cgn1[3] = devstr[3]; cgn1[4] = devstr[4]; cgn1[5] = devstr[5];
  cgn1[6] = cgn[6] + 0.5 * (deps[0] * (cgn1[0] + cgn[0]) + deps[1] * (cgn1[1] + cgn[1]) +
                            deps[2] * (cgn1[2] + cgn[2]) + deps[3] * (cgn1[3] + cgn[3]) +
                            deps[4] * (cgn1[4] + cgn[4]) + deps[5] * (cgn1[5] + cgn[5]));
  history1[3] = state_word(instat);
  // mm01_plastic_work (mm01.f:546-614)
  cgn1[7] = cgn[7]; cgn1[8] = cgn[8];
  if (yield) {
    double dsig[6], dp[6];
    for (int k = 0; k < 6; ++k) dsig[k] = cgn1[k] - cgn[k];
    dp[0] = deps[0] - (dsig[0] - nu * (dsig[1] + dsig[2])) / ym;
    dp[1] = deps[1] - (dsig[1] - nu * (dsig[0] + dsig[2])) / ym;
    dp[2] = deps[2] - (dsig[2] - nu * (dsig[0] + dsig[1])) / ym;
    dp[3] = deps[3] - dsig[3] / shear_mod;
    dp[4] = deps[4] - dsig[4] / shear_mod;
    dp[5] = deps[5] - dsig[5] / shear_mod;
    cgn1[7] = cgn[7] + 0.5 * (dp[0] * (cgn1[0] + cgn[0]) + dp[1] * (cgn1[1] + cgn[1]) +
                              dp[2] * (cgn1[2] + cgn[2]) + dp[3] * (cgn1[3] + cgn[3]) +
                              dp[4] * (cgn1[4] + cgn[4]) + dp[5] * (cgn1[5] + cgn[5]));
    double f1 = (dp[0] - dp[1]) * (dp[0] - dp[1]) + (dp[1] - dp[2]) * (dp[1] - dp[2]) +
                (dp[0] - dp[2]) * (dp[0] - dp[2]);
    double f2 = dp[3] * dp[3] + dp[4] * dp[4] + dp[5] * dp[5];
    double bar = (root2 / 3.0) * std::sqrt(f1 + (3.0 / 2.0) * f2);
    cgn1[8] = cgn[8] + bar;
  }
  // ---- cnst1 (mm01.f:1222-1374) using history1(1,2,4,5) and rtse ----
  const double root2c = 1.414213562;  // sic, mm01.f:1247
  double kn1 = history1[1], hp = history1[4], ldt = history1[0];
  bool yl = state_of(history1[3]) == 1;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) cep[i][j] = 0.0;
  if (!yl) {
    double c1 = (ym / ((1.0 + nu) * (1.0 - 2.0 * nu)));
    double c2 = (1.0 - nu) * c1;
    double c3 = ((1.0 - 2.0 * nu) / 2.0) * c1;
    double c4 = nu * c1;
    cep[0][0] = cep[1][1] = cep[2][2] = c2;
    cep[3][3] = cep[4][4] = cep[5][5] = c3;
    cep[0][1] = cep[0][2] = cep[1][0] = cep[2][0] = cep[1][2] = cep[2][1] = c4;
  } else {
    double g = ym / (2.0 * (1.0 + nu));
    double l = (ym * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu));
    double k = (3.0 * l + 2.0 * g) / 3.0;
    double mrtsq = rtse[0] * rtse[0] + rtse[1] * rtse[1] + rtse[2] * rtse[2] +
                   2.0 * (rtse[3] * rtse[3] + rtse[4] * rtse[4] + rtse[5] * rtse[5]);
    double bb = (root2c * kn1 + (2.0 / 3.0) * (1.0 - beta) * hp * ldt) / std::sqrt(mrtsq);
    double gamma = 1.0 / (1.0 + hp / (3.0 * g));
    double gambar = gamma - 1.0 + bb;
    double gbar = g * bb;
    double albar = k - 2.0 * gbar / 3.0;
    double thbar = 2.0 * g * gambar;
    for (int i = 0; i < 3; ++i) cep[i][i] = (albar + 2.0 * gbar - thbar * (rtse[i] * rtse[i]) / mrtsq);
    for (int i = 3; i < 6; ++i) cep[i][i] = (gbar - thbar * (rtse[i] * rtse[i]) / mrtsq);
    cep[1][0] = (albar - thbar * rtse[0] * rtse[1] / mrtsq);
    cep[2][0] = (albar - thbar * rtse[0] * rtse[2] / mrtsq);
    cep[2][1] = (albar - thbar * rtse[2] * rtse[1] / mrtsq);
    for (int i = 3; i < 6; ++i)
      for (int j = 0; j < i; ++j)
        cep[i][j] = -(thbar * rtse[j] * rtse[i] / mrtsq);
    for (int i = 0; i < 6; ++i)
      for (int j = i + 1; j < 6; ++j) cep[i][j] = cep[j][i];
  }
}

} // namespace orc
