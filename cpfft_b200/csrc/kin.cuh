// cpfft_b200: per-voxel finite-strain kinematics and the exact dP/dF tangent, device side.
// Replaces (per voxel, in registers) rtcmp1/irscp1/ivcmp1/evcmp1_new (polar.f:18-307),
// getrm1+qmply1 (polar.f:680-802, qmply1.f), inv33/mul33/cs2p (drive_eps_sig.f:1017-1224)
// and cep2A_a (cep2A.f:86-284).  3x3 matrices are row-major double[9].
#pragma once
#include "compat.cuh"

CPF_DI void m3_mul(const double* A, const double* B, double* C) {  // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
CPF_DI void m3_mul_tn(const double* A, const double* B, double* C) {  // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
CPF_DI void m3_mul_nt(const double* A, const double* B, double* C) {  // C = A B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
CPF_DI double m3_det(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
         A[2] * (A[3] * A[7] - A[4] * A[6]);
}
// inverse and determinant, cofactor formula as inv33 (drive_eps_sig.f:1017-1107)
CPF_DI double m3_inv(const double* a, double* g) {
  double j1 = a[4] * a[8] - a[5] * a[7];
  double j2 = a[3] * a[8] - a[5] * a[6];
  double j3 = a[3] * a[7] - a[4] * a[6];
  double d = a[0] * j1 - a[1] * j2 + a[2] * j3;
  double id = 1.0 / d;
  g[0] = j1 * id;  g[3] = -j2 * id;  g[6] = j3 * id;
  g[1] = (a[7] * a[2] - a[1] * a[8]) * id;
  g[4] = (a[0] * a[8] - a[6] * a[2]) * id;
  g[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  g[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  g[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  g[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return d;
}

// R = F U^-1 with U^-1 = a (b I + c C + d C^2), C = F^T F; eigenvalues of C by the closed
// form trigonometric solution (polar.f:224-307), invariants of U (polar.f:196-208).
CPF_DI void polar_R(const double* f, double* r) {
  // From F to the discriminant every product and sum is rounded on its own, in source order
  // (CPF_MUL / CPF_ADD / CPF_SUB, never contracted into an FMA).  The discriminant of the cubic
  // cancels to round-off when the principal stretches differ by < 3e-3: the angle phi is then
  // noise that reaches the stress at the order strain^3 ~ 1e-8.  With one fixed rounding
  // sequence -- the same one in oracle/oracle_kin.cpp -- that noise is the same number everywhere,
  // so per-voxel results are comparable to 1e-9 at small strain increments as well.
#define M_(a, b) CPF_MUL(a, b)
#define A_(a, b) CPF_ADD(a, b)
#define S_(a, b) CPF_SUB(a, b)
  const double c0 = A_(A_(M_(f[0], f[0]), M_(f[3], f[3])), M_(f[6], f[6]));   // C11
  const double c1 = A_(A_(M_(f[0], f[1]), M_(f[3], f[4])), M_(f[6], f[7]));   // C12
  const double c2 = A_(A_(M_(f[1], f[1]), M_(f[4], f[4])), M_(f[7], f[7]));   // C22
  const double c3 = A_(A_(M_(f[0], f[2]), M_(f[3], f[5])), M_(f[6], f[8]));   // C13
  const double c4 = A_(A_(M_(f[1], f[2]), M_(f[4], f[5])), M_(f[7], f[8]));   // C23
  const double c5 = A_(A_(M_(f[2], f[2]), M_(f[5], f[5])), M_(f[8], f[8]));   // C33
  double cc0 = c0 * c0 + c1 * c1 + c3 * c3;
  double cc1 = c0 * c1 + c1 * c2 + c3 * c4;
  double cc2 = c1 * c1 + c2 * c2 + c4 * c4;
  double cc3 = c0 * c3 + c1 * c4 + c3 * c5;
  double cc4 = c1 * c3 + c2 * c4 + c4 * c5;
  double cc5 = c3 * c3 + c4 * c4 + c5 * c5;
  const double third = 0.3333333333333333333, oneroot3 = 0.5773502691896258;
  const double de = M_(c1, c4), dd = M_(c1, c1), ee = M_(c4, c4), ff = M_(c3, c3);
  const double m = A_(A_(c0, c2), c5);
  const double k1 = S_(A_(A_(M_(c0, c2), M_(c0, c5)), M_(c2, c5)), A_(A_(dd, ee), ff));
  const double k0 = S_(S_(A_(A_(M_(c5, dd), M_(c0, ee)), M_(c2, ff)), M_(M_(c0, c2), c5)), M_(M_(2.0, c3), de));
  const double p = S_(M_(m, m), M_(3.0, k1));
  const double q = S_(M_(m, S_(p, M_(1.5, k1))), M_(13.5, k0));
  double phi = M_(27.0, A_(M_(M_(M_(0.25, k1), k1), S_(p, k1)), M_(k0, A_(q, M_(6.75, k0)))));
#undef M_
#undef A_
#undef S_
  double sqrtp = sqrt(fabs(p));
  phi = third * atan2(sqrt(fabs(phi)), q);
  double sphi_, cphi_;
  sincos(phi, &sphi_, &cphi_);
  double cphi = sqrtp * cphi_, sphi = oneroot3 * sqrtp * sphi_;
  double e2 = third * (m - cphi);
  double e3 = e2 + sphi, e1 = e2 + cphi;
  e2 = e2 - sphi;
  // ascending order as the reference (polar.f:282-298), so the sums round identically
  { double x;
    if (e2 < e1) { x = e1; e1 = e2; e2 = x; }
    if (e3 < e1) { x = e1; e1 = e3; e3 = x; }
    if (e3 < e2) { x = e2; e2 = e3; e3 = x; } }
  const double lo = sqrt(e1), mid = sqrt(e2), hi = sqrt(e3);
  double iu = lo + mid + hi;
  double iiu = lo * mid + mid * hi + lo * hi;
  double iiiu = lo * mid * hi;
  double a2 = 1.0 / (iiiu * (iu * iiu - iiiu));
  double b2 = iu * iiu * iiu - iiiu * (iu * iu + iiu);
  double cq = -iiiu - iu * (iu * iu - 2.0 * iiu);
  double d2 = iu;
  double u0 = a2 * (b2 + cq * c0 + d2 * cc0);
  double u1 = a2 * (cq * c1 + d2 * cc1);
  double u2 = a2 * (b2 + cq * c2 + d2 * cc2);
  double u3 = a2 * (cq * c3 + d2 * cc3);
  double u4 = a2 * (cq * c4 + d2 * cc4);
  double u5 = a2 * (b2 + cq * c5 + d2 * cc5);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r[3 * i + 0] = f[3 * i] * u0 + f[3 * i + 1] * u1 + f[3 * i + 2] * u3;
    r[3 * i + 1] = f[3 * i] * u1 + f[3 * i + 1] * u2 + f[3 * i + 2] * u4;
    r[3 * i + 2] = f[3 * i] * u3 + f[3 * i + 1] * u4 + f[3 * i + 2] * u5;
  }
}

// Voigt order xx,yy,zz,xy,yz,xz.  sym6 -> full 3x3
CPF_DI void v6_to_m3(const double* v, double* m) {
  m[0] = v[0]; m[4] = v[1]; m[8] = v[2];
  m[1] = m[3] = v[3]; m[5] = m[7] = v[4]; m[2] = m[6] = v[5];
}
// d = R^T D R with engineering shears in and out (getrm1 opt 1 + qmply1)
CPF_DI void unrotate_strain(const double* R, const double* D6eng, double* d6eng) {
  double D[9], T[9], S[9];
  D[0] = D6eng[0]; D[4] = D6eng[1]; D[8] = D6eng[2];
  D[1] = D[3] = 0.5 * D6eng[3]; D[5] = D[7] = 0.5 * D6eng[4]; D[2] = D[6] = 0.5 * D6eng[5];
  m3_mul_tn(R, D, T);
  m3_mul(T, R, S);
  d6eng[0] = S[0]; d6eng[1] = S[4]; d6eng[2] = S[8];
  d6eng[3] = S[1] + S[3]; d6eng[4] = S[5] + S[7]; d6eng[5] = S[2] + S[6];
}

// Kinematics of one voxel (drive_eps_sig.f:203-265): Rh, R, Fh^-1, det Fh and the unrotated
// strain increment uddt (engineering shear).
CPF_DI void voxel_kinematics(const double* fn, const double* fn1, double* Rh, double* R,
                             double* fhinv, double* detFh, double* uddt) {
  double fh[9], df[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { fh[k] = 0.5 * (fn[k] + fn1[k]); df[k] = fn1[k] - fn[k]; }
  polar_R(fh, Rh);
  polar_R(fn1, R);
  *detFh = m3_inv(fh, fhinv);
  double Lm[9];
  m3_mul(df, fhinv, Lm);
  double D6[6] = {Lm[0], Lm[4], Lm[8], Lm[1] + Lm[3], Lm[5] + Lm[7], Lm[2] + Lm[6]};
  unrotate_strain(Rh, D6, uddt);
}

// First Piola-Kirchhoff stress and exact tangent A = dP/dF of one voxel.
// t6: unrotated Cauchy stress (Voigt), C: 6x6 [D] row-major.  Algebraically identical to
// cep2A_a (cep2A.f:86-284) but every rank-one structure of dR/dF, dRh/dF and dL/dF is
// contracted analytically, so the cost is O(81 * const) instead of four 81x9 loop nests.
template <class Cep>
CPF_DI void pk1_and_tangent(const double* fn, const double* fn1, const double* t6, const Cep& C,
                            double* P, double* A /*81, may alias nothing*/,
                            double* __restrict__ outA, int64_t strideA) {
  double Rh[9], R[9], fh[9], df[9], fhinv[9], finv[9], t[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { fh[k] = 0.5 * (fn[k] + fn1[k]); df[k] = fn1[k] - fn[k]; }
  polar_R(fh, Rh);
  polar_R(fn1, R);
  m3_inv(fh, fhinv);
  const double J = m3_inv(fn1, finv);
  v6_to_m3(t6, t);
  double U[9], Y[9], RY[9], RYR[9], sigma[9], Uinv[9], tUinv[9], Rt[9], RtRF[9], tmp[9];
  m3_mul_tn(R, fn1, U);
  double trU = U[0] + U[4] + U[8];
#pragma unroll
  for (int k = 0; k < 9; ++k) Y[k] = ((k % 4 == 0) ? trU : 0.0) - U[k];
  m3_mul(R, t, Rt);
  m3_mul_nt(Rt, R, sigma);
  m3_mul(finv, R, Uinv);
  m3_mul(t, Uinv, tUinv);
  m3_mul(R, Y, RY);
  m3_mul_nt(RY, R, RYR);
  const double y = 1.0 / m3_det(Y);
  m3_mul(Rt, Uinv, RtRF);
  // P = J sigma F^-T (cs2p)
  double SF[9];
  m3_mul_nt(sigma, finv, SF);
#pragma unroll
  for (int k = 0; k < 9; ++k) P[k] = J * SF[k];
  // half-step quantities
  double Uh[9], Yh[9], RYh[9], RYRh[9], G[9], Lm[9], D2[9], E[9], B1[9], B2[9], UU[9], VV[9];
  m3_mul_tn(Rh, fh, Uh);
  double trUh = Uh[0] + Uh[4] + Uh[8];
#pragma unroll
  for (int k = 0; k < 9; ++k) Yh[k] = ((k % 4 == 0) ? trUh : 0.0) - Uh[k];
  m3_mul(Rh, Yh, RYh);
  m3_mul_nt(RYh, Rh, RYRh);
  const double yh2 = 0.5 / m3_det(Yh);
#pragma unroll
  for (int k = 0; k < 9; ++k) G[k] = 0.5 * fhinv[k];
  m3_mul(df, G, Lm);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) D2[3 * i + j] = Lm[3 * i + j] + Lm[3 * j + i];
  m3_mul(D2, Rh, E);
  m3_mul_tn(RYRh, E, B1);
  m3_mul_tn(RYh, E, B2);
  m3_mul_tn(Lm, Rh, tmp);
#pragma unroll
  for (int k = 0; k < 9; ++k) UU[k] = Rh[k] - tmp[k];
  m3_mul(G, Rh, VV);
  // products for the dR/dF terms
  double A1[9], A2[9], A3[9], A4[9], A5[9], A6[9];
  m3_mul(Y, tUinv, A1);
  m3_mul(RY, tUinv, A2);
  m3_mul_nt(Rt, Y, A3);
  m3_mul(finv, RYR, A4);
  m3_mul_nt(Rt, RY, A5);
  m3_mul(finv, RY, A6);
  const double Jy = J * y;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      // dd(i,j) for this (k,l) = (p,q): symmetric, Voigt with doubled shears
      double dd[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
          double v = yh2 * (Yh[3 * l + i] * B1[3 * k + j] - RYh[3 * k + i] * B2[3 * l + j] +
                            Yh[3 * l + j] * B1[3 * k + i] - RYh[3 * k + j] * B2[3 * l + i]) +
                     UU[3 * k + i] * VV[3 * l + j] + VV[3 * l + i] * UU[3 * k + j];
          dd[3 * i + j] = v;
        }
      double dv[6] = {dd[0], dd[4], dd[8], 2.0 * dd[1], 2.0 * dd[5], 2.0 * dd[2]};
      double dt6[6];
#pragma unroll
      for (int v = 0; v < 6; ++v) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 6; ++w) s += C[6 * w + v] * dv[w];
        dt6[v] = s;
      }
      double dtm[9], T1[9], T2[9];
      v6_to_m3(dt6, dtm);
      m3_mul(R, dtm, T1);
      m3_mul(T1, Uinv, T2);
      const double dJ = J * finv[3 * l + k];  // dJdF(k,l) = J finv(l,k)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double v = RtRF[3 * i + j] * dJ - J * SF[3 * i + l] * finv[3 * j + k] +
                     Jy * (RYR[3 * i + k] * A1[3 * l + j] - RY[3 * i + l] * A2[3 * k + j]) +
                     J * T2[3 * i + j] +
                     Jy * (A3[3 * i + l] * A4[3 * j + k] - A5[3 * i + k] * A6[3 * j + l]);
          const int m = 27 * i + 9 * j + 3 * k + l;
          if (outA) outA[(int64_t)m * strideA] = v;
          if (A) A[m] = v;
        }
    }
  }
}
