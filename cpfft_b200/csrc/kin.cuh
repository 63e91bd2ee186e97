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

// Rotation factor R of the polar decomposition F = R U (rtcmp1 / irscp1 / ivcmp1 / evcmp1_new,
// polar.f:18-307).
//
// The reference evaluates a closed form: trigonometric eigenvalues of C = F^T F, invariants of U,
// U^-1 = a (b I + c C + d C^2), R = F U^-1.  In double precision its discriminant cancels to
// round-off when the principal stretches differ by less than a few 1e-3 -- every small load
// increment -- and R comes out with a noise of order strain^3 (up to ~3e-8) that depends on the
// last bit of F and of every intermediate: two correct implementations of that closed form, or
// the same one fed an F that differs by 1e-13, disagree by that much in every stress.  What the
// formula defines (its value in exact arithmetic; oracle/oracle_kin.cpp evaluates it in
// __float128) is the orthogonal polar factor, and that is what is computed here, to ~1e-15, by the
// Newton iteration R <- (R + R^-T) / 2 started from F (Higham): the singular values s of the
// iterate map to (s + 1/s) / 2, so the error squares every trip -- a stretch of 1.5 (50 %
// strain) reaches 1e-16 in 6 trips, the 1.05 of a finished load path in 4; POLAR_ITERS = 8 fixed
// trips, no branch, ~45 flops each.  The result is within the reference's own noise band of
// the reference's result and reproducible to round-off, which lets the parity tests hold
// 1e-9 per voxel at 0.1 % increments (tests/test_oracle_material.py shows both facts).
#define POLAR_ITERS 8
CPF_DI void polar_R(const double* f, double* r) {
  double a[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) a[k] = f[k];
#pragma unroll 1
  for (int it = 0; it < POLAR_ITERS; ++it) {
    // cofactors of a: a^-T = cof(a) / det(a)
    const double c0 = a[4] * a[8] - a[5] * a[7], c1 = a[5] * a[6] - a[3] * a[8], c2 = a[3] * a[7] - a[4] * a[6];
    const double c3 = a[2] * a[7] - a[1] * a[8], c4 = a[0] * a[8] - a[2] * a[6], c5 = a[1] * a[6] - a[0] * a[7];
    const double c6 = a[1] * a[5] - a[2] * a[4], c7 = a[2] * a[3] - a[0] * a[5], c8 = a[0] * a[4] - a[1] * a[3];
    const double hid = 0.5 / (a[0] * c0 + a[1] * c1 + a[2] * c2);
    a[0] = 0.5 * a[0] + hid * c0; a[1] = 0.5 * a[1] + hid * c1; a[2] = 0.5 * a[2] + hid * c2;
    a[3] = 0.5 * a[3] + hid * c3; a[4] = 0.5 * a[4] + hid * c4; a[5] = 0.5 * a[5] + hid * c5;
    a[6] = 0.5 * a[6] + hid * c6; a[7] = 0.5 * a[7] + hid * c7; a[8] = 0.5 * a[8] + hid * c8;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) r[k] = a[k];
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
//
// Register budget.  Written as one pass the 81 outputs keep ~200 doubles live (twice what a thread has): the first
// version spilled 1200 B per thread and moved 32.5 GB of DRAM traffic per 256^3 launch for 20.1 GB of algorithmic
// bytes.  Now two phases with disjoint operand sets: phase 1 forms the geometric part of all 81 entries (dJ/dF,
// dF^-T/dF, dR/dF terms: RtRF, SF, F^-1, RYR, RY, A1..A6) into the scratch tile `S`; phase 2 forms the material part
// J (R dt Uinv), dt = [D] : dd/dF (Yh, B1, RYh, B2, UU, VV, [D], R, Uinv), adds and stores.  `S` is 81 doubles per
// thread in shared memory in the kernel (element m of thread t at S.p[m * PK1_THREADS + t]), a plain array on the host.
#ifndef PK1_THREADS
#define PK1_THREADS 128
#endif
struct Pk1Scratch {
  double* p;
#ifdef __CUDACC__
  CPF_DI double& operator[](int k) const { return p[k * PK1_THREADS]; }
#else
  CPF_DI double& operator[](int k) const { return p[k]; }
#endif
};
// C: [D] as stored, entry 6 * row + col at C[(6 * row + col) * strideC]
CPF_DI void pk1_and_tangent(const double* fn, const double* fn1, const double* t6, const double* __restrict__ C, int64_t strideC,
                            double* P, double* A /*81, may alias nothing*/,
                            double* __restrict__ outA, int64_t strideA, const Pk1Scratch S) {
  double R[9], finv[9], t[9];
  polar_R(fn1, R);
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
  // ---- phase 1: geometric part of the 81 entries -> S ----
  {
    double A1[9], A2[9], A3[9], A4[9], A5[9], A6[9];
    m3_mul(Y, tUinv, A1);
    m3_mul(RY, tUinv, A2);
    m3_mul_nt(Rt, Y, A3);
    m3_mul(finv, RYR, A4);
    m3_mul_nt(Rt, RY, A5);
    m3_mul(finv, RY, A6);
    const double Jy = J * y;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double dJ = J * finv[3 * l + k];  // dJdF(k,l) = J finv(l,k)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            S[27 * i + 9 * j + 3 * k + l] = RtRF[3 * i + j] * dJ - J * SF[3 * i + l] * finv[3 * j + k] +
                                            Jy * (RYR[3 * i + k] * A1[3 * l + j] - RY[3 * i + l] * A2[3 * k + j]) +
                                            Jy * (A3[3 * i + l] * A4[3 * j + k] - A5[3 * i + k] * A6[3 * j + l]);
      }
  }
  // ---- phase 2: material part J (R dt Uinv), dt = [D] : d(d)/dF ----
  // half-step quantities (formed only now: their 108 doubles must not be live during phase 1)
  double Rh[9], fh[9], df[9], fhinv[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { fh[k] = 0.5 * (fn[k] + fn1[k]); df[k] = fn1[k] - fn[k]; }
  polar_R(fh, Rh);
  m3_inv(fh, fhinv);
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
  double Cr[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) Cr[q] = CPF_LDG(C + (int64_t)q * strideC);
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
        for (int w = 0; w < 6; ++w) s += Cr[6 * w + v] * dv[w];
        dt6[v] = s;
      }
      double dtm[9], T1[9], T2[9];
      v6_to_m3(dt6, dtm);
      m3_mul(R, dtm, T1);
      m3_mul(T1, Uinv, T2);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int m = 27 * i + 9 * j + 3 * k + l;
          const double v = S[m] + J * T2[3 * i + j];
          if (outA) outA[(int64_t)m * strideA] = v;
          if (A) A[m] = v;
        }
    }
  }
}
