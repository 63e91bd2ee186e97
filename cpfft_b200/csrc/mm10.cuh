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
//  * |rs/tt|^(n-1) is evaluated ONCE per system per point (the reference calls mm10_slipinc up
//    to three times per evaluation): by straight-line square-and-multiply when n-1 is a small
//    integer, in the residual, which leaves the values in shared memory for the Jacobian;
//  * the residual's slip loop runs in the lattice frame (LF, the default): the stress is mapped
//    once (rs_s = (L^T sigma) . ms0_s), the plastic strain / spin sums are formed on (ms0, qs0)
//    and mapped back once -- both maps are linear, so this is the same algebra in another
//    summation order (equal to round-off);
//  * the 7x7 systems are factorised in shared memory (mm10_lu7_factor / mm10_lu7_solve); the
//    tangent reuses the factors of the last Newton step instead of forming and factorising the
//    Schur complement six times (equal to round-off).
// Control flow that decides iteration counts (tolerances, Armijo test, at least one update
// iteration, lagged Jacobian for the tangent, sub-stepping) follows the reference exactly.
//
// Code layout.  The first version of this kernel was instruction-fetch bound (17 k SASS
// instructions, instruction-cache hit rate 56 %, "no instruction" the top stall reason in ncu).
// Now the stress predictor (6 unknowns) and the coupled update (7 unknowns) of mm10_solve run
// through ONE Newton state machine with one residual site, one Jacobian site and one LU site;
// the predictor pads its system with an identity row / column, which leaves the arithmetic of
// the first six unknowns bit-identical.  Libm functions with long inline expansions sit behind
// non-inlined wrappers.  Round 2 (ncu source page, profiles/r02r_update_hotspots.md): the loop
// is still ~2.6 k instructions (42 KB > the 32 KB L1.5 I-cache), the kernel is bound by issue
// latency at 2 warps per scheduler (255 registers); the changes listed above removed a quarter
// of the executed instructions, the grain-table loads are requested one slip system ahead.
#pragma once
#include "kin.cuh"
#include "material_types.h"

#ifndef MM10_THREADS
#define MM10_THREADS 128
#endif
// unroll factor of the slip-system loops of the residual and the Jacobian (development knob, tools/build_variants.py)
#ifndef MM10_SLIP_UNROLL
#define MM10_SLIP_UNROLL 1
#endif
#ifndef MM10_RESID_UNROLL
#define MM10_RESID_UNROLL MM10_SLIP_UNROLL
#endif
#ifndef MM10_JAC_UNROLL
#define MM10_JAC_UNROLL MM10_SLIP_UNROLL
#endif
#define MM10_PRAGMA_(x) _Pragma(#x)
#define MM10_UNROLL_SLIP_(n) MM10_PRAGMA_(unroll n)
#define MM10_UNROLL_RESID MM10_UNROLL_SLIP_(MM10_RESID_UNROLL)
#define MM10_UNROLL_JAC MM10_UNROLL_SLIP_(MM10_JAC_UNROLL)
// MM10_PREFETCH: the nine grain-table entries of slip system s + 1 are loaded while system s is
// being processed (the loops are not unrolled, so the compiler cannot overlap the load latency
// of one trip with the arithmetic of the previous one by itself).  1: into a second buffer, copied
// at the top of the trip; 2 (default): in the Jacobian, whose geometry step is the only consumer of
// the entry, right after that step into the same registers (no copy).  0: plain loads.
#ifndef MM10_PREFETCH
#define MM10_PREFETCH 2
#endif
// the lattice-frame residual loop is short (per system 9 loads, a 6-term dot product, the power, 9 FMA):
// unrolled, the loads of several systems are in flight together
#ifndef MM10_LF_UNROLL
#define MM10_LF_UNROLL 1
#endif
#define MM10_UNROLL_LF MM10_UNROLL_SLIP_(MM10_LF_UNROLL)

// Per-thread array kept in shared memory, element k of thread t at p[k * blockDim + t]:
// conflict-free, and it takes the Jacobian and the skew-rotation operators out of the
// register budget of the Newton loop.
struct SArr {
  double* p;
  CPF_DI double& operator[](int k) const { return p[k * MM10_THREADS]; }
};
#define MM10_SM_J 0      // 49: Jacobian of the last Newton step (lagged for the tangent)
#define MM10_SM_RWQ 49   // 9 : RW(Rp_n^T)
#define MM10_SM_RWR 58   // 9 : RW(R)
#define MM10_SM_ACC 67   // 39: S (21) and T (18) slip sums of the Jacobian
#define MM10_SMEM_DOUBLES 106
// |rs/tt|^(n-1) of the first MM10_PCACHE systems, left in the `acc` slots by the residual for the
// Jacobian that follows it: mm10_solve always forms the Jacobian at the point of its last
// residual evaluation (the start point or the accepted line-search point), and `acc` is dead
// between the two.  Saves half of the power evaluations; the values are the ones the Jacobian
// would recompute, bit for bit.
#define MM10_PCACHE 39
// MTS kernels: acc[MM10_SM_STASH .. +7) keeps column 7 of the last Jacobian (J12, J22) next to its LU factors --
// their tangent needs the raw entries (mm10_a.f:760-805) -- so the hand-over uses the slots below it
#define MM10_SM_STASH 32
#define MM10_PC(HARD) ((HARD) == MM10_MTS ? MM10_SM_STASH : MM10_PCACHE)

CPF_DNOINLINE double cpf_pow(double x, double y) { return pow(x, y); }
CPF_DNOINLINE double cpf_atan2(double y, double x) { return atan2(y, x); }

CPF_DI double cpf_sgn(double x) { return x >= 0.0 ? 1.0 : -1.0; }  // Fortran sign(one,x)

// x^ie for 0 <= ie <= 64 (x >= 0), else pow(x, fe).  Square-and-multiply from the low bit, written
// without a loop: the chain of squares x^2 .. x^16 is straight-line code and every bit of the
// exponent costs one predicated multiply.  The products are formed in the order of the
// `while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }` loop this replaces (bit-identical results);
// that loop was 13 instructions per bit behind a divergence-safe branch, 15 % of all
// instructions the kernel executed (ncu source page, profiles/r02r_update_hotspots.md).
CPF_DI double cpf_pow_abs(double x, int ie, double fe) {
  if (ie < 0) return cpf_pow(x, fe);
  double r = (ie & 1) ? x : 1.0;
  double b = x * x;
  r = (ie & 2) ? r * b : r;
  b = b * b;
  r = (ie & 4) ? r * b : r;
  b = b * b;
  r = (ie & 8) ? r * b : r;
  b = b * b;
  r = (ie & 16) ? r * b : r;
  if (ie >= 32) {
    b = b * b;
    r = (ie & 32) ? r * b : r;
    b = b * b;
    r = (ie & 64) ? r * b : r;
  }
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

// ---- the 7x7 LU in shared memory ----------------------------------------------------------------
// Dense solve with partial pivoting, the stand-in for DGESV (mm10_solve, mm10_tangent, the lattice strain).
// mm10_lu7_factor: in-place factorisation of the matrix in J in DGETRF's order -- whole-row
// interchanges, unit-lower multipliers below the diagonal -- and the RECIPROCAL of the pivot on the
// diagonal.  Shared memory can be indexed at run time, registers cannot: the row interchange is two
// indexed rows behind a (rare, 1 % per column) branch.  The first version of this kernel kept the
// matrix in 98 registers and interchanged rows by ~530 predicated selects per solve (an `if (i ==
// piv) swap` chain is turned into run-time indexed accesses by the optimiser, which demotes the
// whole matrix to local memory): 17 % of all executed instructions, and the tangent repeated the
// factorisation for each of its six right-hand sides.  Returns the pivot rows, 3 bits each.
// mm10_lu7_solve: b <- A^-1 b from those factors (row interchanges, forward, back substitution).
// The arithmetic on every entry is Gaussian elimination's, operation for operation (l = a_ik *
// (1/a_kk), a_ij -= l a_kj, b_i -= l b_k, x_k = (b_k - sum a_kc x_c) * (1/a_kk)), and (-A)^-1 b ==
// -(A^-1 b) exactly, so the Newton step solves with J and negates.
// (Keeping the row loops rolled takes ~400 instructions out of the Newton loop's instruction-cache
// footprint but measured 3 % slower, profiles/r02u_mm10ab_fp64lat.log.)
CPF_DI int mm10_lu7_factor(SArr J) {
  int pack = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    int piv = k;
    double best = fabs(J[7 * k + k]);
#pragma unroll
    for (int i = k + 1; i < 7; ++i) {
      const double v = fabs(J[7 * i + k]);
      if (v > best) { best = v; piv = i; }
    }
    pack |= piv << (3 * k);
    if (piv != k) {
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const double t = J[7 * k + j];
        J[7 * k + j] = J[7 * piv + j];
        J[7 * piv + j] = t;
      }
    }
    const double inv = 1.0 / J[7 * k + k];
    J[7 * k + k] = inv;
    double u[7];
#pragma unroll
    for (int j = k + 1; j < 7; ++j) u[j] = J[7 * k + j];
#pragma unroll
    for (int i = k + 1; i < 7; ++i) {
      const double l = J[7 * i + k] * inv;
      J[7 * i + k] = l;
#pragma unroll
      for (int j = k + 1; j < 7; ++j) J[7 * i + j] -= l * u[j];
    }
  }
  return pack;
}
CPF_DI void mm10_lu7_solve(SArr J, int pack, double* b) {
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int piv = (pack >> (3 * k)) & 7;
    const double top = b[k];
    double pv = top;
#pragma unroll
    for (int i = k + 1; i < 7; ++i) {
      const double cur = b[i];
      pv = (piv == i) ? cur : pv;
      b[i] = (piv == i) ? top : cur;
    }
    b[k] = pv;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int i = k + 1; i < 7; ++i) b[i] -= J[7 * i + k] * b[k];
#pragma unroll
  for (int k = 6; k >= 0; --k) {
    double sum = b[k];
#pragma unroll
    for (int c2 = k + 1; c2 < 7; ++c2) sum -= J[7 * k + c2] * b[c2];
    b[k] = sum * J[7 * k + k];
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
  SArr J;          // 7x7 Jacobian (shared memory)
  SArr acc;        // 39 Jacobian slip sums (shared memory)
  double sn[6];    // stress at n
  double ttn;      // tau_tilde at n
  double D[6];     // strain increment of the (sub)step
  double dg, tinc, taul;
  // MTS hardening (HARD == 2) only, set by mts_substep(): ur = mu/mu_0, tau_a, and the part of
  // the hardening target that does not depend on the unknowns,
  //   h0 = tau_a (1 - mu/mu_n) + ur (tau_y - tau_y_n) + (mu/mu_n) tau_tilde_n   (mm10_b.f:2103-2108);
  // tau_y / tau_v above then hold the thresholds of the (sub)step (mm10_setup_mts)
  double ur, tau_a, h0;
};
#define MM10_VOCE 1
#define MM10_MTS 2

// Current Schmid vectors of system s.  The reference forms ms = RT2RVE(Rp_n^T) ms0 and
// qs = RT2RVW(Rp_n^T) qs0 (mm10_a.f:867-876).  RT2RVE is the stress-type 6x6 operator, so in
// tensor form ms = V6( Q M~ Q^T ) with M~ the symmetric tensor whose Voigt vector (no shear
// doubling) is ms0, Q = Rp_n^T; that is what is evaluated here (45 FMA, no 6x6 operator).
CPF_DI void mm10_slip_load(const Mm10Ctx& c, int s, double* g) {
  const double* t = c.ms0 + CPF_SLIP_STRIDE * s;
#if CPF_TAB_VEC
  double pad;
  CPF_LDG2(t, g[0], g[1]); CPF_LDG2(t + 2, g[2], g[3]); CPF_LDG2(t + 4, g[4], g[5]); CPF_LDG2(t + 6, g[6], g[7]);
  CPF_LDG2(t + 8, g[8], pad);
  (void)pad;
#else
#pragma unroll
  for (int k = 0; k < 9; ++k) g[k] = CPF_LDG(t + k);
#endif
}
// row i of the grain's rotated stiffness
CPF_DI void mm10_c_row(const Mm10Ctx& c, int i, double* r) {
  const double* t = c.C + 6 * i;
#if CPF_TAB_VEC
  CPF_LDG2(t, r[0], r[1]); CPF_LDG2(t + 2, r[2], r[3]); CPF_LDG2(t + 4, r[4], r[5]);
#else
#pragma unroll
  for (int k = 0; k < 6; ++k) r[k] = CPF_LDG(t + k);
#endif
}
CPF_DI void mm10_slip_geom_v(const Mm10Ctx& c, const double* g, double* ms, double* qs);
CPF_DI void mm10_slip_geom(const Mm10Ctx& c, int s, double* ms, double* qs) {
  double g[9];
  mm10_slip_load(c, s, g);
  mm10_slip_geom_v(c, g, ms, qs);
}
CPF_DI void mm10_slip_geom_v(const Mm10Ctx& c, const double* g, double* ms, double* qs) {
  const double m0 = g[0], m1 = g[1], m2 = g[2], m3 = g[3], m4 = g[4], m5 = g[5];
  const double w0 = g[6], w1 = g[7], w2 = g[8];
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

// The map of mm10_slip_geom, ms = L ms0 with L = RT2RVE(Rp_n^T) (6x6, linear), applied to whole sums instead of
// per system.  mm10_to_lattice: y = L^T sig, so that rs = sig . ms = y . ms0 (Sh = stress tensor with halved
// shears, Y = Q^T Sh Q, shear entries of y doubled).  mm10_from_lattice: out = L a = V6(Q A~ Q^T), A~ the
// symmetric tensor whose Voigt vector (no shear doubling) is a.
CPF_DI void mm10_to_lattice(const Mm10Ctx& c, const double* sig, double* y) {
  const double h3 = 0.5 * sig[3], h4 = 0.5 * sig[4], h5 = 0.5 * sig[5];
  double U[9];   // U = Sh Q
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const double q0 = c.Q[b], q1 = c.Q[3 + b], q2 = c.Q[6 + b];
    U[b] = sig[0] * q0 + h3 * q1 + h5 * q2;
    U[3 + b] = h3 * q0 + sig[1] * q1 + h4 * q2;
    U[6 + b] = h5 * q0 + h4 * q1 + sig[2] * q2;
  }
  // Y = Q^T U, Voigt with doubled shears
  y[0] = c.Q[0] * U[0] + c.Q[3] * U[3] + c.Q[6] * U[6];
  y[1] = c.Q[1] * U[1] + c.Q[4] * U[4] + c.Q[7] * U[7];
  y[2] = c.Q[2] * U[2] + c.Q[5] * U[5] + c.Q[8] * U[8];
  y[3] = 2.0 * (c.Q[0] * U[1] + c.Q[3] * U[4] + c.Q[6] * U[7]);
  y[4] = 2.0 * (c.Q[1] * U[2] + c.Q[4] * U[5] + c.Q[7] * U[8]);
  y[5] = 2.0 * (c.Q[0] * U[2] + c.Q[3] * U[5] + c.Q[6] * U[8]);
}
CPF_DI void mm10_from_lattice(const Mm10Ctx& c, const double* am, double* out) {
  double T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a = c.Q[3 * i], b = c.Q[3 * i + 1], d = c.Q[3 * i + 2];
    T[3 * i + 0] = a * am[0] + b * am[3] + d * am[5];
    T[3 * i + 1] = a * am[3] + b * am[1] + d * am[4];
    T[3 * i + 2] = a * am[5] + b * am[4] + d * am[2];
  }
  out[0] = T[0] * c.Q[0] + T[1] * c.Q[1] + T[2] * c.Q[2];
  out[1] = T[3] * c.Q[3] + T[4] * c.Q[4] + T[5] * c.Q[5];
  out[2] = T[6] * c.Q[6] + T[7] * c.Q[7] + T[8] * c.Q[8];
  out[3] = T[0] * c.Q[3] + T[1] * c.Q[4] + T[2] * c.Q[5];
  out[4] = T[3] * c.Q[6] + T[4] * c.Q[7] + T[5] * c.Q[8];
  out[5] = T[0] * c.Q[6] + T[1] * c.Q[7] + T[2] * c.Q[8];
}

template <int HARD>
CPF_DI double mm10_hfac(const Mm10Ctx& c, double tt, double* hterm_out) {
  if (HARD == MM10_MTS) {
    // ct = 1 - cta / tau_v, cta = (mu_0/mu) (tt - tau_a) - tau_y; the reference raises the signed
    // value to the power voche_m (mm10_b.f:2093-2101); tau_l = 0
    const double cta = (tt - c.tau_a) / c.ur - c.tau_y;
    const double ct = 1.0 - cta / c.tau_v;
    *hterm_out = ct;
    return (c.voche_m == 1.0) ? ct : cpf_pow(ct, c.voche_m);
  }
  const double hterm = 1.0 - (tt - c.tau_y) / c.tau_v + c.taul / (tt - c.tau_y);
  *hterm_out = hterm;
  const double ah = fabs(hterm);
  const double pw = (c.voche_m == 1.0) ? ah : cpf_pow(ah, c.voche_m);
  return pw * cpf_sgn(hterm);
}

// Residual (mm10_formR / formR1 / formR2).  R[0..5] = R1, R[6] = R2 (0 unless want2); returns
// the hardening target h (np1%tt_rate = (h - tt_n)/tinc).
//
// LF = true (development variant, CPFFT_MM10_LF=1, Voce single crystal only): the slip loop runs
// in the lattice frame.  With Sh the stress tensor with halved shears, rs_s = Sh : (Q M~_s Q^T) =
// (Q^T Sh Q) : M~_s, so the stress is rotated once and each system costs a 6-term dot product;
// the plastic strain / spin sums are accumulated on ms0 / qs0 and rotated back once (both maps
// are linear).  Same algebra, different summation order: results agree to round-off, not bit
// for bit, which is why this is not the default before it has been through the GPU suite.
template <int HARD, bool LF = false>
CPF_DI double mm10_resid(const Mm10Ctx& c, const double* sig, double tt, double* R, bool want2) {
  double dbarp[6] = {0, 0, 0, 0, 0, 0}, wq[3] = {0, 0, 0}, sabs = 0.0;
  const double itt = 1.0 / tt, dgtt = c.dg / tt, dif = c.tinc * c.iD_v;
  if (LF) {
    double y[6];
    mm10_to_lattice(c, sig, y);
    double am[6] = {0, 0, 0, 0, 0, 0}, aw[3] = {0, 0, 0};
#if MM10_PREFETCH
    double gn[9];
    mm10_slip_load(c, 0, gn);
#endif
MM10_UNROLL_LF
    for (int s = 0; s < c.nslip; ++s) {
      double m[6], w[3];
#if MM10_PREFETCH
#pragma unroll
      for (int k = 0; k < 6; ++k) m[k] = gn[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) w[k] = gn[6 + k];
      mm10_slip_load(c, (s + 1 < c.nslip) ? s + 1 : s, gn);
#else
      double g9[9];
      mm10_slip_load(c, s, g9);
#pragma unroll
      for (int k = 0; k < 6; ++k) m[k] = g9[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) w[k] = g9[6 + k];
#endif
      const double rs = y[0] * m[0] + y[1] * m[1] + y[2] * m[2] + y[3] * m[3] + y[4] * m[4] + y[5] * m[5];
      const double p = cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0);
      if (s < MM10_PC(HARD)) c.acc[s] = p;
      const double slip = dgtt * p * rs;
      const double f = rs * dif + slip;
#pragma unroll
      for (int k = 0; k < 6; ++k) am[k] += f * m[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) aw[k] += f * w[k];
      sabs += fabs(slip);
    }
    // dbarp = V6(Q A~ Q^T), wq = RWQ aw  (the maps of mm10_slip_geom applied to the sums)
    mm10_from_lattice(c, am, dbarp);
    cpf_mv3(c.RWQ, aw, wq);
  } else
  {
#if MM10_PREFETCH
  double gn[9];
  mm10_slip_load(c, 0, gn);
#endif
MM10_UNROLL_RESID
  for (int s = 0; s < c.nslip; ++s) {
    double ms[6], qs[3];
#if MM10_PREFETCH
    double g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = gn[k];
    mm10_slip_load(c, (s + 1 < c.nslip) ? s + 1 : s, gn);
    mm10_slip_geom_v(c, g, ms, qs);
#else
    mm10_slip_geom(c, s, ms, qs);
#endif
    const double rs = sig[0] * ms[0] + sig[1] * ms[1] + sig[2] * ms[2] + sig[3] * ms[3] + sig[4] * ms[4] + sig[5] * ms[5];
    const double p = cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0);
    if (s < MM10_PC(HARD)) c.acc[s] = p;
    const double slip = dgtt * p * rs;
    const double f = rs * dif + slip;
#pragma unroll
    for (int k = 0; k < 6; ++k) dbarp[k] += f * ms[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) wq[k] += f * qs[k];
    sabs += fabs(slip);
  }
  }
  double wp[3], sw[6], w1[6];
  cpf_mv3(c.RWR, wq, wp);
  cpf_symsw(sig, wp, sw);
#pragma unroll
  for (int k = 0; k < 6; ++k) w1[k] = c.D[k] - dbarp[k];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = 0.0, cr[6];
    mm10_c_row(c, i, cr);
#pragma unroll
    for (int j = 0; j < 6; ++j) s += cr[j] * w1[j];
    R[i] = sig[i] - c.sn[i] - s + 2.0 * sw[i];
  }
  double h = 0.0;
  R[6] = 0.0;
  if (want2) {
    double ht;
    const double hf = mm10_hfac<HARD>(c, tt, &ht);
    if (HARD == MM10_MTS) h = c.h0 + c.theta_0 * c.ur * (hf * sabs);    // mm10_h_mts
    else h = c.ttn + c.theta_0 * (hf * sabs);
    R[6] = tt - h;
  }
  return h;
}

// Jacobian (mm10_formJ) into c.J (7x7 row-major, shared memory).  full = false: J11 only
// (stress predictor), padded with an identity row / column.
// (Forming the slip sums in the lattice frame, as the residual does, and mapping S, T and the four vectors back
// costs what it saves for 12 systems and measured 7 % slower for 48, profiles/r02u_mm10ab_fp64lat.log: not kept.)
template <int HARD>
CPF_DI void mm10_jacobian(const Mm10Ctx& c, const double* sig, double tt, bool full) {
  double dps[6], wqs[3], wqf[3], es[6], sabs = 0.0, ssum = 0.0;
  {
    double S[21], T[18];
#pragma unroll
    for (int k = 0; k < 21; ++k) S[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 18; ++k) T[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { dps[k] = 0.0; es[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { wqs[k] = 0.0; wqf[k] = 0.0; }
    const double itt = 1.0 / tt, dgtt = c.dg / tt, dif = c.tinc * c.iD_v, dgn = c.dg * c.rate_n / tt;
#if MM10_PREFETCH
    double gn[9];
    mm10_slip_load(c, 0, gn);
#endif
MM10_UNROLL_JAC
    for (int s = 0; s < c.nslip; ++s) {
      double ms[6], qs[3];
#if MM10_PREFETCH == 2
      // the geometry is the only consumer of the table entry: the next one is requested right after it, into the
      // same registers, and lands under the ~110 instructions of the slip sums
      mm10_slip_geom_v(c, gn, ms, qs);
      mm10_slip_load(c, (s + 1 < c.nslip) ? s + 1 : s, gn);
#elif MM10_PREFETCH
      double g[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) g[k] = gn[k];
      mm10_slip_load(c, (s + 1 < c.nslip) ? s + 1 : s, gn);
      mm10_slip_geom_v(c, g, ms, qs);
#else
      mm10_slip_geom(c, s, ms, qs);
#endif
      const double rs = sig[0] * ms[0] + sig[1] * ms[1] + sig[2] * ms[2] + sig[3] * ms[3] + sig[4] * ms[4] + sig[5] * ms[5];
      const double p = (s < MM10_PC(HARD)) ? c.acc[s] : cpf_pow_abs(fabs(rs * itt), c.rate_int, c.rate_n - 1.0);
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
        wqs[k] += slip * qs[k];
      }
      const double sp = cpf_sgn(rs) * p;
#pragma unroll
      for (int k = 0; k < 6; ++k) { dps[k] += slip * ms[k]; es[k] += sp * ms[k]; }
      sabs += fabs(slip); ssum += slip;
    }
    // park the sums in shared memory so that the column loop below can index them at run
    // time: acc[0..20] = S packed upper triangle by rows, acc[21 + 6 k + b] = T(k, b)
#pragma unroll
    for (int k = 0; k < 21; ++k) c.acc[k] = S[k];
#pragma unroll
    for (int k = 0; k < 18; ++k) c.acc[21 + k] = T[k];
  }
  // J11 = C S + 2 Lsig (RWR T) + IW(wp) + I, one column per trip
  double rwr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) rwr[k] = c.RWR[k];
#pragma unroll 1
  for (int b = 0; b < 6; ++b) {
    double scol[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const int lo = a < b ? a : b, hi = a < b ? b : a;   // S(a,b) = packed[(lo,hi)]
      scol[a] = c.acc[lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo)];
    }
    const double tcol[3] = {c.acc[21 + b], c.acc[27 + b], c.acc[33 + b]};
    double tc[3], sw[6];
    cpf_mv3(rwr, tcol, tc);
    cpf_symsw(sig, tc, sw);
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double s = 2.0 * sw[a], cr[6];
      mm10_c_row(c, a, cr);
#pragma unroll
      for (int k = 0; k < 6; ++k) s += cr[k] * scol[k];
      c.J[7 * a + b] = s;
    }
  }
  {
    double w[3];
    cpf_mv3(rwr, wqf, w);
    c.J[7 * 0 + 3] += 2.0 * w[2]; c.J[7 * 0 + 5] += -2.0 * w[1];
    c.J[7 * 1 + 3] += 2.0 * w[2]; c.J[7 * 1 + 4] += -2.0 * w[0];
    c.J[7 * 2 + 4] += 2.0 * w[0]; c.J[7 * 2 + 5] += 2.0 * w[1];
    c.J[7 * 3 + 0] += w[2]; c.J[7 * 3 + 1] += -w[2]; c.J[7 * 3 + 4] += -w[1]; c.J[7 * 3 + 5] += w[0];
    c.J[7 * 4 + 1] += w[0]; c.J[7 * 4 + 2] += -w[0]; c.J[7 * 4 + 3] += w[1]; c.J[7 * 4 + 5] += w[2];
    c.J[7 * 5 + 0] += w[1]; c.J[7 * 5 + 2] += -w[1]; c.J[7 * 5 + 3] += w[0]; c.J[7 * 5 + 4] += -w[2];
#pragma unroll
    for (int a = 0; a < 6; ++a) c.J[7 * a + a] += 1.0;
  }
  if (full) {
    // J12 = -(n/tt) [C dps + 2 symSW(sig, RWR wqs)]
    double wc[3], sw[6];
    cpf_mv3(rwr, wqs, wc);
    cpf_symsw(sig, wc, sw);
    const double nt = -c.rate_n / tt;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double s = 2.0 * sw[a], cr[6];
      mm10_c_row(c, a, cr);
#pragma unroll
      for (int k = 0; k < 6; ++k) s += cr[k] * dps[k];
      c.J[7 * a + 6] = nt * s;
    }
    // J21 = -theta0 dg n / tt * hfac * sum sgn(rs) |rs/tt|^(n-1) ms
    double ht;
    const double hf = mm10_hfac<HARD>(c, tt, &ht);
    const double dgn = c.dg * c.rate_n / tt;
    if (HARD == MM10_MTS) {
      // mm10_estress_mts / mm10_ehard_mts (mm10_b.f:2114-2186), tau_l = 0
      const double fac = c.theta_0 * c.ur * dgn * hf;
#pragma unroll
      for (int k = 0; k < 6; ++k) c.J[42 + k] = -(fac * es[k]);
      c.J[48] = 1.0 + c.theta_0 * ((c.voche_m / (c.tau_v * ht) + c.ur * c.rate_n / tt) * hf * sabs);
    } else {
    const double fac = c.theta_0 * dgn * hf;
#pragma unroll
    for (int k = 0; k < 6; ++k) c.J[42 + k] = -(fac * es[k]);
    // J22 (mm10_ehard_voche)
    const double ah = fabs(ht);
    const double pw = (c.voche_m == 1.0) ? ah : cpf_pow(ah, c.voche_m);
    const double A = -1.0 / c.tau_v - c.taul / ((tt - c.tau_y) * (tt - c.tau_y));
    const double etau = (c.voche_m * A * sabs / ah - ssum * c.rate_n / tt * cpf_sgn(ht)) * pw;
    c.J[48] = 1.0 - c.theta_0 * etau;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) { c.J[7 * k + 6] = 0.0; c.J[42 + k] = 0.0; }
    c.J[48] = 1.0;
  }
}

// MTS (mm10_setup_mts, mm10_a.f:2109-2175).  Shear modulus and activation constants at
// temperature T, then the thresholds of a (sub)step from its strain rate dgc = dg / tinc.
struct Mm10Mts {
  double tau_hat_y, tau_hat_v, ky, kv, iq_y, ip_y, iq_v, ip_v, eps_dot_0_y, eps_dot_0_v;
};
CPF_DI void mts_at_temperature(const CpfCryDev& cr, double T, double* mu, Mm10Mts* m) {
  *mu = (T == 0.0) ? cr.mu_0 : cr.mu_0 - cr.D_0 / (exp(cr.T_0 / T) - 1.0);
  m->ky = cr.kby * T / *mu;         // boltz T / (mu b^3 G_0_y)
  m->kv = cr.kbv * T / *mu;
}
CPF_DI void mts_thresholds(const Mm10Mts& m, double dgc, double* tau_y, double* tau_v) {
  if (dgc == 0.0) { *tau_v = m.tau_hat_v; *tau_y = m.tau_hat_y; return; }
  *tau_v = m.tau_hat_v * cpf_pow(1.0 - cpf_pow(m.kv * log(m.eps_dot_0_v / dgc), m.iq_v), m.ip_v);
  *tau_y = m.tau_hat_y * cpf_pow(1.0 - cpf_pow(m.ky * log(m.eps_dot_0_y / dgc), m.iq_y), m.ip_y);
}

// mm10_solve (mm10_a.f:2860-3295): predictor on the stress with extrapolated hardening
// (phase 0, 6 unknowns), then the coupled update (phase 1, 7 unknowns), as one state machine.
// x[7] in/out.  c.J keeps the last Jacobian formed (lagged tangent).  Returns true on failure.
// c.J is left FACTORED (mm10_lu7_factor of the last Jacobian formed), *lu_piv its pivot rows; the
// tangent solves with those factors.
template <int HARD, bool LF = false>
CPF_DI bool mm10_solve(const Mm10Ctx& c, double* x, double cos_ang_ttrate_dt, int* it_pred, int* it_upd,
                       double* h_last, int* lu_piv) {
  const double cc = 1.0e-4, red = 0.5;
  const int mls = 10, mmin = 1;
  double y[7], dx[7], R[7];
#pragma unroll
  for (int k = 0; k < 6; ++k) { y[k] = x[k]; dx[k] = 0.0; R[k] = 0.0; }
  y[6] = x[6] + cos_ang_ttrate_dt; dx[6] = 0.0; R[6] = 0.0;
  int phase = 0, iter = 0, ls = 0;
  bool init = true, fail = false, done = false;
  double alpha = 0.0, ls1 = 0.0, ls2 = 0.0, nR = 0.0, inR = 0.0, inR1 = 0.0, h = 0.0;
  // One loop, one back edge, warp-uniform trip count: lanes that are finished idle until the
  // slowest lane of the warp is done, so the warp reconverges at the top of every trip (a
  // loop with several `continue` edges made the lanes run the body one after another).
  const unsigned lanes = CPF_ACTIVEMASK();
#pragma unroll 1
  while (CPF_ANY_SYNC(lanes, !done)) {
    if (!done) {
      double yt[7], Rt[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) yt[k] = init ? y[k] : y[k] + alpha * dx[k];
      h = mm10_resid<HARD, LF>(c, yt, yt[6], Rt, phase == 1);
      double dot = 0.0;
#pragma unroll
      for (int k = 0; k < 7; ++k) dot += Rt[k] * Rt[k];
      const double nRt = sqrt(dot);
      bool accepted = true;
      if (init) {
#pragma unroll
        for (int k = 0; k < 7; ++k) R[k] = Rt[k];
        nR = nRt; inR = nRt; iter = 0;
        if (phase == 0) inR1 = nRt;
        else if (inR == 0.0) inR = inR1;
        init = false;
      } else if (!((0.5 * dot <= ls1 + ls2 * alpha) || (ls > mls))) {
        alpha = red * alpha; ls = ls + 1;        // Armijo test failed: shorter step, same direction
        accepted = false;
      } else {
        bool nan = false;
#pragma unroll
        for (int k = 0; k < 7; ++k) { y[k] = yt[k]; R[k] = Rt[k]; nan = nan || isnan(yt[k]); }
        nR = nRt;
        iter = iter + 1;
        if ((iter > c.miter) || nan) { fail = true; done = true; }
      }
      if (accepted && !done) {
        const bool go = (phase == 0) ? ((nR > c.atol1) && (nR / inR > c.rtol1))
                                     : (((nR > c.atol) && (nR / inR > c.rtol)) || (iter < mmin));
        if (!go) {
          if (phase == 0) { *it_pred += iter; phase = 1; init = true; }
          else { *it_upd += iter; done = true; }
        } else {
          // Newton step: dx = -J^-1 R with the Armijo data of the line search
          mm10_jacobian<HARD>(c, y, y[6], phase == 1);
          double wv[7];
          dot = 0.0;
#pragma unroll
          for (int k = 0; k < 7; ++k) { dx[k] = R[k]; dot += R[k] * R[k]; }
          ls1 = 0.5 * dot;
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            double sj = 0.0;
#pragma unroll
            for (int i = 0; i < 7; ++i) sj += c.J[7 * i + j] * R[i];
            wv[j] = sj;
          }
          if (HARD == MM10_MTS) {
#pragma unroll
            for (int k = 0; k < 7; ++k) c.acc[MM10_SM_STASH + k] = c.J[7 * k + 6];
          }
          *lu_piv = mm10_lu7_factor(c.J);
          mm10_lu7_solve(c.J, *lu_piv, dx);
#pragma unroll
          for (int k = 0; k < 7; ++k) dx[k] = -dx[k];
          double d = 0.0;
#pragma unroll
          for (int k = 0; k < 7; ++k) d += dx[k] * wv[k];
          ls2 = cc * d;
          alpha = 1.0; ls = 0;
        }
      }
    }
  }
  if (fail) {
    if (phase == 0) *it_pred += iter; else *it_upd += iter;
    return true;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = y[k];
  *h_last = h;
  return false;
}
