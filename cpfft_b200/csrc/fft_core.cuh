// cpfft_b200: register-resident radix-2/3/4/5/8/15/16/17 butterflies and in-place shared-memory FFT
// stages for power-of-two, a few 5-smooth and the odd 15 / 51 / 85 / 255 line lengths (the benchmark
// grids; everything else goes through the generic Stockham path in spectral.cu).
//
// Decomposition of an N-point line, N = R1 R2 [R3]: decimation in frequency, every stage in
// place.  Stage with current sub-transform length Ns and radix R, M = Ns / R: task (q, t)
// (q-th sub-transform, t in [0, M)) owns the R elements  q Ns + m M + t, m = 0..R-1, replaces
// them by their R-point DFT (index m -> output digit k) and multiplies output k by w_Ns^(t k).
// After the last stage position p = k1 (N/R1) + k2 (N/(R1 R2)) + k3 holds X[k1 + R1 k2 + R1 R2 k3]
// ("digit-reversed").  The inverse runs the transposed flow (conjugate twiddle, then the
// conjugate butterfly, stages in reverse order) and takes digit-reversed input back to natural
// order, so a forward -> pointwise -> inverse chain never reorders anything.
//
// The same source compiles for the host (tests/fft_core_host_test.cpp, plain g++) so the
// index algebra is verified on the CPU build box.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define FHD __host__ __device__ __forceinline__
#else
#include <cmath>
#define FHD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif

typedef double2 cplx;

FHD cplx c_add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
FHD cplx c_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
FHD cplx c_mul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FHD cplx c_mulc(cplx a, cplx b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)
FHD cplx c_conj(cplx a) { return make_double2(a.x, -a.y); }
// multiply by -i (DIR = -1, forward) or +i (DIR = +1, inverse)
template <int DIR> FHD cplx c_rot(cplx a) { return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }

template <int R, int DIR> struct Dft;

template <int DIR> struct Dft<1, DIR> { static FHD void run(cplx*) {} };
template <int DIR> struct Dft<2, DIR> {
  static FHD void run(cplx* v) { cplx a = v[0], b = v[1]; v[0] = c_add(a, b); v[1] = c_sub(a, b); }
};
template <int DIR> struct Dft<4, DIR> {
  static FHD void run(cplx* v) {
    const cplx t0 = c_add(v[0], v[2]), t1 = c_sub(v[0], v[2]), t2 = c_add(v[1], v[3]);
    const cplx t3 = c_rot<DIR>(c_sub(v[1], v[3]));
    v[0] = c_add(t0, t2); v[2] = c_sub(t0, t2); v[1] = c_add(t1, t3); v[3] = c_sub(t1, t3);
  }
};
template <int DIR> struct Dft<5, DIR> {
  static FHD void run(cplx* v) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2 pi/5), cos(4 pi/5)
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin(2 pi/5), sin(4 pi/5)
    const cplx a = c_add(v[1], v[4]), b = c_add(v[2], v[3]), c = c_sub(v[1], v[4]), d = c_sub(v[2], v[3]);
    const cplx x0 = v[0];
    const cplx p1 = make_double2(x0.x + c1 * a.x + c2 * b.x, x0.y + c1 * a.y + c2 * b.y);
    const cplx p2 = make_double2(x0.x + c2 * a.x + c1 * b.x, x0.y + c2 * a.y + c1 * b.y);
    // forward: X1 = p1 - i q1, X4 = p1 + i q1, X2 = p2 - i q2, X3 = p2 + i q2
    const cplx q1 = c_rot<DIR>(make_double2(s1 * c.x + s2 * d.x, s1 * c.y + s2 * d.y));
    const cplx q2 = c_rot<DIR>(make_double2(s2 * c.x - s1 * d.x, s2 * c.y - s1 * d.y));
    v[0] = make_double2(x0.x + a.x + b.x, x0.y + a.y + b.y);
    v[1] = c_add(p1, q1); v[4] = c_sub(p1, q1);
    v[2] = c_add(p2, q2); v[3] = c_sub(p2, q2);
  }
};
template <int DIR> struct Dft<3, DIR> {
  static FHD void run(cplx* v) {
    const double s = 0.86602540378443864676;                                  // sin(2 pi/3)
    const cplx a = c_add(v[1], v[2]), d = c_sub(v[1], v[2]), x0 = v[0];
    const cplx p = make_double2(x0.x - 0.5 * a.x, x0.y - 0.5 * a.y);
    const cplx q = c_rot<DIR>(make_double2(s * d.x, s * d.y));                // forward: X1 = p - i s d
    v[0] = c_add(x0, a); v[1] = c_add(p, q); v[2] = c_sub(p, q);
  }
};
// Odd prime radix held in registers (17 for the 255 = 3 x 5 x 17 grid, the reference-faithful size next to
// 256): with a_n = x_n + x_{P-n}, b_n = x_n - x_{P-n} the outputs come in pairs
//   X_k, X_{P-k} = (x_0 + sum_n a_n cos(2 pi n k / P))  -/+  i sum_n b_n sin(2 pi n k / P)   (forward),
// (P-1)^2 / 2 real multiply-adds per complex component instead of (P-1)^2 complex ones.
template <int DIR> struct Dft<17, DIR> {
  static FHD void run(cplx* v) {
    constexpr int P = 17, Hh = 8;
    const double cs[P] = {1.0, 0.932472229404355804573, 0.739008917220659115925, 0.445738355776538267396, 0.0922683594633019952397, -0.273662990072082863539, -0.602634636379256389179, -0.850217135729614152134, -0.982973099683901778282, -0.982973099683901778282, -0.850217135729614152134, -0.602634636379256389179, -0.273662990072082863539, 0.0922683594633019952397, 0.445738355776538267396, 0.739008917220659115925, 0.932472229404355804573};
    const double sn[P] = {0.0, 0.361241666187152948745, 0.673695643646557211713, 0.895163291355062322067, 0.995734176295034521871, 0.961825643172819070409, 0.798017227280239503333, 0.526432162877355800245, 0.183749517816570331574, -0.183749517816570331574, -0.526432162877355800245, -0.798017227280239503333, -0.961825643172819070409, -0.995734176295034521871, -0.895163291355062322067, -0.673695643646557211713, -0.361241666187152948745};
    cplx a[Hh], b[Hh];
#pragma unroll
    for (int n = 1; n <= Hh; ++n) { a[n - 1] = c_add(v[n], v[P - n]); b[n - 1] = c_sub(v[n], v[P - n]); }
    const cplx x0 = v[0];
    cplx sum = x0;
#pragma unroll
    for (int n = 0; n < Hh; ++n) sum = c_add(sum, a[n]);
    v[0] = sum;
#pragma unroll
    for (int k = 1; k <= Hh; ++k) {
      double pr = x0.x, pi = x0.y, qr = 0.0, qi = 0.0;
#pragma unroll
      for (int n = 1; n <= Hh; ++n) {
        const int m = (n * k) % P;
        pr += cs[m] * a[n - 1].x; pi += cs[m] * a[n - 1].y;
        qr += sn[m] * b[n - 1].x; qi += sn[m] * b[n - 1].y;
      }
      const cplx q = c_rot<DIR>(make_double2(qr, qi));
      v[k] = make_double2(pr + q.x, pi + q.y);
      v[P - k] = make_double2(pr - q.x, pi - q.y);
    }
  }
};
// 15 = 3 x 5 (Cooley-Tukey inside the registers, as DftCT below with the 15th roots)
template <int DIR> struct Dft<15, DIR> {
  static FHD void run(cplx* v) {
    constexpr int R1 = 3, R2 = 5, R = 15;
    const double cs[R] = {1.0, 0.913545457642600895502, 0.669130606358858213826, 0.309016994374947424102, -0.1045284632676534714, -0.5, -0.809016994374947424102, -0.978147600733805637929, -0.978147600733805637929, -0.809016994374947424102, -0.5, -0.1045284632676534714, 0.309016994374947424102, 0.669130606358858213826, 0.913545457642600895502};
    const double sn[R] = {0.0, 0.406736643075800207754, 0.743144825477394235015, 0.951056516295153572116, 0.994521895368273336923, 0.866025403784438646764, 0.587785252292473129169, 0.207911690817759337102, -0.207911690817759337102, -0.587785252292473129169, -0.866025403784438646764, -0.994521895368273336923, -0.951056516295153572116, -0.743144825477394235015, -0.406736643075800207754};
    cplx y[R2][R1];
#pragma unroll
    for (int b = 0; b < R2; ++b) {
      cplx t[R1];
#pragma unroll
      for (int a = 0; a < R1; ++a) t[a] = v[R2 * a + b];
      Dft<R1, DIR>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) {
        const int m = (b * k1) % R;
        const cplx w = make_double2(cs[m], DIR < 0 ? -sn[m] : sn[m]);
        y[b][k1] = (b == 0 || k1 == 0) ? t[k1] : c_mul(t[k1], w);
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      cplx t[R2];
#pragma unroll
      for (int b = 0; b < R2; ++b) t[b] = y[b][k1];
      Dft<R2, DIR>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
    }
  }
};
template <int DIR> struct Dft<8, DIR> {
  static FHD void run(cplx* v) {
    const double h = 0.70710678118654752440;
    cplx e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    Dft<4, DIR>::run(e); Dft<4, DIR>::run(o);
    // w8^k o[k]: forward w8 = (1 - i)/sqrt2, inverse its conjugate
    const cplx o1 = DIR < 0 ? make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x))
                            : make_double2(h * (o[1].x - o[1].y), h * (o[1].y + o[1].x));
    const cplx o2 = c_rot<DIR>(o[2]);
    const cplx o3 = DIR < 0 ? make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y))
                            : make_double2(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y));
    v[0] = c_add(e[0], o[0]); v[4] = c_sub(e[0], o[0]);
    v[1] = c_add(e[1], o1);   v[5] = c_sub(e[1], o1);
    v[2] = c_add(e[2], o2);   v[6] = c_sub(e[2], o2);
    v[3] = c_add(e[3], o3);   v[7] = c_sub(e[3], o3);
  }
};
template <int DIR> struct Dft<16, DIR> {
  static FHD void run(cplx* v) {
    // 4 x 4: n = 4 a + b, k = k1 + 4 k2
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
    cplx y[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      cplx t[4] = {v[b], v[4 + b], v[8 + b], v[12 + b]};
      Dft<4, DIR>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < 4; ++k1) y[b][k1] = t[k1];
    }
    // twiddle w16^(b k1); w16^m = cos(m pi/8) -/+ i sin(m pi/8)
    const double wr[10] = {1.0, c1, h, s1, 0.0, -s1, -h, -c1, -1.0, -c1};
    const double wi[10] = {0.0, s1, h, c1, 1.0, c1, h, s1, 0.0, -s1};
#pragma unroll
    for (int b = 1; b < 4; ++b)
#pragma unroll
      for (int k1 = 1; k1 < 4; ++k1) {
        const int m = b * k1;  // <= 9
        const cplx w = make_double2(wr[m], DIR < 0 ? -wi[m] : wi[m]);
        y[b][k1] = c_mul(y[b][k1], w);
      }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      cplx t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
      Dft<4, DIR>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
    }
  }
};

// Composite radices 10 = 5 x 2 and 20 = 5 x 4 (Cooley-Tukey inside the registers):
// n = R2 a + b, k = k1 + R1 k2: R1-point transforms over a, twiddle w_R^(b k1), R2-point over b.
template <int R1, int R2, int DIR> struct DftCT {
  static FHD void run(cplx* v) {
    constexpr int R = R1 * R2;
    // cos / sin (2 pi m / 20), m = 0..19; w_R^m uses every (20 / R)-th entry
    const double c20[20] = {1.0, 0.95105651629515357212, 0.8090169943749474241, 0.58778525229247312917, 0.3090169943749474241,
                            0.0, -0.3090169943749474241, -0.58778525229247312917, -0.8090169943749474241, -0.95105651629515357212,
                            -1.0, -0.95105651629515357212, -0.8090169943749474241, -0.58778525229247312917, -0.3090169943749474241,
                            0.0, 0.3090169943749474241, 0.58778525229247312917, 0.8090169943749474241, 0.95105651629515357212};
    const double s20[20] = {0.0, 0.3090169943749474241, 0.58778525229247312917, 0.8090169943749474241, 0.95105651629515357212,
                            1.0, 0.95105651629515357212, 0.8090169943749474241, 0.58778525229247312917, 0.3090169943749474241,
                            0.0, -0.3090169943749474241, -0.58778525229247312917, -0.8090169943749474241, -0.95105651629515357212,
                            -1.0, -0.95105651629515357212, -0.8090169943749474241, -0.58778525229247312917, -0.3090169943749474241};
    cplx y[R2][R1];
#pragma unroll
    for (int b = 0; b < R2; ++b) {
      cplx t[R1];
#pragma unroll
      for (int a = 0; a < R1; ++a) t[a] = v[R2 * a + b];
      Dft<R1, DIR>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) {
        const int m = (b * k1) % R * (20 / R);
        const cplx w = make_double2(c20[m], DIR < 0 ? -s20[m] : s20[m]);
        y[b][k1] = (b == 0 || k1 == 0) ? t[k1] : c_mul(t[k1], w);
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      cplx t[R2];
#pragma unroll
      for (int b = 0; b < R2; ++b) t[b] = y[b][k1];
      Dft<R2, DIR>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
    }
  }
};
template <int DIR> struct Dft<10, DIR> { static FHD void run(cplx* v) { DftCT<5, 2, DIR>::run(v); } };
template <int DIR> struct Dft<20, DIR> { static FHD void run(cplx* v) { DftCT<5, 4, DIR>::run(v); } };

// compile-time plan of an N-point transform
template <int N> struct FftPlan;
template <> struct FftPlan<8>   { static constexpr int R1 = 8,  R2 = 1,  R3 = 1; };
template <> struct FftPlan<16>  { static constexpr int R1 = 16, R2 = 1,  R3 = 1; };
template <> struct FftPlan<32>  { static constexpr int R1 = 8,  R2 = 4,  R3 = 1; };
template <> struct FftPlan<64>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct FftPlan<128> { static constexpr int R1 = 16, R2 = 8,  R3 = 1; };
template <> struct FftPlan<256> { static constexpr int R1 = 16, R2 = 16, R3 = 1; };
template <> struct FftPlan<512> { static constexpr int R1 = 8,  R2 = 8,  R3 = 8; };
// 5-smooth sizes of the weak-scaling grids (320^3 on 2 GPUs, 400^3 on 4)
template <> struct FftPlan<20>  { static constexpr int R1 = 20, R2 = 1,  R3 = 1; };
template <> struct FftPlan<40>  { static constexpr int R1 = 5,  R2 = 8,  R3 = 1; };
template <> struct FftPlan<80>  { static constexpr int R1 = 5,  R2 = 16, R3 = 1; };
template <> struct FftPlan<100> { static constexpr int R1 = 10, R2 = 10, R3 = 1; };
template <> struct FftPlan<160> { static constexpr int R1 = 10, R2 = 16, R3 = 1; };
template <> struct FftPlan<200> { static constexpr int R1 = 5,  R2 = 5,  R3 = 8; };
template <> struct FftPlan<320> { static constexpr int R1 = 5,  R2 = 8,  R3 = 8; };   // measured: radix 20 spills in the x / y passes
// 400: two radix-20 stages.  Measured at 400^3 on one B200 (profiles/r02i_plan400ab.log, ms per launch, y pass / x pass):
// 5.5.16 3.19 / 2.98 (round 1; its last radix-16 stage spilled 320 B in the forward y pass), 10.10.4 2.50 / 2.96,
// 16.5.5 2.76 / 2.67, 4.10.10 3.11 / 3.15, 20.20 1.70 / 2.45 -- one shared-memory exchange less.
template <> struct FftPlan<400> { static constexpr int R1 = 20, R2 = 20, R3 = 1; };

// odd sizes: 255 = 15 x 17 is the reference-faithful neighbour of 256 (the reference's Ghat is a projection for
// odd N only, FFT_init.f:146-147, 370-375); 15, 51, 85 are its small relatives for the parity tests
template <> struct FftPlan<15>  { static constexpr int R1 = 15, R2 = 1,  R3 = 1; };
// (radix 17 first: the last forward stage of the y pass keeps its outputs in registers, the fewer the better)
template <> struct FftPlan<51>  { static constexpr int R1 = 17, R2 = 3,  R3 = 1; };
template <> struct FftPlan<85>  { static constexpr int R1 = 17, R2 = 5,  R3 = 1; };
template <> struct FftPlan<255> { static constexpr int R1 = 17, R2 = 15, R3 = 1; };

// position (after the forward stages) <-> natural frequency index
template <int N> FHD int fft_natural(int p) {
  typedef FftPlan<N> P;
  constexpr int N1 = N / P::R1, N2 = N1 / P::R2;
  const int k1 = p / N1, r = p - k1 * N1;
  const int k2 = r / N2, k3 = r - k2 * N2;
  return k1 + P::R1 * (k2 + P::R2 * k3);
}
template <int N> FHD int fft_position(int k) {
  typedef FftPlan<N> P;
  constexpr int N1 = N / P::R1, N2 = N1 / P::R2;
  const int k1 = k % P::R1, r = k / P::R1;
  const int k2 = r % P::R2, k3 = r / P::R2;
  return k1 * N1 + k2 * N2 + k3;
}

// One task of a forward (DIF) stage.  A: element accessor with  cplx& A(int i)  semantics given
// through load/store functors so that the first stage can read global memory and the last one
// can write it.  tw[j] = exp(-2 pi i j / NT), TWS = NT / Ns (table stride of this stage).
template <int Ns, int R, int DIR, int TWS, class LoadF, class StoreF>
FHD void fft_stage_dif(int task, const cplx* tw, LoadF ld, StoreF st) {
  constexpr int M = Ns / R;
  const int q = task / M, t = task - q * M;
  const int base = q * Ns + t;
  cplx v[R];
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = ld(base + m * M);
  Dft<R, DIR>::run(v);
  if (M > 1) {
#pragma unroll
    for (int k = 1; k < R; ++k) {
      const cplx w = tw[TWS * t * k];
      v[k] = DIR < 0 ? c_mul(v[k], w) : c_mulc(v[k], w);
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) st(base + k * M, v[k]);
}
// One task of an inverse (DIT, transposed) stage: conjugate twiddle first, then the butterfly.
template <int Ns, int R, int TWS, class LoadF, class StoreF>
FHD void fft_stage_dit_inv(int task, const cplx* tw, LoadF ld, StoreF st) {
  constexpr int M = Ns / R;
  const int q = task / M, t = task - q * M;
  const int base = q * Ns + t;
  cplx v[R];
#pragma unroll
  for (int m = 0; m < R; ++m) v[m] = ld(base + m * M);
  if (M > 1) {
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = c_mulc(v[k], tw[TWS * t * k]);
  }
  Dft<R, +1>::run(v);
#pragma unroll
  for (int k = 0; k < R; ++k) st(base + k * M, v[k]);
}
