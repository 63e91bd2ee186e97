// Host build of cpfft_b200/csrc/fft_core.cuh: checks the butterflies, the in-place stage
// algebra (digit-reversed forward, transposed inverse) and the real <-> half-complex
// packing formulas used by the z passes against a naive DFT.  Built and run by
// tests/test_fft_core_host.py with plain g++ (no GPU needed).
#include "../../cpfft_b200/csrc/fft_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

static double maxerr = 0.0;
static void chk(cplx a, cplx b, double scale) {
  double e = std::fmax(std::fabs(a.x - b.x), std::fabs(a.y - b.y)) / scale;
  if (e > maxerr) maxerr = e;
}
static std::vector<cplx> naive(const std::vector<cplx>& x, int sign) {
  const int N = (int)x.size();
  std::vector<cplx> X(N);
  for (int k = 0; k < N; ++k) {
    long double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      long double a = sign * 2.0L * 3.14159265358979323846264338327950288L * ((long long)k * n % N) / N;
      long double c = cosl(a), s = sinl(a);
      sr += x[n].x * c - x[n].y * s; si += x[n].x * s + x[n].y * c;
    }
    X[k] = make_double2((double)sr, (double)si);
  }
  return X;
}
template <int R> static void test_dft() {
  std::vector<cplx> x(R);
  for (auto& v : x) v = make_double2(drand48() - 0.5, drand48() - 0.5);
  for (int dir : {-1, 1}) {
    cplx v[R];
    for (int i = 0; i < R; ++i) v[i] = x[i];
    if (dir < 0) Dft<R, -1>::run(v); else Dft<R, +1>::run(v);
    auto X = naive(x, dir);
    for (int i = 0; i < R; ++i) chk(v[i], X[i], R);
  }
}

template <int N, int DIR> static void run_dif(std::vector<cplx>& a, const std::vector<cplx>& tw, int tws0) {
  typedef FftPlan<N> P;
  cplx* p = a.data();
  auto ld = [&](int i) { return p[i]; };
  auto st = [&](int i, cplx v) { p[i] = v; };
  for (int t = 0; t < N / P::R1; ++t) fft_stage_dif<N, P::R1, DIR, 1>(t, tw.data(), ld, st);
  constexpr int N1 = N / P::R1;
  if constexpr (P::R2 > 1) for (int t = 0; t < N / P::R2; ++t) fft_stage_dif<N1, P::R2, DIR, N / N1>(t, tw.data(), ld, st);
  constexpr int N2 = N1 / P::R2;
  if constexpr (P::R3 > 1) for (int t = 0; t < N / P::R3; ++t) fft_stage_dif<N2, P::R3, DIR, N / N2>(t, tw.data(), ld, st);
  (void)tws0;
}
template <int N> static void run_dit_inv(std::vector<cplx>& a, const std::vector<cplx>& tw) {
  typedef FftPlan<N> P;
  cplx* p = a.data();
  auto ld = [&](int i) { return p[i]; };
  auto st = [&](int i, cplx v) { p[i] = v; };
  constexpr int N1 = N / P::R1, N2 = N1 / P::R2;
  if constexpr (P::R3 > 1) for (int t = 0; t < N / P::R3; ++t) fft_stage_dit_inv<N2, P::R3, N / N2>(t, tw.data(), ld, st);
  if constexpr (P::R2 > 1) for (int t = 0; t < N / P::R2; ++t) fft_stage_dit_inv<N1, P::R2, N / N1>(t, tw.data(), ld, st);
  for (int t = 0; t < N / P::R1; ++t) fft_stage_dit_inv<N, P::R1, 1>(t, tw.data(), ld, st);
}
static std::vector<cplx> twiddles(int N) {
  std::vector<cplx> tw(N);
  for (int k = 0; k < N; ++k) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * k / N;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  return tw;
}
template <int N> static void test_line() {
  auto tw = twiddles(N);
  std::vector<cplx> x(N);
  for (auto& v : x) v = make_double2(drand48() - 0.5, drand48() - 0.5);
  auto X = naive(x, -1), Xi = naive(x, +1);
  std::vector<cplx> a = x;
  run_dif<N, -1>(a, tw, 1);
  for (int p = 0; p < N; ++p) {
    chk(a[p], X[fft_natural<N>(p)], N);
    if (fft_position<N>(fft_natural<N>(p)) != p) { printf("FAIL position/natural N=%d\n", N); exit(1); }
  }
  run_dit_inv<N>(a, tw);                       // back to natural order, scaled by N
  for (int n = 0; n < N; ++n) chk(make_double2(a[n].x / N, a[n].y / N), x[n], 1.0);
  a = x;
  run_dif<N, +1>(a, tw, 1);                    // inverse transform as DIF (natural in)
  for (int p = 0; p < N; ++p) chk(a[p], Xi[fft_natural<N>(p)], N);
}

// real line of length N through the N/2 complex transform (drop the Nyquist bin), then back
template <int N> static void test_real_pack() {
  constexpr int H = N / 2;
  auto twN = twiddles(N);
  std::vector<cplx> twH(H);
  for (int k = 0; k < H; ++k) twH[k] = twN[2 * k];
  std::vector<double> x(N);
  for (auto& v : x) v = drand48() - 0.5;
  std::vector<cplx> xc(N);
  for (int n = 0; n < N; ++n) xc[n] = make_double2(x[n], 0.0);
  auto X = naive(xc, -1);
  std::vector<cplx> z(H);
  for (int n = 0; n < H; ++n) z[n] = make_double2(x[2 * n], x[2 * n + 1]);
  run_dif<H, -1>(z, twH, 1);
  std::vector<cplx> Xh(H);
  for (int k = 0; k < H; ++k) {                // untangle (k_fz)
    const cplx Zk = z[fft_position<H>(k)], Zm = c_conj(z[fft_position<H>((H - k) % H)]);
    const cplx E = c_add(Zk, Zm), D = c_sub(Zk, Zm);
    const cplx O = c_mul(make_double2(D.y, -D.x), twN[k]);   // -i D w_N^k
    Xh[k] = make_double2(0.5 * (E.x + O.x), 0.5 * (E.y + O.y));
    chk(Xh[k], X[k], N);
  }
  // remove the Nyquist content from x: the inverse below assumes X[N/2] = 0
  std::vector<cplx> Xf(N);
  for (int k = 0; k < N; ++k) Xf[k] = (k == H) ? make_double2(0, 0) : X[k];
  auto xr = naive(Xf, +1);
  std::vector<cplx> zz(H);
  for (int k = 0; k < H; ++k) {                // tangle (k_iz), stored digit-reversed
    const cplx Xk = (k == 0) ? make_double2(Xh[0].x, 0.0) : Xh[k];
    const cplx Xm = (k == 0) ? make_double2(0.0, 0.0) : c_conj(Xh[H - k]);
    const cplx E = c_add(Xk, Xm), D = c_sub(Xk, Xm);
    const cplx O = c_mulc(D, twN[k]);                        // D w_N^-k
    zz[fft_position<H>(k)] = make_double2(E.x - O.y, E.y + O.x);  // E + i O
  }
  run_dit_inv<H>(zz, twH);
  for (int n = 0; n < H; ++n) {
    chk(make_double2(zz[n].x / N, 0), make_double2(xr[2 * n].x / N, 0), 1.0);
    chk(make_double2(zz[n].y / N, 0), make_double2(xr[2 * n + 1].x / N, 0), 1.0);
  }
}

// two real lines of ODD length N through one N-point complex transform (k_fz_odd / k_iz_odd): z = a + i b,
//   A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / (2 i),  bins k = 0 .. (N-1)/2; and back
template <int N> static void test_two_real_lines() {
  constexpr int KB = (N + 1) / 2;
  auto tw = twiddles(N);
  std::vector<cplx> xa(N), xb(N), z(N);
  for (int n = 0; n < N; ++n) {
    xa[n] = make_double2(drand48() - 0.5, 0.0); xb[n] = make_double2(drand48() - 0.5, 0.0);
    z[n] = make_double2(xa[n].x, xb[n].x);
  }
  auto XA = naive(xa, -1), XB = naive(xb, -1);
  run_dif<N, -1>(z, tw, 1);
  std::vector<cplx> A(KB), B(KB);
  for (int k = 0; k < KB; ++k) {
    const cplx Zk = z[fft_position<N>(k)], Zm = c_conj(z[fft_position<N>((N - k) % N)]);
    const cplx E = c_add(Zk, Zm), D = c_sub(Zk, Zm);
    A[k] = make_double2(0.5 * E.x, 0.5 * E.y);
    B[k] = make_double2(0.5 * D.y, -0.5 * D.x);              // D / (2 i)
    chk(A[k], XA[k], N); chk(B[k], XB[k], N);
  }
  std::vector<cplx> zz(N);
  for (int k = 0; k < KB; ++k) {                              // Z[k] = A[k] + i B[k], Z[N-k] = conj A[k] + i conj B[k]
    zz[fft_position<N>(k)] = make_double2(A[k].x - B[k].y, A[k].y + B[k].x);
    if (k > 0) zz[fft_position<N>(N - k)] = make_double2(A[k].x + B[k].y, B[k].x - A[k].y);
  }
  run_dit_inv<N>(zz, tw);
  for (int n = 0; n < N; ++n) {
    chk(make_double2(zz[n].x / N, 0), make_double2(xa[n].x, 0), 1.0);
    chk(make_double2(zz[n].y / N, 0), make_double2(xb[n].x, 0), 1.0);
  }
}

int main() {
  srand48(12345);
  test_dft<2>(); test_dft<4>(); test_dft<5>(); test_dft<8>(); test_dft<10>(); test_dft<16>(); test_dft<20>(); test_dft<3>(); test_dft<15>(); test_dft<17>();
  printf("butterflies maxerr %.3e\n", maxerr);
  test_line<8>(); test_line<16>(); test_line<32>(); test_line<64>(); test_line<128>(); test_line<256>(); test_line<512>(); test_line<20>(); test_line<40>(); test_line<80>(); test_line<100>(); test_line<160>(); test_line<200>(); test_line<320>(); test_line<400>(); test_line<15>(); test_line<51>(); test_line<85>(); test_line<255>();
  printf("lines maxerr %.3e\n", maxerr);
  test_real_pack<16>(); test_real_pack<32>(); test_real_pack<64>(); test_real_pack<128>(); test_real_pack<256>(); test_real_pack<512>(); test_real_pack<40>(); test_real_pack<80>(); test_real_pack<200>(); test_real_pack<320>(); test_real_pack<400>();
  printf("real pack maxerr %.3e\n", maxerr);
  test_two_real_lines<15>(); test_two_real_lines<51>(); test_two_real_lines<85>(); test_two_real_lines<255>();
  printf("two real lines maxerr %.3e\n", maxerr);
  if (!(maxerr < 1e-13)) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
