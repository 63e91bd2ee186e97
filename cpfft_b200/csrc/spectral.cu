// cpfft_b200: the spectral operator  GKF = IFFT3( Ghat(xi) : FFT3( [K4 :] F ) ), G_K_dF.f:11-87.
//
// The reference multiplies by a phase ramp (fftshift), runs nine full complex 3-D FFTs through
// MKL DFTI, contracts real and imaginary parts with a stored N3 x 81 Ghat4 table, runs nine
// inverse FFTs and keeps the real part (G_K_dF.f:101-227, FFT_init.f:272-385).  Because the
// input is real and Ghat is real and even, that equals a real-to-complex transform on the
// un-shifted grid with Ghat evaluated at the signed integer frequency of each bin; this file
// does exactly that with hand-written batched 1-D Stockham FFTs staged in shared memory:
//
//   k_fwd_z   [K4:x contraction fused on load] real z-lines -> half spectrum (kz = 0..N/2)
//   k_fft_y   complex y-lines, forward / inverse, in place
//   k_x_green x forward FFT -> Ghat contraction (computed from integer frequencies, never
//             stored) -> x inverse FFT, in place; works on one tensor ROW (3 components),
//             because Ghat_ijkl = delta_ik xi_j xi_l / |xi|^2 couples only within a row
//   k_inv_z   Hermitian completion, inverse z FFT, real part * scale -> destination field
//
// Even N (not supported correctly by the reference): Ghat = 0 on the
// Nyquist planes, frequencies -N/2+1 .. N/2-1 elsewhere.
#include "common.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>

typedef double2 cplx;

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// Batched Stockham autosort FFT of `nlines` lines of length N held in shared memory
// (line-major, a[line*N + i]).  DIF stage with radix R, sequence length n, stride s:
//   y[q + s (R p + t)] = ( sum_r x[q + s (p + m r)] w_R^{r t} ) w_n^{p t},  m = n / R.
// Every thread produces single outputs, so any radix (including large primes) works.
// Returns the buffer holding the result.  dirsign = -1 forward, +1 inverse (unscaled).
__device__ cplx* fft_lines(cplx* a, cplx* b, int nlines, int N, const int* __restrict__ rad, int nrad,
                           const cplx* __restrict__ tw, int dirsign) {
  int s = 1, n = N;
  const int total = nlines * N;
  for (int st = 0; st < nrad; ++st) {
    const int R = rad[st], m = n / R, twR = N / R;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int line = idx / N, o = idx - line * N;
      const int q = o % s, rest = o / s;
      const int t = rest % R, p = rest / R;
      const cplx* x = a + line * N + q + s * p;
      cplx acc;
      if (R == 2) {
        cplx x0 = x[0], x1 = x[s * m];
        acc = t ? make_double2(x0.x - x1.x, x0.y - x1.y) : make_double2(x0.x + x1.x, x0.y + x1.y);
      } else if (R == 4) {
        cplx x0 = x[0], x1 = x[s * m], x2 = x[2 * s * m], x3 = x[3 * s * m];
        // w_4^{rt}, forward w_4 = -i
        double ar, ai;
        if (t == 0) { ar = x0.x + x1.x + x2.x + x3.x; ai = x0.y + x1.y + x2.y + x3.y; }
        else if (t == 2) { ar = x0.x - x1.x + x2.x - x3.x; ai = x0.y - x1.y + x2.y - x3.y; }
        else {
          // t == 1: x0 + w x1 - x2 - w x3 ; t == 3: x0 - w x1 - x2 + w x3, w = dirsign * i
          const double sg = (t == 1) ? (double)dirsign : -(double)dirsign;
          const double dr = x1.x - x3.x, di = x1.y - x3.y;   // (x1 - x3)
          ar = x0.x - x2.x - sg * di; ai = x0.y - x2.y + sg * dr;
        }
        acc = make_double2(ar, ai);
      } else {
        acc = make_double2(0.0, 0.0);
        int k = 0;  // (r t) mod R
        for (int r = 0; r < R; ++r) {
          cplx w = tw[k * twR];
          w.y *= -(double)dirsign;  // table holds exp(-i..): forward; conj for inverse
          acc = cfma(x[s * m * r], w, acc);
          k += t; if (k >= R) k -= R;
        }
      }
      if (p != 0 && t != 0) {
        cplx w = tw[((long long)p * t * s) % N];
        w.y *= -(double)dirsign;
        acc = cmul(acc, w);
      }
      b[idx] = acc;
    }
    __syncthreads();
    cplx* tmp = a; a = b; b = tmp;
    n = m; s *= R;
  }
  return a;
}

struct SpecArgs {
  int N, Nh, nx, x0;
  int nrad;
  const int* rad;
  const cplx* tw;
  int64_t n3;
};

// ---- forward z: optional K4 contraction, real -> half spectrum ----
__global__ void k_fwd_z(SpecArgs g, const double* __restrict__ src, const double* __restrict__ K4, cplx* __restrict__ spec) {
  extern __shared__ cplx sm[];
  const int N = g.N, Nh = g.Nh;
  cplx* a = sm; cplx* b = sm + 9 * N;
  const int xy = blockIdx.x;  // x * N + y (local x)
  const int64_t base = (int64_t)xy * N, n3 = g.n3;
  for (int z = threadIdx.x; z < N; z += blockDim.x) {
    const int64_t e = base + z;
    double f[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) f[c] = src[c * n3 + e];
    if (K4) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        double t[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) t[j] = K4[(int64_t)(9 * i + j) * n3 + e] * f[j];
        // ddot42n's summation tree (G_K_dF.f:258-264)
        const double v = t[0] + (((t[1] + t[5]) + (t[3] + t[7])) + ((t[2] + t[6]) + (t[4] + t[8])));
        a[i * N + z] = make_double2(v, 0.0);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 9; ++c) a[c * N + z] = make_double2(f[c], 0.0);
    }
  }
  __syncthreads();
  cplx* r = fft_lines(a, b, 9, N, g.rad, g.nrad, g.tw, -1);
  for (int idx = threadIdx.x; idx < 9 * Nh; idx += blockDim.x) {
    const int c = idx / Nh, kz = idx - c * Nh;
    spec[((int64_t)c * g.nx * N + xy) * Nh + kz] = r[c * N + kz];
  }
}

// ---- y lines, in place; one CTA per (comp, x, tile of TZ kz) ----
__global__ void k_fft_y(SpecArgs g, cplx* __restrict__ spec, int TZ, int dirsign) {
  extern __shared__ cplx sm[];
  const int N = g.N, Nh = g.Nh;
  cplx* a = sm; cplx* b = sm + TZ * N;
  const int cx = blockIdx.x;           // c * nx + x
  const int kz0 = blockIdx.y * TZ;
  const int tz = min(TZ, Nh - kz0);
  cplx* base = spec + (int64_t)cx * N * Nh + kz0;
  for (int idx = threadIdx.x; idx < N * TZ; idx += blockDim.x) {
    const int y = idx / TZ, l = idx - y * TZ;
    a[l * N + y] = (l < tz) ? base[(int64_t)y * Nh + l] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  cplx* r = fft_lines(a, b, TZ, N, g.rad, g.nrad, g.tw, dirsign);
  for (int idx = threadIdx.x; idx < N * TZ; idx += blockDim.x) {
    const int y = idx / TZ, l = idx - y * TZ;
    if (l < tz) base[(int64_t)y * Nh + l] = r[l * N + y];
  }
}

__device__ __forceinline__ int signed_freq(int k, int N, bool* nyq) {
  // bin k of an N-point DFT -> signed integer frequency; even-N Nyquist flagged
  const int h = N / 2;
  if ((N & 1) == 0 && k == h) { *nyq = true; return h; }
  return (k <= (N - 1) / 2) ? k : k - N;
}

// ---- x forward, Green contraction, x inverse; one CTA per (y, tile of TZ kz, tensor row) ----
// layout of the operand: spec[((c*NX + x)*NY + y)*Nh + kz] with NX = full N (x lines must be
// complete: single GPU, or the transposed layout of the multi-GPU path where NY = ny local)
__global__ void k_x_green(SpecArgs g, cplx* __restrict__ spec, int TZ, int NY, int y0) {
  extern __shared__ cplx sm[];
  const int N = g.N, Nh = g.Nh;
  const int nl = 3 * TZ;
  cplx* a = sm; cplx* b = sm + nl * N;
  const int y = blockIdx.x, kz0 = blockIdx.y * TZ, row = blockIdx.z;
  const int tz = min(TZ, Nh - kz0);
  for (int idx = threadIdx.x; idx < nl * N; idx += blockDim.x) {
    const int l = idx % TZ, rest = idx / TZ;
    const int x = rest % N, cl = rest / N;
    const int c = 3 * row + cl;
    a[(cl * TZ + l) * N + x] = (l < tz) ? spec[(((int64_t)c * N + x) * NY + y) * Nh + kz0 + l] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  cplx* r = fft_lines(a, b, nl, N, g.rad, g.nrad, g.tw, -1);
  cplx* o = (r == a) ? b : a;
  // Green operator on the row: out_j = xi_j (sum_l tau_l xi_l) / |xi|^2  (FFT_init.f:321-335)
  for (int idx = threadIdx.x; idx < TZ * N; idx += blockDim.x) {
    const int l = idx / N, kx = idx - l * N;
    bool nyq = false;
    const double fx = (double)signed_freq(kx, N, &nyq);
    const double fy = (double)signed_freq(y + y0, N, &nyq);
    const double fz = (double)signed_freq(kz0 + l, N, &nyq);
    const double qq = fx * fx + fy * fy + fz * fz;
    cplx t0 = r[(0 * TZ + l) * N + kx], t1 = r[(1 * TZ + l) * N + kx], t2 = r[(2 * TZ + l) * N + kx];
    double sr = 0.0, si = 0.0;
    if (!(nyq || fabs(qq) <= 1e-10)) {
      const double iq = 1.0 / qq;
      sr = (t0.x * fx + t1.x * fy + t2.x * fz) * iq;
      si = (t0.y * fx + t1.y * fy + t2.y * fz) * iq;
    }
    o[(0 * TZ + l) * N + kx] = make_double2(fx * sr, fx * si);
    o[(1 * TZ + l) * N + kx] = make_double2(fy * sr, fy * si);
    o[(2 * TZ + l) * N + kx] = make_double2(fz * sr, fz * si);
  }
  __syncthreads();
  cplx* r2 = fft_lines(o, r, nl, N, g.rad, g.nrad, g.tw, +1);
  for (int idx = threadIdx.x; idx < nl * N; idx += blockDim.x) {
    const int l = idx % TZ, rest = idx / TZ;
    const int x = rest % N, cl = rest / N;
    const int c = 3 * row + cl;
    if (l < tz) spec[(((int64_t)c * N + x) * NY + y) * Nh + kz0 + l] = r2[(cl * TZ + l) * N + x];
  }
}

// ---- inverse z: Hermitian completion, complex inverse, real part * scale ----
__global__ void k_inv_z(SpecArgs g, const cplx* __restrict__ spec, double* __restrict__ dst, double scale) {
  extern __shared__ cplx sm[];
  const int N = g.N, Nh = g.Nh;
  cplx* a = sm; cplx* b = sm + 9 * N;
  const int xy = blockIdx.x;
  for (int idx = threadIdx.x; idx < 9 * N; idx += blockDim.x) {
    const int c = idx / N, k = idx - c * N;
    const cplx* line = spec + ((int64_t)c * g.nx * N + xy) * Nh;
    cplx v;
    if (k < Nh) v = line[k];
    else { v = line[N - k]; v.y = -v.y; }
    a[idx] = v;
  }
  __syncthreads();
  cplx* r = fft_lines(a, b, 9, N, g.rad, g.nrad, g.tw, +1);
  const int64_t base = (int64_t)xy * N;
  for (int idx = threadIdx.x; idx < 9 * N; idx += blockDim.x) {
    const int c = idx / N, z = idx - c * N;
    dst[c * g.n3 + base + z] = r[idx].x * scale;
  }
}

// ------------------------------------------------------------------------------------------
static void choose_radices(int N, int* rad, int* nrad) {
  int n = N, k = 0;
  while (n % 4 == 0) { rad[k++] = 4; n /= 4; }
  for (int p = 2; p <= n; ++p)
    while (n % p == 0) { rad[k++] = p; n /= p; }
  *nrad = k;
}

static int pick_tz(int N, int lines_per_tz, int Nh) {
  // largest power-of-two tile whose two ping-pong buffers fit in ~200 KB
  int tz = 16;
  while (tz > 1 && (size_t)2 * lines_per_tz * tz * N * sizeof(cplx) > 200 * 1024) tz >>= 1;
  while (tz > 1 && tz / 2 >= Nh) tz >>= 1;
  return tz;
}

int cpf_spectral_init(cpfft_handle* h) {
  const int N = h->N;
  h->fast_pow2 = cpf_pow2_supported(N) && (getenv("CPFFT_GENERIC_FFT") == nullptr);
  {  // development switches for A/B measurements; every setting gives bit-identical results
    const char* e = getenv("CPFFT_IZ_PIPE");
    h->iz_pipe = e ? (e[0] == '0' ? 0 : 1) : 1;
    const char* fx = getenv("CPFFT_CG_FUSE_X");
    h->cg_fuse_x = fx ? (fx[0] != '0') : true;
    const char* l = getenv("CPFFT_IZ_LPC");
    h->iz_lpc = l ? atoi(l) : 8;
    if (h->iz_lpc < 1) h->iz_lpc = 1;
  }
  if (h->fast_pow2) h->Nh = (N & 1) ? (N + 1) / 2 : N / 2;   // even N: the Nyquist bin is never stored (Ghat = 0 there)
  choose_radices(N, h->radices, &h->nrad);
  std::vector<cplx> tw(N);
  for (int k = 0; k < N; ++k) {
    const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)N;
    tw[k] = make_double2((double)cosl(a), (double)sinl(a));
  }
  const size_t spec_elems = (size_t)9 * h->nxloc * N * h->Nh;
  CPF_CUDA(cudaMalloc(&h->tw, sizeof(cplx) * N));
  CPF_CUDA(cudaMemcpy(h->tw, tw.data(), sizeof(cplx) * N, cudaMemcpyHostToDevice));
  CPF_CUDA(cudaMalloc(&h->d_radices, sizeof(int) * 32));
  CPF_CUDA(cudaMemcpy(h->d_radices, h->radices, sizeof(int) * 32, cudaMemcpyHostToDevice));
  CPF_CUDA(cudaMalloc(&h->spec_a, sizeof(cplx) * spec_elems));
  h->spec_b = nullptr; h->spec_c = nullptr;
  if (h->fast_pow2) CPF_CUDA(cudaMalloc(&h->spec_c, sizeof(cplx) * spec_elems));
  if ((size_t)2 * 9 * N * sizeof(cplx) > 227 * 1024) {
    cpf_set_error(h, "grid edge too large for the shared-memory z pass (N <= 806)");
    return CPFFT_ERR_USAGE;
  }
  CPF_CUDA(cudaFuncSetAttribute(k_fwd_z, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CPF_CUDA(cudaFuncSetAttribute(k_inv_z, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CPF_CUDA(cudaFuncSetAttribute(k_fft_y, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CPF_CUDA(cudaFuncSetAttribute(k_x_green, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (h->fast_pow2) return cpf_pow2_init(h);
  return 0;
}

void cpf_spectral_free(cpfft_handle* h) {
  if (h->tw) cudaFree(h->tw);
  if (h->d_radices) cudaFree(h->d_radices);
  if (h->spec_a) cudaFree(h->spec_a);
  if (h->spec_b) cudaFree(h->spec_b);
  if (h->spec_c) cudaFree(h->spec_c);
  h->tw = nullptr; h->d_radices = nullptr; h->spec_a = h->spec_b = h->spec_c = nullptr;
}

int cpf_exchange_fwd(cpfft_handle* h);   // solver.cu (NCCL transposes)
int cpf_exchange_bwd(cpfft_handle* h);

// dst = scale_out * IFFT3( Ghat : FFT3( flgK ? K4:src : src ) ) / N^3
int cpf_apply_G(cpfft_handle* h, const double* src, double* dst, bool flgK, double scale_out) {
  if (h->fast_pow2) return cpf_apply_G_pow2(h, src, dst, flgK, scale_out);
  const int N = h->N, Nh = h->Nh, nx = h->nxloc;
  SpecArgs g;
  g.N = N; g.Nh = Nh; g.nx = nx; g.x0 = h->x0; g.nrad = h->nrad; g.rad = h->d_radices; g.tw = h->tw; g.n3 = h->n3;
  const int threads = (N >= 256) ? 256 : ((N >= 64) ? 128 : 64);
  const size_t sm_z = (size_t)2 * 9 * N * sizeof(cplx);
  int tk = cpf_prof_begin(h, flgK ? CPF_K_FWD_Z_K4 : CPF_K_FWD_Z);
  k_fwd_z<<<nx * N, threads, sm_z, h->stream>>>(g, src, flgK ? h->field[CPFFT_K4] : nullptr, h->spec_a);
  cpf_prof_end(h, tk);
  const int tzy = pick_tz(N, 1, Nh);
  dim3 gy(9 * nx, (Nh + tzy - 1) / tzy);
  const size_t sm_y = (size_t)2 * tzy * N * sizeof(cplx);
  tk = cpf_prof_begin(h, CPF_K_FFT_Y);
  k_fft_y<<<gy, threads, sm_y, h->stream>>>(g, h->spec_a, tzy, -1);
  cpf_prof_end(h, tk);
  const int tzx = pick_tz(N, 3, Nh);
  const size_t sm_x = (size_t)2 * 3 * tzx * N * sizeof(cplx);
  if (h->cfg.world == 1) {
    dim3 gx(N, (Nh + tzx - 1) / tzx, 3);
    tk = cpf_prof_begin(h, CPF_K_X_GREEN);
    k_x_green<<<gx, threads, sm_x, h->stream>>>(g, h->spec_a, tzx, N, 0);
    cpf_prof_end(h, tk);
    h->launches += 5;
  } else {
    int rc = cpf_exchange_fwd(h);   // spec_a (x-slabs) -> spec_b (y-slabs, full x)
    if (rc) return rc;
    const int ny = N / h->cfg.world;
    dim3 gx(ny, (Nh + tzx - 1) / tzx, 3);
    tk = cpf_prof_begin(h, CPF_K_X_GREEN);
    k_x_green<<<gx, threads, sm_x, h->stream>>>(g, h->spec_b, tzx, ny, h->cfg.rank * ny);
    cpf_prof_end(h, tk);
    rc = cpf_exchange_bwd(h);       // spec_b -> spec_a
    if (rc) return rc;
    h->launches += 5;
  }
  tk = cpf_prof_begin(h, CPF_K_FFT_Y);
  k_fft_y<<<gy, threads, sm_y, h->stream>>>(g, h->spec_a, tzy, +1);
  cpf_prof_end(h, tk);
  const double scale = scale_out / ((double)N * (double)N * (double)N);
  tk = cpf_prof_begin(h, CPF_K_INV_Z);
  k_inv_z<<<nx * N, threads, sm_z, h->stream>>>(g, h->spec_a, dst, scale);
  cpf_prof_end(h, tk);
  CPF_CUDA(cudaGetLastError());
  h->n_apply++;
  return 0;
}
