// cpfft_b200: fast path of the spectral operator G_K_dF (G_K_dF.f:11-87) for power-of-two grids
// (N = 16 .. 512) and the 5-smooth weak-scaling grids 320 and 400.  Same mathematics as spectral.cu, built
// from the register butterflies of fft_core.cuh:
//
//   k_fz  : [K4 : x contraction fused on load] two real z lines per CTA, each packed as an N/2
//           complex line (even/odd samples), N/2-point FFT, untangled to bins kz = 0..N/2-1.
//           The Nyquist bin is never stored: the even-N convention zeroes Ghat there.
//   k_fy  : y lines (forward or inverse), tile of TZ consecutive kz per CTA; first stage loads
//           from global memory, last stage stores to it, one shared-memory exchange between.
//   k_fx  : x forward -> Ghat contraction from integer frequencies -> x inverse, one tensor row
//           (3 components) x TZ kz per CTA; the spectrum stays digit-reversed in shared memory
//           between the two transforms, so nothing is reordered.
//   k_iz  : half-spectrum lines -> packed N/2 inverse transform -> two real samples per thread.
//
// Spectrum layout: spec[((c * NX + x) * NY + y) * NZ + kz], NZ = N/2, complex double.
#pragma once
#include "spectral_pow2_decl.cuh"
#include "fft_core.cuh"
#include <algorithm>

struct Pow2Args {
  int nx, x0;    // local x planes (z / y passes) and their global offset
  int NY, y0;    // x pass: local y extent and its global offset (slab-transposed layout)
  int64_t n3;    // local voxels
  const cplx* tw;  // exp(-2 pi i k / N), k < N
};

// Slab <-> pencil transposes fused into the FFT store stages (world > 1): every rank maps the
// spectrum buffers of all ranks (CUDA IPC) and the last butterfly stage of the y pass / of the
// inverse x pass stores each element straight into the rank that owns it -- NVLink stores in
// 128..256-byte runs, no pack / unpack kernels, no staging buffers, the transfer overlaps
// the transform.  peers[r] = base of rank r's destination buffer (own rank: local pointer).
struct PeerPtrs { cplx* p[CPF_MAX_WORLD]; };

// Per-grid tuning: kz tile sizes of the y and x passes (must divide N/2) and the CTA size of
// the z passes.  Measured at 256^3 (tools/time_apply.py, A/B in one run):
//   * x pass: 8 kz per CTA = 2 * 256 * 8 * 16 B = 64 KB, 3 CTAs/SM, 128-byte runs;
//   * z passes: one grid line per CTA beats two (k_iz 0.78 -> 0.70 ms);
//   * giving the whole L1 to shared memory (cudaSharedmemCarveoutMaxShared) slows every pass
//     down (k_fz 1.9 -> 3.0 ms): the streaming loads want the L1.
// KB = stored kz bins per line: N/2 for even N (the Nyquist bin is never stored), (N+1)/2 for odd N (all of
// kz = 0 .. (N-1)/2; an odd grid has no Nyquist frequency and reproduces the reference's Ghat exactly).
template <int N> struct KzBins { static constexpr int value = (N & 1) ? (N + 1) / 2 : N / 2; };
// signed integer frequency of bin k of an N-point transform: 0 .. ceil(N/2)-1, then -floor(N/2) .. -1
template <int N> __device__ __forceinline__ double sfreq(int k) { return (double)((2 * k < N) ? k : k - N); }

template <int N> struct Pow2Cfg {
  static constexpr int H = N / 2;
  static constexpr int TZY = (H < 16 ? H : 16) < (4096 / N) ? (H < 16 ? H : 16) : (4096 / N);
  static constexpr int TZX = (H < 8 ? H : 8) < (2048 / N) ? (H < 8 ? H : 8) : (2048 / N);   // 512: 4
  static constexpr int ZT = (N / 2 + 31) / 32 * 32;   // threads of k_fz / k_iz (N/2 of them work)
};
template <> struct Pow2Cfg<256> { static constexpr int H = 128, TZY = 16, TZX = 8, ZT = 128; };
// 512^3 only runs slab-decomposed: the last stage of the x pass stores over NVLink, where 64-byte runs (TZX = 4,
// what the 64 KB shared-memory rule gives) reach half the link rate of 128-byte runs (round 1, 8 GPUs: k_fx 1.66 ms
// for the 617 MB that k_fyf, with 128-byte runs, moves in 0.84 ms).  TZX = 8: 136 KB of shared memory, one CTA per SM.
template <> struct Pow2Cfg<512> { static constexpr int H = 256, TZY = 8, TZX = 8, ZT = 256; };
template <> struct Pow2Cfg<40> { static constexpr int H = 20, TZY = 10, TZX = 5, ZT = 32; };
template <> struct Pow2Cfg<80> { static constexpr int H = 40, TZY = 8, TZX = 8, ZT = 64; };
template <> struct Pow2Cfg<200> { static constexpr int H = 100, TZY = 10, TZX = 5, ZT = 128; };
template <> struct Pow2Cfg<320> { static constexpr int H = 160, TZY = 16, TZX = 8, ZT = 160; };
template <> struct Pow2Cfg<400> { static constexpr int H = 200, TZY = 8, TZX = 8, ZT = 224; };
// kz tile of the forward y pass when its last stage stores over NVLink (world > 1): runs of TZY x 16 B.  At 4 and 8
// GPUs the 128-byte runs of the HBM-tuned tile reached 55-65 % of the link rate (round 1: 0.75-0.8 ms of link time
// took 1.3-1.4 ms); 256-byte runs, one CTA of 100-130 KB per SM -- this pass waits on the link, not on occupancy.
template <int N> struct ScatterTile { static constexpr int TZY = Pow2Cfg<N>::TZY; };
template <> struct ScatterTile<512> { static constexpr int TZY = 16; };
// CTA size the y / x pass kernels are compiled for: 255 runs 272 -> 288 threads, which leaves its radix-17 butterflies
// 224 registers instead of 128
template <int N> struct YXBound { static constexpr int value = (N == 255) ? 288 : 512; };
// odd grids (z passes: k_fz_odd / k_iz_odd, one voxel per thread): TZY / TZX divide KB = (N+1)/2
template <> struct Pow2Cfg<15>  { static constexpr int H = 7,   TZY = 8,  TZX = 8,  ZT = 32; };
template <> struct Pow2Cfg<51>  { static constexpr int H = 25,  TZY = 13, TZX = 13, ZT = 64; };
template <> struct Pow2Cfg<255> { static constexpr int H = 127, TZY = 16, TZX = 8,  ZT = 256; };
// Resident CTAs per SM the z passes are compiled for: keep >= 512 threads per SM.  Without a
// floor the compiler spends 254 registers per thread on k_fz and a single CTA fits (measured
// at 320^3: occupancy 7.6 %, 56 % of the HBM peak).
template <int N> struct ZOcc { static constexpr int MINB = (512 + Pow2Cfg<N>::ZT - 1) / Pow2Cfg<N>::ZT; };
// resident CTAs the forward z pass is compiled for (development knob -DFZ_N=320 -DFZ_MINB=3, tools/build_variants.py)
#ifdef FZ_MINB
template <int N> struct ZOccF { static constexpr int MINB = (N == FZ_N) ? FZ_MINB : ZOcc<N>::MINB; };
#else
template <int N> struct ZOccF { static constexpr int MINB = ZOcc<N>::MINB; };
// 320: 160 threads x 4 CTAs leave the forward pass 96 registers, too few to keep a K4 row of loads in flight (ncu: 67 % of
// the DRAM peak, long-scoreboard stalls 37 vs 27 at 256^3); 3 CTAs x 128 registers measured 5.51 vs 6.42 ms at 320^3
// (profiles/r02h_fz320ab.log; 2 CTAs x 168 registers: 5.84)
template <> struct ZOccF<320> { static constexpr int MINB = 3; };
#endif
// resident CTAs the inverse z pass is compiled for (development knob IZ_MINB, tools/build_variants.py)
#ifdef IZ_MINB
template <int N> struct ZOccI { static constexpr int MINB = (N == 256) ? IZ_MINB : ZOcc<N>::MINB; };
#else
template <int N> struct ZOccI { static constexpr int MINB = ZOcc<N>::MINB; };
#endif

// ---------------------------------------------------------------------------------------------
// z passes: one grid line (x, y) and its 9 components per CTA, H = N/2 threads, two voxels
// (one packed complex sample) per thread.
//
// Shared memory, both conflict-free for every radix plan:
//   A[i * 9 + c]      the 9 packed lines, component fastest; the in-place FFT stages run on it
//                     with the component as the fastest task index (a quarter warp touches
//                     8-9 consecutive slots of one element index);
//   B[c * HB + k]     natural-order spectrum rows with an odd pitch HB, used for the
//                     (k, H-k) tangling next to the global rows;
//   tw[N]             the twiddle table of the FULL length N (the H-point transform uses
//                     every second entry).
// The last forward stage writes A -> B (digit-reversed position -> natural bin), the first
// inverse stage reads B -> A, so neither side needs a separate reordering pass.
template <int N> struct ZSmem {
  static constexpr int H = N / 2;
  static constexpr int HB = H + 1 + (H & 1);           // odd pitch of the B rows
  static constexpr size_t bytes = sizeof(cplx) * (9 * H + 9 * HB + N);
};

// all stages but the last of the forward transform, in place on A (task = j * 9 + c)
template <int H, class StoreLast>
__device__ __forceinline__ void z_fft_fwd(cplx* A, const cplx* tw, StoreLast store_last) {
  typedef FftPlan<H> P;
  constexpr int N1 = H / P::R1, N2 = N1 / P::R2;
  constexpr bool one = (P::R2 == 1), two = (P::R3 == 1);
  for (int task = threadIdx.x; task < 9 * (H / P::R1); task += blockDim.x) {
    const int j = task / 9, c = task - j * 9;
    if (one) fft_stage_dif<H, P::R1, -1, 2>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { store_last(c, i, v); });
    else fft_stage_dif<H, P::R1, -1, 2>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { A[i * 9 + c] = v; });
  }
  __syncthreads();
  if constexpr (!one) {
    for (int task = threadIdx.x; task < 9 * (H / P::R2); task += blockDim.x) {
      const int j = task / 9, c = task - j * 9;
      if (two) fft_stage_dif<N1, P::R2, -1, 2 * (H / N1)>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { store_last(c, i, v); });
      else fft_stage_dif<N1, P::R2, -1, 2 * (H / N1)>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { A[i * 9 + c] = v; });
    }
    __syncthreads();
  }
  if constexpr (!two) {
    for (int task = threadIdx.x; task < 9 * (H / P::R3); task += blockDim.x) {
      const int j = task / 9, c = task - j * 9;
      fft_stage_dif<N2, P::R3, -1, 2 * (H / N2)>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { store_last(c, i, v); });
    }
    __syncthreads();
  }
}
// inverse (transposed) transform: the first stage loads through load_first(c, position)
// `after_first()` runs once, right after the barrier that ends the stage reading through
// load_first: from there on the source buffer is dead (k_iz_pipe refills it with the next line)
template <int H, class LoadFirst, class AfterFirst>
__device__ __forceinline__ void z_fft_inv(cplx* A, const cplx* tw, LoadFirst load_first, AfterFirst after_first) {
  typedef FftPlan<H> P;
  constexpr int N1 = H / P::R1, N2 = N1 / P::R2;
  constexpr bool one = (P::R2 == 1), two = (P::R3 == 1);
  if (!two) {
    for (int task = threadIdx.x; task < 9 * (H / P::R3); task += blockDim.x) {
      const int j = task / 9, c = task - j * 9;
      fft_stage_dit_inv<N2, P::R3, 2 * (H / N2)>(j, tw, [&](int i) { return load_first(c, i); }, [&](int i, cplx v) { A[i * 9 + c] = v; });
    }
    __syncthreads();
    after_first();
  }
  if (!one) {
    for (int task = threadIdx.x; task < 9 * (H / P::R2); task += blockDim.x) {
      const int j = task / 9, c = task - j * 9;
      if (two) fft_stage_dit_inv<N1, P::R2, 2 * (H / N1)>(j, tw, [&](int i) { return load_first(c, i); }, [&](int i, cplx v) { A[i * 9 + c] = v; });
      else fft_stage_dit_inv<N1, P::R2, 2 * (H / N1)>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { A[i * 9 + c] = v; });
    }
    __syncthreads();
    if (two) after_first();
  }
  for (int task = threadIdx.x; task < 9 * (H / P::R1); task += blockDim.x) {
    const int j = task / 9, c = task - j * 9;
    if (one) fft_stage_dit_inv<H, P::R1, 2>(j, tw, [&](int i) { return load_first(c, i); }, [&](int i, cplx v) { A[i * 9 + c] = v; });
    else fft_stage_dit_inv<H, P::R1, 2>(j, tw, [&](int i) { return A[i * 9 + c]; }, [&](int i, cplx v) { A[i * 9 + c] = v; });
  }
  __syncthreads();
  if (one) after_first();
}
template <int H, class LoadFirst>
__device__ __forceinline__ void z_fft_inv(cplx* A, const cplx* tw, LoadFirst load_first) {
  z_fft_inv<H>(A, tw, load_first, [] {});
}

// MODE 0: transform src.  MODE 1: transform K4 : src (G_K_dF with flgK).  MODE 2: the CG
// direction update p <- r + beta p (FFT_nr3.f:290, MKL dcg) fused in front of MODE 1: src is p
// (read and written), rvec is the residual.  MODE 3: MODE 2 plus the solution update of the
// PREVIOUS iteration, x += alpha p_old with alpha = rr_alpha / *pq (the same expression and
// operands k_cg_update uses), done while p_old is in registers anyway: the separate vector
// pass then only updates the residual (one read of p and one read + write of x less per
// iteration on balance: 72 B / voxel).
template <int N, int MODE>
__global__ void __launch_bounds__(Pow2Cfg<N>::ZT, ZOccF<N>::MINB) k_fz(Pow2Args g, double* __restrict__ src, const double* __restrict__ K4,
                                          cplx* __restrict__ spec, const double* __restrict__ rvec, double beta,
                                          double* __restrict__ xvec, double rr_alpha, const double* __restrict__ pq) {
  typedef ZSmem<N> Z;
  constexpr int H = Z::H, HB = Z::HB;
  extern __shared__ cplx sm[];
  cplx* A = sm;
  cplx* B = sm + 9 * H;
  cplx* tw = B + 9 * HB;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  const int64_t n3 = g.n3;
  const int t = threadIdx.x;
  const int64_t L = blockIdx.x;                         // grid line x * N + y
  const int64_t e0 = L * N + 2 * t;
  if (t < H) {
    double2 f[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) f[c] = *reinterpret_cast<const double2*>(src + c * n3 + e0);
    if (MODE >= 2) {
      double alpha = 0.0;
      if (MODE == 3) alpha = rr_alpha / *pq;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const double2 r = *reinterpret_cast<const double2*>(rvec + c * n3 + e0);
        if (MODE == 3) {
          double2 xv = *reinterpret_cast<const double2*>(xvec + c * n3 + e0);
          xv.x += alpha * f[c].x; xv.y += alpha * f[c].y;
          *reinterpret_cast<double2*>(xvec + c * n3 + e0) = xv;
        }
        f[c] = make_double2(r.x + beta * f[c].x, r.y + beta * f[c].y);
        *reinterpret_cast<double2*>(src + c * n3 + e0) = f[c];
      }
    }
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        double2 a[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) a[j] = *reinterpret_cast<const double2*>(K4 + (int64_t)(9 * i + j) * n3 + e0);
        double ta[9], tb[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) { ta[j] = __dmul_rn(a[j].x, f[j].x); tb[j] = __dmul_rn(a[j].y, f[j].y); }
        // ddot42n's summation tree (G_K_dF.f:258-264)
        const double va = ta[0] + (((ta[1] + ta[5]) + (ta[3] + ta[7])) + ((ta[2] + ta[6]) + (ta[4] + ta[8])));
        const double vb = tb[0] + (((tb[1] + tb[5]) + (tb[3] + tb[7])) + ((tb[2] + tb[6]) + (tb[4] + tb[8])));
        A[t * 9 + i] = make_double2(va, vb);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 9; ++c) A[t * 9 + c] = f[c];
    }
  }
  __syncthreads();
  z_fft_fwd<H>(A, tw, [&](int c, int p, cplx v) { B[c * HB + fft_natural<H>(p)] = v; });
  // untangle: X[k] = (E + w_N^k O), E = (Z[k] + conj Z[H-k]) / 2, O = (Z[k] - conj Z[H-k]) / (2 i)
  const int64_t nxN = (int64_t)g.nx * N;
  for (int idx = threadIdx.x; idx < 9 * H; idx += blockDim.x) {
    const int c = idx / H, k = idx - c * H;
    const cplx* row = B + c * HB;
    const cplx Zk = row[k];
    const cplx Zm = c_conj(row[(k == 0) ? 0 : H - k]);
    const cplx E = c_add(Zk, Zm), D = c_sub(Zk, Zm);
    const cplx O = c_mul(make_double2(D.y, -D.x), tw[k]);
    spec[((int64_t)c * nxN + L) * H + k] = make_double2(0.5 * (E.x + O.x), 0.5 * (E.y + O.y));
  }
}

// DOT: also accumulate sum(dst * pvec) over the CTA's voxels (the p.Ap of CG) into
// partials[blockIdx.x]; fixed summation order, no atomics.
template <int N, bool DOT>
__global__ void __launch_bounds__(Pow2Cfg<N>::ZT, ZOcc<N>::MINB) k_iz(Pow2Args g, const cplx* __restrict__ spec, double* __restrict__ dst, double scale,
                                          const double* __restrict__ pvec, double* __restrict__ partials) {
  typedef ZSmem<N> Z;
  constexpr int H = Z::H, HB = Z::HB;
  extern __shared__ cplx sm[];
  cplx* A = sm;
  cplx* B = sm + 9 * H;
  cplx* tw = B + 9 * HB;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  const int64_t nxN = (int64_t)g.nx * N;
  const int64_t L = blockIdx.x;
  // stage the 9 half-spectrum rows (one coalesced read each), then tangle pairs (k, H-k) in place:
  //   Z'[k] = (X[k] + conj X[H-k]) + i w_N^-k (X[k] - conj X[H-k]),  X[H] = 0,  X[0] real
  for (int idx = threadIdx.x; idx < 9 * H; idx += blockDim.x) {
    const int c = idx / H, k = idx - c * H;
    B[c * HB + k] = spec[((int64_t)c * nxN + L) * H + k];
  }
  __syncthreads();
  constexpr int NP = H / 2 + 1;                 // pairs per row: k = 0 .. H/2 (0 and H/2 are self-paired)
  for (int idx = threadIdx.x; idx < 9 * NP; idx += blockDim.x) {
    const int c = idx / NP, k = idx - c * NP;
    cplx* row = B + c * HB;
    if (k == 0) {
      const double x0 = row[0].x;
      row[0] = make_double2(x0, x0);            // X[0] real, X[H] = 0:  Z'[0] = X0 (1 + i)
    } else {
      const cplx Xk = row[k], Xh = row[H - k];
      {
        const cplx Xm = c_conj(Xh);
        const cplx E = c_add(Xk, Xm), D = c_sub(Xk, Xm);
        const cplx O = c_mulc(D, tw[k]);
        row[k] = make_double2(E.x - O.y, E.y + O.x);
      }
      if (2 * k != H) {
        const cplx Xm = c_conj(Xk);
        const cplx E = c_add(Xh, Xm), D = c_sub(Xh, Xm);
        const cplx O = c_mulc(D, tw[H - k]);
        row[H - k] = make_double2(E.x - O.y, E.y + O.x);
      }
    }
  }
  __syncthreads();
  z_fft_inv<H>(A, tw, [&](int c, int p) { return B[c * HB + fft_natural<H>(p)]; });
  const int t = threadIdx.x;
  const int64_t e0 = L * N + 2 * t;
  double acc = 0.0;
  if (t < H) {
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      const cplx z = A[t * 9 + c];
      const double2 o = make_double2(z.x * scale, z.y * scale);
      *reinterpret_cast<double2*>(dst + c * g.n3 + e0) = o;
      if (DOT) {
        const double2 pv = *reinterpret_cast<const double2*>(pvec + c * g.n3 + e0);
        acc += o.x * pv.x + o.y * pv.y;
      }
    }
  }
  if (DOT) {
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int i = 0; i < Pow2Cfg<N>::ZT / 32; ++i) sum += red[i];
      partials[blockIdx.x] = sum;
    }
  }
}

// Software-pipelined inverse z pass.  k_iz is latency bound (0.42 of the HBM peak at 256^3): a
// CTA loads its 18 KB of spectrum, waits, transforms, stores, and with 4-5 CTAs per SM there
// are stretches where nothing is in flight.  Here a CTA walks `lpc` consecutive grid lines and
// requests the half-spectrum rows of the next line with cp.async (16-byte LDGSTS, no
// registers) as soon as the first butterfly stage has moved the current line from B to A, so
// the loads fly during the remaining stages, the stores and the dot product; the CG direction
// values of the DOT variant are loaded into registers before the transform instead of after
// it.  Shared memory and residency are those of k_iz.  Arithmetic, summation order and the
// per-line partial sums are those of k_iz too, so results are bit-identical
// (tests/test_gpu_spectral.py::test_inverse_z_pass_variants_are_bit_identical).
// Measured at 256^3 (tools/ab_iz.py, profiles/r01g_ab_iz_256.json): 0.724 -> 0.539 ms without
// and 0.831 -> 0.661 ms with the fused dot product; a variant with two B buffers (next line
// requested before the current one is touched, one CTA per SM fewer) measured 0.578 / 0.652 ms
// and was dropped; 4, 8 or 16 lines per CTA make no difference.
// A variant that fetched the nine 2 KB rows with 1-D bulk copies (cp.async.bulk + mbarrier) instead of
// 16-byte LDGSTS measured the same or slower (profiles/r02a_ab_tma64.log) and was removed.
#include <cuda_pipeline.h>
template <int N, bool DOT>
__global__ void __launch_bounds__(Pow2Cfg<N>::ZT, ZOccI<N>::MINB) k_iz_pipe(Pow2Args g, const cplx* __restrict__ spec, double* __restrict__ dst, double scale,
                                                                              const double* __restrict__ pvec, double* __restrict__ partials, int64_t nlines, int lpc) {
  typedef ZSmem<N> Z;
  constexpr int H = Z::H, HB = Z::HB;
  extern __shared__ cplx sm[];
  cplx* A = sm;
  cplx* B = sm + 9 * H;
  cplx* tw = B + 9 * HB;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  const int64_t nxN = (int64_t)g.nx * N;
  const int64_t L0 = (int64_t)blockIdx.x * lpc;          // lpc consecutive grid lines per CTA
  const int nl = (int)((nlines - L0) < lpc ? (nlines - L0) : lpc);
  auto prefetch = [&](int64_t L) {
    for (int idx = threadIdx.x; idx < 9 * H; idx += blockDim.x) {
      const int c = idx / H, k = idx - c * H;
      __pipeline_memcpy_async(B + c * HB + k, spec + ((int64_t)c * nxN + L) * H + k, sizeof(cplx));
    }
    __pipeline_commit();
  };
  prefetch(L0);
  const int t = threadIdx.x;
  constexpr int NP = H / 2 + 1;
  for (int il = 0; il < nl; ++il) {
    const int64_t L = L0 + il;
    __pipeline_wait_prior(0);
    __syncthreads();                                    // line L has landed for every thread (and tw on the first trip)
    const int64_t e0 = L * N + 2 * t;
    double2 pv[9];
    if (DOT && t < H) {
#pragma unroll
      for (int c = 0; c < 9; ++c) pv[c] = *reinterpret_cast<const double2*>(pvec + c * g.n3 + e0);
    }
    for (int idx = threadIdx.x; idx < 9 * NP; idx += blockDim.x) {      // tangle, as in k_iz
      const int c = idx / NP, k = idx - c * NP;
      cplx* row = B + c * HB;
      if (k == 0) {
        const double x0 = row[0].x;
        row[0] = make_double2(x0, x0);
      } else {
        const cplx Xk = row[k], Xh = row[H - k];
        {
          const cplx Xm = c_conj(Xh);
          const cplx E = c_add(Xk, Xm), D = c_sub(Xk, Xm);
          const cplx O = c_mulc(D, tw[k]);
          row[k] = make_double2(E.x - O.y, E.y + O.x);
        }
        if (2 * k != H) {
          const cplx Xm = c_conj(Xk);
          const cplx E = c_add(Xh, Xm), D = c_sub(Xh, Xm);
          const cplx O = c_mulc(D, tw[H - k]);
          row[H - k] = make_double2(E.x - O.y, E.y + O.x);
        }
      }
    }
    __syncthreads();
    z_fft_inv<H>(A, tw, [&](int c, int p) { return B[c * HB + fft_natural<H>(p)]; },
                 [&] { if (il + 1 < nl) prefetch(L + 1); });
    double acc = 0.0;
    if (t < H) {
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const cplx z = A[t * 9 + c];
        const double2 o = make_double2(z.x * scale, z.y * scale);
        *reinterpret_cast<double2*>(dst + c * g.n3 + e0) = o;
        if (DOT) acc += o.x * pv[c].x + o.y * pv[c].y;
      }
    }
    if (DOT) {
      __shared__ double red[32];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      if (lane == 0) red[w] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int i = 0; i < Pow2Cfg<N>::ZT / 32; ++i) sum += red[i];
        partials[L] = sum;
      }
    }
    __syncthreads();                                    // A, red and the B buffer of line L are free again
  }
}

// ---------------------------------------------------------------------------------------------
// z passes for ODD N (15, 51, 255): the reference-faithful grids -- its Ghat is a projection for odd N only
// (FFT_init.f:146-147, 370-375).  A real line of odd length cannot be packed as an N/2-point complex line, so
// the nine real lines of a grid line (x, y) travel as five complex lines z_j = c_{2j} + i c_{2j+1} (j = 0..3) and
// z_4 = c_8, one N-point transform each, and are separated with
//   C_{2j}[k] = (Z[k] + conj Z[N-k]) / 2,   C_{2j+1}[k] = (Z[k] - conj Z[N-k]) / (2 i),   k = 0 .. (N-1)/2.
// One voxel per thread (lines of odd length are not 16-byte aligned); everything else -- the K4 contraction with
// the reference's summation tree, the fused CG direction / solution updates (MODE), the fused p.Ap sums (DOT) --
// is what k_fz / k_iz do.  Shared memory: A[i * 5 + j] (in-place stages), B[j * NB + k] (natural order), tw[N].
template <int N> struct OddZ {
  static constexpr int KB = (N + 1) / 2, NB = N + 1;
  static constexpr int ZT = (N + 31) / 32 * 32;
  static constexpr int MINB = (512 + ZT - 1) / ZT > 8 ? 8 : (512 + ZT - 1) / ZT;
  static constexpr size_t bytes = sizeof(cplx) * (5 * N + 5 * NB + N);
};

template <int N, int MODE>
__global__ void __launch_bounds__(OddZ<N>::ZT, OddZ<N>::MINB) k_fz_odd(Pow2Args g, double* __restrict__ src, const double* __restrict__ K4,
                                              cplx* __restrict__ spec, const double* __restrict__ rvec, double beta,
                                              double* __restrict__ xvec, double rr_alpha, const double* __restrict__ pq) {
  typedef FftPlan<N> P;
  constexpr int KB = OddZ<N>::KB, NB = OddZ<N>::NB, N1 = N / P::R1;
  static_assert(P::R3 == 1, "odd plans have at most two stages");
  extern __shared__ cplx sm[];
  cplx* A = sm;
  cplx* B = sm + 5 * N;
  cplx* tw = B + 5 * NB;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  const int64_t n3 = g.n3;
  const int t = threadIdx.x;
  const int64_t L = blockIdx.x;                         // grid line x * N + y
  const int64_t e = L * N + t;
  if (t < N) {
    double f[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) f[c] = src[c * n3 + e];
    if (MODE >= 2) {
      double alpha = 0.0;
      if (MODE == 3) alpha = rr_alpha / *pq;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const double r = rvec[c * n3 + e];
        if (MODE == 3) xvec[c * n3 + e] += alpha * f[c];
        f[c] = r + beta * f[c];
        src[c * n3 + e] = f[c];
      }
    }
    double v[10];
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        double ta[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) ta[j] = __dmul_rn(K4[(int64_t)(9 * i + j) * n3 + e], f[j]);
        // ddot42n's summation tree (G_K_dF.f:258-264)
        v[i] = ta[0] + (((ta[1] + ta[5]) + (ta[3] + ta[7])) + ((ta[2] + ta[6]) + (ta[4] + ta[8])));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 9; ++c) v[c] = f[c];
    }
    v[9] = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) A[t * 5 + j] = make_double2(v[2 * j], v[2 * j + 1]);
  }
  __syncthreads();
  auto to_B = [&](int c, int p, cplx val) { B[c * NB + fft_natural<N>(p)] = val; };
  for (int task = threadIdx.x; task < 5 * (N / P::R1); task += blockDim.x) {
    const int j = task / 5, c = task - j * 5;
    if (P::R2 == 1) fft_stage_dif<N, P::R1, -1, 1>(j, tw, [&](int i) { return A[i * 5 + c]; }, [&](int i, cplx val) { to_B(c, i, val); });
    else fft_stage_dif<N, P::R1, -1, 1>(j, tw, [&](int i) { return A[i * 5 + c]; }, [&](int i, cplx val) { A[i * 5 + c] = val; });
  }
  __syncthreads();
  if constexpr (P::R2 > 1) {
    for (int task = threadIdx.x; task < 5 * (N / P::R2); task += blockDim.x) {
      const int j = task / 5, c = task - j * 5;
      fft_stage_dif<N1, P::R2, -1, N / N1>(j, tw, [&](int i) { return A[i * 5 + c]; }, [&](int i, cplx val) { to_B(c, i, val); });
    }
    __syncthreads();
  }
  const int64_t nxN = (int64_t)g.nx * N;
  for (int idx = threadIdx.x; idx < 9 * KB; idx += blockDim.x) {
    const int c = idx / KB, k = idx - c * KB;
    const cplx* row = B + (c >> 1) * NB;
    const cplx Zk = row[k];
    const cplx Zm = c_conj(row[(k == 0) ? 0 : N - k]);
    cplx X;
    if ((c & 1) == 0) { const cplx E = c_add(Zk, Zm); X = make_double2(0.5 * E.x, 0.5 * E.y); }
    else { const cplx D = c_sub(Zk, Zm); X = make_double2(0.5 * D.y, -0.5 * D.x); }       // D / (2 i)
    spec[((int64_t)c * nxN + L) * KB + k] = X;
  }
}

template <int N, bool DOT>
__global__ void __launch_bounds__(OddZ<N>::ZT, OddZ<N>::MINB) k_iz_odd(Pow2Args g, const cplx* __restrict__ spec, double* __restrict__ dst, double scale,
                                              const double* __restrict__ pvec, double* __restrict__ partials) {
  typedef FftPlan<N> P;
  constexpr int KB = OddZ<N>::KB, NB = OddZ<N>::NB, N1 = N / P::R1;
  extern __shared__ cplx sm[];
  cplx* A = sm;                 // also the staging area S[c * KB + k] of the nine half-spectrum rows (9 KB <= 5 N)
  cplx* B = sm + 5 * N;
  cplx* tw = B + 5 * NB;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  const int64_t nxN = (int64_t)g.nx * N;
  const int64_t L = blockIdx.x;
  for (int idx = threadIdx.x; idx < 9 * KB; idx += blockDim.x) {
    const int c = idx / KB, k = idx - c * KB;
    A[idx] = spec[((int64_t)c * nxN + L) * KB + k];
  }
  __syncthreads();
  // Z_j[k] = C_{2j}[k] + i C_{2j+1}[k],  Z_j[N-k] = conj C_{2j}[k] + i conj C_{2j+1}[k];  C[0] is real
  for (int idx = threadIdx.x; idx < 5 * KB; idx += blockDim.x) {
    const int j = idx / KB, k = idx - j * KB;
    const cplx Xa = A[(2 * j) * KB + k];
    const cplx Xb = (j < 4) ? A[(2 * j + 1) * KB + k] : make_double2(0.0, 0.0);
    if (k == 0) B[j * NB] = make_double2(Xa.x, Xb.x);
    else {
      B[j * NB + k] = make_double2(Xa.x - Xb.y, Xa.y + Xb.x);
      B[j * NB + N - k] = make_double2(Xa.x + Xb.y, Xb.x - Xa.y);
    }
  }
  __syncthreads();
  auto from_B = [&](int c, int p) { return B[c * NB + fft_natural<N>(p)]; };
  if constexpr (P::R2 > 1) {
    for (int task = threadIdx.x; task < 5 * (N / P::R2); task += blockDim.x) {
      const int j = task / 5, c = task - j * 5;
      fft_stage_dit_inv<N1, P::R2, N / N1>(j, tw, [&](int i) { return from_B(c, i); }, [&](int i, cplx val) { A[i * 5 + c] = val; });
    }
    __syncthreads();
  }
  for (int task = threadIdx.x; task < 5 * (N / P::R1); task += blockDim.x) {
    const int j = task / 5, c = task - j * 5;
    if (P::R2 == 1) fft_stage_dit_inv<N, P::R1, 1>(j, tw, [&](int i) { return from_B(c, i); }, [&](int i, cplx val) { A[i * 5 + c] = val; });
    else fft_stage_dit_inv<N, P::R1, 1>(j, tw, [&](int i) { return A[i * 5 + c]; }, [&](int i, cplx val) { A[i * 5 + c] = val; });
  }
  __syncthreads();
  const int t = threadIdx.x;
  const int64_t e = L * N + t;
  double acc = 0.0;
  if (t < N) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const cplx z = A[t * 5 + j];
      const double o0 = z.x * scale;
      dst[(2 * j) * g.n3 + e] = o0;
      if (DOT) acc += o0 * pvec[(2 * j) * g.n3 + e];
      if (j < 4) {
        const double o1 = z.y * scale;
        dst[(2 * j + 1) * g.n3 + e] = o1;
        if (DOT) acc += o1 * pvec[(2 * j + 1) * g.n3 + e];
      }
    }
  }
  if (DOT) {
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int i = 0; i < OddZ<N>::ZT / 32; ++i) sum += red[i];
      partials[blockIdx.x] = sum;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// y passes.  Shared memory s[i * TZ + l] (+ twiddles), tile of TZ consecutive kz per CTA.
//
// The Green operator of a tensor row r is rank one (see k_fx): the x pass needs only
//   T0 = F(t_r0)   and   B = xi_y F(t_r1) + xi_z F(t_r2)
// and returns only  U0 = IFFTx(xi_x s)  and  W = IFFTx(s)  (out_r1 = xi_y W, out_r2 = xi_z W).
// So between the forward y pass and the inverse y pass the spectrum carries 6 lines instead of
// 9 (slots 3r and 3r+1 of the 9-slot buffers; slot 3r+2 is unused): one third less HBM
// traffic in those passes and one third less NVLink traffic in both slab transposes.

// all stages of one y line tile; `ld(i, l)` supplies the input, `fin(it, cnt, i, l, v)` receives
// the natural-order result of task iteration `it` (cnt-th output of that butterfly)
template <int N, int DIR, int TZ, class Ld, class Fin>
__device__ __forceinline__ void y_line_fft(cplx* s, const cplx* tw, Ld ld, Fin fin) {
  typedef FftPlan<N> P;
  constexpr int N1 = N / P::R1, N2 = N1 / P::R2;
  constexpr bool single = (P::R2 == 1), two = (P::R3 == 1);
  constexpr int RL = single ? P::R1 : (two ? P::R2 : P::R3);
  const int ntl = TZ * (N / RL);                       // tasks of the last stage
  if constexpr (single) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int task = threadIdx.x + it * blockDim.x;
      if (task < ntl) {
        const int j = task / TZ, l = task - j * TZ;
        int cnt = 0;
        fft_stage_dif<N, P::R1, DIR, 1>(j, tw, [&](int i) { return ld(i, l); }, [&](int i, cplx v) { fin(it, cnt++, fft_natural<N>(i), l, v); });
      }
    }
  } else {
  for (int task = threadIdx.x; task < TZ * (N / P::R1); task += blockDim.x) {
    const int j = task / TZ, l = task - j * TZ;
    fft_stage_dif<N, P::R1, DIR, 1>(j, tw, [&](int i) { return ld(i, l); }, [&](int i, cplx v) { s[i * TZ + l] = v; });
  }
  __syncthreads();
  if constexpr (!two) {
    for (int task = threadIdx.x; task < TZ * (N / P::R2); task += blockDim.x) {
      const int j = task / TZ, l = task - j * TZ;
      fft_stage_dif<N1, P::R2, DIR, N / N1>(j, tw, [&](int i) { return s[i * TZ + l]; }, [&](int i, cplx v) { s[i * TZ + l] = v; });
    }
    __syncthreads();
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {                     // the launch guarantees ntl <= 2 * blockDim
    const int task = threadIdx.x + it * blockDim.x;
    if (task < ntl) {
      const int j = task / TZ, l = task - j * TZ;
      int cnt = 0;
      if constexpr (two)
        fft_stage_dif<N1, P::R2, DIR, N / N1>(j, tw, [&](int i) { return s[i * TZ + l]; }, [&](int i, cplx v) { fin(it, cnt++, fft_natural<N>(i), l, v); });
      else
        fft_stage_dif<N2, P::R3, DIR, N / N2>(j, tw, [&](int i) { return s[i * TZ + l]; }, [&](int i, cplx v) { fin(it, cnt++, fft_natural<N>(i), l, v); });
    }
  }
  }
}

// forward y pass, grid = (6 * nx, NZ / TZ): line slot ls = blockIdx.x / nx, row r = ls / 2.
//   ls even: component 3r   -> slot 3r
//   ls odd : components 3r+1 and 3r+2 -> slot 3r+1 = xi_y(ky) F(t_r1) + xi_z(kz) F(t_r2)
// in place in spec (x-slab layout), or SCATTER: to the y-slab layout [slot][x global][y local][kz]
// of the rank that owns y (forward slab transpose fused into the store).
// One tile (bx, by) of the forward y pass: bx = line slot * nx + local x plane, by = kz tile.
template <int N, bool SCATTER>
__device__ __forceinline__ void fyf_tile(const Pow2Args& g, cplx* __restrict__ spec, const PeerPtrs& peers, cplx* s, const cplx* tw,
                                         int bx, int by) {
  typedef FftPlan<N> P;
  constexpr int H = KzBins<N>::value, TZ = SCATTER ? ScatterTile<N>::TZY : Pow2Cfg<N>::TZY;
  constexpr int RL = (P::R2 == 1) ? P::R1 : ((P::R3 == 1) ? P::R2 : P::R3);
  const int ls = bx / g.nx, xl = bx - ls * g.nx;
  const int row = ls >> 1, kz0 = by * TZ;
  const int slot = 3 * row + (ls & 1);
  const int ny = g.NY;                                   // SCATTER: y planes per rank
  const int64_t plane = (int64_t)N * H;                  // one (component, x) plane of the x-slab layout
  const cplx* G1 = spec + ((int64_t)slot * g.nx + xl) * plane + kz0;
  cplx* Gout = spec + ((int64_t)slot * g.nx + xl) * plane + kz0;
  const int64_t sbase = ((int64_t)slot * N + (g.x0 + xl)) * ny * H + kz0;
  auto out = [&](int y, int l, cplx v) {
    if (SCATTER) {
      const int pr = y / ny, yl = y - pr * ny;
      peers.p[pr][sbase + (int64_t)yl * H + l] = v;
    } else {
      Gout[(int64_t)y * H + l] = v;
    }
  };
  if ((ls & 1) == 0) {
    y_line_fft<N, -1, TZ>(s, tw, [&](int i, int l) { return G1[(int64_t)i * H + l]; },
                      [&](int, int, int y, int l, cplx v) { out(y, l, v); });
  } else {
    cplx keep[2][RL];
    y_line_fft<N, -1, TZ>(s, tw, [&](int i, int l) { return G1[(int64_t)i * H + l]; },
                      [&](int it, int cnt, int y, int, cplx v) {
                        const double fy = sfreq<N>(y);
                        keep[it][cnt] = make_double2(fy * v.x, fy * v.y);
                      });
    __syncthreads();                                     // the tile buffer is reused for component 3r+2
    const cplx* G2 = G1 + (int64_t)g.nx * plane;
    y_line_fft<N, -1, TZ>(s, tw, [&](int i, int l) { return G2[(int64_t)i * H + l]; },
                      [&](int it, int cnt, int y, int l, cplx v) {
                        const double fz = (double)(kz0 + l);
                        out(y, l, make_double2(keep[it][cnt].x + fz * v.x, keep[it][cnt].y + fz * v.y));
                      });
  }
}

template <int N, bool SCATTER>
__global__ void __launch_bounds__(YXBound<N>::value) k_fyf(Pow2Args g, cplx* __restrict__ spec, PeerPtrs peers) {
  constexpr int TZ = SCATTER ? ScatterTile<N>::TZY : Pow2Cfg<N>::TZY;
  extern __shared__ cplx sm[];
  cplx* s = sm;
  cplx* tw = sm + N * TZ;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  __syncthreads();
  fyf_tile<N, SCATTER>(g, spec, peers, s, tw, blockIdx.x, blockIdx.y);
}
// inverse y pass, grid = (9 * nx, NZ / TZ), out of place: component c = 3r + m reads line slot
// 3r (m = 0) or 3r+1 scaled by xi_y(ky) (m = 1) / xi_z(kz) (m = 2) and writes component c of
// `dst` (x-slab layout, all 9 components again).
template <int N>
__global__ void __launch_bounds__(YXBound<N>::value) k_fyi(Pow2Args g, const cplx* __restrict__ src, cplx* __restrict__ dst) {
  constexpr int H = KzBins<N>::value, TZ = Pow2Cfg<N>::TZY;
  extern __shared__ cplx sm[];
  cplx* s = sm;
  cplx* tw = sm + N * TZ;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  __syncthreads();
  const int c = blockIdx.x / g.nx, xl = blockIdx.x - c * g.nx;
  const int row = c / 3, m = c - 3 * row, kz0 = blockIdx.y * TZ;
  const int slot = 3 * row + (m ? 1 : 0);
  const int64_t plane = (int64_t)N * H;
  const cplx* Gin = src + ((int64_t)slot * g.nx + xl) * plane + kz0;
  cplx* Gout = dst + ((int64_t)c * g.nx + xl) * plane + kz0;
  y_line_fft<N, +1, TZ>(s, tw, [&](int i, int l) {
                      const cplx v = Gin[(int64_t)i * H + l];
                      if (m == 0) return v;
                      const double f = (m == 1) ? sfreq<N>(i) : (double)(kz0 + l);
                      return make_double2(f * v.x, f * v.y);
                    },
                    [&](int, int, int y, int l, cplx v) { Gout[(int64_t)y * H + l] = v; });
}

// ---------------------------------------------------------------------------------------------
// x pass: forward, Green operator, inverse.  grid = (NY, NZ / TZ, 3 tensor rows).
//
// The Green operator of a tensor row is rank one, out_j = xi_j s with
// s = (xi_x t0 + xi_y t1 + xi_z t2) / |xi|^2 (FFT_init.f:321-335), and xi_y, xi_z are constant
// along an x line.  So only TWO lines per (y, kz) are transformed instead of three:
//   forward :  A = FFTx(T0),  B = FFTx(xi_y T1 + xi_z T2)    (line slots 3r, 3r+1 from k_fyf)
//   Green   :  s = (xi_x A + B) / |xi|^2;   A <- xi_x s,  B <- s
//   inverse :  U0 = IFFTx(A) -> slot 3r,  W = IFFTx(B) -> slot 3r+1   (k_fyi expands W)
// Shared memory s[(line * N + i) * TZ + l].  The spectrum stays digit-reversed between the two
// transforms.  SCATTER: the inverse-transformed lines (natural x order) go back to the x-slab
// layout [c][x local][y global][kz] of the rank that owns x (backward transpose fused in).
template <int N, bool SCATTER>
__global__ void __launch_bounds__(YXBound<N>::value) k_fx(Pow2Args g, cplx* __restrict__ spec, PeerPtrs peers) {
  typedef FftPlan<N> P;
  constexpr int H = KzBins<N>::value, TZ = Pow2Cfg<N>::TZX;
  constexpr int N1 = N / P::R1, N2 = N1 / P::R2;
  extern __shared__ cplx sm[];
  cplx* s = sm;
  cplx* tw = sm + 2 * N * TZ;
  for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = g.tw[i];
  __syncthreads();
  const int y = blockIdx.x, kz0 = blockIdx.y * TZ, row = blockIdx.z;
  const int64_t xs = (int64_t)g.NY * H;                       // stride between x planes
  const int64_t cs = (int64_t)N * xs;                         // stride between components
  cplx* G = spec + ((int64_t)(3 * row) * N * g.NY + y) * H + kz0;   // + cl * cs + x * xs + l
  const int ky = y + g.y0;
  const double fy = sfreq<N>(ky);
  // ---- forward: stage 1 from global memory ----
  for (int task = threadIdx.x; task < 2 * TZ * (N / P::R1); task += blockDim.x) {
    const int l = task % TZ, r = task / TZ;
    const int j = r % (N / P::R1), ln = r / (N / P::R1);
    cplx* sc = s + ln * N * TZ + l;
    const cplx* gc = G + l;
    const cplx* gl = gc + (int64_t)ln * cs;
    fft_stage_dif<N, P::R1, -1, 1>(j, tw, [&](int i) { return gl[(int64_t)i * xs]; }, [&](int i, cplx v) { sc[i * TZ] = v; });
  }
  __syncthreads();
  if (P::R2 > 1) {
    for (int task = threadIdx.x; task < 2 * TZ * (N / P::R2); task += blockDim.x) {
      const int l = task % TZ, r = task / TZ;
      const int j = r % (N / P::R2), ln = r / (N / P::R2);
      cplx* sc = s + ln * N * TZ + l;
      fft_stage_dif<N1, P::R2, -1, N / N1>(j, tw, [&](int i) { return sc[i * TZ]; }, [&](int i, cplx v) { sc[i * TZ] = v; });
    }
    __syncthreads();
  }
  if (P::R3 > 1) {
    for (int task = threadIdx.x; task < 2 * TZ * (N / P::R3); task += blockDim.x) {
      const int l = task % TZ, r = task / TZ;
      const int j = r % (N / P::R3), ln = r / (N / P::R3);
      cplx* sc = s + ln * N * TZ + l;
      fft_stage_dif<N2, P::R3, -1, N / N2>(j, tw, [&](int i) { return sc[i * TZ]; }, [&](int i, cplx v) { sc[i * TZ] = v; });
    }
    __syncthreads();
  }
  // ---- Green operator: zero at xi = 0 and on the Nyquist planes (even-N convention) ----
  for (int idx = threadIdx.x; idx < N * TZ; idx += blockDim.x) {
    const int l = idx % TZ, p = idx / TZ;
    const int kx = fft_natural<N>(p);
    const double fx = sfreq<N>(kx), fz = (double)(kz0 + l);
    const double qq = fx * fx + fy * fy + fz * fz;
    // even N: Ghat = 0 on the kx / ky Nyquist planes (the kz one is not stored); odd N has none
    const bool zero = ((N & 1) == 0 && (2 * kx == N || 2 * ky == N)) || (fabs(qq) <= 1e-10);
    cplx* a = s + p * TZ + l;
    const cplx A = a[0], B = a[N * TZ];
    double sr = 0.0, si = 0.0;
    if (!zero) {
      const double iq = 1.0 / qq;
      sr = (fx * A.x + B.x) * iq;
      si = (fx * A.y + B.y) * iq;
    }
    a[0] = make_double2(fx * sr, fx * si);
    a[N * TZ] = make_double2(sr, si);
  }
  __syncthreads();
  // ---- inverse: transposed flow, last stage stores to global memory in natural order ----
  if (P::R3 > 1) {
    for (int task = threadIdx.x; task < 2 * TZ * (N / P::R3); task += blockDim.x) {
      const int l = task % TZ, r = task / TZ;
      const int j = r % (N / P::R3), ln = r / (N / P::R3);
      cplx* sc = s + ln * N * TZ + l;
      fft_stage_dit_inv<N2, P::R3, N / N2>(j, tw, [&](int i) { return sc[i * TZ]; }, [&](int i, cplx v) { sc[i * TZ] = v; });
    }
    __syncthreads();
  }
  if (P::R2 > 1) {
    for (int task = threadIdx.x; task < 2 * TZ * (N / P::R2); task += blockDim.x) {
      const int l = task % TZ, r = task / TZ;
      const int j = r % (N / P::R2), ln = r / (N / P::R2);
      cplx* sc = s + ln * N * TZ + l;
      fft_stage_dit_inv<N1, P::R2, N / N1>(j, tw, [&](int i) { return sc[i * TZ]; }, [&](int i, cplx v) { sc[i * TZ] = v; });
    }
    __syncthreads();
  }
  for (int task = threadIdx.x; task < 2 * TZ * (N / P::R1); task += blockDim.x) {
    const int l = task % TZ, r = task / TZ;
    const int j = r % (N / P::R1), ln = r / (N / P::R1);
    cplx* sc = s + ln * N * TZ + l;
    const int c0 = 3 * row;
    // destination of element (component c, x plane i) of this CTA's (y, kz0 + l)
    auto put = [&](int c, int i, cplx v) {
      if (SCATTER) {
        const int q = i / g.nx, xl = i - q * g.nx;       // owner of x plane i
        peers.p[q][(((int64_t)c * g.nx + xl) * N + ky) * H + kz0 + l] = v;
      } else {
        G[(int64_t)(c - c0) * cs + (int64_t)i * xs + l] = v;
      }
    };
    fft_stage_dit_inv<N, P::R1, 1>(j, tw, [&](int i) { return sc[i * TZ]; }, [&](int i, cplx v) { put(c0 + ln, i, v); });
  }
}

// ---------------------------------------------------------------------------------------------
int cpf_exchange_fwd(cpfft_handle* h);   // solver.cu (NCCL transposes, fallback without peer mapping)
int cpf_rank_barrier(cpfft_handle* h);   // solver.cu: stream-ordered barrier over all ranks
int cpf_exchange_bwd(cpfft_handle* h);

template <int N>
int apply_pow2(cpfft_handle* h, double* src, double* dst, bool flgK, double scale_out, const CgFuse* cg) {
  constexpr bool ODD = (N & 1) != 0;
  constexpr int H = KzBins<N>::value, TZY = Pow2Cfg<N>::TZY, TZX = Pow2Cfg<N>::TZX, ZT = ODD ? OddZ<N>::ZT : Pow2Cfg<N>::ZT;
  static_assert(H % TZY == 0 && H % TZX == 0, "kz tiles must divide the stored bins");
  typedef FftPlan<N> P;
  const int nx = h->nxloc, world = h->cfg.world;
  Pow2Args g;
  g.nx = nx; g.x0 = h->x0; g.NY = N; g.y0 = 0; g.n3 = h->n3; g.tw = h->tw;
  PeerPtrs none = {};
  const size_t sm_z = ODD ? OddZ<N>::bytes : ZSmem<N>::bytes;
  const size_t sm_y = sizeof(cplx) * (N * TZY + N), sm_x = sizeof(cplx) * (2 * N * TZX + N);
  const unsigned zgrid = (unsigned)(nx * N);
  int tk = cpf_prof_begin(h, flgK ? CPF_K_FWD_Z_K4 : CPF_K_FWD_Z);
  if constexpr (ODD) {
    if (cg && cg->update_p && cg->x) k_fz_odd<N, 3><<<zgrid, ZT, sm_z, h->stream>>>(g, src, h->field[CPFFT_K4], h->spec_a, cg->r, cg->beta, cg->x, cg->rr_alpha, cg->pq);
    else if (cg && cg->update_p) k_fz_odd<N, 2><<<zgrid, ZT, sm_z, h->stream>>>(g, src, h->field[CPFFT_K4], h->spec_a, cg->r, cg->beta, nullptr, 0.0, nullptr);
    else if (flgK) k_fz_odd<N, 1><<<zgrid, ZT, sm_z, h->stream>>>(g, src, h->field[CPFFT_K4], h->spec_a, nullptr, 0.0, nullptr, 0.0, nullptr);
    else k_fz_odd<N, 0><<<zgrid, ZT, sm_z, h->stream>>>(g, src, nullptr, h->spec_a, nullptr, 0.0, nullptr, 0.0, nullptr);
  } else {
    if (cg && cg->update_p && cg->x) k_fz<N, 3><<<zgrid, ZT, sm_z, h->stream>>>(g, src, h->field[CPFFT_K4], h->spec_a, cg->r, cg->beta, cg->x, cg->rr_alpha, cg->pq);
    else if (cg && cg->update_p) k_fz<N, 2><<<zgrid, ZT, sm_z, h->stream>>>(g, src, h->field[CPFFT_K4], h->spec_a, cg->r, cg->beta, nullptr, 0.0, nullptr);
    else if (flgK) k_fz<N, 1><<<zgrid, ZT, sm_z, h->stream>>>(g, src, h->field[CPFFT_K4], h->spec_a, nullptr, 0.0, nullptr, 0.0, nullptr);
    else k_fz<N, 0><<<zgrid, ZT, sm_z, h->stream>>>(g, src, nullptr, h->spec_a, nullptr, 0.0, nullptr, 0.0, nullptr);
  }
  cpf_prof_end(h, tk);
  constexpr int Rm12 = P::R2 > 1 ? (P::R1 < P::R2 ? P::R1 : P::R2) : P::R1;
  constexpr int RminY = P::R3 > 1 ? (Rm12 < P::R3 ? Rm12 : P::R3) : Rm12;
  constexpr int ty0 = (TZY * (N / RminY) + 31) / 32 * 32, tx0 = (2 * TZX * (N / RminY) + 31) / 32 * 32;
  constexpr int thr_y = ty0 > YXBound<N>::value ? YXBound<N>::value : ty0;
  constexpr int thr_x = tx0 > YXBound<N>::value ? YXBound<N>::value : tx0;
  const dim3 gy(9 * nx, H / TZY);
  const dim3 gyf(6 * nx, H / TZY);
  // forward y pass with the slab transpose fused in: its own kz tile (ScatterTile)
  constexpr int TZYS = ScatterTile<N>::TZY;
  static_assert(H % TZYS == 0, "kz tiles must divide the stored bins");
  constexpr int tys0 = (TZYS * (N / RminY) + 31) / 32 * 32;
  constexpr int thr_ys = tys0 > YXBound<N>::value ? YXBound<N>::value : tys0;
  const size_t sm_ys = sizeof(cplx) * (N * TZYS + N);
  const dim3 gyfs(6 * nx, H / TZYS);
  if (world == 1) {
    tk = cpf_prof_begin(h, CPF_K_FFT_Y);
    k_fyf<N, false><<<gyf, thr_y, sm_y, h->stream>>>(g, h->spec_a, none);
    cpf_prof_end(h, tk);
    const dim3 gx(N, H / TZX, 3);
    tk = cpf_prof_begin(h, CPF_K_X_GREEN);
    k_fx<N, false><<<gx, thr_x, sm_x, h->stream>>>(g, h->spec_a, none);
    cpf_prof_end(h, tk);
  } else {
    const int ny = N / world;
    Pow2Args gt = g;
    gt.NY = ny; gt.y0 = h->cfg.rank * ny;
    const dim3 gx(ny, H / TZX, 3);
    if (h->p2p) {
      PeerPtrs pa, pb;
      for (int r = 0; r < CPF_MAX_WORLD; ++r) { pa.p[r] = h->peer_spec_a[r]; pb.p[r] = h->peer_spec_b[r]; }
      Pow2Args gs = g;
      gs.NY = ny;                                  // y planes per rank, for the scatter
      // A variant that ran this NVLink-bound pass chunk by chunk on a second stream under the forward z pass of the next
      // chunk of x planes (ordinary or persistent grid, 37..296 CTAs, 4 or 8 chunks) measured 1-4 % SLOWER on 2 GPUs
      // (profiles/r02d_mgpu2_pipeline.log, r02f_mgpu2_pipeline.log): k_fz is latency-bound, it loses throughput in
      // proportion to the CTA slots the y pass holds while it waits on the link, so nothing is hidden.  Moving the
      // transfer to the copy engines instead (y pass in place or into a staging layout, 3-D or contiguous peer copies
      // on a second stream under the z / y passes of the next chunk; the copies alone run at 0.9-1.2 TB/s,
      // profiles/r02l_xfer_probe.log) also lost: -4 % on 2 GPUs, -3 % on 8 (profiles/r02m_*, r02n_*): the chunked z pass
      // runs 5-8 % slower and the last chunk's copies stay exposed.  Both removed; the fused peer stores stay.
      tk = cpf_prof_begin(h, CPF_K_FFT_Y);
      k_fyf<N, true><<<gyfs, thr_ys, sm_ys, h->stream>>>(gs, h->spec_a, pb);     // -> every rank's spec_b
      cpf_prof_end(h, tk);
      int rc = cpf_rank_barrier(h); if (rc) return rc;
      tk = cpf_prof_begin(h, CPF_K_X_GREEN);
      k_fx<N, true><<<gx, thr_x, sm_x, h->stream>>>(gt, h->spec_b, pa);        // -> every rank's spec_a
      cpf_prof_end(h, tk);
      rc = cpf_rank_barrier(h); if (rc) return rc;
    } else {
      tk = cpf_prof_begin(h, CPF_K_FFT_Y);
      k_fyf<N, false><<<gyf, thr_y, sm_y, h->stream>>>(g, h->spec_a, none);
      cpf_prof_end(h, tk);
      int rc = cpf_exchange_fwd(h);
      if (rc) return rc;
      tk = cpf_prof_begin(h, CPF_K_X_GREEN);
      k_fx<N, false><<<gx, thr_x, sm_x, h->stream>>>(gt, h->spec_b, none);
      cpf_prof_end(h, tk);
      rc = cpf_exchange_bwd(h);
      if (rc) return rc;
    }
  }
  tk = cpf_prof_begin(h, CPF_K_FFT_Y);
  k_fyi<N><<<gy, thr_y, sm_y, h->stream>>>(g, h->spec_a, h->spec_c);           // 6 line slots -> 9 components
  cpf_prof_end(h, tk);
  const double scale = scale_out / ((double)N * (double)N * (double)N);
  tk = cpf_prof_begin(h, CPF_K_INV_Z);
  if constexpr (ODD) {
    if (cg) { k_iz_odd<N, true><<<zgrid, ZT, sm_z, h->stream>>>(g, h->spec_c, dst, scale, src, h->d_partials); const_cast<CgFuse*>(cg)->nparts = (int)zgrid; }
    else k_iz_odd<N, false><<<zgrid, ZT, sm_z, h->stream>>>(g, h->spec_c, dst, scale, nullptr, nullptr);
  } else {
  if (h->iz_pipe) {
    const int lpc = h->iz_lpc;
    const unsigned pgrid = (zgrid + lpc - 1) / lpc;
    if (cg) const_cast<CgFuse*>(cg)->nparts = (int)zgrid;
    if (cg) k_iz_pipe<N, true><<<pgrid, ZT, sm_z, h->stream>>>(g, h->spec_c, dst, scale, src, h->d_partials, (int64_t)zgrid, lpc);
    else k_iz_pipe<N, false><<<pgrid, ZT, sm_z, h->stream>>>(g, h->spec_c, dst, scale, nullptr, nullptr, (int64_t)zgrid, lpc);
  } else if (cg) { k_iz<N, true><<<zgrid, ZT, sm_z, h->stream>>>(g, h->spec_c, dst, scale, src, h->d_partials); const_cast<CgFuse*>(cg)->nparts = (int)zgrid; }
  else k_iz<N, false><<<zgrid, ZT, sm_z, h->stream>>>(g, h->spec_c, dst, scale, nullptr, nullptr);
  }
  cpf_prof_end(h, tk);
  h->launches += 5;
  CPF_CUDA(cudaGetLastError());
  h->n_apply++;
  return 0;
}

template <int N>
int init_pow2(cpfft_handle* h) {
  constexpr int TZY = Pow2Cfg<N>::TZY, TZX = Pow2Cfg<N>::TZX;
  const size_t sm_y = sizeof(cplx) * (N * TZY + N), sm_x = sizeof(cplx) * (2 * N * TZX + N);
#define CPF_SMEM_ATTR(kern, bytes) \
  CPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));
  if constexpr ((N & 1) != 0) {
    const size_t sm_z = OddZ<N>::bytes;
    CPF_SMEM_ATTR((k_fz_odd<N, 0>), sm_z);
    CPF_SMEM_ATTR((k_fz_odd<N, 1>), sm_z);
    CPF_SMEM_ATTR((k_fz_odd<N, 2>), sm_z);
    CPF_SMEM_ATTR((k_fz_odd<N, 3>), sm_z);
    CPF_SMEM_ATTR((k_iz_odd<N, true>), sm_z);
    CPF_SMEM_ATTR((k_iz_odd<N, false>), sm_z);
  } else {
    const size_t sm_z = ZSmem<N>::bytes;
    CPF_SMEM_ATTR((k_fz<N, 0>), sm_z);
    CPF_SMEM_ATTR((k_fz<N, 1>), sm_z);
    CPF_SMEM_ATTR((k_fz<N, 2>), sm_z);
    CPF_SMEM_ATTR((k_fz<N, 3>), sm_z);
    CPF_SMEM_ATTR((k_iz<N, true>), sm_z);
    CPF_SMEM_ATTR((k_iz<N, false>), sm_z);
    CPF_SMEM_ATTR((k_iz_pipe<N, true>), sm_z);
    CPF_SMEM_ATTR((k_iz_pipe<N, false>), sm_z);
  }
  CPF_SMEM_ATTR((k_fyf<N, false>), sm_y);
  CPF_SMEM_ATTR((k_fyf<N, true>), sizeof(cplx) * (N * ScatterTile<N>::TZY + N));
  CPF_SMEM_ATTR((k_fyi<N>), sm_y);
  CPF_SMEM_ATTR((k_fx<N, false>), sm_x);
  CPF_SMEM_ATTR((k_fx<N, true>), sm_x);
  return 0;
}
