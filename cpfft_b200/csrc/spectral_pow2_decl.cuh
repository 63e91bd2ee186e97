// cpfft_b200: interface between the dispatchers of spectral_pow2.cu and the per-grid instantiations of the fast
// spectral path.  The kernels and their launch code are templates on the grid size N (spectral_pow2_impl.cuh); the 14
// supported sizes are instantiated in three translation units (spectral_pow2_g1/g2/g3.cu) that build.py compiles side by
// side -- one file took 107 s to compile, the longest pole of the build.
#pragma once
#include "common.cuh"

// cg != nullptr: the operator application of one CG iteration, q = G K4 p, with the direction
// update (update_p) and the p.q partial sums fused into the z passes.
// x != nullptr: also the pending solution update x += (rr_alpha / *pq) p_old (k_fz MODE 3).
struct CgFuse { const double* r; double beta; bool update_p; int nparts; double* x; double rr_alpha; const double* pq; };


// q = G (K4 :) p of one grid size: the five passes of G_K_dF (spectral_pow2_impl.cuh)
template <int N> int apply_pow2(cpfft_handle* h, double* src, double* dst, bool flgK, double scale_out, const CgFuse* cg);
// shared-memory opt-in of the kernels of one grid size
template <int N> int init_pow2(cpfft_handle* h);

#define CPF_POW2_SIZES_G1(X) X(16) X(32) X(64) X(128) X(256)
#define CPF_POW2_SIZES_G2(X) X(512) X(40) X(80) X(200)
#define CPF_POW2_SIZES_G3(X) X(320) X(400) X(15) X(51) X(255)       // 15, 51, 255: the reference-faithful odd grids
#define CPF_POW2_SIZES(X) CPF_POW2_SIZES_G1(X) CPF_POW2_SIZES_G2(X) CPF_POW2_SIZES_G3(X)
#define CPF_POW2_INSTANTIATE(N)                                                                             \
  template int apply_pow2<N>(cpfft_handle*, double*, double*, bool, double, const CgFuse*);                 \
  template int init_pow2<N>(cpfft_handle*);
