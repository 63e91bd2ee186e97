// cpfft_b200: the mm10 sweep kernels -- (one crystal per point | Taylor point) x (Voce | MTS), each compiled twice:
// crystal constants per voxel from the crystal table, or (suffix _u) from the kernel parameters when the whole model
// uses one crystal-library entry (UpdArgs::uni_cry).  *_lf*: residual slip loop in the lattice frame
// (mm10_resid<.., LF = true>), the default; CPFFT_MM10_LF=0 selects the sample-frame loop.
// Twelve instantiations of a 10 k-instruction kernel: they are spread over three translation units
// (material.cu, material_taylor.cu, material_mts.cu) that build.py compiles side by side.
#pragma once
#include "common.cuh"
#ifndef UPD_THREADS
#define UPD_THREADS 128
#endif
#ifndef MM10_MIN_CTAS
#define MM10_MIN_CTAS 2     // 255 registers; 3 CTAs (168 registers) spill and run 1.5x slower
#endif
#define MM10_THREADS UPD_THREADS
#define PK1_THREADS UPD_THREADS
#include "update.cuh"

#define MM10_KERNEL_DECL(name) __global__ void name(const __grid_constant__ UpdArgs a);
#define MM10_KERNEL(name, MULTI, HARD, LF, UNI)                                                       \
  __global__ void __launch_bounds__(UPD_THREADS, MM10_MIN_CTAS) name(const __grid_constant__ UpdArgs a) { \
    extern __shared__ double mm10_sm[];                                                               \
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;                                 \
    if (e >= a.n3) return;                                                                            \
    upd_mm10_voxel<MULTI, HARD, LF, UNI>(a, e, mm10_sm + threadIdx.x);                                \
  }

MM10_KERNEL_DECL(k_update_mm10) MM10_KERNEL_DECL(k_update_mm10_u) MM10_KERNEL_DECL(k_update_mm10_lf) MM10_KERNEL_DECL(k_update_mm10_lf_u)
MM10_KERNEL_DECL(k_update_mm10_taylor) MM10_KERNEL_DECL(k_update_mm10_taylor_u)
MM10_KERNEL_DECL(k_update_mm10_taylor_lf) MM10_KERNEL_DECL(k_update_mm10_taylor_lf_u)
MM10_KERNEL_DECL(k_update_mm10_mts) MM10_KERNEL_DECL(k_update_mm10_mts_u)
MM10_KERNEL_DECL(k_update_mm10_taylor_mts) MM10_KERNEL_DECL(k_update_mm10_taylor_mts_u)
