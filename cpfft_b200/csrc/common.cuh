// cpfft_b200: shared declarations for the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/cpfft_b200.h"
#include "material_types.h"

#define CPF_MAX_WORLD 8   // one NVSwitch domain

#define CPF_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      cpf_set_error(h, std::string(#call) + ": " + cudaGetErrorString(e_));               \
      return CPFFT_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

// kernel classes for the built-in CUDA-event profiler (cpfft_profile_*)
enum CpfKernelClass {
  CPF_K_UPDATE_MM01 = 0, CPF_K_UPDATE_MM10, CPF_K_PK1_TANGENT, CPF_K_FWD_Z, CPF_K_FWD_Z_K4, CPF_K_FFT_Y,
  CPF_K_X_GREEN, CPF_K_INV_Z, CPF_K_VECTOR, CPF_K_EXCHANGE,
  CPF_K_UPDATE_MM10_EL,      // mm10 sweeps with iter == 0 (elastic predictor, rstgp1.f:870-877): no local Newton solve
  CPF_K_NUM
};
struct CpfProfEvt { cudaEvent_t a, b; int cls; };

struct cpfft_handle {
  cpfft_config cfg;
  // profiler
  bool prof_on;
  std::vector<CpfProfEvt> prof_live;
  std::vector<CpfProfEvt> prof_pool;
  double prof_ms[CPF_K_NUM]; int64_t prof_cnt[CPF_K_NUM];
  int N, Nh;                 // Nh = stored kz bins: N/2+1, or N/2 on the power-of-two path
  bool fast_pow2;            // spectral_pow2.cu handles this grid
  bool cg_fuse_x;            // CG: solution update x += alpha p fused into the next forward z pass (k_fz MODE 3)
  int iz_lpc;                // grid lines per CTA of k_iz_pipe
  int iz_pipe;               // inverse z pass: 1 k_iz_pipe with cp.async (default), 0 k_iz (CPFFT_IZ_PIPE)
  int nxloc, x0;             // local slab
  int64_t n3;                // local voxels
  int H;                     // history comps
  cudaStream_t stream;
  std::string err;
  std::string log;           // the reference's step / iteration lines of the last cpfft_FFT_nr3 call
  int64_t launches;
  // fields
  double* field[CPFFT_NUM_FIELDS];
  int ncomp[CPFFT_NUM_FIELDS];
  // model
  std::vector<cpfft_material> mats;
  std::vector<cpfft_crystal> crys;
  CpfMatDev* d_mats; CpfCryDev* d_crys;
  int32_t* d_matidx;         // per voxel 0-based material
  int32_t* d_grain;          // (ncmax, n3) grain-table entry per voxel and crystal
  int32_t* d_grain_cry;      // per grain-table entry: crystal library index
  double* d_grains;          // grain table
  int ngrains;
  bool has_mm01, has_mm10;
  bool has_taylor;                    // some cp material has n_crystals > 1 per point
  bool mm10_kern[2][3];               // [Taylor point][hardening law 1 Voce, 2 MTS]: update kernels to launch
  int uni_cry; CpfCryDev cr0;         // >= 0: every grain uses this crystal-library entry (constants passed as kernel parameters)
  CpfHistLayout L;
  int32_t* d_fail; int32_t* d_liters;
  int* d_failcnt;            // {mm10 local failures since reset, failures of the last sweep}
  int64_t n_fail, n_fail_final;
  // spectral work
  double2* spec_a; double2* spec_b;      // half-spectrum buffers [9][nx][N][Nh] (x slabs) / [9][N][ny][Nh] (y slabs)
  double2* spec_c;                       // fast path: output of the inverse y pass, input of the inverse z pass
  double2* tw;                           // twiddles exp(-2 pi i k / N)
  int radices[32]; int nrad;
  int* d_radices;
  double* work9;                         // 9*n3 scratch (K4:x)
  // reductions
  double* d_partials; double* d_scalars; double* h_scalars; int nblocks_red;
  // step-loop state (FFT_nr3 locals that persist between calls)
  double barF[9], barF_t[9], P_bar[9], C_homo[81];
  bool have_chomo;
  int next_step;
  bool committed;            // cpfft_update ran and no sweep since: the n+1 names alias the n buffers
  int cg_truncated;          // CG counts dropped by the last cpfft_FFT_nr3 because cg_cap was too small
  // stats
  double t_pcg, t_sig; int64_t n_apply, n_sweep, n_cg;
  // nccl
  void* nccl_comm; void* nccl_lib;
  double2* xchg_send; double2* xchg_recv;
  // peer-mapped spectrum buffers of every rank (CUDA IPC), own rank = local pointers
  bool p2p;
  double2* peer_spec_a[CPF_MAX_WORLD]; double2* peer_spec_b[CPF_MAX_WORLD];
};

void cpf_set_error(cpfft_handle* h, const std::string& s);
// CUDA-event bracket around one kernel launch (no-ops unless profiling is enabled)
int cpf_prof_begin(cpfft_handle* h, int cls);
void cpf_prof_end(cpfft_handle* h, int token);

// material.cu
int cpf_material_setup(cpfft_handle* h, const int32_t* matlist, int ncmax, const double* angles,
                       const int32_t* crystal_ids);
int cpf_launch_update(cpfft_handle* h, int step, int iter);

// spectral.cu
int cpf_spectral_init(cpfft_handle* h);
void cpf_spectral_free(cpfft_handle* h);
int cpf_apply_G(cpfft_handle* h, const double* src, double* dst, bool flgK, double scale_out);
// spectral_pow2.cu
bool cpf_pow2_supported(int N);
int cpf_pow2_init(cpfft_handle* h);
int cpf_apply_G_pow2(cpfft_handle* h, const double* src, double* dst, bool flgK, double scale_out);
int cpf_cg_apply_pow2(cpfft_handle* h, double* p, double* q, const double* r, double beta, bool update_p, int* nparts,
                      double* x, double rr_alpha, const double* pq);

// solver.cu
double* cpf_field_ptr(cpfft_handle* h, int f);
int cpf_dot(cpfft_handle* h, const double* x, const double* y, int64_t n, double* out);
