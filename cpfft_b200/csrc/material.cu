// cpfft_b200: the per-voxel stress update sweep (drive_eps_sig.f:16-339), one thread per voxel.
//
//   k_update_mm01 / k_update_mm10 : kinematics -> material model -> scatter of the n+1 state
//                                   (rstgp1.f dispatch, rplstr.f:62-85 scatter rules)
//   k_pk1_tangent                 : P = J sigma F^-T and A = dP/dF (cs2p + gptns1 + cep2A)
//
// All state is structure-of-arrays field[comp * n3 + voxel]; consecutive threads touch
// consecutive voxels so every load/store is coalesced.
#include "material_kernels.cuh"
#include "material_tables.hpp"


__global__ void __launch_bounds__(UPD_THREADS) k_update_mm01(UpdArgs a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n3) return;
  upd_mm01_voxel(a, e);
}

// one crystal per material point, Voce hardening: the benchmark's kernels (the Taylor-point and MTS kernels are
// compiled in material_taylor.cu / material_mts.cu, side by side with this file)
MM10_KERNEL(k_update_mm10, false, MM10_VOCE, false, false)
MM10_KERNEL(k_update_mm10_u, false, MM10_VOCE, false, true)
MM10_KERNEL(k_update_mm10_lf, false, MM10_VOCE, true, false)
MM10_KERNEL(k_update_mm10_lf_u, false, MM10_VOCE, true, true)

__global__ void __launch_bounds__(UPD_THREADS) k_pk1_tangent(const double* Fn, const double* Fn1, const double* urcs_n1,
                                                              const double* cep, double* Pn1, double* K4, int64_t n3) {
  extern __shared__ double pk1_sm[];        // 81 doubles per thread: the geometric part of dP/dF between the two phases
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  Pk1Scratch S; S.p = pk1_sm + threadIdx.x;
  upd_pk1_voxel(Fn, Fn1, urcs_n1, cep, Pn1, K4, n3, e, S);
}

// ------------------------------------------------------------------------------------------
int cpf_material_setup(cpfft_handle* h, const int32_t* matlist, int ncmax, const double* angles,
                       const int32_t* crystal_ids) {
  const int64_t n3 = h->n3;
  CpfMatTables T;
  std::string err;
  const int rc = cpf_build_material_tables(h->mats, h->crys, matlist, ncmax, angles, crystal_ids, n3, T, err);
  if (rc) { cpf_set_error(h, err); return rc; }
  h->has_mm01 = T.has_mm01; h->has_mm10 = T.has_mm10; h->ngrains = T.ngrains; h->L = T.L;
  h->has_taylor = T.has_taylor;
  // one crystal-library entry behind every grain: its constants travel as kernel parameters
  h->uni_cry = -1;
  if (T.has_mm10 && !T.gcry.empty()) {
    h->uni_cry = T.gcry[0];
    for (int32_t ci : T.gcry) if (ci != h->uni_cry) { h->uni_cry = -1; break; }
    if (h->uni_cry >= 0) h->cr0 = T.cd[h->uni_cry];
  }
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) h->mm10_kern[i][j] = T.kern[i][j];
  const int H = T.H;
  // (re)allocate history fields
  for (int f : {CPFFT_HIST_N, CPFFT_HIST_N1}) {
    if (h->field[f]) cudaFree(h->field[f]);
    h->field[f] = nullptr; h->ncomp[f] = H;
    CPF_CUDA(cudaMalloc(&h->field[f], sizeof(double) * (size_t)H * n3));
    CPF_CUDA(cudaMemsetAsync(h->field[f], 0, sizeof(double) * (size_t)H * n3, h->stream));
  }
  h->H = H;
  auto upload = [&](void** dst, const void* src, size_t bytes) -> int {
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    if (bytes == 0) return 0;
    if (cudaMalloc(dst, bytes) != cudaSuccess) return 1;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess;
  };
  if (upload((void**)&h->d_mats, T.md.data(), sizeof(CpfMatDev) * T.md.size()) ||
      upload((void**)&h->d_crys, T.cd.data(), sizeof(CpfCryDev) * T.cd.size()) ||
      upload((void**)&h->d_matidx, T.midx.data(), sizeof(int32_t) * n3) ||
      upload((void**)&h->d_grain, T.gidx.data(), sizeof(int32_t) * T.gidx.size()) ||
      upload((void**)&h->d_grain_cry, T.gcry.data(), sizeof(int32_t) * T.gcry.size()) ||
      upload((void**)&h->d_grains, T.gtab.data(), sizeof(double) * T.gtab.size())) {
    cpf_set_error(h, "device allocation failed in cpfft_set_voxels");
    return CPFFT_ERR_CUDA;
  }
  return 0;
}

int cpf_launch_update(cpfft_handle* h, int step, int iter) {
  h->committed = false;      // this sweep writes the n+1 buffers: the names stop aliasing the n state (cpfft_update)
  UpdArgs a;
  a.Fn = h->field[CPFFT_FN]; a.Fn1 = h->field[CPFFT_FN1];
  a.urcs_n = h->field[CPFFT_URCS_N]; a.urcs_n1 = h->field[CPFFT_URCS_N1];
  a.eps_n = h->field[CPFFT_EPS_N]; a.eps_n1 = h->field[CPFFT_EPS_N1];
  a.rot_n1 = h->field[CPFFT_ROT_N1];
  a.hist_n = h->field[CPFFT_HIST_N]; a.hist_n1 = h->field[CPFFT_HIST_N1];
  a.cep = h->field[CPFFT_CEP];
  a.matidx = h->d_matidx; a.grain = h->d_grain; a.grain_cry = h->d_grain_cry; a.mats = h->d_mats; a.crys = h->d_crys; a.grains = h->d_grains;
  a.fail = h->d_fail; a.liters = h->d_liters; a.failcnt = h->d_failcnt;
  a.n3 = h->n3; a.step = step; a.iter = iter; a.dt = h->cfg.tstep; a.L = h->L;
  const int64_t n3 = h->n3;
  const unsigned grid = (unsigned)((n3 + UPD_THREADS - 1) / UPD_THREADS);
  if (h->has_mm01) {
    const int tk = cpf_prof_begin(h, CPF_K_UPDATE_MM01);
    k_update_mm01<<<grid, UPD_THREADS, 0, h->stream>>>(a); h->launches++;
    cpf_prof_end(h, tk);
  }
  {
    // one kernel per (one crystal | Taylor point) x (Voce | MTS) combination present in the model
    typedef void (*Kern)(const UpdArgs);
    // residual slip loop of the Voce single-crystal kernel in the lattice frame (default; CPFFT_MM10_LF=0: sample frame)
    static const bool lf = [] { const char* v = getenv("CPFFT_MM10_LF"); return !(v && v[0] == '0'); }();
    static const bool nouni = [] { const char* v = getenv("CPFFT_MM10_UNI"); return v && v[0] == '0'; }();
    const bool uni = h->uni_cry >= 0 && !nouni;
    a.uni_cry = uni ? 1 : 0;
    a.cr0 = h->cr0;
    const Kern kerns_t[2][2][3] = {{{nullptr, lf ? k_update_mm10_lf : k_update_mm10, k_update_mm10_mts},
                                    {nullptr, lf ? k_update_mm10_taylor_lf : k_update_mm10_taylor, k_update_mm10_taylor_mts}},
                                   {{nullptr, lf ? k_update_mm10_lf_u : k_update_mm10_u, k_update_mm10_mts_u},
                                    {nullptr, lf ? k_update_mm10_taylor_lf_u : k_update_mm10_taylor_u, k_update_mm10_taylor_mts_u}}};
    const auto& kerns = kerns_t[uni ? 1 : 0];
    const size_t smem = sizeof(double) * MM10_SMEM_DOUBLES * UPD_THREADS;
    for (int multi = 0; multi < 2; ++multi)
      for (int hard = 1; hard <= 2; ++hard) {
        if (!h->mm10_kern[multi][hard]) continue;
        CPF_CUDA(cudaFuncSetAttribute(kerns[multi][hard], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tk = cpf_prof_begin(h, iter == 0 ? CPF_K_UPDATE_MM10_EL : CPF_K_UPDATE_MM10);
        kerns[multi][hard]<<<grid, UPD_THREADS, smem, h->stream>>>(a); h->launches++;
        cpf_prof_end(h, tk);
      }
  }
  const int tk = cpf_prof_begin(h, CPF_K_PK1_TANGENT);
  const size_t sm_pk1 = sizeof(double) * 81 * UPD_THREADS;
  CPF_CUDA(cudaFuncSetAttribute(k_pk1_tangent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pk1));
  k_pk1_tangent<<<grid, UPD_THREADS, sm_pk1, h->stream>>>(a.Fn, a.Fn1, a.urcs_n1, a.cep, h->field[CPFFT_PN1], h->field[CPFFT_K4], n3);
  cpf_prof_end(h, tk);
  h->launches++;
  CPF_CUDA(cudaGetLastError());
  return 0;
}
