// Host build of the per-voxel material code of the CUDA library: the SAME source that nvcc
// compiles into k_update_mm01 / k_update_mm10 / k_pk1_tangent (cpfft_b200/csrc/update.cuh,
// mm10.cuh, mm01.cuh, kin.cuh, material_tables.hpp), compiled by plain g++ through the macros
// of compat.cuh and run voxel by voxel.  TEST INFRASTRUCTURE ONLY: it lets the `-m "not gpu"`
// suite check the kernels' arithmetic and control flow against the CPU oracle on the build
// box; the product path never links it.  Built by tests/host_kernels.py.
#define MM10_THREADS 1
#include "../../cpfft_b200/csrc/update.cuh"
#include "../../cpfft_b200/csrc/material_tables.hpp"

#include <cstdio>
#include <vector>

struct mh_model {
  CpfMatTables T;
  std::string err;
  int64_t n3;
  double dt;
  // SoA state, field[comp * n3 + voxel] as on the device
  std::vector<double> Fn, Fn1, Pn1, K4, urcs_n, urcs_n1, eps_n, eps_n1, rot_n1, hist_n, hist_n1, cep;
  std::vector<int32_t> fail, liters;
  int failcnt[2];
  bool lattice_frame = false;   // CPFFT_MM10_LF variant of the Voce single-crystal kernel
};

extern "C" {

mh_model* mh_create(int64_t n3, int nmat, const cpfft_material* mats, int ncry, const cpfft_crystal* crys,
                    const int32_t* matlist, int ncmax, const double* angles, const int32_t* crystal_ids, double dt) {
  mh_model* m = new mh_model;
  m->n3 = n3; m->dt = dt;
  std::vector<cpfft_material> vm(mats, mats + nmat);
  std::vector<cpfft_crystal> vc(crys, crys + ncry);
  if (cpf_build_material_tables(vm, vc, matlist, ncmax, angles, crystal_ids, n3, m->T, m->err)) {
    std::fprintf(stderr, "mh_create: %s\n", m->err.c_str());
    delete m;
    return nullptr;
  }
  const size_t n = (size_t)n3;
  m->Fn.assign(9 * n, 0.0);
  for (int d : {0, 4, 8}) for (size_t e = 0; e < n; ++e) m->Fn[d * n + e] = 1.0;   // FFT_init.f:157-161
  m->Fn1 = m->Fn;
  m->Pn1.assign(9 * n, 0.0); m->K4.assign(81 * n, 0.0);
  m->urcs_n.assign(9 * n, 0.0); m->urcs_n1.assign(9 * n, 0.0);
  m->eps_n.assign(6 * n, 0.0); m->eps_n1.assign(6 * n, 0.0);
  m->rot_n1.assign(9 * n, 0.0);
  for (int d : {0, 4, 8}) for (size_t e = 0; e < n; ++e) m->rot_n1[d * n + e] = 1.0;
  m->hist_n.assign((size_t)m->T.H * n, 0.0); m->hist_n1.assign((size_t)m->T.H * n, 0.0);
  m->cep.assign(36 * n, 0.0);
  m->fail.assign(n, 0); m->liters.assign(2 * n, 0);
  m->failcnt[0] = m->failcnt[1] = 0;
  return m;
}
void mh_destroy(mh_model* m) { delete m; }
void mh_set_lattice_frame(mh_model* m, int on) { m->lattice_frame = on != 0; }
int mh_hist_size(mh_model* m) { return m->T.H; }
int mh_ngrains(mh_model* m) { return m->T.ngrains; }

double* mh_field(mh_model* m, const char* name) {
  const std::string s(name);
  if (s == "Fn") return m->Fn.data();
  if (s == "Fn1") return m->Fn1.data();
  if (s == "Pn1") return m->Pn1.data();
  if (s == "K4") return m->K4.data();
  if (s == "urcs_n") return m->urcs_n.data();
  if (s == "urcs_n1") return m->urcs_n1.data();
  if (s == "eps_n") return m->eps_n.data();
  if (s == "eps_n1") return m->eps_n1.data();
  if (s == "rot_n1") return m->rot_n1.data();
  if (s == "hist_n") return m->hist_n.data();
  if (s == "hist_n1") return m->hist_n1.data();
  if (s == "cep") return m->cep.data();
  return nullptr;
}
int32_t* mh_fail_flags(mh_model* m) { return m->fail.data(); }
int32_t* mh_local_iters(mh_model* m) { return m->liters.data(); }

// one drive_eps_sig sweep: what cpf_launch_update launches, voxel by voxel
int mh_drive_eps_sig(mh_model* m, int step, int iter) {
  UpdArgs a;
  a.Fn = m->Fn.data(); a.Fn1 = m->Fn1.data();
  a.urcs_n = m->urcs_n.data(); a.urcs_n1 = m->urcs_n1.data();
  a.eps_n = m->eps_n.data(); a.eps_n1 = m->eps_n1.data();
  a.rot_n1 = m->rot_n1.data();
  a.hist_n = m->hist_n.data(); a.hist_n1 = m->hist_n1.data();
  a.cep = m->cep.data();
  a.matidx = m->T.midx.data(); a.grain = m->T.gidx.data(); a.grain_cry = m->T.gcry.data();
  a.mats = m->T.md.data(); a.crys = m->T.cd.data(); a.grains = m->T.gtab.data();
  a.fail = m->fail.data(); a.liters = m->liters.data(); a.failcnt = m->failcnt;
  a.n3 = m->n3; a.step = step; a.iter = iter; a.dt = m->dt; a.L = m->T.L;
  // one crystal-library entry behind every grain: the product launches the _u kernels (constants from the kernel parameters)
  bool uni = m->T.has_mm10 && !m->T.gcry.empty();
  for (int32_t ci : m->T.gcry) uni = uni && ci == m->T.gcry[0];
  a.uni_cry = uni ? 1 : 0;
  std::memset(&a.cr0, 0, sizeof(a.cr0));
  if (uni) a.cr0 = m->T.cd[m->T.gcry[0]];
  m->failcnt[1] = 0;
  const int64_t n3 = m->n3;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t e = 0; e < n3; ++e) {
    const CpfMatDev& mp = a.mats[a.matidx[e]];
    if (mp.type == 1) upd_mm01_voxel(a, e);
    else if (mp.type == 10) {
      double sm[MM10_SMEM_DOUBLES];
      if (mp.hard == MM10_MTS) {
        if (mp.ncry > 1) upd_mm10_voxel<true, MM10_MTS>(a, e, sm);
        else upd_mm10_voxel<false, MM10_MTS>(a, e, sm);
      } else if (mp.ncry > 1 && m->lattice_frame && uni) upd_mm10_voxel<true, MM10_VOCE, true, true>(a, e, sm);   // k_update_mm10_taylor_lf_u
      else if (mp.ncry > 1 && m->lattice_frame) upd_mm10_voxel<true, MM10_VOCE, true>(a, e, sm);                  // k_update_mm10_taylor_lf
      else if (mp.ncry > 1) upd_mm10_voxel<true, MM10_VOCE>(a, e, sm);
      else if (m->lattice_frame && uni) upd_mm10_voxel<false, MM10_VOCE, true, true>(a, e, sm);                             // k_update_mm10_lf_u
      else if (m->lattice_frame) upd_mm10_voxel<false, MM10_VOCE, true>(a, e, sm);                                          // k_update_mm10_lf
      else upd_mm10_voxel<false, MM10_VOCE>(a, e, sm);
    }
    // even voxels: [D] in registers (the kernel's default), odd voxels: the memory-resident variant
    double pk1_scratch[81];
    Pk1Scratch S; S.p = pk1_scratch;
    upd_pk1_voxel(a.Fn, a.Fn1, a.urcs_n1, a.cep, m->Pn1.data(), m->K4.data(), n3, e, S);
  }
  return m->failcnt[1];
}

// building blocks of the mm10 kernels, exported for direct tests
// A (7x7 row-major) x = b by the kernels' shared-memory LU (mm10_lu7_factor / mm10_lu7_solve); returns the packed pivot rows
int mh_lu7(const double* A, double* b) {
  double J[49];
  for (int k = 0; k < 49; ++k) J[k] = A[k];
  SArr Ja; Ja.p = J;                       // MM10_THREADS = 1 on the host: element k at J[k]
  const int piv = mm10_lu7_factor(Ja);
  mm10_lu7_solve(Ja, piv, b);
  return piv;
}
double mh_pow_abs(double x, int ie, double fe) { return cpf_pow_abs(x, ie, fe); }

// update.f:75-106 (history, eps, urcs) -- on the device a pointer swap in cpfft_commit_step
void mh_update(mh_model* m) {
  m->hist_n = m->hist_n1; m->eps_n = m->eps_n1; m->urcs_n = m->urcs_n1;
}

}  // extern "C"
