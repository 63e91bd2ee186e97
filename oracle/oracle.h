/*
 * CPFFT ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (C++17 / OpenMP) of the hot path of maranGit/CPFFT:
 *   FFT_nr3 / fftPcg / NBC_update        (src/FFT_nr3.f)
 *   G_K_dF / fftfem3d / ifftfem3d / ddot42n (src/G_K_dF.f)
 *   formG / formfftshift                 (src/FFT_init.f:272-385)
 *   tangent_homo                         (src/tangent_homo.f)
 *   drive_eps_sig / do_nleps_block       (src/drive_eps_sig.f)
 *   rtcmp1 .. getrm1                     (src/polar.f)
 *   cep2A                                (src/cep2A.f)
 *   mm01 + cnst1                         (src/mm01.f)
 *   mm10 (Voce and MTS hardening, NR solver, one or several crystals per point)
 *                                        (src/mm10_a.f, src/mm10_b.f)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (cpfft_b200/) never links, imports or calls it.
 *
 * PARITY: the reference ships no golden vectors, no tests and cannot be built
 * here (ifort + MKL).  PINNED by outputs of the reference's own source, executed
 * statement by statement by tools/fortran_subset.py (fixture
 * tests/golden/reference_vectors.npz, tests/test_reference_vectors.py): rtcmp1
 * (polar.f), getrm1, cep2A_a, mm10_rotation_matrix, mm10_RT2RVE / RT2RVW,
 * mm10_symSW, formG (odd N), ddot42n and the whole operator G_K_dF (DFTI by
 * numpy), mm01 + cnst1, mm10_setup / mm10_formR / mm10_formJ, and
 * mm10_solve_crystal end to end (converged state, tangent, Newton iteration
 * counts, failure flags; Voce and MTS, fcc and bcc48), one voxel through
 * do_nleps_block's sequence, and A WHOLE JOB (tests/golden/reference_global.npz,
 * tests/test_reference_global.py): FFT_nr3, fftPcg, NBC_update, tangent_homo,
 * G_K_dF and the wrapper mm10 executed on a 3^3 polycrystal under mixed boundary
 * conditions -- the oracle's solver reproduces it with the CG iteration count of
 * every Newton-loop solve, the sweep and the outer-iteration counts identical.
 * NOT the reference's text in that run: MKL's closed RCI CG (restated from its
 * documentation) and the gather / scatter of the block driver.  Beyond the pin
 * the oracle is held by derived identities
 * (tests/test_oracle_*.py, tests/test_py_mm10.py): Green-operator projection
 * identities, independent numpy restatements of G_K_dF and of the crystal
 * update, finite-difference checks of cep2A / cnst1 / the local Jacobian /
 * mm10_tangent, MTS == Voce in the degenerate case, homogeneous-deck behaviour
 * -- and by the slot for a maintainer's ifort run (tests/golden/reference_run/).
 */
#ifndef CPFFT_ORACLE_H
#define CPFFT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SLIP 48
#define ORC_MAX_CRYSTALS_PER_POINT 64

/* one entry of the crystal library c_array (mod_crystals.f:142-214), Voce subset */
typedef struct {
  int32_t slip_type;    /* 1 fcc, 2 bcc, 3 single, 6 roters, 7 bcc12 (12 or 1 systems), 8 bcc48 (mod_crystals.f:164-172) */
  int32_t elastic_type; /* 1 = isotropic, 2 = cubic        (mod_crystals.f:173-176) */
  int32_t h_type;       /* 1 = voce, 2 = mts                                      */
  int32_t alter_mode;   /* mm10_a.f:2073                                          */
  int32_t miter;        /* mod_crystals.f:398                                     */
  int32_t pad_;
  double e, nu, mu;
  double harden_n;      /* rate_n */
  double theta_0, tau_y, tau_v, voche_m, iD_v;
  double eps_dot_0_y;   /* gamma_bar synonym (incrystal.f:217) */
  double k_0, burgers;
  double atol, atol1, rtol, rtol1;
  /* MTS hardening (h_type 2; mod_crystals.f:256-275 defaults, incrystal.f:165-236 keywords) */
  double tau_a, tau_hat_y, g_0_y, tau_hat_v, g_0_v, p_y, q_y, p_v, q_v;
  double boltzman, eps_dot_0_v, mu_0, D_0, T_0;
} orc_crystal;

/* one material (inmat.f:97-133 for bilinear, :176-298 for cp) */
typedef struct {
  int32_t type;      /* 1 = bilinear (mm01), 10 = crystal plasticity (mm10) */
  int32_t crystal;   /* cp: 1-based crystal number                          */
  float e, nu, beta, tan_e, yld_pt; /* REAL*4 matprp slots 1,2,3,4,5 (mod_fft.f:20) */
  int32_t n_crystals; /* cp: crystals per material point, imatprp(101) (inmat.f:201-204); 0 = 1 */
} orc_material;

typedef struct orc_model orc_model;

/* build a model; matlist is 1-based material number per voxel (e = x*N*N + y*N + z),
 * angles = Kocks (psi,theta,phi) degrees per voxel (ignored for mm01 voxels) */
orc_model* orc_create(int N, int nmat, const orc_material* mats, int ncry,
                      const orc_crystal* crys, const int32_t* matlist,
                      const double* angles);
/* polycrystalline material points (n_crystals > 1, Taylor average mm10_a.f:112-197):
 * angles (N3, ncmax, 3) and 1-based crystal numbers (N3, ncmax) of every crystal of every voxel
 * (angle_input / crystal_input of read_crystal_data, mod_crystals.f:2111-2210); crystal_ids may
 * be NULL = the material's own crystal (crystal_input single).  A voxel uses the first
 * n_crystals entries of its material.  Resizes the history. */
int orc_set_taylor(orc_model*, int ncmax, const double* angles, const int32_t* crystal_ids);
void orc_destroy(orc_model*);
void orc_set_params(orc_model*, double tolNR, double tolPCG, int maxIter, double tstep);
void orc_set_threads(int nthreads);

int  orc_hist_size(const orc_model*);
/* raw array access (SoA column-major (N3, ncomp) like the reference modules) */
double* orc_Fn(orc_model*);   double* orc_Fn1(orc_model*);
double* orc_Pn(orc_model*);   double* orc_Pn1(orc_model*);
double* orc_K4(orc_model*);   double* orc_dFm(orc_model*);  double* orc_b(orc_model*);
/* per-voxel AoS state, row-major (N3, nvals) */
double* orc_hist_n(orc_model*);  double* orc_hist_n1(orc_model*);
double* orc_urcs_n(orc_model*);  double* orc_urcs_n1(orc_model*);
double* orc_eps_n(orc_model*);   double* orc_eps_n1(orc_model*);
double* orc_rot_n1(orc_model*);
int32_t* orc_fail_flags(orc_model*);
int32_t* orc_local_iters(orc_model*); /* (N3,2): mm10 predictor / update NR iterations of last sweep */

/* hot-path entry points, names follow the reference */
int  orc_drive_eps_sig(orc_model*, int step, int iter);
void orc_G_K_dF(orc_model*, const double* F, double* GKF, int flgK);
int  orc_fftPcg(orc_model*, const double* b, double* x, double tol, int* iters, double* relres);
/* the same loop left after `cap` iterations without an error, and the model's work counters
 * {G_K_dF applications, sweeps, CG iterations} / seconds {pcg, sig-eps}: for the bounded CPU
 * sample of bench.py --impl reference */
int  orc_fftPcg_capped(orc_model*, const double* b, double* x, double tol, int cap, int* iters, double* relres);
void orc_counters(const orc_model*, int64_t* c3, double* t2);
/* process-wide: arithmetic of rtcmp1 (polar.f).  0 = double, the literal restatement, which carries the
 * reference's own small-strain noise (<= ~3e-8 on R); 1 = the same formulas in __float128, rounded to
 * double at the end (default: the value the reference's algorithm defines) */
void orc_set_polar_precision(int quad);
int  orc_get_polar_precision(void);
int  orc_tangent_homo(orc_model*, double* C_homo);
void orc_update(orc_model*);
void orc_mean_P(orc_model*, double* Pbar);

/* whole FFT_nr3 step loop.  BC_all is (nstep,9) row-major; outputs sized nstep
 * (nr_iters, Pbar(nstep,9)), cg_iters sized nstep*cg_cap (per-Newton-solve CG counts,
 * -1 terminated per step), returns 0 or reference-equivalent error code. */
int  orc_FFT_nr3(orc_model*, int nstep, const double* BC_all, const int32_t* isNBC,
                 int32_t* nr_iters, int32_t* cg_iters, int cg_cap, double* Pbar,
                 double* bucket_seconds /* [3]: pcg, sig-eps, total */,
                 int64_t* counters /* [5]: G_K_dF applies, drive sweeps, cg iterations, local mm10 failures (all sweeps), failures in the last sweep */);

/* same, continuing from load step `first_step` (1-based) with the state kept in the model;
 * BC_all holds the rows of the steps to run */
int  orc_FFT_nr3_from(orc_model*, int first_step, int nstep, const double* BC_all, const int32_t* isNBC,
                      int32_t* nr_iters, int32_t* cg_iters, int cg_cap, double* Pbar,
                      double* bucket_seconds, int64_t* counters);

/* unit-level probes used by the pinning tests */
void orc_rtcmp1(const double* F9_rowmajor, double* R9_rowmajor);
void orc_cep2A(const double* Fn, const double* Fn1, const double* t6, const double* cep36_colmajor,
               double* A81);
void orc_point_update(orc_model*, int voxel, int step, int iter, const double* Fn, const double* Fn1,
                      double* P9, double* A81);
void orc_formG_entry(int N, int ii, int jj, int kk, double* G81);
void orc_crystal_stiffness(const orc_crystal*, double* C36_colmajor);
void orc_slip_table(int slip_type, int* nslip, double* b, double* n);
void orc_mm10_residual_jacobian(const orc_crystal* c, const double* angles_deg, const double* D6, double dt,
                                const double* x7, const double* n_stress6, double n_tau_tilde, double* R7, double* J49);

#ifdef __cplusplus
}
#endif
#endif
