/*
 * cpfft_b200 -- C ABI of the B200-native CPFFT hot path.
 *
 * The reference (maranGit/CPFFT) has no FFI: its "interface" is a set of fixed-form
 * Fortran subroutines operating on module-global arrays.  Each export below replaces one
 * of those entry points (file:line given) and keeps its argument meaning and error
 * behaviour; all state lives on the GPU behind an opaque handle.  A Fortran host binds
 * these through ISO_C_BINDING (see INTEGRATION.md and cpfft_b200/fortran/cpfft_iso_c.f90).
 *
 * Conventions
 *   - every call returns int: 0 ok; >0 a reference-equivalent fatal condition (the reference
 *     prints to unit `out` and stops, mpi_code.f:15-40); <0 a CUDA/NCCL/usage error.
 *     cpfft_last_error() returns the text.
 *   - voxel index e = x*N*N + y*N + z (0-based; FFT_init.f:311-318); 9-vectors are
 *     11,12,13,21,..,33; 81-vectors are 27i+9j+3k+l (FFT_init.f:283-304).
 *   - with world > 1 the grid is slab-decomposed along x: rank r owns
 *     x in [r*N/world, (r+1)*N/world); host arrays passed in/out are the LOCAL slab.
 *   - non re-entrant per handle, one host thread per GPU (the reference calls all of these
 *     from its serial master thread).
 */
#ifndef CPFFT_B200_H
#define CPFFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cpfft_handle cpfft_handle;

/* error codes > 0: where the reference would `die_abort` */
enum {
  CPFFT_OK = 0,
  CPFFT_ERR_NEWTON = 1,     /* FFT_nr3.f:116  Newton loop does not converge           */
  CPFFT_ERR_CG = 2,         /* FFT_nr3.f:335  fftPcg failed to converge in 1000 its   */
  CPFFT_ERR_TOL = 3,        /* FFT_nr3.f:248  improper CG tolerance                   */
  CPFFT_ERR_STRESS_BC = 4,  /* FFT_nr3.f:153  prescribed stress cannot be reached     */
  CPFFT_ERR_PBAR = 5,       /* FFT_nr3.f:418  P_bar update failed (singular 9x9)      */
  CPFFT_ERR_MATERIAL = 6,   /* mm10_a.f:2811  mm10 implicit solution failed (reserved: local
                               failures are counted, see cpfft_material_failures)        */
  CPFFT_ERR_CUDA = -1, CPFFT_ERR_USAGE = -2, CPFFT_ERR_NCCL = -3
};

/* crystal library entry, c_array(n) of mod_crystals.f:142-214 (Voce subset) */
typedef struct {
  int32_t slip_type;    /* 1 fcc, 2 bcc, 3 single, 6 roters, 7 bcc12, 8 bcc48 (mod_crystals.f:164-172) */
  int32_t elastic_type; /* 1 isotropic, 2 cubic (mod_crystals.f:173) */
  int32_t h_type;       /* 1 voce, 2 mts (incrystal.f:305-331)       */
  int32_t alter_mode;   /* mm10_a.f:2073                             */
  int32_t miter;        /* mod_crystals.f:398                        */
  int32_t pad_;
  double e, nu, mu, harden_n, theta_0, tau_y, tau_v, voche_m, iD_v, eps_dot_0_y, k_0, burgers;
  double atol, atol1, rtol, rtol1;
  /* MTS hardening, h_type 2 (mm10_a.f:2109-2175, mm10_b.f:2080-2345; defaults
     mod_crystals.f:256-275, deck keywords incrystal.f:165-236) */
  double tau_a, tau_hat_y, g_0_y, tau_hat_v, g_0_v, p_y, q_y, p_v, q_v;
  double boltzman, eps_dot_0_v, mu_0, D_0, T_0;
} cpfft_crystal;

/* material table entry: matprp slots of inmat.f:97-133 (REAL*4 on purpose) / :176-298 */
typedef struct {
  int32_t type;     /* 1 bilinear (mm01), 10 crystal plasticity (mm10) */
  int32_t crystal;  /* cp: 1-based crystal number                      */
  float e, nu, beta, tan_e, yld_pt;
  int32_t n_crystals; /* cp: crystals per material point, imatprp(101) (inmat.f:201-204);
                         0 or 1 = one crystal, > 1 = Taylor average (mm10_a.f:112-197) */
} cpfft_material;

typedef struct {
  int32_t N;        /* grid points per edge ("number of grid", FFT_finite_3d.f:88-93) */
  int32_t device;   /* CUDA device ordinal                                            */
  int32_t rank, world;
  int32_t maxIter;  /* indypm.f:30-58 */
  int32_t pad_;
  double tolNR, tolPCG, tstep;
} cpfft_config;

/* device-resident fields addressable by id */
typedef enum {
  CPFFT_FN = 0, CPFFT_FN1, CPFFT_PN, CPFFT_PN1, CPFFT_DFM, CPFFT_B,   /* mod_fft.f:56-58, 9 comps */
  CPFFT_CG_P, CPFFT_CG_AP, CPFFT_CG_R,                                /* tmpPcg cols 1-3          */
  CPFFT_K4,                                                           /* 81 comps                  */
  CPFFT_URCS_N, CPFFT_URCS_N1,                                        /* 9  (mod_eleblocks.f:60-94)*/
  CPFFT_EPS_N, CPFFT_EPS_N1,                                          /* 6                         */
  CPFFT_ROT_N1,                                                       /* 9                         */
  CPFFT_HIST_N, CPFFT_HIST_N1,                                        /* cpfft_hist_size() comps   */
  CPFFT_CEP,                                                          /* 36, [D] of last sweep     */
  CPFFT_NUM_FIELDS
} cpfft_field;

typedef enum {
  CPFFT_LAYOUT_SOA = 0,   /* (ncomp, N3loc): the reference's column-major (N3, ncomp)            */
  CPFFT_LAYOUT_AOS = 1    /* (N3loc, ncomp): the per-block (nvals, ngp=1, span) storage, blocks
                             concatenated in voxel order (mod_eleblocks.f:60-94, dupstr.f:326-349) */
} cpfft_layout;

/* ---- life cycle: FFT_init / fftAllocate (FFT_init.f:16-258) ---- */
int  cpfft_create(const cpfft_config* cfg, cpfft_handle** out);
void cpfft_destroy(cpfft_handle* h);
const char* cpfft_last_error(const cpfft_handle* h);

/* ---- model data: what inmat/incrystal/inelem leave in matprp, c_array, matList ---- */
int cpfft_set_materials(cpfft_handle* h, int nmat, const cpfft_material* mats, int ncry,
                        const cpfft_crystal* crys);
/* matlist: 1-based material per local voxel; angles: (N3loc,3) Kocks degrees */
int cpfft_set_voxels(cpfft_handle* h, const int32_t* matlist, const double* angles_deg);
/* polycrystalline material points (materials with n_crystals > 1): what read_crystal_data
 * leaves in angle_input / crystal_input (mod_crystals.f:2111-2210).  angles_deg (N3loc, ncmax, 3),
 * crystal_ids (N3loc, ncmax) 1-based crystal numbers or NULL = the material's crystal_type
 * (crystal_input single); a voxel uses the first n_crystals entries of its material.  The
 * history then holds the common block followed by n_crystals per-crystal blocks
 * (mm10_set_sizes_special, mm10_a.f:640-641). */
int cpfft_set_voxels_taylor(cpfft_handle* h, const int32_t* matlist, int ncmax,
                            const double* angles_deg, const int32_t* crystal_ids);
int cpfft_set_params(cpfft_handle* h, double tolNR, double tolPCG, int maxIter, double tstep);
int cpfft_hist_size(const cpfft_handle* h);
int64_t cpfft_local_voxels(const cpfft_handle* h);

/* ---- hot path, one export per reference entry point ---- */
int cpfft_drive_eps_sig(cpfft_handle* h, int step, int iter);          /* drive_eps_sig.f:16    */
int cpfft_G_K_dF(cpfft_handle* h, cpfft_field src, cpfft_field dst, int flgK); /* G_K_dF.f:11   */
int cpfft_fftPcg(cpfft_handle* h, cpfft_field b, cpfft_field x, double tol,
                 int* iters, double* relres);                          /* FFT_nr3.f:214         */
int cpfft_tangent_homo(cpfft_handle* h, double C_homo[81]);            /* tangent_homo.f:11     */
int cpfft_mean_P(cpfft_handle* h, double Pbar[9]);                     /* FFT_nr3.f:127-134     */
int cpfft_update(cpfft_handle* h);                  /* update.f:75-106 + dcopy FFT_nr3.f:174-175:
                                                       n/n+1 history buffers exchanged, Fn/Pn copied */

/* The whole step loop of FFT_nr3 (FFT_nr3.f:14-200) with zero host traffic inside:
 * BC_all (nstep,9) row-major cumulative table (inlod.f:57-63), isNBC[9].
 * Outputs (may be NULL): nr_iters[nstep] Newton iterations per step, cg_iters[nstep*cg_cap]
 * CG iteration counts of the successive solves of the step (-1 terminated),
 * Pbar[nstep*9], seconds[3] = {pcg bucket, sig-eps bucket, total} (thyme.f buckets 1,2),
 * counters[5] = {G_K_dF applications, drive_eps_sig sweeps, CG iterations, mm10 local solver
 * failures over all sweeps, failures in the final (converged) sweeps of the steps}.
 * first_step > 1 continues a previous call (state is kept in the handle). */
int cpfft_FFT_nr3(cpfft_handle* h, int nstep, const double* BC_all, const int32_t* isNBC,
                  int32_t* nr_iters, int32_t* cg_iters, int cg_cap, double* Pbar,
                  double* seconds, int64_t* counters);

/* The lines the reference prints to unit `out` during FFT_nr3 (formats 1000-1003,
 * FFT_nr3.f:195-199: step banner, "Initial residual", "Iteration i residual", "Stress
 * iteration"), for the steps of the last cpfft_FFT_nr3 call; the host prints them verbatim. */
const char* cpfft_step_log(const cpfft_handle* h);
/* next_step: number of the load step the next cpfft_FFT_nr3 call starts with (the `step` loop
 * variable of FFT_nr3.f:51); cg_truncated: CG counts the last call could not return because
 * cg_cap was too small (0 = the cg_iters rows are complete). */
int cpfft_step_counter(const cpfft_handle* h, int* next_step, int* cg_truncated);

/* ---- host <-> device movement of whole fields ---- */
int cpfft_field_ncomp(const cpfft_handle* h, cpfft_field f);
int cpfft_upload(cpfft_handle* h, cpfft_field f, const double* host, cpfft_layout layout);
int cpfft_download(cpfft_handle* h, cpfft_field f, double* host, cpfft_layout layout);
int cpfft_download_fail_flags(cpfft_handle* h, int32_t* flags);   /* per voxel, mm10 local solve */
int cpfft_download_local_iters(cpfft_handle* h, int32_t* iters2); /* (N3loc,2) predictor/update  */
/* mm10 material_cut_step events (mm10_a.f:2811 ">>> Warning: mm10 implicit solution failed"):
 * the reference prints and carries on with an un-updated block; here the point keeps its n
 * state with the elastic tangent, and the events are counted (summed over ranks). */
int cpfft_material_failures(cpfft_handle* h, int64_t* total, int64_t* last_sweep);

/* ---- multi-GPU (new; the reference's mpi_code.f is all stubs) ---- */
int cpfft_nccl_unique_id(void* id128);                   /* rank 0: create, then broadcast */
int cpfft_nccl_init(cpfft_handle* h, const void* id128); /* all ranks                       */
/* how the slab <-> pencil transposes of G_K_dF run: 0 single GPU, 1 NCCL send/recv with
 * pack/unpack kernels, 2 peer-mapped stores fused into the FFT store stages (CUDA IPC) */
int cpfft_exchange_mode(const cpfft_handle* h);

/* ---- measurement helpers ---- */
int cpfft_synchronize(cpfft_handle* h);
void* cpfft_stream(cpfft_handle* h);                     /* cudaStream_t the kernels run on */
int64_t cpfft_kernel_launches(const cpfft_handle* h);    /* launches since create           */
/* CUDA events recorded on the launching stream around every kernel, summed per kernel class;
 * the analogue of the reference's thyme() buckets (thyme.f:15-47) at kernel granularity */
int cpfft_profile_enable(cpfft_handle* h, int on);
int cpfft_profile_reset(cpfft_handle* h);
int cpfft_profile_classes(void);
const char* cpfft_profile_name(int cls);
int cpfft_profile_get(cpfft_handle* h, int cls, double* ms, int64_t* count);
/* measured FP64 FMA issue peak of the device (TFLOP/s): roofline denominator of the material
 * update kernels (MEASURED_PEAKS.json has HBM and bf16 figures only) */
int cpfft_fp64_peak(cpfft_handle* h, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
