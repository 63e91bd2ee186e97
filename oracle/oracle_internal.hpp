// CPFFT ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Internal declarations shared by the oracle translation units.
#pragma once
#include "oracle.h"
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdio>

namespace orc {

typedef double M33[3][3];
typedef double M66[6][6];

// ---- kinematics (polar.f, drive_eps_sig.f helpers) ----
void rtcmp1(const M33 f, M33 r);                       // polar.f:18-45
void set_polar_precision(int quad);                    // 0: literal double arithmetic, 1: __float128 (default)
int get_polar_precision();
void getrm1(M66 q, const M33 r, int opt);              // polar.f:680-802
void qmply1(const M66 q, const double* m1, double* m2);// qmply1.f:15-36
void inv33(const M33 jac, M33 gama, double* dj);       // drive_eps_sig.f:1017-1107
void mul33(const M33 a, const M33 b, double* c6);      // drive_eps_sig.f:1110-1166
void cs2p(const double* cs, const M33 finv, double detF, double* P9); // drive_eps_sig.f:1182-1224
void cep2A_a(const M33 Fn, const M33 t, const M66 C, const M33 Rh, double detFnh,
             const M33 fnhinv, const M33 R, const M33 Fn1, const M33 finv, double detFn1,
             double* dPdF81);                          // cep2A.f:86-284

// ---- mm01 (mm01.f) ----
struct Mm01Props { double ym, nu, beta, tan_e, yld, hprime; };
// history(11), cgn(9) -> cgn1(9), history1(11), cep(6,6); returns yield flag
void mm01_point(int step, const Mm01Props& p, double* history /*may be reset at step 1*/,
                double* cgn /*slots 8,9 reset at step 1*/, const double* deps, double* cgn1,
                double* history1, M66 cep);

// ---- mm10 (mm10_a.f / mm10_b.f), Voce ----
struct CrystalLib {
  orc_crystal in;
  int nslip;
  double bi[ORC_MAX_SLIP][3], ni[ORC_MAX_SLIP][3];
  M66 elast_stiff;
};
void finalize_crystal(CrystalLib& c);                  // mod_crystals.f:414-1931

struct HistLayout {                                    // mm10_d.f:25-360
  int use_max, nslip, num_hard;
  int cep, gradfe, R, work, slipsum;                   // 0-based starts of common terms
  int c_stress, c_euler, c_Rp, c_D, c_eps, c_slipinc, c_tt, c_u, c_ttrate, c_ep, c_ed;
  int len_u, len_slip, total;
};
HistLayout mm10_history_layout(int nslip_max, int num_hard_max);

// one point with `ncrystals` crystals (Taylor average, mm10_a.f:112-197); crystals[c] and
// angles_deg[3 c ..] describe crystal c; the history holds the common block followed by
// ncrystals per-crystal blocks.  returns 0 ok, 1 = material_cut_step (a local solve failed)
int mm10_point(int step, int iter, int ncrystals, const CrystalLib* const* crystals, const double* angles_deg,
               const HistLayout& L, double dt, const double* rot_n1_colmajor,
               const double* uddt, double* history_n, double* history_np1,
               const double* urcs_n, double* urcs_n1, int* local_iters /*[2], summed over crystals*/);

} // namespace orc
