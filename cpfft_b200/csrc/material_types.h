// cpfft_b200: plain-data types of the material stage, shared by the CUDA translation units and the
// host build of the per-voxel code (tests/native/material_host.cpp).  No CUDA types here.
#pragma once
#include <stdint.h>

#define CPF_MAX_SLIP 48

// mm10 history layout (offsets, 0-based) -- mm10_d.f:137-331
struct CpfHistLayout {
  int use_max, nslip, num_hard;
  int cep, gradfe, R, work, slipsum;
  int c_stress, c_euler, c_Rp, c_D, c_eps, c_slipinc, c_tt, c_u, c_ttrate, c_ep, c_ed;
  int len_u, len_slip, total;
};

// per-material constants in device memory
struct CpfMatDev {
  int type, crystal;        // crystal: 0-based index into crystal table
  int ncry, hard;           // crystals per material point (imatprp(101)), > 1 = Taylor average; hardening law of its crystals (1 Voce, 2 MTS)
  double ym, nu, beta, tan_e, yld, hprime;  // mm01 (REAL*4 promoted, drive_eps_sig.f:486-521)
};

// per-crystal constants (Voce), device
struct CpfCryDev {
  int nslip, alter_mode, miter, rate_int;   // rate_int: harden_n-1 if small integer else -1
  double rate_n, theta_0, tau_y, tau_v, voche_m, iD_v, eps_dot_0_y, k_0, burgers;
  double atol, atol1, rtol, rtol1;
  // MTS (hard == 2): mu(T) = mu_0 - D_0 / (exp(T_0 / T) - 1); kby / kbv = boltz / (b^3 G_0_y / G_0_v),
  // so that boltz T / (mu b^3 G_0) = kb T / mu; inverse exponents of the thermal activation law
  int hard, pad_;
  double tau_a, mu_0, D_0, T_0, tau_hat_y, tau_hat_v, kby, kbv, iq_y, ip_y, iq_v, ip_v, p_y, q_y, p_v, q_v, eps_dot_0_v;
};

// per-grain (unique crystal+orientation) table entry, device, doubles:
//   g (9, row-major), rotated stiffness C (36, row-major 6x6), per slip system ms0[6] (engineering-shear Schmid
//   vector) + qs0[3] (skew vector), the Kocks angles in degrees (3).
// CPF_TAB_VEC = 1: C and the slip entries start on 16-byte boundaries and a slip entry is padded to 10 doubles, so the
// kernels fetch them with 16-byte loads (half the load instructions of the hot loops); 0: packed layout, 8-byte loads.
#ifndef CPF_TAB_VEC
#define CPF_TAB_VEC 1
#endif
#define CPF_GRAIN_G 0
#if CPF_TAB_VEC
#define CPF_SLIP_STRIDE 10
#define CPF_GRAIN_C 10
#define CPF_GRAIN_B 46
#define CPF_GRAIN_ANG (CPF_GRAIN_B + CPF_SLIP_STRIDE * CPF_MAX_SLIP)
#define CPF_GRAIN_STRIDE (CPF_GRAIN_ANG + 4)     // even: every entry of the table starts 16-byte aligned
#else
#define CPF_SLIP_STRIDE 9
#define CPF_GRAIN_C 9
#define CPF_GRAIN_B 45
#define CPF_GRAIN_ANG (CPF_GRAIN_B + CPF_SLIP_STRIDE * CPF_MAX_SLIP)   // Kocks angles (degrees) of the grain
#define CPF_GRAIN_STRIDE (CPF_GRAIN_ANG + 3)
#endif


// mm10_d.f:137-331
inline CpfHistLayout cpf_hist_layout(int nslip, int num_hard) {
  CpfHistLayout L;
  L.use_max = (num_hard == 48 || nslip == 48) ? 1 : 0;
  L.nslip = nslip; L.num_hard = num_hard;
  const int lc5 = L.use_max ? 48 : nslip;
  L.cep = 0; L.gradfe = 36; L.R = 63; L.work = 72; L.slipsum = 75;
  const int common = 75 + lc5;
  const int l6 = L.use_max ? 48 : nslip, l7 = L.use_max ? 48 : num_hard, l8 = L.use_max ? 48 : 15,
            l9 = L.use_max ? 48 : num_hard;
  L.len_slip = l6; L.len_u = l8;
  L.c_stress = common; L.c_euler = L.c_stress + 6; L.c_Rp = L.c_euler + 3; L.c_D = L.c_Rp + 9;
  L.c_eps = L.c_D + 6; L.c_slipinc = L.c_eps + 6; L.c_tt = L.c_slipinc + l6; L.c_u = L.c_tt + l7;
  L.c_ttrate = L.c_u + l8; L.c_ep = L.c_ttrate + l9; L.c_ed = L.c_ep + 6;
  L.total = L.c_ed + 6;
  return L;
}

