// cpfft_b200: bilinear (mixed isotropic/kinematic hardening) Mises model, one thread per
// voxel.  Device-side replacement of mm01 + mm01_init + mm01_simple1 + mm01_sig_final +
// mm01_plastic_work (mm01.f:28-784) and the consistent tangent cnst1 (mm01.f:1222-1374),
// isothermal path (dtemps == 0 in this code base, rstgp1.f:513).
// The numerical constants are the reference's literals, including the truncated ones.
#pragma once
#include "kin.cuh"
#include "material_types.h"

CPF_DI double mm01_state_word(int s) { return CPF_LL2D((long long)(unsigned)s); }
CPF_DI int mm01_state_of(double d) { return (int)(CPF_D2LL(d) & 0xffffffffLL); }

// hn[11]: history at n (re-initialised when step == 1, mm01.f:141-144); sn[9]: urcs at n;
// de[6]: unrotated strain increment; outputs s1[9], h1[11], cep[36] (row-major, symmetric).
CPF_DI void mm01_update(int step, const CpfMatDev& mp, double* hn, double* sn, const double* de,
                        double* s1, double* h1, double* cep) {
  const double ym = mp.ym, nu = mp.nu, beta = mp.beta, hp = mp.hprime, yld = mp.yld;
  const double root2 = 1.414213562373095;
  if (step == 1) {  // mm01_set_history
    const double kn = yld / 1.73205080756888;
    hn[0] = 0.0; hn[1] = kn; hn[2] = 0.0; hn[3] = mm01_state_word(3); hn[4] = hp;
#pragma unroll
    for (int k = 5; k < 11; ++k) hn[k] = 0.0;
    sn[7] = 0.0; sn[8] = 0.0;
  }
  // trial state (mm01_init)
  const double dvol = de[0] + de[1] + de[2];
  const double em = dvol / 3.0;
  const double gn = ym / 2.0 / (1.0 + nu);
  const double ee1 = (sn[0] - nu * (sn[1] + sn[2])) / ym;
  const double ee2 = (sn[1] - nu * (sn[0] + sn[2])) / ym;
  const double ee3 = (sn[2] - nu * (sn[0] + sn[1])) / ym;
  const double evol1 = ee1 + ee2 + ee3 + dvol;
  const double emn = (ee1 + ee2 + ee3) / 3.0;
  double e[6];
  e[0] = (ee1 - emn) + (de[0] - em);
  e[1] = (ee2 - emn) + (de[1] - em);
  e[2] = (ee3 - emn) + (de[2] - em);
  e[3] = sn[3] / gn + de[3];
  e[4] = sn[4] / gn + de[4];
  e[5] = sn[5] / gn + de[5];
  const double G = ym / (2.0 * (1.0 + nu));
  double dse[6], rt[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) dse[k] = 2.0 * G * e[k];
#pragma unroll
  for (int k = 3; k < 6; ++k) dse[k] = G * e[k];
  const double hbi = beta * hp, hbk = (1.0 - beta) * hp, hbkn = (1.0 - beta) * hn[4];
  const double kbar = (yld + hbi * hn[2]) / 1.7320508075688;  // sic (mm01.f:362)
  double lk = 1.0;
  if (fabs(hbkn) > 0.000001) lk = hbk / hbkn;
#pragma unroll
  for (int k = 0; k < 6; ++k) rt[k] = dse[k] - hn[5 + k] * lk;
  double mr = sqrt(rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2] +
                   2.0 * (rt[3] * rt[3] + rt[4] * rt[4] + rt[5] * rt[5]));
  const double yf = mr - root2 * kbar;
#pragma unroll
  for (int k = 0; k < 6; ++k) rt[k] = dse[k] - hn[5 + k];
  mr = sqrt(rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2] +
            2.0 * (rt[3] * rt[3] + rt[4] * rt[4] + rt[5] * rt[5]));
  const bool yield = yf >= 0.0000001 * root2 * kbar;
  double dev[6];
  if (yield) {  // mm01_simple1
    const double ldt = (mr - root2 * kbar) / (0.666666666666667 * (3.0 * G + hp));
    const double k1 = kbar + (root2 / 3.0) * hbi * ldt;
    h1[0] = ldt; h1[1] = k1; h1[2] = hn[2] + ldt * 0.816496580927; h1[4] = hp;
    const double c1 = 0.666666666666667 * hbk * ldt / mr, c2 = root2 * k1 / mr;
#pragma unroll
    for (int k = 0; k < 6; ++k) { h1[5 + k] = hn[5 + k] + c1 * rt[k]; dev[k] = h1[5 + k] + c2 * rt[k]; }
  } else {
    h1[0] = 0.0; h1[1] = kbar; h1[2] = hn[2]; h1[4] = hp;
#pragma unroll
    for (int k = 0; k < 6; ++k) { h1[5 + k] = hn[5 + k] * lk; dev[k] = rt[k] + hn[5 + k]; }
  }
  h1[3] = mm01_state_word(yield ? 1 : 3);
  // mm01_sig_final
  const double sm = evol1 * (3.0 * ym * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)) + 2.0 * G) / 3.0;
  s1[0] = dev[0] + sm; s1[1] = dev[1] + sm; s1[2] = dev[2] + sm;
  s1[3] = dev[3]; s1[4] = dev[4]; s1[5] = dev[5];
  s1[6] = sn[6] + 0.5 * (de[0] * (s1[0] + sn[0]) + de[1] * (s1[1] + sn[1]) + de[2] * (s1[2] + sn[2]) +
                         de[3] * (s1[3] + sn[3]) + de[4] * (s1[4] + sn[4]) + de[5] * (s1[5] + sn[5]));
  // mm01_plastic_work
  s1[7] = sn[7]; s1[8] = sn[8];
  if (yield) {
    double ds[6], dp[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) ds[k] = s1[k] - sn[k];
    dp[0] = de[0] - (ds[0] - nu * (ds[1] + ds[2])) / ym;
    dp[1] = de[1] - (ds[1] - nu * (ds[0] + ds[2])) / ym;
    dp[2] = de[2] - (ds[2] - nu * (ds[0] + ds[1])) / ym;
    dp[3] = de[3] - ds[3] / G; dp[4] = de[4] - ds[4] / G; dp[5] = de[5] - ds[5] / G;
    s1[7] = sn[7] + 0.5 * (dp[0] * (s1[0] + sn[0]) + dp[1] * (s1[1] + sn[1]) + dp[2] * (s1[2] + sn[2]) +
                           dp[3] * (s1[3] + sn[3]) + dp[4] * (s1[4] + sn[4]) + dp[5] * (s1[5] + sn[5]));
    const double f1 = (dp[0] - dp[1]) * (dp[0] - dp[1]) + (dp[1] - dp[2]) * (dp[1] - dp[2]) +
                      (dp[0] - dp[2]) * (dp[0] - dp[2]);
    const double f2 = dp[3] * dp[3] + dp[4] * dp[4] + dp[5] * dp[5];
    s1[8] = sn[8] + (root2 / 3.0) * sqrt(f1 + (3.0 / 2.0) * f2);
  }
  // consistent tangent (cnst1); root2 there is the literal 1.414213562
#pragma unroll
  for (int k = 0; k < 36; ++k) cep[k] = 0.0;
  if (!yield) {
    const double c1 = ym / ((1.0 + nu) * (1.0 - 2.0 * nu));
    const double c2 = (1.0 - nu) * c1, c3 = ((1.0 - 2.0 * nu) / 2.0) * c1, c4 = nu * c1;
    cep[0] = cep[7] = cep[14] = c2;
    cep[21] = cep[28] = cep[35] = c3;
    cep[1] = cep[2] = cep[6] = cep[12] = cep[8] = cep[13] = c4;
  } else {
    const double l = (ym * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu));
    const double kb = (3.0 * l + 2.0 * G) / 3.0;
    const double mq = rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2] +
                      2.0 * (rt[3] * rt[3] + rt[4] * rt[4] + rt[5] * rt[5]);
    const double bb = (1.414213562 * h1[1] + (2.0 / 3.0) * (1.0 - beta) * h1[4] * h1[0]) / sqrt(mq);
    const double gam = 1.0 / (1.0 + h1[4] / (3.0 * G));
    const double gbar = G * bb, albar = kb - 2.0 * gbar / 3.0, thbar = 2.0 * G * (gam - 1.0 + bb);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double v = -(thbar * rt[j] * rt[i] / mq);
        if (i == j) v = ((i < 3) ? (albar + 2.0 * gbar) : gbar) - thbar * (rt[i] * rt[i]) / mq;
        else if (i < 3) v = albar - thbar * rt[j] * rt[i] / mq;
        cep[6 * i + j] = v; cep[6 * j + i] = v;
      }
  }
}
