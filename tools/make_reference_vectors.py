#!/usr/bin/env python
"""Golden vectors produced by the reference's OWN source: leaf routines of /root/reference/src/*.f executed by the
Fortran-subset interpreter tools/fortran_subset.py on seeded inputs, saved as tests/golden/reference_vectors.npz.

    python tools/make_reference_vectors.py            # needs /root/reference (this container); writes the fixture

tests/test_reference_vectors.py (which needs neither /root/reference nor this script) holds the oracle, the numpy
restatements of tests/py_mm10.py and the spectral restatement to these outputs.  This is the pin the oracle has: the
reference cannot be compiled here (ifort + MKL), but these routines of it can be run, statement by statement, in the
order and precision the source states.  Routines (reference file:line of the subroutine statement):

  polar.f:18    rtcmp1  (+ irscp1 :58, ivcmp1 :133, evcmp1_new :226)   R of F = R U, closed form        -> K2
  polar.f:680   getrm1                                                   6x6 rotation operators, opt 1-3 -> K3
  cep2A.f:86    cep2A_a (+ multiply33, transpose33, ddot44, det33)       dP/dF from [D], sigma, F        -> K4
  mm10_a.f:1287 mm10_rotation_matrix                                     Kocks angles -> g               -> M2
  mm10_a.f:1400 mm10_rt2rve, :1461 mm10_rt2rvw                           slip-vector rotation operators  -> M4
  mm10_b.f      mm10_symSW                                               sym(S W) in Voigt form          -> M5
  FFT_init.f:272 formG                                                   Green operator table, odd N     -> G3
  G_K_dF.f:241  ddot42n                                                  K4 : x with its summation tree  -> G1
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import fortran_subset as F  # noqa: E402

REF = "/root/reference/src/"
FILES = ["param_def", "mod_crystals.f", "polar.f", "cep2A.f", "mm10_a.f", "mm10_b.f", "FFT_init.f", "G_K_dF.f"]


def interpreter():
    it = F.Interpreter()
    it.add_constants(open(REF + "param_def").read())
    mc = open(REF + "mod_crystals.f").read()
    i0 = mc.index("      module mm10_constants")
    it.add_constants(mc[i0:mc.index("      end module", i0)])
    for f in FILES[2:]:
        it.load(open(REF + f).read())
    return it


def rand_rotation(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    return q * np.sign(np.linalg.det(q))


def main():
    it = interpreter()
    mx = it.consts["mxvl"]
    rng = np.random.default_rng(20240607)
    out = {}

    # ---- rtcmp1: small / moderate / large stretches, with and without a superposed rotation
    Fs, amps = [], []
    for amp in (1e-4, 1e-3, 1e-2, 5e-2, 0.2, 0.5):
        for rot in (False, True):
            for _ in range(2):
                Fm = np.eye(3) + amp * rng.standard_normal((3, 3))
                Fs.append(rand_rotation(rng) @ Fm if rot else Fm)
                amps.append(amp)
    Fs = np.array(Fs)
    out["polar_amp"] = np.array(amps)
    n = len(Fs)
    fb = np.zeros((mx, 3, 3), order="F"); rb = np.zeros((mx, 3, 3), order="F")
    fb[:n] = Fs
    it.call("rtcmp1", n, fb, rb)
    out["polar_F"], out["polar_R"] = Fs, rb[:n].copy()

    # ---- getrm1: the three operator flavours for the rotations just computed
    qs = []
    for opt in (1, 2, 3):
        qb = np.zeros((mx, 6, 6), order="F")
        it.call("getrm1", n, qb, rb, opt)
        qs.append(qb[:n].copy())
    out["getrm1_q"] = np.array(qs)                       # (3 opts, n, 6, 6)

    # ---- cep2A_a
    cases = []
    out["cep2A_amp"] = np.repeat([1e-3, 2e-2, 0.1], 3)
    for amp in (1e-3, 2e-2, 0.1):
        for _ in range(3):
            Fn = np.eye(3) + amp * rng.standard_normal((3, 3))
            Fn1 = Fn + 0.3 * amp * rng.standard_normal((3, 3))
            t6 = 300.0 * rng.standard_normal(6)
            A = rng.standard_normal((6, 6)); C66 = 1e4 * (A @ A.T) + 8e4 * np.eye(6)
            Fnh = 0.5 * (Fn + Fn1)
            fb[:] = 0; rb[:] = 0
            fb[0], fb[1] = Fnh, Fn1
            it.call("rtcmp1", 2, fb, rb)
            Rh, R = rb[0].copy(), rb[1].copy()
            t33 = np.array([[t6[0], t6[3], t6[5]], [t6[3], t6[1], t6[4]], [t6[5], t6[4], t6[2]]])
            dPdF = np.zeros(81)
            it.call("cep2a_a", np.asfortranarray(Fn), np.asfortranarray(t33), np.asfortranarray(C66), np.asfortranarray(Rh),
                    float(np.linalg.det(Fnh)), np.asfortranarray(np.linalg.inv(Fnh)), np.asfortranarray(R), np.asfortranarray(Fn1),
                    np.asfortranarray(np.linalg.inv(Fn1)), float(np.linalg.det(Fn1)), dPdF)
            cases.append((Fn, Fn1, t6, C66, dPdF))
    for k, name in enumerate(("cep2A_Fn", "cep2A_Fn1", "cep2A_t6", "cep2A_C66", "cep2A_dPdF")):
        out[name] = np.array([c[k] for c in cases])

    # ---- mm10 rotation helpers
    rots = np.array([rand_rotation(rng) for _ in range(6)])
    rve, rvw = [], []
    for Q in rots:
        a, b = np.zeros((6, 6), order="F"), np.zeros((3, 3), order="F")
        it.call("mm10_rt2rve", np.asfortranarray(Q), a); it.call("mm10_rt2rvw", np.asfortranarray(Q), b)
        rve.append(a.copy()); rvw.append(b.copy())
    out["mm10_rt"], out["mm10_rt2rve"], out["mm10_rt2rvw"] = rots, np.array(rve), np.array(rvw)
    angs = np.array([[0, 0, 0], [30, 45, 60], [90, 90, 90], [12.5, 133.0, 271.25], [359, 1, 180], [45, 0, 45], [200, 70, 10], [77, 150, 300]], dtype=float)
    gs = []
    for a3 in angs:
        g = np.zeros((3, 3), order="F")
        it.call("mm10_rotation_matrix", a3.copy(), "kocks", "degrees", g, 6)
        gs.append(g.copy())
    out["kocks_angles"], out["kocks_g"] = angs, np.array(gs)
    S, W, SW = rng.standard_normal((6, 6)) * 100.0, rng.standard_normal((6, 3)) * 1e-3, []
    for s6, w3 in zip(S, W):
        sw = np.zeros(6)
        it.call("mm10_symsw", s6.copy(), w3.copy(), sw)
        SW.append(sw.copy())
    out["symsw_s"], out["symsw_w"], out["symsw_sw"] = S, W, np.array(SW)

    # ---- formG for odd grids (the sizes at which the reference's table is a projection)
    for N in (3, 5, 7):
        G = np.zeros((N ** 3, 81), order="F")
        it2 = F.Interpreter(it.consts); it2.units = it.units
        it2.module_vars = dict(n=N, nhalf=(N + 1) // 2, n3=N ** 3, ndim1=3, ndim2=9, ghat4=G)     # FFT_init.f:146
        it2.call("formg")
        out[f"formG_{N}"] = np.ascontiguousarray(G)

    # ---- ddot42n
    nv = 7
    A4 = np.asfortranarray(rng.standard_normal((nv, 81)) * 1e5); B2 = np.asfortranarray(rng.standard_normal((nv, 9)) * 1e-3)
    C2, tmp = np.zeros((nv, 9), order="F"), np.zeros((nv, 9), order="F")
    it.call("ddot42n", A4, B2, C2, tmp, nv)
    out["ddot42_A4"], out["ddot42_B2"], out["ddot42_C2"] = np.ascontiguousarray(A4), np.ascontiguousarray(B2), np.ascontiguousarray(C2)

    prov = "; ".join(f"{f} sha256 {hashlib.sha256(open(REF + f, 'rb').read()).hexdigest()[:16]}" for f in FILES)
    out["provenance"] = np.array("maranGit/CPFFT src: " + prov + "; executed by tools/fortran_subset.py (tools/make_reference_vectors.py)")
    dst = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
