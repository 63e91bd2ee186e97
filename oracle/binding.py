"""ctypes binding of the CPU oracle (oracle/oracle.h).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def oracle_lib_path() -> str:
    return os.path.join(_HERE, "build", "liboracle_cpfft.so")


def build_oracle(force: bool = False) -> str:
    """Compile the C++ restatement with the Makefile next to this file."""
    path = oracle_lib_path()
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h", ".inc"))]
    stale = (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return path


def _lib():
    global _LIB
    if _LIB is None:
        path = oracle_lib_path()
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, ip, dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_taylor.argtypes = [C.c_void_p, C.c_int, dp, ip]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_hist_size.argtypes = [C.c_void_p]
        for name in ("Fn", "Fn1", "Pn", "Pn1", "K4", "dFm", "b", "hist_n", "hist_n1", "urcs_n",
                     "urcs_n1", "eps_n", "eps_n1", "rot_n1"):
            f = getattr(L, "orc_" + name)
            f.restype = dp
            f.argtypes = [C.c_void_p]
        L.orc_fail_flags.restype = ip
        L.orc_fail_flags.argtypes = [C.c_void_p]
        L.orc_local_iters.restype = ip
        L.orc_local_iters.argtypes = [C.c_void_p]
        L.orc_drive_eps_sig.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_G_K_dF.argtypes = [C.c_void_p, dp, dp, C.c_int]
        L.orc_fftPcg.argtypes = [C.c_void_p, dp, dp, C.c_double, ip, dp]
        L.orc_fftPcg_capped.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_int, ip, dp]
        L.orc_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), dp]
        L.orc_tangent_homo.argtypes = [C.c_void_p, dp]
        L.orc_update.argtypes = [C.c_void_p]
        L.orc_mean_P.argtypes = [C.c_void_p, dp]
        L.orc_FFT_nr3.argtypes = [C.c_void_p, C.c_int, dp, ip, ip, ip, C.c_int, dp, dp,
                                  C.POINTER(C.c_int64)]
        L.orc_FFT_nr3_from.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, ip, ip, ip, C.c_int, dp, dp,
                                       C.POINTER(C.c_int64)]
        L.orc_rtcmp1.argtypes = [dp, dp]
        L.orc_getrm1.argtypes = [dp, C.c_int, dp]
        L.orc_ddot42_point.argtypes = [dp, dp, dp]
        L.orc_set_polar_precision.argtypes = [C.c_int]
        L.orc_cep2A.argtypes = [dp, dp, dp, dp, dp]
        L.orc_point_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp]
        L.orc_formG_entry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_crystal_stiffness.argtypes = [C.c_void_p, dp]
        L.orc_slip_table.argtypes = [C.c_int, ip, dp, dp]
        L.orc_mm10_residual_jacobian.argtypes = [C.c_void_p, dp, dp, C.c_double, dp, dp, C.c_double, dp, dp]
        _LIB = L
    return _LIB


_DP = C.POINTER(C.c_double)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Oracle:
    """CPU restatement of the reference hot path for one ``Problem`` (cpfft_b200.problem)."""

    CG_CAP = 256

    def __init__(self, prob, threads: int = 0, polar: str = "quad"):
        """polar: arithmetic of the reference's polar decomposition, process-wide: "quad" (default) evaluates
        polar.f's closed form in __float128 -- the value the algorithm defines; "double" is the literal
        restatement with the reference's own small-strain round-off noise (<= ~3e-8 on R)."""
        L = _lib()
        L.orc_set_polar_precision(1 if polar == "quad" else 0)
        self.L, self.prob = L, prob
        if threads:
            L.orc_set_threads(threads)
        mats, crys = prob.material_pods(), prob.crystal_pods()
        ml = np.ascontiguousarray(prob.matlist, dtype=np.int32)
        nc, ang, ids = prob.taylor_tables()
        ang1 = np.ascontiguousarray(ang[:, 0, :])
        self.h = L.orc_create(prob.N, len(prob.materials), C.addressof(mats), len(prob.crystals),
                              C.addressof(crys), _ip(ml), _dp(ang1))
        if prob.taylor:
            rc = L.orc_set_taylor(self.h, nc, _dp(ang), _ip(ids) if ids is not None else None)
            if rc:
                raise ValueError(f"orc_set_taylor: inconsistent crystal tables (code {rc})")
        L.orc_set_params(self.h, prob.tolNR, prob.tolPCG, prob.maxIter, prob.tstep)
        self.N, self.N3 = prob.N, prob.N3
        self.H = L.orc_hist_size(self.h)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def _view(self, name, shape):
        ptr = getattr(self.L, "orc_" + name)(self.h)
        return np.ctypeslib.as_array(ptr, shape=shape)

    # SoA fields: (ncomp, N3) views of the reference's column-major (N3, ncomp) arrays
    @property
    def Fn(self): return self._view("Fn", (9, self.N3))
    @property
    def Fn1(self): return self._view("Fn1", (9, self.N3))
    @property
    def Pn1(self): return self._view("Pn1", (9, self.N3))
    @property
    def K4(self): return self._view("K4", (81, self.N3))
    @property
    def dFm(self): return self._view("dFm", (9, self.N3))
    @property
    def hist_n(self): return self._view("hist_n", (self.N3, self.H))
    @property
    def hist_n1(self): return self._view("hist_n1", (self.N3, self.H))
    @property
    def urcs_n(self): return self._view("urcs_n", (self.N3, 9))
    @property
    def urcs_n1(self): return self._view("urcs_n1", (self.N3, 9))
    @property
    def eps_n1(self): return self._view("eps_n1", (self.N3, 6))
    @property
    def rot_n1(self): return self._view("rot_n1", (self.N3, 9))
    @property
    def local_iters(self):
        return np.ctypeslib.as_array(self.L.orc_local_iters(self.h), shape=(self.N3, 2))

    def drive_eps_sig(self, step, it):
        return self.L.orc_drive_eps_sig(self.h, step, it)

    def G_K_dF(self, F, flgK):
        F = np.ascontiguousarray(F, dtype=np.float64)
        out = np.empty_like(F)
        self.L.orc_G_K_dF(self.h, _dp(F), _dp(out), int(bool(flgK)))
        return out

    def fftPcg(self, b, tol):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros_like(b)
        it = C.c_int32(0)
        rr = C.c_double(0)
        rc = self.L.orc_fftPcg(self.h, _dp(b), _dp(x), tol, C.byref(it), C.byref(rr))
        return rc, x, it.value, rr.value

    def tangent_homo(self):
        Ch = np.zeros(81)
        rc = self.L.orc_tangent_homo(self.h, _dp(Ch))
        return rc, Ch

    def update(self):
        self.L.orc_update(self.h)

    def mean_P(self):
        p = np.zeros(9)
        self.L.orc_mean_P(self.h, _dp(p))
        return p

    def FFT_nr3(self, nstep=None):
        prob = self.prob
        bc = prob.BC_all()
        nstep = prob.nstep if nstep is None else nstep
        bc = np.ascontiguousarray(bc[:nstep])
        nbc = np.ascontiguousarray(prob.isNBC, dtype=np.int32)
        nr = np.zeros(nstep, dtype=np.int32)
        cg = np.full((nstep, self.CG_CAP), -1, dtype=np.int32)
        pbar = np.zeros((nstep, 9))
        buckets = np.zeros(3)
        counters = np.zeros(5, dtype=np.int64)
        rc = self.L.orc_FFT_nr3(self.h, nstep, _dp(bc), _ip(nbc), _ip(nr), _ip(cg), self.CG_CAP,
                                _dp(pbar), _dp(buckets), counters.ctypes.data_as(C.POINTER(C.c_int64)))
        cg_lists = [list(row[:list(row).index(-1)]) if -1 in row else list(row) for row in cg]
        return dict(rc=rc, nr_iters=nr, cg_iters=cg_lists, Pbar=pbar, buckets=buckets,
                    counters=counters)

    # unit probes
    @staticmethod
    def rtcmp1(F):
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(9)
        R = np.zeros(9)
        _lib().orc_rtcmp1(_dp(F), _dp(R))
        return R.reshape(3, 3)

    @staticmethod
    def kinematics_probe(Fn, Fn1, ur6):
        a = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (Fn, Fn1, ur6)]
        R, uddt, P = np.zeros(9), np.zeros(6), np.zeros(9)
        L = _lib()
        L.orc_kinematics_probe.argtypes = [_DP] * 6
        L.orc_kinematics_probe(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(R), _dp(uddt), _dp(P))
        return R.reshape(3, 3), uddt, P

    @staticmethod
    def getrm1(R, opt):
        R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
        q = np.zeros(36)
        _lib().orc_getrm1(_dp(R), int(opt), _dp(q))
        return q.reshape(6, 6)

    @staticmethod
    def ddot42_point(A81, B9):
        a, b = (np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (A81, B9))
        c = np.zeros(9)
        _lib().orc_ddot42_point(_dp(a), _dp(b), _dp(c))
        return c

    @staticmethod
    def cep2A(Fn, Fn1, t6, cep):
        a = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (Fn, Fn1, t6)]
        c = np.asfortranarray(cep, dtype=np.float64).ravel(order="F").copy()
        A = np.zeros(81)
        _lib().orc_cep2A(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(c), _dp(A))
        return A

    def point_update(self, voxel, step, it, Fn, Fn1):
        a = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (Fn, Fn1)]
        P, A = np.zeros(9), np.zeros(81)
        self.L.orc_point_update(self.h, voxel, step, it, _dp(a[0]), _dp(a[1]), _dp(P), _dp(A))
        return P, A

    @staticmethod
    def formG_entry(N, i, j, k):
        G = np.zeros(81)
        _lib().orc_formG_entry(N, i, j, k, _dp(G))
        return G

    @staticmethod
    def crystal_stiffness(crystal):
        pod = crystal.pod()
        Cm = np.zeros(36)
        _lib().orc_crystal_stiffness(C.addressof(pod), _dp(Cm))
        return Cm.reshape(6, 6, order="F")

    @staticmethod
    def mm10_residual_jacobian(crystal, angles, D6, dt, x7, n_stress, n_tt):
        """R(x) (7,) and J(x) (7,7) of the local Newton system of one crystal (mm10_formR / mm10_formJ)"""
        pod = crystal.pod()
        a = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (angles, D6, x7, n_stress)]
        R, J = np.zeros(7), np.zeros(49)
        _lib().orc_mm10_residual_jacobian(C.addressof(pod), _dp(a[0]), _dp(a[1]), float(dt), _dp(a[2]), _dp(a[3]),
                                          float(n_tt), _dp(R), _dp(J))
        return R, J.reshape(7, 7)

    @staticmethod
    def mm10_residual_jacobian_rot(crystal, angles, D6, dt, x7, n_stress, n_tt, Rp, R):
        """as mm10_residual_jacobian, with the plastic rotation Rp_n and the polar rotation R of the step; also returns
        the current Schmid vectors ms (nslip, 6), qs, qc (nslip, 3) of mm10_setup"""
        pod = crystal.pod()
        a = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (angles, D6, x7, n_stress, Rp, R)]
        Rv, J = np.zeros(7), np.zeros(49)
        ms, qs, qc = np.zeros(48 * 6), np.zeros(48 * 3), np.zeros(48 * 3)
        L = _lib()
        L.orc_mm10_residual_jacobian_rot.argtypes = [C.c_void_p, _DP, _DP, C.c_double, _DP, _DP, C.c_double, _DP, _DP, _DP, _DP, _DP, _DP, _DP]
        L.orc_mm10_residual_jacobian_rot(C.addressof(pod), _dp(a[0]), _dp(a[1]), float(dt), _dp(a[2]), _dp(a[3]), float(n_tt),
                                         _dp(a[4]), _dp(a[5]), _dp(Rv), _dp(J), _dp(ms), _dp(qs), _dp(qc))
        return Rv, J.reshape(7, 7), ms.reshape(48, 6), qs.reshape(48, 3), qc.reshape(48, 3)

    @staticmethod
    def mm10_crystal_probe(crystal, angles, dt, R, D6, it, n_state):
        """the whole update of one crystal from an explicit n state (orc_mm10_crystal_probe): dict of the n+1 state"""
        pod = crystal.pod()
        a = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (angles, R, D6, n_state)]
        out, iters = np.zeros(137), np.zeros(2, dtype=np.int32)
        L = _lib()
        L.orc_mm10_crystal_probe.argtypes = [C.c_void_p, _DP, C.c_double, _DP, _DP, C.c_int, _DP, _DP, C.POINTER(C.c_int32)]
        fail = L.orc_mm10_crystal_probe(C.addressof(pod), _dp(a[0]), float(dt), _dp(a[1]), _dp(a[2]), int(it), _dp(a[3]), _dp(out), _ip(iters))
        return dict(fail=int(fail), iters=iters.copy(), stress=out[:6], tt=out[6], tt_rate=out[7], tangent=out[8:44].reshape(6, 6),
                    Rp=out[44:53].reshape(3, 3), euler=out[53:56], eps=out[56:62], slip_incs=out[62:110], u=out[110:125],
                    ep=out[125:131], ed=out[131:137])

    @staticmethod
    def slip_table(slip_type):
        n = C.c_int32(0)
        b, nn = np.zeros(48 * 3), np.zeros(48 * 3)
        _lib().orc_slip_table(slip_type, C.byref(n), _dp(b), _dp(nn))
        return b[:3 * n.value].reshape(-1, 3), nn[:3 * n.value].reshape(-1, 3)
