"""Host-side mirror of the reference's hot-path entry points on top of the C ABI
(include/cpfft_b200.h).  Method names and argument meaning follow the Fortran subroutines:
``drive_eps_sig(step, iiter)`` (drive_eps_sig.f:16), ``G_K_dF(F, GKF, flgK)`` (G_K_dF.f:11),
``fftPcg(b, x, tol)`` (FFT_nr3.f:214), ``tangent_homo`` (tangent_homo.f:11), ``update``
(update.f:75) and ``FFT_nr3`` (FFT_nr3.f:14).  Errors the reference turns into
``die_abort`` are raised as :class:`CpfftError` carrying the same message.

There is no CPU fallback: if the CUDA library is missing or no GPU is present the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .problem import Problem

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIELDS = ["FN", "FN1", "PN", "PN1", "DFM", "B", "CG_P", "CG_AP", "CG_R", "K4", "URCS_N", "URCS_N1",
          "EPS_N", "EPS_N1", "ROT_N1", "HIST_N", "HIST_N1", "CEP"]
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}
SOA, AOS = 0, 1

ERRORS = {1: "Newton loop does not converge", 2: "fftPcg failed to converge", 3: "improper CG tolerance",
          4: "prescribed stress cannot be reached", 5: "P_bar update failed", 6: "mm10 implicit solution failed",
          -1: "CUDA error", -2: "usage error", -3: "NCCL error"}


class CpfftError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cpfft_b200 error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("N", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("maxIter", C.c_int32), ("pad_", C.c_int32),
                ("tolNR", C.c_double), ("tolPCG", C.c_double), ("tstep", C.c_double)]


EXPORTS = [
    "cpfft_create", "cpfft_destroy", "cpfft_last_error", "cpfft_set_materials", "cpfft_set_voxels",
    "cpfft_set_voxels_taylor",
    "cpfft_set_params", "cpfft_hist_size", "cpfft_local_voxels", "cpfft_drive_eps_sig", "cpfft_G_K_dF",
    "cpfft_fftPcg", "cpfft_tangent_homo", "cpfft_mean_P", "cpfft_update", "cpfft_FFT_nr3", "cpfft_step_log", "cpfft_step_counter",
    "cpfft_field_ncomp", "cpfft_upload", "cpfft_download", "cpfft_download_fail_flags",
    "cpfft_download_local_iters", "cpfft_material_failures", "cpfft_nccl_unique_id", "cpfft_nccl_init", "cpfft_exchange_mode", "cpfft_synchronize",
    "cpfft_stream", "cpfft_kernel_launches", "cpfft_profile_enable", "cpfft_profile_reset",
    "cpfft_profile_classes", "cpfft_profile_name", "cpfft_profile_get", "cpfft_fp64_peak",
]


def library_path() -> str:
    # CPFFT_B200_LIB: development override (kernel-variant experiments); still a CUDA build
    return os.environ.get("CPFFT_B200_LIB") or os.path.join(_HERE, "libcpfft_b200.so")


def load_library():
    """dlopen the in-tree CUDA library; never falls back to anything else."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.cpfft_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.cpfft_destroy.argtypes = [vp]
    L.cpfft_destroy.restype = None
    L.cpfft_last_error.argtypes = [vp]
    L.cpfft_last_error.restype = C.c_char_p
    L.cpfft_set_materials.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    L.cpfft_set_voxels.argtypes = [vp, ip, dp]
    L.cpfft_set_voxels_taylor.argtypes = [vp, ip, C.c_int, dp, ip]
    L.cpfft_set_params.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_double]
    L.cpfft_hist_size.argtypes = [vp]
    L.cpfft_local_voxels.argtypes = [vp]
    L.cpfft_local_voxels.restype = C.c_int64
    L.cpfft_drive_eps_sig.argtypes = [vp, C.c_int, C.c_int]
    L.cpfft_G_K_dF.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.cpfft_fftPcg.argtypes = [vp, C.c_int, C.c_int, C.c_double, ip, dp]
    L.cpfft_tangent_homo.argtypes = [vp, dp]
    L.cpfft_mean_P.argtypes = [vp, dp]
    L.cpfft_update.argtypes = [vp]
    L.cpfft_FFT_nr3.argtypes = [vp, C.c_int, dp, ip, ip, ip, C.c_int, dp, dp, C.POINTER(C.c_int64)]
    L.cpfft_step_log.argtypes = [vp]
    L.cpfft_step_log.restype = C.c_char_p
    L.cpfft_step_counter.argtypes = [vp, ip, ip]
    L.cpfft_field_ncomp.argtypes = [vp, C.c_int]
    L.cpfft_upload.argtypes = [vp, C.c_int, dp, C.c_int]
    L.cpfft_download.argtypes = [vp, C.c_int, dp, C.c_int]
    L.cpfft_download_fail_flags.argtypes = [vp, ip]
    L.cpfft_download_local_iters.argtypes = [vp, ip]
    L.cpfft_material_failures.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.cpfft_nccl_unique_id.argtypes = [vp]
    L.cpfft_nccl_init.argtypes = [vp, vp]
    L.cpfft_exchange_mode.argtypes = [vp]
    L.cpfft_synchronize.argtypes = [vp]
    L.cpfft_stream.argtypes = [vp]
    L.cpfft_stream.restype = vp
    L.cpfft_kernel_launches.argtypes = [vp]
    L.cpfft_kernel_launches.restype = C.c_int64
    L.cpfft_profile_enable.argtypes = [vp, C.c_int]
    L.cpfft_profile_reset.argtypes = [vp]
    L.cpfft_profile_name.argtypes = [C.c_int]
    L.cpfft_profile_name.restype = C.c_char_p
    L.cpfft_profile_get.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int64)]
    L.cpfft_fp64_peak.argtypes = [vp, dp]
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Solver:
    """One GPU's share of a CPFFT analysis (the whole grid when ``world == 1``)."""

    CG_CAP = 256     # CG solves returned per load step (stress-BC loops run (maxIter + 2) Newton loops at most)

    def __init__(self, prob: Problem, device: int = 0, rank: int = 0, world: int = 1, nccl_id: bytes | None = None,
                 local_slab: bool = False):
        self.L = load_library()
        self.prob = prob
        self.N, self.rank, self.world = prob.N, rank, world
        cfg = Config(N=prob.N, device=device, rank=rank, world=world, maxIter=prob.maxIter,
                     tolNR=prob.tolNR, tolPCG=prob.tolPCG, tstep=prob.tstep)
        self.h = C.c_void_p()
        rc = self.L.cpfft_create(C.byref(cfg), C.byref(self.h))
        self._check(rc)
        self.n3 = int(self.L.cpfft_local_voxels(self.h))
        mats, crys = prob.material_pods(), prob.crystal_pods()
        self._check(self.L.cpfft_set_materials(self.h, len(prob.materials), C.addressof(mats),
                                               len(prob.crystals), C.addressof(crys)))
        lo, hi = (0, self.n3) if local_slab else (rank * self.n3, (rank + 1) * self.n3)
        assert len(prob.matlist) >= hi, "matlist does not cover this rank's slab"
        ml = np.ascontiguousarray(prob.matlist[lo:hi], dtype=np.int32)
        if prob.taylor:       # n_crystals > 1 per material point and / or crystal_input file
            nc = prob.ncmax
            ang = np.ascontiguousarray(np.asarray(prob.angles, dtype=np.float64).reshape(-1, nc, 3)[lo:hi])
            ids = None if prob.crystal_ids is None else \
                np.ascontiguousarray(np.asarray(prob.crystal_ids, dtype=np.int32).reshape(-1, nc)[lo:hi])
            self._check(self.L.cpfft_set_voxels_taylor(self.h, _ip(ml), nc, _dp(ang), _ip(ids) if ids is not None else None))
        else:
            ang = np.ascontiguousarray(prob.angles[lo:hi], dtype=np.float64)
            self._check(self.L.cpfft_set_voxels(self.h, _ip(ml), _dp(ang)))
        self.H = self.L.cpfft_hist_size(self.h)
        if world > 1:
            if nccl_id is None:
                raise ValueError("world > 1 needs the broadcast NCCL unique id")
            buf = C.create_string_buffer(nccl_id, 128)
            self._check(self.L.cpfft_nccl_init(self.h, buf))

    @staticmethod
    def nccl_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.cpfft_nccl_unique_id(buf)
        if rc:
            raise CpfftError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def _check(self, rc):
        if rc != 0:
            msg = self.L.cpfft_last_error(self.h).decode() if self.h else ""
            raise CpfftError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.cpfft_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fields ----
    def ncomp(self, name):
        return self.L.cpfft_field_ncomp(self.h, FIELD_ID[name])

    def download(self, name, layout=SOA):
        nc = self.ncomp(name)
        out = np.empty((nc, self.n3) if layout == SOA else (self.n3, nc))
        self._check(self.L.cpfft_download(self.h, FIELD_ID[name], _dp(out), layout))
        return out

    def upload(self, name, arr, layout=SOA):
        nc = self.ncomp(name)
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.size == nc * self.n3
        self._check(self.L.cpfft_upload(self.h, FIELD_ID[name], _dp(a), layout))

    def fail_flags(self):
        f = np.zeros(self.n3, dtype=np.int32)
        self._check(self.L.cpfft_download_fail_flags(self.h, _ip(f)))
        return f

    def local_iters(self):
        f = np.zeros((self.n3, 2), dtype=np.int32)
        self._check(self.L.cpfft_download_local_iters(self.h, _ip(f)))
        return f

    def material_failures(self):
        """(total, last sweep) mm10 local-solver failures, summed over ranks"""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.L.cpfft_material_failures(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- hot path, reference names ----
    def drive_eps_sig(self, step, iiter):
        self._check(self.L.cpfft_drive_eps_sig(self.h, step, iiter))

    def G_K_dF(self, src, dst, flgK):
        self._check(self.L.cpfft_G_K_dF(self.h, FIELD_ID[src], FIELD_ID[dst], int(bool(flgK))))

    def fftPcg(self, b, x, tol):
        it, rr = C.c_int32(0), C.c_double(0)
        self._check(self.L.cpfft_fftPcg(self.h, FIELD_ID[b], FIELD_ID[x], tol, C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def tangent_homo(self):
        Ch = np.zeros(81)
        self._check(self.L.cpfft_tangent_homo(self.h, _dp(Ch)))
        return Ch

    def mean_P(self):
        p = np.zeros(9)
        self._check(self.L.cpfft_mean_P(self.h, _dp(p)))
        return p

    def update(self):
        self._check(self.L.cpfft_update(self.h))

    def FFT_nr3(self, nstep=None, first=0):
        """Run ``nstep`` load steps starting after the ones already taken by this solver."""
        prob = self.prob
        nstep = prob.nstep - first if nstep is None else nstep
        if first != self.next_step() - 1:
            raise CpfftError(-2, f"FFT_nr3(first={first}): the handle's next load step is {self.next_step()}")
        bc = np.ascontiguousarray(prob.BC_all()[first:first + nstep])
        nbc = np.ascontiguousarray(prob.isNBC, dtype=np.int32)
        nr = np.zeros(nstep, dtype=np.int32)
        cg = np.full((nstep, self.CG_CAP), -1, dtype=np.int32)
        pbar = np.zeros((nstep, 9))
        sec = np.zeros(3)
        cnt = np.zeros(5, dtype=np.int64)
        rc = self.L.cpfft_FFT_nr3(self.h, nstep, _dp(bc), _ip(nbc), _ip(nr), _ip(cg), self.CG_CAP, _dp(pbar),
                                  _dp(sec), cnt.ctypes.data_as(C.POINTER(C.c_int64)))
        self._check(rc)
        trunc = C.c_int32(0)
        self._check(self.L.cpfft_step_counter(self.h, None, C.byref(trunc)))
        if trunc.value:
            raise CpfftError(-2, f"{trunc.value} CG iteration counts did not fit cg_cap = {self.CG_CAP}")
        cg_lists = [list(r[:list(r).index(-1)]) if -1 in r else list(r) for r in cg]
        return dict(rc=rc, nr_iters=nr, cg_iters=cg_lists, Pbar=pbar, buckets=sec, counters=cnt,
                    log=self.L.cpfft_step_log(self.h).decode())

    def next_step(self):
        """number of the load step the next FFT_nr3 call starts with (1-based, FFT_nr3.f:51)"""
        n = C.c_int32(0)
        self._check(self.L.cpfft_step_counter(self.h, C.byref(n), None))
        return n.value

    def profile(self, on=True):
        self._check(self.L.cpfft_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.L.cpfft_profile_reset(self.h))

    def profile_table(self):
        """{kernel class: (total ms, launches)} from CUDA events on the launching stream."""
        out = {}
        for c in range(self.L.cpfft_profile_classes()):
            ms, cnt = C.c_double(0), C.c_int64(0)
            self._check(self.L.cpfft_profile_get(self.h, c, C.byref(ms), C.byref(cnt)))
            out[self.L.cpfft_profile_name(c).decode()] = (ms.value, cnt.value)
        return out

    def fp64_peak(self):
        """measured FP64 FMA peak of this GPU in TFLOP/s"""
        v = C.c_double(0)
        self._check(self.L.cpfft_fp64_peak(self.h, C.byref(v)))
        return v.value

    def upload_ptr(self, name, ptr, layout=SOA):
        """upload from a raw host pointer (e.g. pinned memory)"""
        self._check(self.L.cpfft_upload(self.h, FIELD_ID[name], C.cast(ptr, C.POINTER(C.c_double)), layout))

    def download_ptr(self, name, ptr, layout=SOA):
        self._check(self.L.cpfft_download(self.h, FIELD_ID[name], C.cast(ptr, C.POINTER(C.c_double)), layout))

    def synchronize(self):
        self._check(self.L.cpfft_synchronize(self.h))

    def stream(self):
        return self.L.cpfft_stream(self.h)

    def exchange_mode(self):
        """0 single GPU, 1 NCCL send/recv transposes, 2 peer stores fused into the FFT kernels"""
        return int(self.L.cpfft_exchange_mode(self.h))

    def kernel_launches(self):
        return int(self.L.cpfft_kernel_launches(self.h))
