"""Host build of the CUDA library's per-voxel material code (tests/native/material_host.cpp):
the same source nvcc compiles into k_update_mm01 / k_update_mm10 / k_pk1_tangent, compiled by
g++ and run voxel by voxel.  TEST INFRASTRUCTURE ONLY -- it gives the CPU suite a check of the
kernels' arithmetic and control flow against the oracle; the product never loads it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "native", "material_host.cpp")
_OUT = os.path.join(_HERE, "native", "build", "libmaterial_host.so")
_LIB = None


def build(force: bool = False) -> str:
    csrc = os.path.join(_ROOT, "cpfft_b200", "csrc")
    deps = [_SRC, os.path.join(_ROOT, "include", "cpfft_b200.h")] + \
           [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h", ".hpp"))]
    stale = (not os.path.exists(_OUT)) or any(os.path.getmtime(d) > os.path.getmtime(_OUT) for d in deps)
    if force or stale:
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        # same -march as the oracle (FMA contraction on, like nvcc's default)
        subprocess.check_call(["g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-w", "-o", _OUT, _SRC])
    return _OUT


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.mh_create.restype = C.c_void_p
        L.mh_create.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, ip, C.c_int, dp, ip, C.c_double]
        L.mh_destroy.argtypes = [C.c_void_p]
        L.mh_hist_size.argtypes = [C.c_void_p]
        L.mh_ngrains.argtypes = [C.c_void_p]
        L.mh_field.restype = dp
        L.mh_field.argtypes = [C.c_void_p, C.c_char_p]
        L.mh_fail_flags.restype = ip
        L.mh_fail_flags.argtypes = [C.c_void_p]
        L.mh_local_iters.restype = ip
        L.mh_local_iters.argtypes = [C.c_void_p]
        L.mh_drive_eps_sig.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mh_update.argtypes = [C.c_void_p]
        L.mh_set_lattice_frame.argtypes = [C.c_void_p, C.c_int]
        L.mh_lu7.argtypes = [dp, dp]
        L.mh_pow_abs.restype = C.c_double
        L.mh_pow_abs.argtypes = [C.c_double, C.c_int, C.c_double]
        _LIB = L
    return _LIB


class HostKernels:
    """Material stage of one ``Problem`` on the host build of the device code.  Fields are the
    device layout: (ncomp, N3) structure-of-arrays views."""

    NCOMP = {"Fn": 9, "Fn1": 9, "Pn1": 9, "K4": 81, "urcs_n": 9, "urcs_n1": 9, "eps_n": 6, "eps_n1": 6,
             "rot_n1": 9, "cep": 36}

    def __init__(self, prob, lattice_frame=True):   # the product's default (CPFFT_MM10_LF unset)
        L = _lib()
        self.L, self.prob, self.N3 = L, prob, prob.N3
        mats, crys = prob.material_pods(), prob.crystal_pods()
        ml = np.ascontiguousarray(prob.matlist, dtype=np.int32)
        nc, ang, ids = prob.taylor_tables()
        ip = C.POINTER(C.c_int32)
        self.h = L.mh_create(prob.N3, len(prob.materials), C.addressof(mats), len(prob.crystals), C.addressof(crys),
                             ml.ctypes.data_as(ip), nc, ang.ctypes.data_as(C.POINTER(C.c_double)),
                             ids.ctypes.data_as(ip) if ids is not None else None, prob.tstep)
        if not self.h:
            raise RuntimeError("mh_create failed")
        self.H = L.mh_hist_size(self.h)
        L.mh_set_lattice_frame(self.h, int(lattice_frame))

    def __del__(self):
        try:
            self.L.mh_destroy(self.h)
        except Exception:
            pass

    def field(self, name):
        ncomp = self.H if name.startswith("hist") else self.NCOMP[name]
        return np.ctypeslib.as_array(self.L.mh_field(self.h, name.encode()), shape=(ncomp, self.N3))

    def __getattr__(self, name):
        if name in HostKernels.NCOMP or name in ("hist_n", "hist_n1"):
            return self.field(name)
        raise AttributeError(name)

    @property
    def local_iters(self):
        return np.ctypeslib.as_array(self.L.mh_local_iters(self.h), shape=(self.N3, 2))

    @property
    def fail_flags(self):
        return np.ctypeslib.as_array(self.L.mh_fail_flags(self.h), shape=(self.N3,))

    def drive_eps_sig(self, step, it):
        return self.L.mh_drive_eps_sig(self.h, step, it)

    def update(self):
        self.L.mh_update(self.h)
