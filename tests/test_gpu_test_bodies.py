"""The bodies of the GPU tests that had not run on a GPU when they were written
(tests/test_zz_gpu_new_features.py), executed on the CPU with stand-ins behind the Solver interface (the host
build of the kernel source / the oracle): checks the TEST code -- indices, shapes, tolerances, helper calls --
so that a failure on the GPU box points at the kernels, not at the test."""
import numpy as np
import pytest

import test_zz_gpu_new_features as T


class _KernelSourceSolver:
    """tests/host_kernels.HostKernels behind the Solver calls of test_crystal_and_grid_variants"""
    def __init__(self, p):
        from host_kernels import HostKernels
        self.k = HostKernels(p)
        self.H = self.k.H

    def drive_eps_sig(self, step, it):
        return self.k.drive_eps_sig(step, it)

    def upload(self, name, F):
        getattr(self.k, {"FN1": "Fn1", "FN": "Fn"}[name])[:] = F

    def download(self, name):
        return np.array(getattr(self.k, {"PN1": "Pn1", "K4": "K4"}[name]))

    def local_iters(self):
        return np.array(self.k.local_iters)

    def update(self):
        self.k.update()


class _OracleSolver:
    """the oracle behind the Solver calls of test_fast_path_matches_plain_fft_statement"""
    def __init__(self, p):
        from oracle import Oracle
        self.o, self.buf = Oracle(p), {}

    def upload(self, name, A):
        if name == "FN1":
            self.o.Fn1[:] = A
        else:
            self.buf[name] = np.array(A)

    def drive_eps_sig(self, step, it):
        return self.o.drive_eps_sig(step, it)

    def download(self, name):
        return np.array(self.o.K4) if name == "K4" else self.buf[name]

    def G_K_dF(self, src, dst, flgK):
        self.buf[dst] = self.o.G_K_dF(self.buf[src], flgK)


@pytest.mark.parametrize("kind", ["voce_m_2", "bcc48", "mixed_materials", "crystal_file_single", "mts", "mts_taylor"])
def test_body_of_the_gpu_variant_test(oracle_built, kind):
    from host_kernels import build
    from oracle import Oracle
    build()
    T.test_crystal_and_grid_variants((_KernelSourceSolver, Oracle), kind)


def test_body_of_the_gpu_spectral_test(oracle_built):
    T.test_fast_path_matches_plain_fft_statement((_OracleSolver, None), 16)
