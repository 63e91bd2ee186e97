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

    FIELDS = {"FN1": "Fn1", "FN": "Fn", "PN1": "Pn1", "K4": "K4", "HIST_N": "hist_n", "HIST_N1": "hist_n1", "URCS_N": "urcs_n",
              "URCS_N1": "urcs_n1", "EPS_N": "eps_n", "EPS_N1": "eps_n1"}

    def upload(self, name, F):
        f = getattr(self.k, self.FIELDS[name])
        f[:] = np.asarray(F).reshape(f.shape)

    def download(self, name, layout=0):
        a = np.array(getattr(self.k, self.FIELDS[name]))
        return np.ascontiguousarray(a.T) if layout == 1 else a

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


# ---- tests/test_zzz_gpu_reference_fixtures.py: the reference-executed fixtures through the Solver interface

import test_zzz_gpu_reference_fixtures as R


@pytest.fixture(scope="module")
def kernel_source_solver(oracle_built):
    from host_kernels import build
    build()
    return _KernelSourceSolver


@pytest.mark.parametrize("k", range(5))
def test_body_of_the_gpu_voxel_test(kernel_source_solver, k):
    R.voxel_case(kernel_source_solver, k)


@pytest.mark.parametrize("name", ["taylor", "mts", "bcc48"])
def test_body_of_the_gpu_wrapper_test(kernel_source_solver, name):
    R.wrapper_case(kernel_source_solver, name)


@pytest.mark.parametrize("job,k", [("", 1), ("", 2), ("", 3), ("m01_", 1), ("m01_", 2), ("m01_", 3)])
def test_body_of_the_gpu_job_sweep_test(kernel_source_solver, job, k):
    R.job_sweep(kernel_source_solver, job, k)


@pytest.mark.parametrize("deck_name,job,k", [("test_mm10.in", "deck_", 1), ("test_mm10.in", "deck_", 3), ("test_mm10.in", "deck_", 10),
                                             ("test_mm01.in", "deck01_", 1), ("test_mm01.in", "deck01_", 10),
                                             ("mts_mm10.in", "deckmts_", 2), ("mts_mm10.in", "deckmts_", 4),
                                             ("taylor_mm10.in", "decktaylor_", 2), ("taylor_mm10.in", "decktaylor_", 5)])
def test_kernel_source_on_the_shipped_decks_as_the_reference_ran_them(kernel_source_solver, deck_name, job, k):
    """the body of tests/test_zzz_gpu_reference_fixtures.py::test_shipped_deck_sweeps_through_the_c_abi with the host build of the
    kernel source: the kernels' own text on the state the REFERENCE'S source passed through when it ran its shipped decks"""
    R.deck_sweep(kernel_source_solver, deck_name, job, k)
