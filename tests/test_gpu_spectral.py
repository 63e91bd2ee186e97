"""The spectral operator on the GPU (both the generic Stockham path and the power-of-two fast
path of spectral_pow2.cu) against the oracle at sizes it finishes in seconds, and through
size-independent projection identities at the benchmark size (256^3)."""
import os

import numpy as np
import pytest

from helpers import relerr, TOL_VOXEL, TOL_MACRO, compare_mm10_history, assert_same_cg_counts
from test_oracle_spectral import _toy_problem, _grad_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(oracle_built):
    from cpfft_b200 import Solver
    from oracle import Oracle
    return Solver, Oracle


def _hetero_state(p, s, o, amp=0.02, seed=3):
    rng = np.random.default_rng(seed)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += amp * rng.standard_normal((9, p.N3))
    s.upload("FN1", F); o.Fn1[:] = F
    s.drive_eps_sig(1, 1); o.drive_eps_sig(1, 1)
    # same operator on both sides: the tangents themselves agree to ~1e-11 only (the closed-form
    # eigenvalues of the polar decomposition amplify round-off), which is tested elsewhere
    assert relerr(s.download("K4"), o.K4) <= TOL_VOXEL
    s.upload("K4", o.K4)
    return rng


@pytest.mark.parametrize("N", [9, 12, 15, 16, 32, 40, 64, 80])
@pytest.mark.parametrize("flgK", [0, 1])
def test_G_K_dF_matches_oracle(libs, N, flgK):
    """odd N = reference-faithful; even N = documented Nyquist-zero convention; 16/32/64 (radix
    2/4/8/16) and 40/80 (radix 5) run through the fast path, 9/12/15 through the generic one."""
    Solver, Oracle = libs
    p = _toy_problem(N)
    s, o = Solver(p), Oracle(p, threads=8)
    rng = _hetero_state(p, s, o)
    x = rng.standard_normal((9, p.N3))
    s.upload("DFM", x)
    s.G_K_dF("DFM", "B", flgK)
    ref = o.G_K_dF(x, flgK)
    assert relerr(s.download("B"), ref) <= 1e-13


def test_G_K_dF_128_matches_oracle(libs):
    Solver, Oracle = libs
    N = 128
    p = _toy_problem(N)
    s, o = Solver(p), Oracle(p, threads=16)
    rng = _hetero_state(p, s, o)
    x = rng.standard_normal((9, p.N3))
    s.upload("DFM", x)
    s.G_K_dF("DFM", "B", 1)
    assert relerr(s.download("B"), o.G_K_dF(x, 1)) <= 1e-13


@pytest.mark.parametrize("N", [32, 200])
def test_generic_and_fast_path_agree(libs, N):
    """the same even grid through both implementations (CPFFT_GENERIC_FFT forces the generic one);
    200 = 5 x 5 x 8 is the three-stage radix-5 plan family of the 400^3 weak-scaling grid"""
    Solver, _ = libs
    p = _toy_problem(N)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, p.N3))
    out = []
    for force in (False, True):
        if force:
            os.environ["CPFFT_GENERIC_FFT"] = "1"
        try:
            s = Solver(p)
            s.drive_eps_sig(1, 0)
            s.upload("DFM", x)
            s.G_K_dF("DFM", "B", 1)
            out.append(s.download("B"))
            s.close()
        finally:
            os.environ.pop("CPFFT_GENERIC_FFT", None)
    assert relerr(out[0], out[1]) <= 1e-13


@pytest.mark.parametrize("N", [200, 256, 320])
def test_projection_identities_at_benchmark_size(libs, N):
    """Ghat:grad(u) = grad(u), Ghat:const = 0, idempotence and self-adjointness at 256^3, where
    the oracle is too slow: the operator is an orthogonal projection at any size."""
    Solver, _ = libs
    p = _toy_problem(N)
    s = Solver(p)
    rng = np.random.default_rng(N)
    g = _grad_field(N, rng, kmax=N // 2 - 1)
    s.upload("DFM", g); s.G_K_dF("DFM", "B", 0)
    assert relerr(s.download("B"), g) <= 1e-12
    const = np.repeat(rng.standard_normal((9, 1)), N ** 3, axis=1)
    s.upload("DFM", const); s.G_K_dF("DFM", "B", 0)
    assert np.abs(s.download("B")).max() <= 1e-12
    x = rng.standard_normal((9, N ** 3))
    s.upload("DFM", x); s.G_K_dF("DFM", "B", 0)
    Gx = s.download("B")
    s.G_K_dF("B", "DFM", 0)
    assert relerr(s.download("DFM"), Gx) <= 1e-12
    y = rng.standard_normal((9, N ** 3))
    s.upload("DFM", y); s.G_K_dF("DFM", "B", 0)
    Gy = s.download("B")
    assert abs((Gx * y).sum() - (x * Gy).sum()) <= 1e-10 * np.sqrt((x * x).sum() * (y * y).sum())


@pytest.mark.parametrize("N,grains", [(16, 20), (15, 20)])
def test_polycrystal_steps_match_oracle(libs, N, grains):
    """the synthetic fcc/Voce Voronoi polycrystal of the benchmark at a size the oracle runs in
    seconds: identical Newton and CG iteration counts, stress-strain curve, fields, history."""
    from cpfft_b200.polycrystal import polycrystal
    Solver, Oracle = libs
    p = polycrystal(N, ngrains=grains)
    s, o = Solver(p), Oracle(p, threads=8)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(nstep=4), o.FFT_nr3(nstep=4)
    assert ro["rc"] == 0
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    assert_same_cg_counts(rs["cg_iters"], ro["cg_iters"])
    assert int(rs["counters"][3]) == int(ro["counters"][3]) == 0
    scale = np.abs(ro["Pbar"]).max()
    errs = {"Pbar": np.abs(rs["Pbar"] - ro["Pbar"]).max() / scale, "P": relerr(s.download("PN1"), o.Pn1),
            "F": relerr(s.download("FN1"), o.Fn1)}
    print("polycrystal parity", N, errs)
    # F and the macroscopic curve meet the north-star tolerances; the per-voxel stress at 0.1 %
    # strain increments is limited by the round-off noise floor of the reference's own polar
    # decomposition (tests/test_oracle_material.py::test_stress_noise_floor_...): 5e-8
    assert errs["Pbar"] <= TOL_MACRO, errs
    assert errs["F"] <= TOL_VOXEL, errs
    assert errs["P"] <= 5e-8, errs
    compare_mm10_history(s.download("HIST_N", 1)[:, :o.H], o.hist_n, 12, tol=5e-8)


def test_polycrystal_stress_bc_matches_oracle(libs):
    """uniaxial tension with P_yy = P_zz = 0 on the benchmark polycrystal (N = 16): the
    stress-BC loop, tangent_homo (9 CG solves through the fused fast path) and NBC_update."""
    from cpfft_b200.polycrystal import polycrystal
    Solver, Oracle = libs
    p = polycrystal(16, ngrains=20, stress_bc=True)
    s, o = Solver(p), Oracle(p, threads=8)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(nstep=3), o.FFT_nr3(nstep=3)
    assert ro["rc"] == 0
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    assert_same_cg_counts(rs["cg_iters"], ro["cg_iters"], slack=1)
    scale = np.abs(ro["Pbar"]).max()
    assert np.abs(rs["Pbar"] - ro["Pbar"]).max() / scale <= 1e-9
    assert np.abs(rs["Pbar"][:, [4, 8]]).max() <= 1e-5 * scale      # the prescribed stresses are met
    assert relerr(s.download("FN1"), o.Fn1) <= TOL_VOXEL


@pytest.mark.parametrize("N", [16, 32, 40, 64, 80, 128])
def test_kernel_variants_are_bit_identical(libs, N, monkeypatch):
    """The software-pipelined inverse z pass (k_iz_pipe, default) against k_iz (CPFFT_IZ_PIPE=0),
    and the CG solution update fused into the next forward z pass (k_fz MODE 3, default) against the
    separate vector pass (CPFFT_CG_FUSE_X=0): same arithmetic in the same order, so G_K_dF and a CG
    solve must agree bit for bit.  The switches are read when the handle is created."""
    from test_oracle_spectral import _toy_problem
    Solver, _ = libs
    p = _toy_problem(N)
    rng = np.random.default_rng(N)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.002 * rng.standard_normal((9, p.N3))        # elastic, heterogeneous: K4 stays positive definite
    x = rng.standard_normal((9, p.N3))
    res = []
    for pipe, fuse in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
        monkeypatch.setenv("CPFFT_IZ_PIPE", pipe)
        monkeypatch.setenv("CPFFT_CG_FUSE_X", fuse)
        s = Solver(p)
        s.upload("FN1", F)
        s.drive_eps_sig(1, 1)
        s.upload("DFM", x)
        s.G_K_dF("DFM", "B", 1)
        g = s.download("B")
        it, rr = s.fftPcg("B", "DFM", 1e-8)
        res.append((g, s.download("DFM"), it, rr))
        del s
    assert 3 < res[0][2] < 200
    for g, sol, it, rr in res[1:]:
        assert np.array_equal(g, res[0][0])
        assert it == res[0][2] and rr == res[0][3]
        assert np.array_equal(sol, res[0][1])
