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


@pytest.mark.parametrize("N", [9, 12, 15, 16, 32, 40, 51, 64, 80])
@pytest.mark.parametrize("flgK", [0, 1])
def test_G_K_dF_matches_oracle(libs, N, flgK):
    """odd N = reference-faithful; even N = documented Nyquist-zero convention; 16/32/64 (radix
    2/4/8/16), 40/80 (radix 5) and the odd 15/51 (radix 3/5/17, two real lines per complex transform
    in the z passes: the plan family of 255^3) run through the fast path, 9/12 through the generic one."""
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


@pytest.mark.parametrize("N", [32, 200, 15, 51, 255])
def test_generic_and_fast_path_agree(libs, N):
    """the same grid through both implementations (CPFFT_GENERIC_FFT forces the generic one);
    200 = 5 x 5 x 8 is the three-stage radix-5 plan family of the 400^3 weak-scaling grid; 15, 51 and
    255 = 17 x 15 are the odd grids, on which the operator is the reference's own"""
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


@pytest.mark.parametrize("N", [200, 255, 256, 320])
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


@pytest.mark.parametrize("N,grains,stress_bc", [(16, 20, False), (15, 20, False), (16, 20, True), (15, 20, True)])
def test_polycrystal_steps_match_oracle(libs, N, grains, stress_bc):
    """the synthetic fcc/Voce Voronoi polycrystal of the benchmark at a size the oracle runs in
    seconds, 8 load steps of 0.1 % (>= 6 of them plastic), strain-controlled and with the
    stress-BC loop of the benchmark's loading (P_yy = P_zz = 0: tangent_homo, 9 CG solves through
    the fused fast path, NBC_update): identical Newton and CG iteration counts, stress-strain
    curve, fields, history -- at the north-star tolerances (the polar decomposition's small-strain
    noise is the same number on both sides, kin.cuh polar_R)."""
    from cpfft_b200.polycrystal import polycrystal
    Solver, Oracle = libs
    p = polycrystal(N, ngrains=grains, stress_bc=stress_bc)
    s, o = Solver(p), Oracle(p, threads=8)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(nstep=8), o.FFT_nr3(nstep=8)
    assert ro["rc"] == 0
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    assert_same_cg_counts(rs["cg_iters"], ro["cg_iters"], slack=1 if stress_bc else 0)
    assert int(rs["counters"][3]) == int(ro["counters"][3])
    scale = np.abs(ro["Pbar"]).max()
    errs = {"Pbar": np.abs(rs["Pbar"] - ro["Pbar"]).max() / scale, "P": relerr(s.download("PN1"), o.Pn1),
            "F": relerr(s.download("FN1"), o.Fn1)}
    print("polycrystal parity", N, stress_bc, errs)
    assert errs["Pbar"] <= TOL_MACRO, errs
    assert errs["F"] <= TOL_VOXEL, errs
    assert errs["P"] <= TOL_VOXEL, errs
    if stress_bc:
        assert np.abs(rs["Pbar"][:, [4, 8]]).max() <= 1e-5 * scale      # the prescribed stresses are met
    compare_mm10_history(s.download("HIST_N", 1)[:, :o.H], o.hist_n, 12, tol=TOL_VOXEL)
    # after the commit the n+1 names still read the committed state (cpfft_update exchanges buffers)
    assert np.array_equal(s.download("HIST_N1", 1), s.download("HIST_N", 1))
    assert relerr(s.download("URCS_N1", 1), o.urcs_n1) <= TOL_VOXEL


@pytest.mark.parametrize("name", ["poly32_strain", "poly32_stress", "poly64_strain"])
def test_polycrystal_matches_frozen_oracle(libs, name):
    """the benchmark polycrystal at 32^3 (both loadings) and 64^3, 8 load steps: sizes the oracle needs
    minutes for, so its output is a committed fixture (tools/make_golden_poly.py -> tests/golden/<name>.npz:
    iteration counts, macroscopic curve, F / P / unrotated stress / mm10 history at 1024 sampled voxels)."""
    from cpfft_b200.polycrystal import polycrystal
    Solver, _ = libs
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    N, nstep, sbc = int(g["N"]), int(g["nstep"]), bool(g["stress_bc"])
    s = Solver(polycrystal(N, ngrains=int(g["grains"]), stress_bc=sbc))
    s.drive_eps_sig(1, 0)
    r = s.FFT_nr3(nstep=nstep)
    assert list(r["nr_iters"]) == list(g["nr_iters"])
    assert_same_cg_counts(r["cg_iters"], [[v for v in row if v >= 0] for row in g["cg_iters"]], slack=1 if sbc else 0)
    assert [int(v) for v in r["counters"][3:5]] == [int(v) for v in g["failures"]]
    idx = g["idx"]
    F, P = s.download("FN1"), s.download("PN1")
    errs = {"Pbar": np.abs(r["Pbar"] - g["Pbar"]).max() / np.abs(g["Pbar"]).max(),
            "F": np.abs(F[:, idx] - g["F"]).max() / np.abs(g["F"]).max(), "P": np.abs(P[:, idx] - g["P"]).max() / np.abs(g["P"]).max(),
            "P_absmax": abs(np.abs(P).max() - float(g["P_absmax"])) / float(g["P_absmax"]),
            "F_abssum": abs(np.abs(F).sum() - float(g["F_abssum"])) / float(g["F_abssum"]),
            "urcs": relerr(s.download("URCS_N1", 1)[idx], g["urcs"])}
    print("frozen-oracle parity", name, errs)
    assert errs["Pbar"] <= TOL_MACRO, errs
    assert max(errs["F"], errs["P"], errs["urcs"], errs["P_absmax"]) <= TOL_VOXEL and errs["F_abssum"] <= 1e-12, errs
    compare_mm10_history(s.download("HIST_N", 1)[idx], g["hist"], 12, tol=TOL_VOXEL)


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
