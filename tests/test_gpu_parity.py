"""GPU <-> oracle parity through the C ABI on the reference's own decks (N = 7)."""
import numpy as np
import pytest

from helpers import (deck, relerr, TOL_VOXEL, TOL_MACRO, mm10_variant, stress_bc_variant,
                     compare_mm10_history, assert_same_cg_counts, NSLIP)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(oracle_built):
    from cpfft_b200 import Solver
    from oracle import Oracle
    return Solver, Oracle


def _compare_state(s, o, hist=True, tol=TOL_VOXEL):
    errs = {
        "Pn1": relerr(s.download("PN1"), o.Pn1),
        "K4": relerr(s.download("K4"), o.K4),
        "urcs_n1": relerr(s.download("URCS_N1", 1), o.urcs_n1),
        "eps_n1": relerr(s.download("EPS_N1", 1), o.eps_n1),
    }
    if hist:
        hg = s.download("HIST_N1", 1)[:, :o.H]
        if any(m.type == 10 for m in o.prob.materials):
            nslip = NSLIP[o.prob.crystals[0].slip_type]
            errs.update({"hist." + k: v for k, v in compare_mm10_history(hg, o.hist_n1, nslip, tol).items()})
        else:
            errs["hist_n1"] = relerr(hg, o.hist_n1)
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"parity violated: {bad} (all: {errs})"
    return errs


@pytest.mark.parametrize("name", ["test_mm01.in", "test_mm10.in"])
def test_initial_sweep(libs, name):
    """drive_eps_sig(1,0) at F = I: P = 0, elastic K4 (FFT_finite_3d.f:145)."""
    Solver, Oracle = libs
    p = deck(name)
    s, o = Solver(p), Oracle(p)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    _compare_state(s, o)


@pytest.mark.parametrize("name", ["test_mm01.in", "test_mm10.in"])
def test_sweep_on_perturbed_F(libs, name):
    """one nonlinear sweep (iter = 1) on a heterogeneous, finite deformation field."""
    Solver, Oracle = libs
    p = deck(name)
    s, o = Solver(p), Oracle(p)
    rng = np.random.default_rng(7)
    amp = 0.05 if name == "test_mm01.in" else 0.004
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += amp * rng.standard_normal((9, p.N3))
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    s.upload("FN1", F); o.Fn1[:] = F
    for it in (0, 1, 2):
        s.drive_eps_sig(1, it); assert o.drive_eps_sig(1, it) == 0
        _compare_state(s, o)
    if name == "test_mm10.in":
        assert np.array_equal(s.local_iters(), o.local_iters)


@pytest.mark.parametrize("name", ["test_mm01.in", "test_mm10.in"])
@pytest.mark.parametrize("flgK", [0, 1])
def test_G_K_dF(libs, name, flgK):
    Solver, Oracle = libs
    p = deck(name)
    s, o = Solver(p), Oracle(p)
    rng = np.random.default_rng(11)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.01 * rng.standard_normal((9, p.N3))
    s.upload("FN1", F); o.Fn1[:] = F
    s.drive_eps_sig(1, 1); o.drive_eps_sig(1, 1)
    x = rng.standard_normal((9, p.N3))
    s.upload("DFM", x)
    s.G_K_dF("DFM", "B", flgK)
    ref = o.G_K_dF(x, flgK)
    assert relerr(s.download("B"), ref) <= 1e-12


@pytest.mark.parametrize("name", ["test_mm01.in", "test_mm10.in"])
def test_full_deck(libs, name):
    """whole FFT_nr3 step loop: identical Newton counts, stress-strain curve, final state."""
    Solver, Oracle = libs
    p = deck(name)
    s, o = Solver(p), Oracle(p)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(), o.FFT_nr3()
    assert ro["rc"] == 0
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    assert_same_cg_counts(rs["cg_iters"], ro["cg_iters"])
    scale = np.abs(ro["Pbar"]).max()
    assert np.abs(rs["Pbar"] - ro["Pbar"]).max() / scale <= TOL_MACRO
    _compare_state(s, o)
    assert relerr(s.download("FN1"), o.Fn1) <= TOL_VOXEL
    if name == "test_mm01.in":
        assert relerr(s.download("HIST_N", 1)[:, :o.H], o.hist_n) <= TOL_VOXEL
    else:
        compare_mm10_history(s.download("HIST_N", 1)[:, :o.H], o.hist_n, 48)


def test_homogeneous_single_crystal(libs):
    """test_mm10.in with angle2.in: uniform fields, zero fluctuation."""
    Solver, Oracle = libs
    p = mm10_variant("angle2.in")
    s, o = Solver(p), Oracle(p)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(nstep=4), o.FFT_nr3(nstep=4)
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    P = s.download("PN1")
    assert np.abs(P - P.mean(axis=1, keepdims=True)).max() <= 1e-9 * np.abs(P).max()
    assert np.abs(rs["Pbar"] - ro["Pbar"]).max() / np.abs(ro["Pbar"]).max() <= TOL_MACRO


def test_stress_bc_deck(libs):
    """derived deck with P_yy = P_zz = 0: exercises tangent_homo + NBC_update."""
    Solver, Oracle = libs
    p = stress_bc_variant(deck("test_mm01.in"))
    s, o = Solver(p), Oracle(p)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(nstep=3), o.FFT_nr3(nstep=3)
    assert ro["rc"] == 0
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    assert np.abs(rs["Pbar"] - ro["Pbar"]).max() / np.abs(ro["Pbar"]).max() <= 1e-9
    assert relerr(s.download("FN1"), o.Fn1) <= TOL_VOXEL


def test_mm10_local_failure_points(libs):
    """Two voxels of the 256^3 polycrystal whose local Newton solve fails (sub-stepping
    exhausted) at load step 2 / global iteration 4 -- captured on the GPU, stored in
    tests/golden/mm10_fail_points.npz.  Oracle and GPU must fail on exactly the same points with
    the same iteration counts and leave the same defined state (n state + elastic tangent)."""
    import os
    from cpfft_b200.polycrystal import polycrystal
    Solver, Oracle = libs
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "mm10_fail_points.npz"))
    nb = d["Fn"].shape[1]
    p = polycrystal(2, ngrains=8)
    p.angles[:nb] = d["angles"]
    s, o = Solver(p), Oracle(p)
    H = o.H
    hist = np.zeros((H, p.N3)); urcs = np.zeros((9, p.N3)); eps = np.zeros((6, p.N3))
    Fn = np.zeros((9, p.N3)); Fn[[0, 4, 8]] = 1.0
    Fn1 = Fn.copy()
    # the remaining 6 voxels: copies of point 0's state with a benign (small) increment
    for v in range(p.N3):
        src = v if v < nb else 0
        hist[:, v] = d["hist_n"][:H, src]; urcs[:, v] = d["urcs_n"][:, src]; eps[:, v] = d["eps_n"][:, src]
        Fn[:, v] = d["Fn"][:, src]
        Fn1[:, v] = d["Fn1"][:, src] if v < nb else d["Fn"][:, src] + 0.05 * (d["Fn1"][:, src] - d["Fn"][:, src])
    if nb < p.N3:
        p.angles[nb:] = d["angles"][0]
        s, o = Solver(p), Oracle(p)
    for name, arr in (("HIST_N", hist), ("URCS_N", urcs), ("EPS_N", eps), ("FN", Fn), ("FN1", Fn1)):
        s.upload(name, arr)
    o.hist_n[:] = hist.T; o.urcs_n[:] = urcs.T; o._view("eps_n", (o.N3, 6))[:] = eps.T
    o.Fn[:] = Fn; o.Fn1[:] = Fn1
    step, it = int(d["step"]), int(d["iter"])
    s.drive_eps_sig(step, it)
    nfail = o.drive_eps_sig(step, it)
    assert nfail == nb
    flags = s.fail_flags()
    assert flags.sum() == nb and flags[:nb].all()
    assert s.material_failures() == (nb, nb)
    assert np.array_equal(s.local_iters(), o.local_iters)
    assert np.array_equal(s.local_iters()[:nb], d["liters"])
    _compare_state(s, o)


def test_cli_runs_a_deck_end_to_end(tmp_path):
    """python -m cpfft_b200 deck: the reference's log lines (FFT_nr3.f:195-199) on stdout and
    its flat-text result files (ouresult.f) for the steps of `output results steps ...`."""
    import os
    import re
    import subprocess
    import sys
    from helpers import DECKS, CLI_MM10_FILES
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "cpfft_b200", os.path.join(DECKS, "test_mm10.in"), "--outdir", str(tmp_path),
                          "--steps", "4"], capture_output=True, text=True, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    txt = out.stdout
    assert len(re.findall(r"^     Now starting step:\s+\d+$", txt, flags=re.M)) == 4
    assert len(re.findall(r"^       Initial residual\s+[-0-9.]+D[+-]\d\d$", txt, flags=re.M)) == 4
    its = re.findall(r"^       Iteration\s+(\d+)\s+residual\s+([-0-9.]+D[+-]\d\d)$", txt, flags=re.M)
    assert len(its) == 1 + 1 + 3 + 3          # Newton iterations of steps 1-4 of test_mm10.in
    assert sorted(os.listdir(tmp_path)) == CLI_MM10_FILES
    rows = open(tmp_path / "wes00004_text").read().splitlines()[7:]
    assert len(rows) == 343 and all(len(r) == 26 * 15 for r in rows)


@pytest.mark.parametrize("deck_name", ["test_mm10.in", "test_mm01.in"])
def test_cli_matches_reference_run(deck_name, tmp_path):
    """the CUDA path against flat result files of a real reference run, when a maintainer has provided them under
    tests/golden/reference_run/<deck>/ (recipe there); skips otherwise -- the external pin is open."""
    import os
    import subprocess
    import sys
    from helpers import DECKS
    from test_reference_run_slot import SLOT, reference_files, compare_flat
    files = reference_files(deck_name)
    if not files:
        pytest.skip(f"no reference run under {SLOT}/{deck_name}: parity unpinned")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nstep = max(int(f[3:8]) for f in files)
    out = subprocess.run([sys.executable, "-m", "cpfft_b200", os.path.join(DECKS, deck_name), "--outdir", str(tmp_path),
                          "--steps", str(nstep)], capture_output=True, text=True, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    for f in files:
        compare_flat(str(tmp_path / f), os.path.join(SLOT, deck_name, f), f)


@pytest.mark.parametrize("mixed", [False, True])
def test_taylor_points(libs, mixed):
    """polycrystalline material points (mm10 n_crystals = 3, Taylor average mm10_a.f:112-197;
    `mixed`: fcc and bcc48 crystals in one point through crystal_input file): sweeps along a
    prescribed heterogeneous path with commits, then a full FFT_nr3 solve."""
    from cpfft_b200.polycrystal import taylor_polycrystal
    Solver, Oracle = libs
    p = taylor_polycrystal(5, ncrystals=3, ngrains=12, nstep=4, mixed=mixed)
    s, o = Solver(p), Oracle(p)
    assert s.H == o.H
    nslip = 48 if mixed else 12
    rng = np.random.default_rng(3)
    G = rng.standard_normal((9, p.N3))
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    for step in (1, 2, 3):
        for it, frac in ((0, 0.9), (1, 1.0)):
            F = I + 0.002 * (step - 1 + frac) * (bar + 0.3 * G)
            s.upload("FN1", F); o.Fn1[:] = F
            s.drive_eps_sig(step, it); assert o.drive_eps_sig(step, it) == 0
            assert np.array_equal(s.local_iters(), o.local_iters)
            for name, ref in (("PN1", o.Pn1), ("K4", o.K4)):
                assert relerr(s.download(name), ref) <= TOL_VOXEL
            compare_mm10_history(s.download("HIST_N1", 1)[:, :o.H], o.hist_n1, nslip, TOL_VOXEL, ncrystals=3)
        s.upload("FN", F); o.Fn[:] = F
        s.update(); o.update()
    assert o.local_iters.sum() > 0
    # full solve from a fresh state
    s, o = Solver(p), Oracle(p)
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    rs, ro = s.FFT_nr3(), o.FFT_nr3()
    assert ro["rc"] == 0
    assert list(rs["nr_iters"]) == list(ro["nr_iters"])
    assert_same_cg_counts(rs["cg_iters"], ro["cg_iters"], slack=1)
    assert np.abs(rs["Pbar"] - ro["Pbar"]).max() / np.abs(ro["Pbar"]).max() <= TOL_MACRO
    assert relerr(s.download("FN1"), o.Fn1) <= TOL_VOXEL
