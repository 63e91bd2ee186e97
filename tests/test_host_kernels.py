"""The CUDA library's per-voxel material code, compiled for the host (tests/host_kernels.py,
tests/native/material_host.cpp), against the CPU oracle: the same comparisons the `-m gpu`
suite makes through the C ABI (tests/test_gpu_parity.py), made on the build box.  What this
cannot see is anything CUDA-specific (launch geometry, shared-memory addressing, FMA
contraction) -- the GPU suite covers that."""
import os

import numpy as np
import pytest

from helpers import deck, relerr, TOL_VOXEL, compare_mm10_history, NSLIP

# Small-strain tolerance: the reference's closed-form polar decomposition (polar.f:224-307) loses
# digits when the principal stretches differ by < 3e-3; two builds of the SAME source then differ
# by a few 1e-9 in R and in everything rotated by it (measured for the oracle itself in
# tests/test_oracle_material.py::test_stress_noise_floor_..., bound 5e-8; the GPU polycrystal
# test uses the same bound).  Local Newton iteration counts are compared exactly.
TOL_SMALL_STRAIN = 1.0e-9


@pytest.fixture(scope="module")
def libs(oracle_built):
    from host_kernels import HostKernels, build
    from oracle import Oracle
    build()
    return HostKernels, Oracle


def _compare_state(k, o, tol=TOL_VOXEL):
    errs = {
        "Pn1": relerr(k.Pn1, o.Pn1),
        "K4": relerr(k.K4, o.K4),
        "urcs_n1": relerr(k.urcs_n1.T, o.urcs_n1),
        "eps_n1": relerr(k.eps_n1.T, o.eps_n1),
    }
    hk = k.hist_n1.T[:, :o.H]
    if any(m.type == 10 for m in o.prob.materials):
        nslip = max(NSLIP[c.slip_type] for c in o.prob.crystals)
        ncry = max(m.n_crystals for m in o.prob.materials if m.type == 10)
        errs.update({"hist." + n: v for n, v in compare_mm10_history(hk, o.hist_n1, nslip, tol, ncry).items()})
    else:
        errs["hist_n1"] = relerr(hk, o.hist_n1)
    bad = {n: v for n, v in errs.items() if not v <= tol}
    assert not bad, f"parity violated: {bad} (all: {errs})"
    return errs


@pytest.mark.parametrize("name", ["test_mm01.in", "test_mm10.in"])
def test_initial_sweep(libs, name):
    """drive_eps_sig(1,0) at F = I: P = 0, elastic K4 (FFT_finite_3d.f:145)."""
    HostKernels, Oracle = libs
    p = deck(name)
    k, o = HostKernels(p), Oracle(p)
    assert k.H == o.H
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    _compare_state(k, o)


@pytest.mark.parametrize("name", ["test_mm01.in", "test_mm10.in"])
def test_sweep_on_perturbed_F(libs, name):
    """nonlinear sweeps (iter = 0, 1, 2) on a heterogeneous, finite deformation field."""
    HostKernels, Oracle = libs
    p = deck(name)
    k, o = HostKernels(p), Oracle(p)
    rng = np.random.default_rng(7)
    amp = 0.05 if name == "test_mm01.in" else 0.004
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += amp * rng.standard_normal((9, p.N3))
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    k.Fn1[:] = F; o.Fn1[:] = F
    for it in (0, 1, 2):
        assert k.drive_eps_sig(1, it) == 0 and o.drive_eps_sig(1, it) == 0
        _compare_state(k, o, TOL_VOXEL if name == "test_mm01.in" else TOL_SMALL_STRAIN)
    if name == "test_mm10.in":
        assert np.array_equal(k.local_iters, o.local_iters)


@pytest.mark.parametrize("kind", ["fcc_polycrystal", "test_mm10.in", "test_mm01.in", "taylor_fcc", "taylor_mixed"])
def test_load_path_with_commits(libs, kind):
    """four load steps along a prescribed heterogeneous deformation path, two sweeps per step,
    n <- n+1 commits in between (update.f:75-106): the history written by one step is the
    input of the next, so layout or scatter mistakes accumulate and show."""
    from cpfft_b200.polycrystal import polycrystal, taylor_polycrystal
    HostKernels, Oracle = libs
    if kind == "fcc_polycrystal":
        p = polycrystal(6, ngrains=20)
    elif kind.startswith("taylor"):       # n_crystals = 3 per material point (mm10_a.f:112-197)
        p = taylor_polycrystal(5, ncrystals=3, ngrains=12, mixed=(kind == "taylor_mixed"))
    else:
        p = deck(kind)
    k, o = HostKernels(p), Oracle(p)
    rng = np.random.default_rng(3)
    G = rng.standard_normal((9, p.N3))               # fixed direction of the fluctuation
    G[[0, 4, 8]] -= G[[0, 4, 8]].mean(axis=0)        # roughly isochoric
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
    amp = 0.01 if kind == "test_mm01.in" else 0.002
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    iters = []
    for step in range(1, 5):
        for it, frac in ((0, 0.9), (1, 1.0)):
            F = I + amp * (step - 1 + frac) * (bar + 0.3 * G)
            k.Fn1[:] = F; o.Fn1[:] = F
            assert k.drive_eps_sig(step, it) == 0 and o.drive_eps_sig(step, it) == 0
            _compare_state(k, o, TOL_SMALL_STRAIN)
            if kind != "test_mm01.in":
                assert np.array_equal(k.local_iters, o.local_iters)
                iters.append(int(o.local_iters.sum()))
        k.Fn[:] = k.Fn1; o.Fn[:] = o.Fn1
        k.update(); o.update()
        assert relerr(k.hist_n.T[:, :o.H], o.hist_n) <= TOL_SMALL_STRAIN
    if kind != "test_mm01.in":
        assert iters[-1] > 0                          # the path reaches the plastic regime


def test_mm10_local_failure_points(libs):
    """the captured failing points of the 256^3 polycrystal (tests/golden/mm10_fail_points.npz):
    same failing set, same local iteration counts, same defined fallback state."""
    from cpfft_b200.polycrystal import polycrystal
    HostKernels, Oracle = libs
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "mm10_fail_points.npz"))
    nb = d["Fn"].shape[1]
    p = polycrystal(2, ngrains=8)
    p.angles[:nb] = d["angles"]
    p.angles[nb:] = d["angles"][0]
    k, o = HostKernels(p), Oracle(p)
    H = o.H
    for v in range(p.N3):
        src = v if v < nb else 0
        k.hist_n[:H, v] = d["hist_n"][:H, src]; k.urcs_n[:, v] = d["urcs_n"][:, src]; k.eps_n[:, v] = d["eps_n"][:, src]
        k.Fn[:, v] = d["Fn"][:, src]
        k.Fn1[:, v] = d["Fn1"][:, src] if v < nb else d["Fn"][:, src] + 0.05 * (d["Fn1"][:, src] - d["Fn"][:, src])
    o.hist_n[:] = k.hist_n.T[:, :H]; o.urcs_n[:] = k.urcs_n.T; o._view("eps_n", (o.N3, 6))[:] = k.eps_n.T
    o.Fn[:] = k.Fn; o.Fn1[:] = k.Fn1
    step, it = int(d["step"]), int(d["iter"])
    assert k.drive_eps_sig(step, it) == nb
    assert o.drive_eps_sig(step, it) == nb
    assert k.fail_flags.sum() == nb and k.fail_flags[:nb].all()
    assert np.array_equal(k.local_iters, o.local_iters)
    assert np.array_equal(k.local_iters[:nb], d["liters"])
    _compare_state(k, o)


def _variant_problem(kind):
    """Edge-case variants of the fcc / bcc48 Voce crystal and of the grid make-up."""
    import copy
    from cpfft_b200.polycrystal import polycrystal
    from cpfft_b200.problem import Material
    p = polycrystal(4, ngrains=9)
    c = p.crystals[0]
    if kind == "cubic_elasticity":           # elastic_type 2: mu independent of e, nu (mod_crystals.f:1840-1860)
        c.elastic_type = 2; c.mu = 60000.0
    elif kind == "voce_m_2":                 # |h|^m with m != 1: the pow path of mm10_h_voche
        c.voche_m = 2.0
    elif kind == "rate_exponent_7p5":        # non-integer harden_n: pow instead of repeated squaring
        c.harden_n = 7.5
    elif kind == "diffusion":                # iD_v != 0: rs*dt*iD_v slip term (mm10_b.f:1805-1835)
        c.iD_v = 2.0e-7
    elif kind == "alter_mode":               # dg = gamma_bar * dt (mm10_a.f:2073)
        c.alter_mode = 1; c.eps_dot_0_y = 4.0e-5; c.harden_n = 5.0; p.tstep = 10.0
    elif kind == "bcc48":
        c.slip_type = 8
    elif kind in ("bcc", "single", "roters", "bcc12"):       # the other cubic slip families (mod_crystals.f:516-811)
        c.slip_type = {"bcc": 2, "single": 3, "roters": 6, "bcc12": 7}[kind]
    elif kind == "mixed_materials":          # mm01 and mm10 voxels in one grid (blocks break at material changes)
        p.materials.append(Material(name="iso", type=1, e=70000.0, nu=0.33, beta=0.5, tan_e=2000.0, yld_pt=150.0))
        p.matlist = p.matlist.copy(); p.matlist[::3] = 2
    elif kind in ("mts", "mts_athermal", "mts_voce_m_2"):     # `hardening mts` (mm10_setup_mts, mm10_b.f:2080-2345)
        from test_oracle_mts import mts_crystal
        kw = {"mts": {}, "mts_athermal": dict(boltzman=0.0, D_0=0.0), "mts_voce_m_2": dict(voche_m=2.0)}[kind]
        p.crystals = [mts_crystal(**kw)]
    elif kind == "mts_and_voce":             # one MTS and one Voce material in the same grid: two kernels
        from test_oracle_mts import mts_crystal
        p.crystals.append(mts_crystal())
        p.materials.append(Material(name="mts", type=10, crystal=2))
        p.matlist = p.matlist.copy(); p.matlist[1::2] = 2
    elif kind == "mts_taylor":               # Taylor points whose crystals follow the MTS law
        from test_oracle_mts import mts_crystal
        from cpfft_b200.polycrystal import taylor_polycrystal
        p = taylor_polycrystal(4, ncrystals=2, ngrains=9)
        p.crystals = [mts_crystal()]
    elif kind == "crystal_file_single":      # n_crystals 1 with `crystal_input file`: crystal number per voxel
        c2 = copy.copy(c); c2.tau_y = 60.0; c2.theta_0 = 300.0; c2.slip_type = 8
        p.crystals.append(c2)
        p.materials[0].crystal_input = 2; p.materials[0].crystal = 0
        ids = np.ones((p.N3, 1), dtype=np.int32); ids[::2] = 2
        p.crystal_ids = ids
    elif kind == "two_crystal_types":        # two library crystals, two cp materials
        c2 = copy.copy(c); c2.tau_y = 60.0; c2.theta_0 = 300.0; c2.e = 120000.0; c2.mu = 120000.0 / 2.6
        p.crystals.append(c2)
        p.materials.append(Material(name="soft", type=10, crystal=2))
        p.matlist = p.matlist.copy(); p.matlist[1::2] = 2
    else:
        raise KeyError(kind)
    return p


VARIANTS = ["cubic_elasticity", "voce_m_2", "rate_exponent_7p5", "diffusion", "alter_mode", "bcc48", "bcc", "single", "roters",
            "bcc12", "mixed_materials",
            "two_crystal_types", "crystal_file_single", "mts", "mts_athermal", "mts_voce_m_2", "mts_and_voce", "mts_taylor"]


@pytest.mark.parametrize("kind", VARIANTS)
def test_crystal_and_grid_variants(libs, kind):
    """three load steps (0.3 % strain each, well into plastic flow) with two sweeps per step"""
    HostKernels, Oracle = libs
    p = _variant_problem(kind)
    k, o = HostKernels(p), Oracle(p)
    assert k.H == o.H
    rng = np.random.default_rng(11)
    G = rng.standard_normal((9, p.N3))
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45; bar[1] = 0.3
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    total = 0
    for step in range(1, 4):
        for it, frac in ((0, 0.8), (1, 1.0)):
            F = I + 0.003 * (step - 1 + frac) * (bar + 0.25 * G)
            k.Fn1[:] = F; o.Fn1[:] = F
            assert k.drive_eps_sig(step, it) == o.drive_eps_sig(step, it)
            assert np.array_equal(k.local_iters, o.local_iters)
            assert np.array_equal(k.fail_flags, np.ctypeslib.as_array(o.L.orc_fail_flags(o.h), shape=(o.N3,)))
            errs = {"Pn1": relerr(k.Pn1, o.Pn1), "K4": relerr(k.K4, o.K4), "urcs": relerr(k.urcs_n1.T, o.urcs_n1)}
            assert max(errs.values()) <= TOL_SMALL_STRAIN, (kind, step, it, errs)
            total += int(o.local_iters.sum())
        k.Fn[:] = k.Fn1; o.Fn[:] = o.Fn1
        k.update(); o.update()
    assert relerr(k.hist_n.T[:, :o.H], o.hist_n) <= 50 * TOL_SMALL_STRAIN     # u(12..14) diagnostics amplify by n
    assert total > 0


@pytest.mark.parametrize("law", ["voce", "mts"])
def test_large_increment_triggers_substepping(libs, law):
    """a 3 % strain increment in one sweep: mm10_solve_strup halves the step (mm10_a.f:2759-2843);
    iteration counts (which include the failed attempts) and results must agree.  MTS: the
    sub-step state is set up at temperature 297 (step + frac) (n%temp = 0, mm10_a.f:2769)."""
    from cpfft_b200.polycrystal import polycrystal
    HostKernels, Oracle = libs
    p = polycrystal(4, ngrains=9)
    if law == "mts":
        from test_oracle_mts import mts_crystal
        p.crystals = [mts_crystal(miter=8)]       # a solve that needs more than 8 iterations is cut
    k, o = HostKernels(p), Oracle(p)
    rng = np.random.default_rng(2)
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.5; bar[5] = 0.4
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    F1 = I + 0.004 * (bar + 0.2 * rng.standard_normal((9, p.N3)))
    for F, step in ((F1, 1), (I + 0.034 * (bar + 0.2 * rng.standard_normal((9, p.N3))), 2)):
        for it in (0, 1):
            k.Fn1[:] = F; o.Fn1[:] = F
            nf_k, nf_o = k.drive_eps_sig(step, it), o.drive_eps_sig(step, it)
            assert nf_k == nf_o
            if law == "voce":
                assert np.array_equal(k.local_iters, o.local_iters)
            else:
                # miter = 8 makes the stress PREDICTOR fail; the reference then still runs the update
                # phase before it cuts the step (mm10_a.f:2944-2960, fail stays set), the kernel cuts
                # at once -- same result, but the update-iteration counter of such attempts differs
                assert np.array_equal(k.local_iters[:, 0], o.local_iters[:, 0])
        k.Fn[:] = F; o.Fn[:] = F
        k.update(); o.update()
    # more update iterations than a plain solve needs / than miter allows: sub-steps were taken somewhere
    assert (o.local_iters[:, 1].max() > 12) if law == "voce" else (o.local_iters[:, 0].max() > 8)
    ok = np.ctypeslib.as_array(o.L.orc_fail_flags(o.h), shape=(o.N3,)) == 0
    assert relerr(k.urcs_n.T[ok], o.urcs_n[ok]) <= TOL_SMALL_STRAIN


def _hybrid_FFT_nr3(k, o, prob, nstep):
    """FFT_nr3's strain-controlled step / Newton loop (FFT_nr3.f:51-185) in Python: material sweeps by
    the host build of the KERNEL source, spectral operator and CG by the oracle."""
    bc = prob.BC_all()
    barF = np.array([1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0])
    nr, pbar = [], []
    k.drive_eps_sig(1, 0)
    for step in range(1, nstep + 1):
        dF = np.repeat((bc[step - 1] - barF)[:, None], prob.N3, axis=1)
        Fnorm = np.linalg.norm(k.Fn1)
        k.Fn1[:] += dF
        o.K4[:] = k.K4
        rc, x, it, rr = o.fftPcg(-o.G_K_dF(dF, 1), prob.tolPCG)
        assert rc == 0
        k.Fn1[:] += x
        res, it_nr = 1.0, 0
        while res > prob.tolNR:
            k.drive_eps_sig(step, it_nr)
            o.K4[:] = k.K4
            rc, x, it, rr = o.fftPcg(-o.G_K_dF(np.ascontiguousarray(k.Pn1), 0), prob.tolPCG)
            assert rc == 0
            k.Fn1[:] += x
            res = np.linalg.norm(x) / Fnorm
            it_nr += 1
            assert it_nr < prob.maxIter
        k.drive_eps_sig(step, it_nr)
        pbar.append(k.Pn1.mean(axis=1))
        k.Fn[:] = k.Fn1
        k.update()
        barF = bc[step - 1].copy()
        nr.append(it_nr)
    return nr, np.array(pbar)


@pytest.mark.parametrize("name,nstep", [("test_mm10.in", 6), ("test_mm01.in", 4), ("taylor_mm10.in", 5), ("mts_mm10.in", 4)])
def test_full_solve_with_kernel_source(libs, name, nstep):
    """whole load steps with the kernels' own source doing the material sweeps: Newton iteration
    counts and the homogenised stress must be those of the pure oracle run (1e-10)"""
    HostKernels, Oracle = libs
    p = deck(name)
    o_ref = Oracle(p)
    o_ref.drive_eps_sig(1, 0)
    r = o_ref.FFT_nr3(nstep=nstep)
    assert r["rc"] == 0
    k, o = HostKernels(p), Oracle(p)
    nr, pbar = _hybrid_FFT_nr3(k, o, p, nstep)
    assert nr == [int(v) for v in r["nr_iters"]]
    # mts_mm10.in: slowly converging Newton loop, round-off differences of the two
    # implementations are amplified to 2e-10 on the curve; the other decks meet 1e-10
    assert np.abs(pbar - r["Pbar"]).max() / np.abs(r["Pbar"]).max() <= (1e-9 if name == "mts_mm10.in" else 1e-10)
    assert relerr(k.Fn1, o_ref.Fn1) <= (1e-8 if name == "mts_mm10.in" else 1e-9)


def test_lattice_frame_residual_variant(libs):
    """the residual slip loop exists twice (mm10.cuh mm10_resid<.., LF>): in the lattice frame (the
    product's default and the default of HostKernels, i.e. what every other test here runs) and in the
    sample frame (CPFFT_MM10_LF=0, this test).  Same algebra in another summation order: same results
    to round-off and the SAME local iteration counts along a plastic load path with a large,
    sub-stepped increment, and the same Newton counts / curve on the test deck."""
    from cpfft_b200.polycrystal import polycrystal
    HostKernels, Oracle = libs
    p = polycrystal(6, ngrains=20)
    k, o = HostKernels(p, lattice_frame=False), Oracle(p)
    rng = np.random.default_rng(3)
    G = rng.standard_normal((9, p.N3)); G[[0, 4, 8]] -= G[[0, 4, 8]].mean(axis=0)
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    plastic = 0
    for step, amp in enumerate((0.002, 0.004, 0.006, 0.03), start=1):     # last step: 2.4 % increment
        for it, frac in ((0, 0.9), (1, 1.0)):
            F = I + amp * frac * (bar + 0.3 * G)
            k.Fn1[:] = F; o.Fn1[:] = F
            assert k.drive_eps_sig(step, it) == o.drive_eps_sig(step, it)
            ok = np.ctypeslib.as_array(o.L.orc_fail_flags(o.h), shape=(o.N3,)) == 0
            assert relerr(k.urcs_n1.T[ok], o.urcs_n1[ok]) <= TOL_SMALL_STRAIN
            assert relerr(k.K4[:, ok], o.K4[:, ok]) <= TOL_SMALL_STRAIN
            assert np.array_equal(k.local_iters, o.local_iters)
            plastic += int(o.local_iters[:, 1].sum())
        k.Fn[:] = k.Fn1; o.Fn[:] = o.Fn1
        k.update(); o.update()
    assert plastic > 0 and o.local_iters[:, 1].max() > 12          # plastic, and sub-steps were taken
    # whole solve on the reference's deck
    p = deck("test_mm10.in")
    o_ref = Oracle(p); o_ref.drive_eps_sig(1, 0)
    r = o_ref.FFT_nr3(nstep=6)
    nr, pbar = _hybrid_FFT_nr3(HostKernels(p, lattice_frame=False), Oracle(p), p, 6)
    assert nr == [int(v) for v in r["nr_iters"]]
    assert np.abs(pbar - r["Pbar"]).max() / np.abs(r["Pbar"]).max() <= 1e-10


def test_taylor_point_with_failing_crystals(libs):
    """a Taylor point whose crystals are the captured failing states (tests/golden/mm10_fail_points.npz):
    crystal 0 fails, crystal 1 (the same state with a benign increment history) may or may not; the
    point is flagged, the failed crystal keeps its n state and elastic tangent, the averages and the
    summed iteration counts agree between the kernel source and the oracle."""
    from cpfft_b200.polycrystal import taylor_polycrystal
    from helpers import mm10_layout
    HostKernels, Oracle = libs
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "mm10_fail_points.npz"))
    nb = d["Fn"].shape[1]
    p = taylor_polycrystal(2, ncrystals=2, ngrains=8)
    ang = np.asarray(p.angles).copy()
    for v in range(p.N3):
        ang[v, 0] = d["angles"][v % nb]
        ang[v, 1] = d["angles"][(v + 1) % nb]
    p.angles = ang
    k, o = HostKernels(p), Oracle(p)
    L = mm10_layout(12)
    common, per = L["stress"][0], L["total"] - L["stress"][0]
    assert k.H == o.H == common + 2 * per
    H1 = d["hist_n"].shape[0]
    for v in range(p.N3):
        src0, src1 = v % nb, (v + 1) % nb
        k.hist_n[:common, v] = d["hist_n"][:common, src0]
        k.hist_n[common:common + per, v] = d["hist_n"][common:common + per, src0]
        k.hist_n[common + per:common + 2 * per, v] = d["hist_n"][common:common + per, src1]
        k.urcs_n[:, v] = d["urcs_n"][:, src0]; k.eps_n[:, v] = d["eps_n"][:, src0]
        k.Fn[:, v] = d["Fn"][:, src0]
        k.Fn1[:, v] = d["Fn1"][:, src0]
    o.hist_n[:] = k.hist_n.T[:, :o.H]; o.urcs_n[:] = k.urcs_n.T; o._view("eps_n", (o.N3, 6))[:] = k.eps_n.T
    o.Fn[:] = k.Fn; o.Fn1[:] = k.Fn1
    step, it = int(d["step"]), int(d["iter"])
    nf_k, nf_o = k.drive_eps_sig(step, it), o.drive_eps_sig(step, it)
    assert nf_k == nf_o == p.N3                       # crystal 0 of every point fails
    assert np.array_equal(k.local_iters, o.local_iters)
    assert k.fail_flags.all()
    assert relerr(k.urcs_n1.T, o.urcs_n1) <= TOL_SMALL_STRAIN
    assert relerr(k.K4, o.K4) <= TOL_SMALL_STRAIN
    compare_mm10_history(k.hist_n1.T[:, :o.H], o.hist_n1, 12, TOL_SMALL_STRAIN, ncrystals=2)
    # the failed crystal's block holds its n state
    a, b = L["stress"]
    assert np.array_equal(k.hist_n1[a:b], k.hist_n[a:b])


def test_material_table_errors(libs, capfd):
    """usage errors of cpfft_set_voxels[_taylor] are caught when the tables are built (material_tables.hpp)"""
    import copy
    from cpfft_b200.polycrystal import polycrystal, taylor_polycrystal
    from test_oracle_mts import mts_crystal
    HostKernels, _ = libs
    # crystals of one cp material with different hardening laws
    p = taylor_polycrystal(2, ncrystals=2, ngrains=4, mixed=True)
    p.crystals[1] = mts_crystal()
    with pytest.raises(RuntimeError):
        HostKernels(p)
    assert "share one hardening law" in capfd.readouterr().err
    # a voxel that names a crystal outside the library
    p = taylor_polycrystal(2, ncrystals=2, ngrains=4, mixed=True)
    p.crystal_ids = p.crystal_ids.copy(); p.crystal_ids[3, 1] = 7
    with pytest.raises(RuntimeError):
        HostKernels(p)
    assert "undefined crystal" in capfd.readouterr().err
    # unsupported hardening law / slip family
    for field, val, msg in (("h_type", 4, "hardening law"), ("slip_type", 9, "slip_type")):
        p = polycrystal(2, ngrains=4)
        c = copy.copy(p.crystals[0]); setattr(c, field, val); p.crystals = [c]
        with pytest.raises(RuntimeError):
            HostKernels(p)
        assert msg in capfd.readouterr().err


def test_shared_memory_lu_against_numpy():
    """mm10_lu7_factor / mm10_lu7_solve (the 7x7 solve of the Newton step, the tangent and the lattice strain): random
    matrices that need row interchanges in every column pattern, against numpy's LAPACK solve, and the pivot rows
    against a plain partial-pivoting elimination (first maximum wins, as in DGETRF)."""
    import ctypes as C
    from host_kernels import _lib
    L = _lib()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(11)
    swaps = 0
    for trial in range(400):
        A = rng.standard_normal((7, 7))
        if trial % 4 == 0:
            A += 8.0 * np.eye(7)                       # diagonally dominant: no interchange at all
        if trial % 4 == 1:
            A[:, rng.integers(0, 7)] *= 1e3            # one badly scaled column
        b = rng.standard_normal(7)
        x = b.copy()
        piv = L.mh_lu7(np.ascontiguousarray(A).ctypes.data_as(dp), x.ctypes.data_as(dp))
        ref = np.linalg.solve(A, b)
        assert np.abs(x - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()) * np.linalg.cond(A)
        # pivot rows of a textbook elimination
        M = A.copy(); want = 0
        for k in range(7):
            p = k + int(np.argmax(np.abs(M[k:, k])))
            want |= p << (3 * k)
            M[[k, p]] = M[[p, k]]
            for i in range(k + 1, 7):
                M[i, k:] -= M[i, k] / M[k, k] * M[k, k:]
        assert piv == want
        swaps += sum(((piv >> (3 * k)) & 7) != k for k in range(7))
    assert swaps > 400                                 # the interchange path was exercised


def test_integer_power_matches_repeated_squaring():
    """cpf_pow_abs: the straight-line square-and-multiply equals the loop it replaced bit for bit for every
    exponent 0..64, and falls back to pow() for non-integer exponents"""
    from host_kernels import _lib
    L = _lib()
    rng = np.random.default_rng(5)
    for x in list(rng.uniform(0.0, 1.3, 40)) + [0.0, 1.0]:
        for ie in range(65):
            r, b, e = 1.0, float(x), ie
            while e:
                if e & 1:
                    r *= b
                b *= b
                e >>= 1
            assert L.mh_pow_abs(float(x), ie, float(ie)) == r
    assert abs(L.mh_pow_abs(0.7, -1, 18.5) - 0.7 ** 18.5) <= 1e-15


def test_mts_with_48_systems_partial_handover(libs):
    """MTS hardening on the 48-system bcc family: the residual hands |rs/tt|^(n-1) to the Jacobian for the first
    32 systems only (slots 32..38 of the shared `acc` tile hold the stashed J12 / J22 of the factored Jacobian,
    mm10.cuh MM10_SM_STASH), the other 16 are recomputed -- same results and local iteration counts as the oracle."""
    import dataclasses
    from cpfft_b200.polycrystal import polycrystal, workload_variant
    HostKernels, Oracle = libs
    p = workload_variant(polycrystal(5, ngrains=12), "mts", 12)
    p.crystals = [dataclasses.replace(p.crystals[0], slip_type=8)]
    k, o = HostKernels(p), Oracle(p)
    rng = np.random.default_rng(3)
    G = rng.standard_normal((9, p.N3)); G[[0, 4, 8]] -= G[[0, 4, 8]].mean(axis=0)
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    k.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    plastic = 0
    for step, amp in enumerate((0.001, 0.002, 0.003, 0.01), start=1):
        for it, frac in ((0, 0.9), (1, 1.0)):
            F = I + amp * frac * (bar + 0.3 * G)
            k.Fn1[:] = F; o.Fn1[:] = F
            assert k.drive_eps_sig(step, it) == o.drive_eps_sig(step, it)
            ok = np.ctypeslib.as_array(o.L.orc_fail_flags(o.h), shape=(o.N3,)) == 0
            assert relerr(k.urcs_n1.T[ok], o.urcs_n1[ok]) <= TOL_SMALL_STRAIN
            assert relerr(k.K4[:, ok], o.K4[:, ok]) <= TOL_SMALL_STRAIN
            assert np.array_equal(k.local_iters, o.local_iters)
            plastic += int(o.local_iters[:, 1].sum())
        k.Fn[:] = k.Fn1; o.Fn[:] = o.Fn1
        k.update(); o.update()
    assert plastic > 0
