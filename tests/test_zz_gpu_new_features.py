"""GPU tests of the features written after the round-1 GPU budget was spent (MTS hardening, crystal / grid
variants incl. the extra cubic slip families, direct numpy comparison of the fast spectral path).  They
were verified on the CPU through the host build of the kernel source (tests/test_host_kernels.py) and
are collected in this file, which sorts last, so that a surprise here cannot hide the results of the
established GPU tests under `pytest -x`."""
import numpy as np
import pytest

from helpers import relerr, deck, TOL_MACRO, assert_same_cg_counts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(oracle_built):
    from cpfft_b200 import Solver
    from oracle import Oracle
    return Solver, Oracle


def _variant_kinds():
    from test_host_kernels import VARIANTS
    return VARIANTS


@pytest.mark.parametrize("kind", _variant_kinds())
def test_crystal_and_grid_variants(libs, kind):
    """edge-case variants of the Voce crystal and of the grid make-up (tests/test_host_kernels.py runs
    the same cases on the host build of the kernel source): three load steps, two sweeps each.
    Local Newton counts: pow / log / exp of the device and of the host libm differ in the last bit, so
    a point sitting exactly on a convergence threshold may take an iteration more or less; allowed on
    at most 10 % of the 64 points (the deck-level GPU tests assert exact equality)."""
    from test_host_kernels import _variant_problem
    Solver, Oracle = libs
    p = _variant_problem(kind)
    s, o = Solver(p), Oracle(p)
    assert s.H == o.H
    rng = np.random.default_rng(11)
    G = rng.standard_normal((9, p.N3))
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45; bar[1] = 0.3
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    s.drive_eps_sig(1, 0); o.drive_eps_sig(1, 0)
    for step in range(1, 4):
        for it, frac in ((0, 0.8), (1, 1.0)):
            F = I + 0.003 * (step - 1 + frac) * (bar + 0.25 * G)
            s.upload("FN1", F); o.Fn1[:] = F
            s.drive_eps_sig(step, it); o.drive_eps_sig(step, it)
            d = np.abs(s.local_iters() - o.local_iters)
            assert d.max() <= 2 and (d > 0).mean() <= 0.10, (kind, step, it, d.max(), (d > 0).mean())
            same = (d.sum(axis=1) == 0)
            for name, ref in (("PN1", o.Pn1), ("K4", o.K4)):
                got = s.download(name)
                assert relerr(got[:, same], ref[:, same]) <= 1e-9, (kind, step, it, name)
        s.upload("FN", F); o.Fn[:] = F
        s.update(); o.update()
    assert o.local_iters.sum() > 0


@pytest.mark.parametrize("N", [16, 32, 64])
def test_fast_path_matches_plain_fft_statement(libs, N):
    """the GPU operator against numpy's FFT directly (plain statement of the even-N convention,
    tests/test_oracle_spectral.py::conv_G_K_dF), not through the oracle"""
    from test_oracle_spectral import _toy_problem, conv_G_K_dF
    Solver, _ = libs
    p = _toy_problem(N)
    s = Solver(p)
    rng = np.random.default_rng(N)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.002 * rng.standard_normal((9, p.N3))
    s.upload("FN1", F)
    s.drive_eps_sig(1, 1)
    K4 = s.download("K4")
    x = rng.standard_normal((9, p.N3))
    s.upload("DFM", x)
    for flgK in (0, 1):
        s.G_K_dF("DFM", "B", flgK)
        want = conv_G_K_dF(N, x, K4 if flgK else None)
        assert relerr(s.download("B"), want) <= 1e-12, (N, flgK)


def test_mts_deck_matches_golden():
    """the derived MTS deck against the committed oracle fixture (tests/golden/deck_results.json)"""
    import json
    import os
    from cpfft_b200 import Solver
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "deck_results.json")))["mts_mm10.in"]
    s = Solver(deck("mts_mm10.in"))
    s.drive_eps_sig(1, 0)
    r = s.FFT_nr3()
    assert [int(v) for v in r["nr_iters"]] == gold["nr_iters"]
    assert_same_cg_counts(r["cg_iters"], gold["cg_iters"], 2)
    ref = np.array(gold["Pbar"])
    assert np.abs(r["Pbar"] - ref).max() / np.abs(ref).max() <= 1e-9
    assert abs(np.abs(s.download("PN1")).max() / gold["P_absmax"] - 1.0) <= 1e-7


def test_lattice_frame_variant_in_subprocess():
    """CPFFT_MM10_LF=0 CPFFT_MM10_UNI=0 (k_update_mm10: residual slip loop in the sample frame, crystal
    constants from the crystal table instead of the kernel parameters; the switches are read once per
    process, hence the subprocess): same Newton / CG counts and the same curve as the default kernel
    (k_update_mm10_lf_u) on the reference's crystal-plasticity deck -- the golden fixture."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ("import json, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from helpers import deck\nfrom cpfft_b200 import Solver\n"
            "s = Solver(deck('test_mm10.in')); s.drive_eps_sig(1, 0); r = s.FFT_nr3()\n"
            "print('RESULT ' + json.dumps({'nr': [int(v) for v in r['nr_iters']], 'cg': [[int(v) for v in row] for row in r['cg_iters']],"
            " 'Pbar': r['Pbar'].tolist(), 'launches': s.kernel_launches()}))\n") % (here, os.path.dirname(here))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CPFFT_MM10_LF="0", CPFFT_MM10_UNI="0"), capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads([l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    gold = json.load(open(os.path.join(here, "golden", "deck_results.json")))["test_mm10.in"]
    assert got["launches"] > 0
    assert got["nr"] == gold["nr_iters"]
    assert_same_cg_counts(got["cg"], gold["cg_iters"], 1)
    ref = np.array(gold["Pbar"])
    assert np.abs(np.array(got["Pbar"]) - ref).max() / np.abs(ref).max() <= 1e-10
