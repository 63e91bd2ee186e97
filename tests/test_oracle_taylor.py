"""Polycrystalline material points (mm10 with n_crystals > 1: Taylor average over the crystals of
a point, mm10_a.f:112-197, history layout mm10_d.f / mm10_a.f:640-641) in the oracle.  Pinned by
what the reference's own loop implies: the point result is the arithmetic mean of independent
single-crystal updates under the same deformation, and each crystal's history block is what
that single-crystal update would have written."""
import numpy as np
import pytest

from helpers import relerr, mm10_layout


@pytest.fixture(scope="module")
def Oracle(oracle_built):
    from oracle import Oracle
    return Oracle


def _drive(o, F_path):
    """prescribed deformation path: per step two sweeps (iter 0, 1), then commit"""
    o.drive_eps_sig(1, 0)
    for step, F in enumerate(F_path, start=1):
        for it in (0, 1):
            o.Fn1[:] = F
            o.drive_eps_sig(step, it)
        o.Fn[:] = o.Fn1
        o.update()


def _path(n3, nsteps=3, seed=5, amp=0.003):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((9, n3))
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45
    I = np.zeros((9, n3)); I[[0, 4, 8]] = 1.0
    return [I + amp * s * (bar + 0.3 * G) for s in range(1, nsteps + 1)]


@pytest.mark.parametrize("mixed", [False, True])
def test_point_is_mean_of_single_crystal_updates(Oracle, mixed):
    from cpfft_b200.polycrystal import taylor_polycrystal
    from cpfft_b200.problem import Problem
    nc = 3
    p = taylor_polycrystal(4, ncrystals=nc, ngrains=6, mixed=mixed)
    o = Oracle(p)
    path = _path(p.N3)
    _drive(o, path)
    nslip = 48 if mixed else 12
    L = mm10_layout(nslip)
    common = L["stress"][0]
    per = L["total"] - common
    assert o.H == common + nc * per
    # independent single-crystal problems, one per crystal slot, SAME history layout rules
    singles = []
    for c in range(nc):
        q = Problem(**{k: getattr(p, k) for k in p.__dataclass_fields__})
        q.materials = [type(p.materials[0])(**{**p.materials[0].__dict__, "n_crystals": 1, "crystal_input": 1,
                                               "crystal": int(p.crystal_ids[0, c]) if mixed else 1})]
        q.angles = np.ascontiguousarray(p.angles[:, c, :]); q.crystal_ids = None
        s = Oracle(q)
        _drive(s, path)
        singles.append(s)
    sig = np.mean([s.urcs_n1[:, :6] for s in singles], axis=0)
    assert relerr(o.urcs_n1[:, :6], sig) <= 1e-13
    for k in (6, 7, 8):                                   # work, plastic work, eq. plastic strain
        acc = np.mean([s.urcs_n1[:, k] for s in singles], axis=0)
        assert relerr(o.urcs_n1[:, k], acc) <= 1e-12
    cep = np.mean([s.hist_n1[:, 0:36] for s in singles], axis=0)
    assert relerr(o.hist_n1[:, 0:36], cep) <= 1e-13
    # K4 / P are functions of the averaged stress and tangent only
    assert np.abs(o.Pn1).max() > 0
    # per-crystal blocks: identical to the single-crystal histories (their own layout may be
    # narrower when the mixed point forces the 48-system layout on an fcc crystal)
    for c, s in enumerate(singles):
        Ls = mm10_layout(48 if (mixed and c % 2 == 1) else 12)
        blk = o.hist_n1[:, common + c * per: common + (c + 1) * per]
        for name in ("stress", "euler", "Rp", "D", "eps", "ep", "ed"):
            a0, a1 = L[name][0] - common, L[name][1] - common
            b0, b1 = Ls[name]
            assert relerr(blk[:, a0:a1], s.hist_n1[:, b0:b1]) <= 1e-13, (c, name)
        a0 = L["tau_tilde"][0] - common; b0 = Ls["tau_tilde"][0]
        assert relerr(blk[:, a0], s.hist_n1[:, b0]) <= 1e-13
        ns = 48 if (mixed and c % 2 == 1) else 12
        a0 = L["slipinc"][0] - common; b0 = Ls["slipinc"][0]
        assert relerr(blk[:, a0:a0 + ns], s.hist_n1[:, b0:b0 + ns]) <= 1e-13
    # accumulated slip: mean of the crystals' slip increments, system by system (mm10_a.f:232,291-292)
    ssum = np.zeros((p.N3, L["slipsum"][1] - L["slipsum"][0]))
    for c, s in enumerate(singles):
        Ls = mm10_layout(48 if (mixed and c % 2 == 1) else 12)
        w = Ls["slipsum"][1] - Ls["slipsum"][0]
        ssum[:, :w] += s.hist_n1[:, Ls["slipsum"][0]:Ls["slipsum"][1]]
    assert relerr(o.hist_n1[:, L["slipsum"][0]:L["slipsum"][1]], ssum / nc) <= 1e-12
    assert o.local_iters.sum() == sum(s.local_iters.sum() for s in singles)


def test_identical_crystals_reduce_to_the_single_crystal_point(Oracle):
    """n_crystals copies of one crystal: stress and tangent of the point are bit-for-bit those of
    the single-crystal point when n_crystals is a power of two (exact sums and divisions)."""
    from cpfft_b200.polycrystal import polycrystal, taylor_polycrystal
    p1 = polycrystal(3, ngrains=5)
    p4 = taylor_polycrystal(3, ncrystals=4, ngrains=5)
    p4.angles = np.ascontiguousarray(np.repeat(p1.angles[:, None, :], 4, axis=1))
    a, b = Oracle(p1), Oracle(p4)
    path = _path(p1.N3)
    _drive(a, path); _drive(b, path)
    assert np.array_equal(a.urcs_n1, b.urcs_n1)
    assert np.array_equal(a.Pn1, b.Pn1) and np.array_equal(a.K4, b.K4)
    assert np.array_equal(4 * a.local_iters, b.local_iters)


def test_full_solve_with_taylor_points(Oracle):
    """FFT_nr3 on a small Taylor polycrystal converges with sane iteration counts and a softer
    response than the elastic one."""
    from cpfft_b200.polycrystal import taylor_polycrystal
    p = taylor_polycrystal(5, ncrystals=2, ngrains=8, nstep=4)
    o = Oracle(p)
    o.drive_eps_sig(1, 0)
    r = o.FFT_nr3(nstep=4)
    assert r["rc"] == 0 and all(1 <= n <= p.maxIter for n in r["nr_iters"])
    sxx = r["Pbar"][:, 0]                      # 0.25 % strain per step: yields in the second step
    assert sxx[0] > 0 and 0 < (sxx[3] - sxx[2]) < 0.6 * sxx[0]
