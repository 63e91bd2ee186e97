"""Pinning the oracle's spectral operator (oracle_solver.cpp: orc_G_K_dF, green_entry, phase
ramp) without the reference binary:

  * an INDEPENDENT numpy restatement that follows the reference literally -- stored Ghat4 table
    (FFT_init.f:272-340), fftshift phase ramp coeffs1/2 (FFT_init.f:354-385), full complex 3-D
    transforms with split re/im contraction and "keep the real part" (G_K_dF.f:11-227);
  * the projection identities the Green operator must satisfy for odd N;
  * the documented even-N convention (Nyquist planes zeroed; the reference itself is wrong
    for even N).
"""
import numpy as np
import pytest

from helpers import deck


@pytest.fixture(scope="module")
def Oracle(oracle_built):
    from oracle import Oracle
    return Oracle


# ---------------------------------------------------------------------------------------------
# literal numpy restatement of the reference
def ref_indices():
    """indices(4,81) of formG: column m = 27 i + 9 j + 3 k + l (FFT_init.f:283-304)."""
    m = np.arange(81)
    return m // 27, (m // 9) % 3, (m // 3) % 3, m % 3


def ref_formG(N):
    Nhalf = (N + 1) // 2 if N % 2 == 1 else N // 2 + 1
    r = np.arange(1, N + 1) - Nhalf
    q = np.stack(np.meshgrid(r, r, r, indexing="ij"), axis=-1).reshape(-1, 3).astype(float)  # ii slowest, kk fastest
    qq = (q * q).sum(axis=1)
    G = np.zeros((N ** 3, 81))
    i, j, k, l = ref_indices()
    for m in range(81):
        if i[m] == k[m]:
            G[:, m] = q[:, j[m]] * q[:, l[m]]
    zero = np.abs(qq) <= 1e-10
    G[zero] = 0.0
    G[~zero] /= qq[~zero, None]
    return G


def ref_shift(N):
    Nhalf = (N + 1) // 2 if N % 2 == 1 else N // 2 + 1
    NhN = Nhalf * 2.0 * np.pi / N
    a = np.arange(N)
    s = a[:, None, None] + a[None, :, None] + a[None, None, :]
    return np.cos(NhN * s).ravel(), -np.sin(NhN * s).ravel()


def ref_ddot42n(A4, B2):
    """C2(:,i) = sum_j A4(:,9 i + j) B2(:,j) (G_K_dF.f:241-268); A4 (N3,81), B2 (N3,9)."""
    return np.einsum("eij,ej->ei", A4.reshape(-1, 9, 9), B2)


def ref_G_K_dF(N, F, K4=None):
    """F (9,N3) -> (9,N3), exactly the data flow of G_K_dF.f:11-87."""
    c1, c2 = ref_shift(N)
    G = ref_formG(N)
    x = F.T if K4 is None else ref_ddot42n(K4.T, F.T)         # (N3, 9)
    re, im = np.empty_like(x), np.empty_like(x)
    for c in range(9):                                         # fftfem3d
        z = np.fft.fftn((x[:, c] * c1 + 1j * x[:, c] * c2).reshape(N, N, N))
        re[:, c], im[:, c] = z.real.ravel(), z.imag.ravel()
    gre, gim = ref_ddot42n(G, re), ref_ddot42n(G, im)          # Ghat is real
    out = np.empty_like(x)
    for c in range(9):                                         # ifftfem3d: re*c1 + im*c2
        z = np.fft.ifftn((gre[:, c] + 1j * gim[:, c]).reshape(N, N, N))
        out[:, c] = z.real.ravel() * c1 + z.imag.ravel() * c2
    return out.T


# ---------------------------------------------------------------------------------------------
def _toy_problem(N):
    from cpfft_b200.problem import Problem, Material
    mats = [Material(name="a", type=1, e=12000.0, nu=0.3, beta=0.5, tan_e=1000.0, yld_pt=100.0),
            Material(name="b", type=1, e=24000.0, nu=0.3, beta=0.5, tan_e=1000.0, yld_pt=200.0)]
    rng = np.random.default_rng(N)
    ml = rng.integers(1, 3, N ** 3).astype(np.int32)
    return Problem(N=N, materials=mats, crystals=[], matlist=ml, angles=np.zeros((N ** 3, 3)),
                   FP_max=np.zeros(9), mults=np.ones(1))


@pytest.mark.parametrize("N", [3, 5, 7, 9])
def test_formG_entries_match_reference_table(Oracle, N):
    G = ref_formG(N)
    rng = np.random.default_rng(1)
    for _ in range(40):
        ii, jj, kk = rng.integers(0, N, 3)
        e = (ii * N + jj) * N + kk
        assert np.array_equal(Oracle.formG_entry(N, int(ii), int(jj), int(kk)), G[e])


@pytest.mark.parametrize("N", [5, 7, 9, 11, 13, 15, 21, 25])      # radix 3 / 5 passes, generic prime passes and their products
def test_G_K_dF_matches_numpy_restatement(Oracle, N):
    p = _toy_problem(N)
    o = Oracle(p)
    rng = np.random.default_rng(3)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.02 * rng.standard_normal((9, p.N3))
    o.Fn1[:] = F
    o.drive_eps_sig(1, 1)                      # heterogeneous K4
    x = rng.standard_normal((9, p.N3))
    for flgK in (0, 1):
        got = o.G_K_dF(x, flgK)
        want = ref_G_K_dF(N, x, o.K4.copy() if flgK else None)
        assert np.abs(got - want).max() <= 2e-13 * np.abs(want).max()


def test_G_K_dF_on_shipped_deck(Oracle):
    """N = 7 bicrystal of test_mm10.in, K4 from a real mm10 sweep."""
    p = deck("test_mm10.in")
    o = Oracle(p)
    rng = np.random.default_rng(5)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.002 * rng.standard_normal((9, p.N3))
    o.drive_eps_sig(1, 0)
    o.Fn1[:] = F
    assert o.drive_eps_sig(1, 1) == 0
    x = rng.standard_normal((9, p.N3))
    want = ref_G_K_dF(7, x, o.K4.copy())
    assert np.abs(o.G_K_dF(x, 1) - want).max() <= 2e-13 * np.abs(want).max()


def _grad_field(N, rng, kmax):
    """gradient of a periodic, band-limited displacement field, F_ij = d u_i / d X_j"""
    a = np.arange(N)
    X = np.stack(np.meshgrid(a, a, a, indexing="ij"), axis=0).astype(float)  # X[0] = x index (slowest)
    grad = np.zeros((3, 3, N, N, N))
    for _ in range(6):
        k = rng.integers(-kmax, kmax + 1, 3)
        if not k.any():
            continue
        ph = 2 * np.pi * (k[0] * X[0] + k[1] * X[1] + k[2] * X[2]) / N + rng.uniform(0, 2 * np.pi)
        amp = rng.standard_normal(3)
        for i in range(3):
            for j in range(3):
                grad[i, j] += amp[i] * (2 * np.pi * k[j] / N) * np.cos(ph)
    return grad.reshape(9, -1)


@pytest.mark.parametrize("N", [5, 7, 6, 8, 12, 16, 20, 24])       # even: radix 2 / 4 passes
def test_green_projection_identities(Oracle, N):
    """Ghat:grad(u) = grad(u), Ghat:const = 0, idempotence, self-adjointness.  For even N the
    identities hold on fields without Nyquist content (documented convention)."""
    o = Oracle(_toy_problem(N))
    rng = np.random.default_rng(N)
    g = _grad_field(N, rng, kmax=(N - 1) // 2)
    assert np.abs(o.G_K_dF(g, 0) - g).max() <= 1e-12 * np.abs(g).max()
    const = np.repeat(rng.standard_normal((9, 1)), N ** 3, axis=1)
    assert np.abs(o.G_K_dF(const, 0)).max() <= 1e-12
    x, y = rng.standard_normal((9, N ** 3)), rng.standard_normal((9, N ** 3))
    Gx, Gy = o.G_K_dF(x, 0), o.G_K_dF(y, 0)
    assert np.abs(o.G_K_dF(Gx, 0) - Gx).max() <= 1e-12 * np.abs(Gx).max()
    assert abs((Gx * y).sum() - (x * Gy).sum()) <= 1e-11 * np.abs(x).max() * np.abs(y).max() * N ** 3


def test_reference_table_is_broken_for_even_N():
    """the literal reference operator is not a projection for even N, which
    is why even grids use the corrected convention instead of the reference's table."""
    rng = np.random.default_rng(0)
    for N, ok in ((5, True), (6, False)):
        g = _grad_field(N, rng, kmax=1)
        err = np.abs(ref_G_K_dF(N, g) - g).max() / np.abs(g).max()
        assert (err <= 1e-12) == ok


def conv_G_K_dF(N, F, K4=None):
    """The documented even-N convention stated WITHOUT the phase ramp: plain 3-D FFT, signed integer
    frequencies xi = -N/2+1 .. N/2-1, Ghat_ijkl = delta_ik xi_j xi_l / |xi|^2, zero at xi = 0 and on the
    three Nyquist planes.  (For odd N this is algebraically the reference's operator.)"""
    x = F.T if K4 is None else ref_ddot42n(K4.T, F.T)
    xi1 = np.fft.fftfreq(N, d=1.0 / N)
    xi = np.stack(np.meshgrid(xi1, xi1, xi1, indexing="ij"), axis=-1).reshape(-1, 3)
    qq = (xi * xi).sum(axis=1)
    nyq = (np.abs(xi) == N / 2).any(axis=1) if N % 2 == 0 else np.zeros(len(xi), bool)
    zero = nyq | (qq == 0)
    X = np.stack([np.fft.fftn(x[:, c].reshape(N, N, N)).ravel() for c in range(9)], axis=1).reshape(-1, 3, 3)
    # out_ij = xi_j (sum_l xi_l X_il) / |xi|^2
    s = np.einsum("eil,el->ei", X, xi)
    s[~zero] /= qq[~zero, None]
    s[zero] = 0.0
    Y = (s[:, :, None] * xi[:, None, :]).reshape(-1, 9)
    out = np.stack([np.fft.ifftn(Y[:, c].reshape(N, N, N)).real.ravel() for c in range(9)], axis=1)
    return out.T


@pytest.mark.parametrize("N", [5, 6, 8, 9, 12, 16])
def test_G_K_dF_matches_plain_fft_statement_of_the_convention(Oracle, N):
    p = _toy_problem(N)
    o = Oracle(p)
    rng = np.random.default_rng(N)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.02 * rng.standard_normal((9, p.N3))
    o.Fn1[:] = F
    o.drive_eps_sig(1, 1)
    x = rng.standard_normal((9, p.N3))
    for flgK in (0, 1):
        got = o.G_K_dF(x, flgK)
        want = conv_G_K_dF(N, x, o.K4.copy() if flgK else None)
        assert np.abs(got - want).max() <= 5e-13 * np.abs(want).max(), (N, flgK)


def test_tangent_homo_is_the_derivative_of_the_mean_stress(Oracle):
    """tangent_homo (tangent_homo.f:11-73): C_homo = d P_bar / d F_bar.  Checked by finite differences
    of whole FFT_nr3 solves of a two-phase elastic block under F_bar = I + delta e_kl."""
    from cpfft_b200.problem import Problem
    base = _toy_problem(5)
    o = Oracle(base)
    o.drive_eps_sig(1, 0)
    rc, C = o.tangent_homo()
    assert rc == 0
    C = C.reshape(9, 9)                        # row ij, column kl (FFT_init.f:283-304 ordering)
    delta = 1.0e-6
    fd = np.zeros((9, 9))
    for kl in range(9):
        cols = []
        for sgn in (+1.0, -1.0):
            p = Problem(**{k: getattr(base, k) for k in base.__dataclass_fields__})
            p.FP_max = np.zeros(9); p.FP_max[kl] = sgn * delta
            p.mults = np.ones(1); p.tolNR = 1.0e-9; p.tolPCG = 1.0e-12; p.maxIter = 20
            q = Oracle(p)
            q.drive_eps_sig(1, 0)
            r = q.FFT_nr3()
            assert r["rc"] == 0
            cols.append(r["Pbar"][0])
        fd[:, kl] = (cols[0] - cols[1]) / (2.0 * delta)
    assert np.abs(C - fd).max() <= 2e-5 * np.abs(fd).max(), np.abs(C - fd).max() / np.abs(fd).max()
    # a heterogeneous block is softer than the volume average of the phase tangents (Voigt bound)
    voigt = o.K4.mean(axis=1).reshape(9, 9)
    assert C[0, 0] < voigt[0, 0] and C[0, 0] > 0.5 * voigt[0, 0]


def test_stress_boundary_conditions_are_met(Oracle):
    """NBC_update + the stress-BC loop (FFT_nr3.f:127-165, 375-433): the prescribed mean stresses are
    reached to the loop's own stop test, and the free strains contract (uniaxial tension)."""
    from helpers import stress_bc_variant
    p = stress_bc_variant(deck("test_mm01.in"))
    o = Oracle(p)
    o.drive_eps_sig(1, 0)
    r = o.FFT_nr3(nstep=3)
    assert r["rc"] == 0
    for P in r["Pbar"]:
        assert np.sqrt(P[4] ** 2 + P[8] ** 2) <= p.tolNR * np.linalg.norm(P)
        assert P[0] > 0
    Fbar = o.Fn1.mean(axis=1)
    assert Fbar[0] > 1.0 and Fbar[4] < 1.0 and Fbar[8] < 1.0


def test_fftPcg_matches_a_textbook_cg(Oracle):
    """fftPcg (FFT_nr3.f:214-360, MKL RCI dcg with the user stop test ||r|| <= tol ||b|| or ||r|| <= tol
    checked before every iteration): same iteration count and solution as a plain numpy CG on the
    same operator."""
    p = _toy_problem(7)
    o = Oracle(p)
    rng = np.random.default_rng(2)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.002 * rng.standard_normal((9, p.N3))
    o.Fn1[:] = F
    o.drive_eps_sig(1, 1)
    b = -o.G_K_dF(rng.standard_normal((9, p.N3)), 1)
    tol = 1e-9
    rc, x, it, rr = o.fftPcg(b, tol)
    assert rc == 0
    xs = np.zeros_like(b); r = b.copy(); pvec = None; rr_old = None; k = 0
    nb = np.linalg.norm(b)
    while True:
        rn = np.linalg.norm(r)
        if rn <= tol * nb or rn <= tol:
            break
        rho = float((r * r).sum())
        pvec = r.copy() if pvec is None else r + (rho / rr_old) * pvec
        q = o.G_K_dF(pvec, 1)
        alpha = rho / float((pvec * q).sum())
        xs += alpha * pvec; r -= alpha * q
        rr_old = rho; k += 1
        assert k < 1000
    assert k == it and 5 < it < 200
    assert np.abs(xs - x).max() <= 1e-10 * np.abs(x).max()
    assert abs(rr - rn / nb) <= 1e-6 * rr
