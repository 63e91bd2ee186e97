"""Physical properties of the crystal-plasticity update that hold whatever the transcription: frame
indifference under a superposed rigid rotation and invariance under the cubic symmetry group of the
lattice.  They pin conventions (rotation operators, Kocks angles, slip tables, Voigt orderings) from
outside the reference's text; run on the CPU oracle AND on the host build of the kernel source."""
import numpy as np
import pytest

from helpers import relerr, mm10_layout
from py_mm10 import kocks

TOL = 5.0e-8          # small-strain noise floor of the closed-form polar decomposition (tests/test_host_kernels.py)


@pytest.fixture(scope="module", params=["oracle", "kernel_source"])
def Impl(request, oracle_built):
    if request.param == "oracle":
        from oracle import Oracle
        return Oracle
    from host_kernels import HostKernels, build
    build()
    return HostKernels


def _fields(m):
    """(P (9,n), K4 (81,n), unrotated Cauchy stress (n,6), history (n,H)) whichever implementation"""
    u, h = np.asarray(m.urcs_n1), np.asarray(m.hist_n1)
    n = m.N3
    u = u if u.shape[0] == n else u.T
    h = h if h.shape[0] == n else h.T
    return np.array(m.Pn1), np.array(m.K4), np.array(u[:, :6]), np.array(h)


def _rot(axis, angle):
    from scipy.spatial.transform import Rotation
    return Rotation.from_rotvec(angle * np.asarray(axis, float) / np.linalg.norm(axis)).as_matrix()


def _path(p, nstep=3, amp=0.003, seed=5):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((9, p.N3))
    bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45; bar[1] = 0.3
    I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
    return [I + amp * k * (bar + 0.25 * G) for k in range(1, nstep + 1)]


def _lmul(Q, F):
    """Q F voxel by voxel, F as (9, n) row-major tensors"""
    return np.einsum("im,mjn->ijn", Q, F.reshape(3, 3, -1)).reshape(9, -1)


def _run(Impl, p, path, Q=None):
    m = Impl(p)
    m.drive_eps_sig(1, 0)
    if Q is not None:
        m.Fn[:] = _lmul(Q, np.array(m.Fn))
    out = []
    for step, F in enumerate(path, start=1):
        Fq = F if Q is None else _lmul(Q, F)
        for it in (0, 1):
            m.Fn1[:] = Fq
            assert m.drive_eps_sig(step, it) == 0
        out.append(_fields(m))
        m.Fn[:] = Fq
        m.update()
    return out


@pytest.mark.parametrize("kind", ["mm01_plastic", "mm10_elastic", "mm10_plastic"])
def test_frame_indifference(Impl, kind):
    """superposed rigid rotation Q of the whole path (F_n and F_n+1): the unrotated stress and the
    history are unchanged, P' = Q P, K4'_ijkl = Q_im Q_kn K4_mjnl  (drive_eps_sig.f:203-295, cep2A.f).
    Exact for the bilinear model and for the crystal as long as it does not slip.  With slip the reference's
    crystal update is only approximately frame indifferent: the plastic spin of the residual is formed
    with qc = RW(R Rp_n^T) qs0 (mm10_a.f:867-876, mm10_b.f:1304-1340), i.e. it carries the rotation R
    of the step into an equation written in the unrotated frame; the effect is of order
    |sigma| / |C| * |Wp| and is bounded here, not asserted away -- the oracle and the kernels reproduce
    the reference, they do not repair it."""
    from cpfft_b200.polycrystal import polycrystal
    if kind == "mm01_plastic":
        from helpers import deck
        p = deck("test_mm01.in")
        path = _path(p, amp=0.01)
    else:
        p = polycrystal(3, ngrains=6)
        path = _path(p, amp=0.0001 if kind == "mm10_elastic" else 0.003)
    Q = _rot([1.0, -2.0, 0.5], 0.9)
    ref, rot = _run(Impl, p, path), _run(Impl, p, path, Q)
    # mm10_elastic: strains of 1e-4 .. 3e-4 (below first slip), where the closed-form polar decomposition and
    # the dR/dF terms of the tangent built on it are noisier still (measured 8e-8 on the stress, 4e-7 on K4)
    tol = {"mm10_plastic": 1.0e-3, "mm10_elastic": 2.0e-6}.get(kind, TOL)
    for (P, K, u, h), (Pq, Kq, uq, hq) in zip(ref, rot):
        assert relerr(uq, u) <= tol
        assert relerr(Pq, _lmul(Q, P)) <= tol
        K4 = K.reshape(3, 3, 3, 3, -1)
        assert relerr(Kq.reshape(3, 3, 3, 3, -1), np.einsum("im,kn,mjnlv->ijklv", Q, Q, K4)) <= \
            (5 * tol if kind == "mm10_plastic" else tol)       # measured 2e-4 (stress) and 1.5e-3 (tangent) at 0.9 rad
    if kind == "mm01_plastic":
        assert np.abs(ref[-1][3][:, 0]).max() > 0          # accumulated plastic strain: the path yields
        assert relerr(rot[-1][3], ref[-1][3]) <= TOL
    else:
        a, b = mm10_layout(12)["slipsum"]
        slip = np.abs(ref[-1][3][:, a:b]).max()
        assert (slip < 1e-8) if kind == "mm10_elastic" else (slip > 1e-5), slip


def _angles_of(G):
    """Kocks angles (degrees) of the orientation matrix G (numerical inverse of mm10_rotation_matrix)"""
    from scipy.optimize import least_squares
    best = None
    rng = np.random.default_rng(0)
    for _ in range(40):
        r = least_squares(lambda a: (kocks(a) - G).ravel(), rng.uniform(-180.0, 180.0, 3), xtol=1e-15, ftol=1e-15, gtol=1e-15)
        if best is None or r.cost < best.cost:
            best = r
        if best.cost < 1e-26:
            break
    assert best.cost < 1e-24, best.cost
    return best.x


@pytest.mark.parametrize("slip_type", [1, 8])
def test_cubic_symmetry_of_the_lattice(Impl, slip_type):
    """g and S^T g (S a rotation of the cube group) describe the same fcc / bcc lattice: stresses, tangents
    and accumulated slip must agree, the slip increments are permuted (mod_crystals.f slip tables,
    mm10_rotation_matrix mm10_a.f:1287-1345, drive_eps_sig.f:975-986)"""
    from cpfft_b200.polycrystal import polycrystal
    def problem():
        # local Newton solves converged to round-off: with the default tolerances (rtol 5e-5) two orderings of
        # the same slip systems stop at iterates that differ by 2e-7
        q = polycrystal(2, ngrains=8, slip_type=slip_type)
        for k, v in dict(atol=1e-7, atol1=1e-7, rtol=1e-9, rtol1=1e-9, miter=60).items():
            setattr(q.crystals[0], k, v)
        return q
    p = problem()
    path = _path(p, nstep=3, amp=0.002)
    ref = _run(Impl, p, path)
    nslip = {1: 12, 8: 48}[slip_type]
    for S in (_rot([0, 0, 1], np.pi / 2), _rot([1, 1, 1], 2 * np.pi / 3), _rot([1, 1, 0], np.pi)):
        q = problem()
        q.angles = np.array([_angles_of(S.T @ kocks(a)) for a in p.angles])
        sym = _run(Impl, q, path)
        for (P, K, u, h), (Ps, Ks, us, hs) in zip(ref, sym):
            assert relerr(us, u) <= TOL and relerr(Ps, P) <= TOL
            # the tangent comes from the LAGGED Jacobian of the last local iteration (mm10_a.f:1137-1142), i.e. it
            # depends on the iterate before convergence, which the ordering of the systems moves (measured 7e-5)
            assert relerr(Ks, K) <= 1.0e-3
            # accumulated slip and the slip increments of the step: same multiset of magnitudes per voxel
            L = mm10_layout(nslip)
            for name in ("slipsum", "slipinc"):
                c0, c1 = L[name]
                a, b = np.sort(np.abs(h[:, c0:c1]), axis=1), np.sort(np.abs(hs[:, c0:c1]), axis=1)
                assert a.max() > 0 and np.abs(a - b).max() <= 1e-6 * a.max(), name
