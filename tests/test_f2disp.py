"""cpfft_b200/f2disp.py (DCT-I solution of the reference's least-squares displacement recovery) against a literal
assembly + dense solve of the same normal equations (tests/py_f2disp.py), and against closed forms."""
import numpy as np
import pytest

from cpfft_b200.f2disp import f2disp, rhs, node_coordinates
from py_f2disp import f2disp_literal


@pytest.mark.parametrize("N,lengths", [(2, (1.0, 1.0, 1.0)), (3, (1.0, 1.0, 1.0)), (4, (2.0, 1.0, 0.5)), (5, (1.0, 3.0, 1.0))])
def test_matches_literal_normal_equations(N, lengths):
    rng = np.random.default_rng(N)
    F = np.zeros((9, N ** 3)); F[[0, 4, 8]] = 1.0
    F += 0.05 * rng.standard_normal(F.shape)
    u_ref, A, b_ref = f2disp_literal(F, N, lengths)
    b = rhs(F, N, lengths).reshape(3, -1).T
    assert np.abs(b - b_ref).max() <= 1e-13 * np.abs(b_ref).max()
    assert np.abs(b.sum(axis=0)).max() <= 1e-12 * np.abs(b).max()              # compatible right-hand side
    u = f2disp(F, N, lengths)
    assert np.abs(u - u_ref).max() <= 1e-10 * np.abs(u_ref).max()
    assert np.abs(u[0]).max() == 0.0                                           # node 1 is pinned (f2disp.f:171)


def test_homogeneous_gradient_is_reproduced_exactly():
    """F constant: the trilinear mesh carries x = F X, u = (F - I) X"""
    N, lengths = 6, (1.0, 2.0, 1.5)
    Fc = np.array([[1.02, 0.01, 0.0], [0.03, 0.97, -0.02], [0.0, 0.015, 1.01]])
    F = np.repeat(Fc.reshape(9, 1), N ** 3, axis=1)
    X = node_coordinates(N, lengths)
    u = f2disp(F, N, lengths)
    assert np.abs(u - X @ (Fc - np.eye(3)).T).max() <= 1e-13


def test_node_order_is_x_fastest():
    X = node_coordinates(2, (1.0, 1.0, 1.0))
    assert X.shape == (27, 3)
    assert np.allclose(X[1], (0.5, 0.0, 0.0)) and np.allclose(X[3], (0.0, 0.5, 0.0)) and np.allclose(X[9], (0.0, 0.0, 0.5))


def test_nodal_displacement_file(tmp_path):
    """wnd#####_text of oudisp (ouresult.f:56-124): nodal-results header, 3e15.6 per node, node order x fastest;
    the deck's `sizes of x_direction ...` card gives the mesh lengths (FFT_finite_3d.f:97-114)"""
    from helpers import deck
    from cpfft_b200.results import write_step, flat_name
    p = deck("test_mm10.in")
    assert p.lengths == (100.0, 100.0, 100.0)
    assert flat_name("displacements", 3) == "wnd00003_text"
    N = p.N
    Fc = np.array([[1.01, 0.0, 0.0], [0.0, 0.997, 0.0], [0.0, 0.0, 0.997]])
    F = np.repeat(Fc.reshape(9, 1), N ** 3, axis=1)
    write_step(str(tmp_path), 3, np.zeros((N ** 3, 9)), np.zeros((N ** 3, 6)), p.name, N, Fn1=F, lengths=p.lengths)
    lines = open(tmp_path / "wnd00003_text").read().splitlines()
    assert lines[1].startswith("#  WARP3D nodal results: displacements")
    assert lines[3] == f"#  Model nodes, elements: {(N + 1) ** 3:8d}{N ** 3:8d}" and lines[5] == "#  Load(time) step:        3"
    rows = lines[7:]
    assert len(rows) == (N + 1) ** 3 and all(len(r) == 45 for r in rows)
    u = np.array([[float(r[15 * k:15 * k + 15]) for k in range(3)] for r in rows])
    assert np.abs(u[0]).max() == 0.0
    assert abs(u[N, 0] - 1.0) <= 1e-6 and abs(u[-1, 0] - 1.0) <= 1e-6            # 1 % of l_x = 100 at the x = l_x face
    assert abs(u[-1, 1] + 0.3) <= 1e-6 and abs(u[-1, 2] + 0.3) <= 1e-6
    assert (tmp_path / "wes00003_text").exists() and (tmp_path / "wee00003_text").exists()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_layered_stretch_is_integrated_along_the_right_axis(axis):
    """F_aa varying with the voxel coordinate a only (a compatible field the trilinear mesh represents exactly):
    u_a(node) = sum of h (F_aa - 1) over the layers below it, constant across the other two directions.  Pins
    the pairing of the solver's voxel order (x slowest, FFT_init.f:311-318) with the mesh (elements z fastest,
    nodes x fastest, oumodel.f:735-745, 932-955)."""
    N, lengths = 5, (1.0, 1.0, 1.0)
    stretch = 1.0 + 0.01 * np.arange(1, N + 1)
    idx = np.indices((N, N, N))[axis].ravel()                     # voxel order: x slowest, z fastest
    F = np.zeros((9, N ** 3)); F[[0, 4, 8]] = 1.0
    F[4 * axis] = stretch[idx]
    u = f2disp(F, N, lengths).T.reshape(3, N + 1, N + 1, N + 1)   # [c, kz, jy, ix]
    want = np.concatenate([[0.0], np.cumsum((stretch - 1.0) / N)])
    shape = [1, 1, 1]; shape[2 - axis] = N + 1                    # node arrays are [kz, jy, ix]
    assert np.abs(u[axis] - want.reshape(shape)).max() <= 1e-13
    others = [c for c in range(3) if c != axis]
    assert np.abs(u[others]).max() <= 1e-13
    u_ref, _, _ = f2disp_literal(F, N, lengths)
    assert np.abs(u.reshape(3, -1).T - u_ref).max() <= 1e-12


def test_model_description_file(tmp_path):
    """`output model "<file>"`: oumodel_flat text form (ouneut.f:19-200): header, 2i9 counts, 3e25.13 per node,
    i4,i4,27i8 per element (Patran hex = type 8, block 1, 8 incidences, zero fill); node / element numbering of
    blkgen (oumodel.f:693-745, 932-955)"""
    from helpers import deck
    from cpfft_b200.results import write_model, element_incidences
    p = deck("test_mm10.in")
    assert p.model_file == "RM_model_flat"
    N = 3
    path = write_model(str(tmp_path / p.model_file), N, (3.0, 6.0, 9.0), p.name)
    assert path.endswith("RM_model_flat.text")
    lines = open(path).read().splitlines()
    assert lines[0] == "#" and lines[1] == "#  Structure: test" and lines[3].startswith("#  Created: ")
    assert lines[5].startswith("#  Convention: Patran element type and node ordering") and lines[6] == "#"
    assert lines[7] == f"{64:9d}{27:9d}"
    nodes, elems = lines[8:8 + 64], lines[8 + 64:]
    assert len(elems) == 27 and all(len(r) == 75 for r in nodes) and all(len(r) == 8 + 27 * 8 for r in elems)
    xyz = np.array([[float(r[25 * k:25 * k + 25]) for k in range(3)] for r in nodes])
    assert np.allclose(xyz[1], (1.0, 0.0, 0.0)) and np.allclose(xyz[4], (0.0, 2.0, 0.0)) and np.allclose(xyz[16], (0.0, 0.0, 3.0))
    assert np.allclose(xyz[-1], (3.0, 6.0, 9.0))
    first = [int(elems[0][4 * k:4 * k + 4]) for k in range(2)] + [int(elems[0][8 + 8 * k:16 + 8 * k]) for k in range(27)]
    assert first == [8, 1, 1, 2, 6, 5, 17, 18, 22, 21] + [0] * 19
    inc = element_incidences(N)
    assert inc[1].tolist() == [17, 18, 22, 21, 33, 34, 38, 37]           # element 2 is the next one in z (z fastest)
    assert inc[-1].max() == 64
    # every brick is right-handed with positive volume on this numbering
    X = xyz[inc - 1]                                                       # (nel, 8, 3)
    vol = np.einsum("ei,ei->e", np.cross(X[:, 1] - X[:, 0], X[:, 3] - X[:, 0]), X[:, 4] - X[:, 0])
    assert np.allclose(vol, 1.0 * 2.0 * 3.0)
