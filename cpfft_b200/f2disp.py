"""Nodal displacements from the voxel deformation gradients: the reference's ``f2disp`` (src/f2disp.f:16-185),
the post-processing step behind the ``wnd#####_text`` result file (``oudisp``, src/ouresult.f:56-124).  Host-side
output code, not part of the GPU hot path.

What the reference computes: on the (N+1)^3-node mesh of trilinear 8-node bricks that ``blkgen`` lays over the box
[0,l_x] x [0,l_y] x [0,l_z] (src/oumodel.f:44-62, 178-330), the current nodal positions
x (one scalar problem per component c) that fit the element gradients in the least-squares sense,

    A x_c = b_c,   A = sum_e int_e B^T B        (8-point Gauss, ``form_BTB`` f2disp.f:202-445)
                   b_c = sum_e V_e B_e(0)^T F_e^T(:, c)   (1-point Gauss, f2disp.f:104-131)

with node 1 removed (x = 0 there), solved by PARDISO; the displacement is u = x - X (f2disp.f:171-176).  Nodes are
numbered x fastest (``vblkn``, oumodel.f:932-955), elements z fastest (``vblke`` loops i, j, k with k innermost,
oumodel.f:735-745), i.e. element ``ii`` is voxel ``ii`` of the solver (FFT_init.f:311-318) and takes its deformation
gradient (f2disp.f:64-66, 119-127).

Not reproduced: ``form_BTB`` stores the upper triangle and adds A_local(kk, ll), kk <= ll, at (row of node kk, column of
node ll) (f2disp.f:428-439); for the local pairs (3,4) and (7,8) the global numbers are in the other order, the column
search falls through and the value lands on the first entry of the next row.  Those two entries couple nodes that
share an x-edge; on cubic cells (l_x = l_y = l_z, every shipped deck) the edge-neighbour entry of the trilinear
stiffness is exactly zero, so the slip is harmless there and the reference solves the equations written above.  For
unequal cell lengths it does not, and what it solves instead is an accident of its storage scheme.

How it is solved here: on the uniform mesh A = Kx (x) My (x) Mz + Mx (x) Ky (x) Mz + Mx (x) My (x) Kz with the 1-D
linear-element stiffness K = tridiag(-1, 2, -1)/h (corner entries 1/h) and mass M = h tridiag(1, 4, 1)/6 (corners
2h/6).  Both satisfy K v_k = lambda_k D v_k, M v_k = mu_k D v_k for the DCT-I vectors v_k(j) = cos(pi k j / N),
D = diag(1/2, 1, ..., 1, 1/2), lambda_k = (2 - 2 cos(pi k/N))/h, mu_k = h (4 + 2 cos(pi k/N))/6, so

    x = DCT1( DCT1(b / D3) / (8 N^3 sigma) ),   sigma_abc = lambda_a mu_b mu_c + mu_a lambda_b mu_c + mu_a mu_b lambda_c

exactly (sigma_000 = 0 is the constant null vector; b sums to zero, and the constant is fixed by x(node 1) = 0).
O(N^3 log N) instead of a sparse factorisation; tests/test_f2disp.py checks it against a literal assembly + dense
solve of the same normal equations.  No output of the reference exists to pin the file against."""
from __future__ import annotations

import numpy as np


def node_coordinates(N: int, lengths=(1.0, 1.0, 1.0)) -> np.ndarray:
    """(nnode, 3) reference coordinates in the mesh's node order (x fastest, oumodel.f:923-955)."""
    ax = [np.linspace(0.0, float(l), N + 1) for l in lengths]
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def rhs(Fn1: np.ndarray, N: int, lengths=(1.0, 1.0, 1.0)) -> np.ndarray:
    """b (3, N+1, N+1, N+1) indexed [component, kz, jy, ix]: sum over the 8 elements of a node of
    V_e dN/dX_J(centre) F_cJ (f2disp.f:104-142).  Fn1: (9, N^3), row-major F per voxel, voxels in solver order."""
    h = [float(l) / N for l in lengths]
    V = h[0] * h[1] * h[2]
    # voxel / element index ii = x N^2 + y N + z  ->  [c, J, ek, ej, ei] like the node arrays
    F = np.asarray(Fn1, dtype=np.float64).reshape(3, 3, N, N, N).transpose(0, 1, 4, 3, 2)
    b = np.zeros((3, N + 1, N + 1, N + 1))
    for sz in (0, 1):
        for sy in (0, 1):
            for sx in (0, 1):
                g = [(2 * sx - 1) / (4.0 * h[0]), (2 * sy - 1) / (4.0 * h[1]), (2 * sz - 1) / (4.0 * h[2])]
                contrib = V * (g[0] * F[:, 0] + g[1] * F[:, 1] + g[2] * F[:, 2])
                b[:, sz:sz + N, sy:sy + N, sx:sx + N] += contrib
    return b


def _spectrum(N: int, h: float):
    th = np.pi * np.arange(N + 1) / N
    return (2.0 - 2.0 * np.cos(th)) / h, h * (4.0 + 2.0 * np.cos(th)) / 6.0


def solve_positions(b: np.ndarray, N: int, lengths=(1.0, 1.0, 1.0)) -> np.ndarray:
    """x (3, N+1, N+1, N+1) with A x_c = b_c and x(node 1) = 0, by the DCT-I diagonalisation above."""
    from scipy.fft import dctn
    h = [float(l) / N for l in lengths]
    (lx, mx), (ly, my), (lz, mz) = _spectrum(N, h[0]), _spectrum(N, h[1]), _spectrum(N, h[2])
    sigma = (mz[:, None, None] * my[None, :, None] * lx[None, None, :] +
             mz[:, None, None] * ly[None, :, None] * mx[None, None, :] +
             lz[:, None, None] * my[None, :, None] * mx[None, None, :])
    sigma[0, 0, 0] = 1.0
    d = np.ones(N + 1); d[0] = d[N] = 0.5
    D3 = d[:, None, None] * d[None, :, None] * d[None, None, :]
    x = np.empty_like(b)
    for c in range(3):
        B = dctn(b[c] / D3, type=1) / (8.0 * float(N) ** 3 * sigma)
        B[0, 0, 0] = 0.0
        x[c] = dctn(B, type=1)
        x[c] -= x[c, 0, 0, 0]                       # node 1 is the removed (pinned) equation, f2disp.f:134-136
    return x


def f2disp(Fn1: np.ndarray, N: int, lengths=(1.0, 1.0, 1.0)) -> np.ndarray:
    """(nnode, 3) nodal displacements u = x - X in node order; u(node 1) = 0 (f2disp.f:171-176)."""
    x = solve_positions(rhs(Fn1, N, lengths), N, lengths)
    return x.reshape(3, -1).T - node_coordinates(N, lengths)
