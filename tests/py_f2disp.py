"""Literal restatement of the reference's f2disp normal equations (src/f2disp.f:16-185, form_BTB :202-445) for small
N: element loop, trilinear 8-node brick shape functions, 8-point Gauss rule for A = sum int B^T B, 1-point rule for
b = sum V B(0)^T F^T, node 1 removed, dense solve.  TEST INFRASTRUCTURE: the checker of cpfft_b200/f2disp.py."""
import numpy as np

_SIGNS = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], dtype=float)   # local nodes


def _dshape(xi):
    """dN_k/dxi_J (8, 3) of the trilinear brick at natural coordinates xi"""
    s = _SIGNS
    f = 1.0 + s * xi
    return 0.125 * np.stack([s[:, 0] * f[:, 1] * f[:, 2], f[:, 0] * s[:, 1] * f[:, 2], f[:, 0] * f[:, 1] * s[:, 2]], axis=1)


def f2disp_literal(Fn1, N, lengths=(1.0, 1.0, 1.0)):
    n1 = N + 1
    nn = n1 ** 3
    h = np.array(lengths, float) / N
    node = lambda i, j, k: i + n1 * j + n1 * n1 * k                       # x fastest (oumodel.f:923-955)
    X = np.zeros((nn, 3))
    for k in range(n1):
        for j in range(n1):
            for i in range(n1):
                X[node(i, j, k)] = (i * h[0], j * h[1], k * h[2])
    A = np.zeros((nn, nn))
    b = np.zeros((nn, 3))
    gp = 1.0 / np.sqrt(3.0)
    e = 0
    for ei in range(N):                                                 # elements: z fastest (vblke, oumodel.f:735-745)
        for ej in range(N):
            for ek in range(N):
                ids = [node(ei + (int(s[0]) + 1) // 2, ej + (int(s[1]) + 1) // 2, ek + (int(s[2]) + 1) // 2) for s in _SIGNS]
                xe = X[ids]                                             # (8, 3)
                for q in _SIGNS * gp:                                   # 8-point rule, weights 1
                    dN = _dshape(q)
                    J = dN.T @ xe                                       # dX/dxi
                    B = dN @ np.linalg.inv(J)                           # dN/dX (8, 3)
                    A[np.ix_(ids, ids)] += np.linalg.det(J) * (B @ B.T)
                dN = _dshape(np.zeros(3))
                J = dN.T @ xe
                B = dN @ np.linalg.inv(J)
                F = np.asarray(Fn1)[:, e].reshape(3, 3)                 # element e <- voxel e (f2disp.f:64-66, 119-127)
                b[ids] += 8.0 * np.linalg.det(J) * (B @ F.T)            # weight 8 of the 1-point rule
                e += 1
    x = np.zeros((nn, 3))
    x[1:] = np.linalg.solve(A[1:, 1:], b[1:])                           # node 1 removed
    return x - X, A, b
