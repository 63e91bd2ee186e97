"""Result files in the reference's flat-text format (src/ouresult.f): per output step one ``wnd#####_text``
(nodal displacements recovered from F, ``oudisp`` ouresult.f:56-124, three ``e15.6`` values per node), one
``wes#####_text`` (unrotated Cauchy stresses) and one ``wee#####_text`` (strains) file with the
``ouddpa_flat_header`` header (ouresult.f:335-398) and one ``30e15.6`` record per element
(ouresult.f:296-298; 11 + 15 values per stress record, 7 + 15 per strain record, of which the
first 6 carry data -- the reference zero-fills the rest, ouresult.f:170-172, 208-215)."""
from __future__ import annotations

import math
import os
import time

import numpy as np


def fortran_e(x: float, width: int = 15, digits: int = 6) -> str:
    """Fortran ``Ew.d`` edit descriptor: 0.123457E+03 right-justified in ``width``."""
    if abs(x) < 1.0e-30:            # "zero small values to prevent 3-digit exponents" (ouresult.f:289)
        x = 0.0
    if x == 0.0:
        body = "0." + "0" * digits + "E+00"
    else:
        e = int(math.floor(math.log10(abs(x)))) + 1
        m = abs(x) / 10.0 ** e
        r = int(round(m * 10 ** digits))
        if r >= 10 ** digits:
            r //= 10
            e += 1
        body = ("-" if x < 0 else "") + "0." + f"{r:0{digits}d}" + "E" + ("-" if e < 0 else "+") + f"{abs(e):02d}"
    return body.rjust(width)


def flat_name(kind: str, step: int) -> str:
    """``wes`` / ``wee`` / ``wnd`` + step (i5.5) + ``_text`` (ouresult.f:440, 704-711)."""
    return {"stresses": "wes", "strains": "wee", "displacements": "wnd"}[kind] + f"{step:05d}_text"


def write_flat(path: str, kind: str, step: int, values: np.ndarray, structure: str = "", nnode: int = 0):
    """values: (nelem, 6) in the reference's Voigt order xx, yy, zz, xy, yz, xz."""
    nvals = {"stresses": 11 + 15, "strains": 7 + 15}[kind]
    nelem = values.shape[0]
    with open(path, "w") as f:
        f.write("#\n")
        f.write(f"#  WARP3D element results: {kind:<15s}\n")
        f.write(f"#  Structure name: {structure[:8]:<8s}\n")
        f.write(f"#  Model nodes, elements: {nnode:8d}{nelem:8d}\n")
        f.write(f"#  {time.strftime('%a %b %e %H:%M:%S %Y'):<24s}\n")
        f.write(f"#  Load(time) step: {step:8d}\n")
        f.write("#\n")
        zero = fortran_e(0.0)
        pad = zero * (nvals - 6)
        for row in values:
            f.write("".join(fortran_e(float(v)) for v in row[:6]) + pad + "\n")


def write_nodal(path: str, step: int, u: np.ndarray, structure: str = "", nelem: int = 0):
    """u: (nnode, 3) nodal displacements in node order (ouresult.f:104-113, format 930 = 3e15.6)."""
    with open(path, "w") as f:
        f.write("#\n")
        f.write(f"#  WARP3D nodal results: {'displacements':<15s}\n")
        f.write(f"#  Structure name: {structure[:8]:<8s}\n")
        f.write(f"#  Model nodes, elements: {u.shape[0]:8d}{nelem:8d}\n")
        f.write(f"#  {time.strftime('%a %b %e %H:%M:%S %Y'):<24s}\n")
        f.write(f"#  Load(time) step: {step:8d}\n")
        f.write("#\n")
        for row in u:
            f.write("".join(fortran_e(float(v)) for v in row[:3]) + "\n")


def write_step(outdir: str, step: int, urcs_n1: np.ndarray, eps_n1: np.ndarray, structure: str = "", N: int = 0,
               Fn1: np.ndarray | None = None, lengths=(1.0, 1.0, 1.0)):
    """urcs_n1 (nelem, >= 6), eps_n1 (nelem, 6): the AoS block order of the reference.  Fn1 (9, nelem): also the
    nodal displacement file, recovered from the deformation gradients as the reference's f2disp does."""
    os.makedirs(outdir, exist_ok=True)
    nnode = (N + 1) ** 3 if N else 0
    if Fn1 is not None and N:
        from .f2disp import f2disp
        write_nodal(os.path.join(outdir, flat_name("displacements", step)), step, f2disp(Fn1, N, lengths), structure, N ** 3)
    write_flat(os.path.join(outdir, flat_name("stresses", step)), "stresses", step, urcs_n1[:, :6], structure, nnode)
    write_flat(os.path.join(outdir, flat_name("strains", step)), "strains", step, eps_n1[:, :6], structure, nnode)


def element_incidences(N: int) -> np.ndarray:
    """(N^3, 8) 1-based node numbers of the 8-node bricks ``blkgen`` generates (oumodel.f:693-745): elements in the
    solver's voxel order (x slowest, z fastest), nodes numbered x fastest, local order counter-clockwise bottom
    face then top face."""
    n1 = N + 1
    i, j, k = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    base = (1 + i + n1 * j + n1 * n1 * k).ravel()                       # node (i, j, k)
    off = np.array([0, 1, 1 + n1, n1, n1 * n1, 1 + n1 * n1, 1 + n1 + n1 * n1, n1 + n1 * n1])
    return base[:, None] + off[None, :]


def write_model(path: str, N: int, lengths=(1.0, 1.0, 1.0), structure: str = "") -> str:
    """The model description file of ``output model "<name>"`` (``ModelOut`` oumodel.f:11-120 -> ``oumodel_flat``
    ouneut.f:19-200, text form, Patran convention): header, ``2i9`` node / element counts, ``3e25.13`` coordinates
    per node, ``i4,i4,27i8`` per element (Patran type 8 = hex, block 1, 8 incidences, zero fill).  Returns the
    file name (``.text`` appended when the name has no extension, ouneut.f:67-74)."""
    from .f2disp import node_coordinates
    if not (path.endswith(".text") or path.endswith(".str")):
        path += ".text"
    X = node_coordinates(N, lengths)
    inc = element_incidences(N)
    with open(path, "w") as f:
        f.write("#\n#  " + ("Structure: " + structure.strip()).rstrip() + "\n")
        f.write("#\n#  Created: " + f"{time.strftime('%b %e %Y'):<12s}" + "  " + time.strftime("%H:%M:%S") + "\n#\n")
        f.write("#  Convention: " + f"{'Patran element type and node ordering':<40s}" + "\n#\n")
        f.write(f"{X.shape[0]:9d}{inc.shape[0]:9d}\n")
        for row in X:
            f.write("".join(fortran_e(float(v), 25, 13) for v in row) + "\n")
        tail = "".join(f"{0:8d}" for _ in range(19))
        for row in inc:
            f.write(f"{8:4d}{1:4d}" + "".join(f"{int(v):8d}" for v in row) + tail + "\n")
    return path
