"""Plain-data description of a CPFFT analysis: grid, materials, crystal library, per-voxel
material/orientation maps and the load table.  This mirrors what the reference keeps in
module ``fft`` (src/mod_fft.f:13-79), ``crystal_data`` (src/mod_crystals.f:137-227) and the
``matprp`` slots filled by src/inmat.f, i.e. everything the hot path reads after input.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

HARDENING = {"voce": 1, "voche": 1, "mts": 2}     # incrystal.f:305-331 (supported subset)
SLIP_TYPES = {"fcc": 1, "bcc": 2, "single": 3, "roters": 6, "bcc12": 7, "bcc48": 8}   # mod_crystals.f:164-172 (cubic families; hcp6 / hcp18 not supported)
ELASTIC_TYPES = {"isotropic": 1, "cubic": 2}  # mod_crystals.f:173-176
COMPONENTS = ["xx", "xy", "xz", "yx", "yy", "yz", "zx", "zy", "zz"]  # inlodcase.f:29-139


class CrystalPOD(C.Structure):
    """C layout shared by ``cpfft_crystal`` (include/cpfft_b200.h) and the test oracle."""
    _fields_ = [
        ("slip_type", C.c_int32), ("elastic_type", C.c_int32), ("h_type", C.c_int32),
        ("alter_mode", C.c_int32), ("miter", C.c_int32), ("pad_", C.c_int32),
        ("e", C.c_double), ("nu", C.c_double), ("mu", C.c_double), ("harden_n", C.c_double),
        ("theta_0", C.c_double), ("tau_y", C.c_double), ("tau_v", C.c_double),
        ("voche_m", C.c_double), ("iD_v", C.c_double), ("eps_dot_0_y", C.c_double),
        ("k_0", C.c_double), ("burgers", C.c_double),
        ("atol", C.c_double), ("atol1", C.c_double), ("rtol", C.c_double), ("rtol1", C.c_double),
        # MTS hardening (h_type 2)
        ("tau_a", C.c_double), ("tau_hat_y", C.c_double), ("g_0_y", C.c_double), ("tau_hat_v", C.c_double),
        ("g_0_v", C.c_double), ("p_y", C.c_double), ("q_y", C.c_double), ("p_v", C.c_double), ("q_v", C.c_double),
        ("boltzman", C.c_double), ("eps_dot_0_v", C.c_double), ("mu_0", C.c_double), ("D_0", C.c_double),
        ("T_0", C.c_double),
    ]


class MaterialPOD(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("crystal", C.c_int32),
        ("e", C.c_float), ("nu", C.c_float), ("beta", C.c_float), ("tan_e", C.c_float),
        ("yld_pt", C.c_float), ("n_crystals", C.c_int32),
    ]


@dataclass
class Crystal:
    """Defaults are ``initialize_new_crystal`` (mod_crystals.f:231-410); the single-precision
    literals there are promoted by ifort ``-fpconstant`` (src/makefile:20), so they are doubles."""
    slip_type: int = 1
    elastic_type: int = 1
    h_type: int = 1
    alter_mode: int = 0
    miter: int = 30
    e: float = 69000.0
    nu: float = 0.33
    mu: float = 69000.0 / 2.0 / 1.33
    harden_n: float = 20.0
    theta_0: float = 100.0
    tau_y: float = 0.0
    tau_v: float = 0.0
    voche_m: float = 1.0
    iD_v: float = 0.0
    eps_dot_0_y: float = 1.0e10
    k_0: float = 0.0
    burgers: float = 2.87e-7
    atol: float = 1.0e-5
    atol1: float = 1.0e-5
    rtol: float = 5.0e-5
    rtol1: float = 1.0e-5
    # MTS hardening, h_type 2 (defaults mod_crystals.f:256-275)
    tau_a: float = 0.0
    tau_hat_y: float = -1.0
    g_0_y: float = -1.0
    tau_hat_v: float = -1.0
    g_0_v: float = -1.0
    p_y: float = 0.5
    q_y: float = 2.0
    p_v: float = 0.5
    q_v: float = 2.0
    boltzman: float = 1.3806e-20
    eps_dot_0_v: float = 1.0e10
    mu_0: float = 69000.0 / 2.0 / 1.33
    D_0: float = 0.0
    T_0: float = 294.0

    def pod(self) -> CrystalPOD:
        p = CrystalPOD()
        for name, _ in CrystalPOD._fields_:
            if name != "pad_":
                setattr(p, name, getattr(self, name))
        return p


@dataclass
class Material:
    """type 1 = bilinear (mm01), 10 = crystal plasticity (mm10).  The bilinear properties
    live in REAL*4 ``matprp`` slots (mod_fft.f:20, inmat.f:100-127): kept as float32."""
    name: str = "mat"
    type: int = 1
    crystal: int = 0          # cp: crystal_type (1-based), crystal_input single
    n_crystals: int = 1       # cp: crystals per material point, Taylor average (inmat.f:201-204)
    crystal_input: int = 1    # 1 single (crystal_type), 2 file (per voxel and crystal)
    e: float = 0.0
    nu: float = 0.0
    beta: float = 0.0
    tan_e: float = 0.0
    yld_pt: float = 0.0
    orientation_file: str = ""
    angles: tuple = (0.0, 0.0, 0.0)
    orientation_input: int = 1  # 1 single, 2 file
    angle_scale: float = 1.0    # deck reader: factor to degrees (`angle_type radians`)

    def pod(self) -> MaterialPOD:
        p = MaterialPOD()
        p.type, p.crystal, p.n_crystals = self.type, self.crystal, self.n_crystals
        p.e, p.nu, p.beta, p.tan_e, p.yld_pt = (np.float32(self.e), np.float32(self.nu),
                                                 np.float32(self.beta), np.float32(self.tan_e),
                                                 np.float32(self.yld_pt))
        return p


@dataclass
class Problem:
    N: int
    materials: List[Material]
    crystals: List[Crystal]
    matlist: np.ndarray                      # (N3,) int32, 1-based material number per voxel
    angles: np.ndarray                       # (N3,3) Kocks degrees per voxel; (N3,ncmax,3) with n_crystals > 1
    FP_max: np.ndarray = field(default_factory=lambda: np.zeros(9))
    isNBC: np.ndarray = field(default_factory=lambda: np.zeros(9, dtype=np.int32))
    mults: np.ndarray = field(default_factory=lambda: np.zeros(0))
    tolNR: float = 1.0e-5
    tolPCG: float = 1.0e-10
    maxIter: int = 10
    tstep: float = 1.0
    name: str = ""                           # `project` card (stname)
    out_steps: tuple = ()                    # `output results steps <list>` (oudriv.f:82-166)
    lengths: tuple = (1.0, 1.0, 1.0)         # `sizes of x_direction .. ` l_x, l_y, l_z (output mesh only)
    model_file: str = ""                     # `output model "<file>"`: mesh description file (oumodel.f)
    crystal_ids: np.ndarray = None           # (N3,ncmax) 1-based crystal numbers (crystal_input file) or None

    @property
    def N3(self) -> int:
        return self.N ** 3

    @property
    def ncmax(self) -> int:
        """crystals per material point the per-voxel tables are sized for (1 = the classic case)"""
        return 1 if np.ndim(self.angles) == 2 else int(np.shape(self.angles)[1])

    @property
    def taylor(self) -> bool:
        return any(m.type == 10 and m.n_crystals > 1 for m in self.materials) or self.crystal_ids is not None

    def taylor_tables(self):
        """(ncmax, angles (N3,ncmax,3) float64, crystal_ids (N3,ncmax) int32 or None), C-contiguous"""
        nc = self.ncmax
        nv = len(self.matlist)              # N3, or the voxels of one rank's slab
        ang = np.ascontiguousarray(np.asarray(self.angles, dtype=np.float64).reshape(nv, nc, 3))
        ids = None if self.crystal_ids is None else \
            np.ascontiguousarray(np.asarray(self.crystal_ids, dtype=np.int32).reshape(nv, nc))
        return nc, ang, ids

    @property
    def nstep(self) -> int:
        return len(self.mults)

    def BC_all(self) -> np.ndarray:
        """Cumulative load table (inlod.f:57-63), shape (nstep, 9)."""
        bc = np.cumsum(np.outer(self.mults, self.FP_max), axis=0)
        for d in (0, 4, 8):
            if not self.isNBC[d]:
                bc[:, d] += 1.0
        return np.ascontiguousarray(bc)

    def material_pods(self):
        arr = (MaterialPOD * len(self.materials))()
        for i, m in enumerate(self.materials):
            arr[i] = m.pod()
        return arr

    def crystal_pods(self):
        n = max(1, len(self.crystals))
        arr = (CrystalPOD * n)()
        for i, c in enumerate(self.crystals):
            arr[i] = c.pod()
        return arr
