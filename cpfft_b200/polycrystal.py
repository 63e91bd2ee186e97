"""Deterministic synthetic Voronoi polycrystal (BASELINE.json configs[3,4]).

seed 20240607; G seeds uniform in [0,1)^3; periodic Voronoi tessellation evaluated at the voxel
centres (i+1/2)/N; one orientation per grain, uniform on SO(3), given as Kocks angles in
degrees (psi, phi uniform in [0,360), cos(theta) uniform in [-1,1]); voxel
e = x*N*N + y*N + z.  One crystal-plasticity material: fcc, isotropic e 200000 nu 0.3, Voce
harden_n 20 theta_0 100 voce_m 1 tau_v 100 tau_y 100, alter_mode off, time step 1.
Loading: F_xx +1 %, F_yy = F_zz -0.3 % in 10 equal steps (pure-strain variant, the benchmark), or
with ``stress_bc`` uniaxial tension F_xx +1 %, P_yy = P_zz = 0 (tangent_homo + NBC_update path).
"""
from __future__ import annotations

import numpy as np

from .problem import Crystal, Material, Problem

SEED = 20240607


def grain_map(N: int, ngrains: int = 1000, seed: int = SEED, x_range=None) -> np.ndarray:
    """grain index per voxel (optionally only for the x-planes in ``x_range``)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    seeds = rng.random((ngrains, 3))
    tree = cKDTree(seeds, boxsize=1.0)
    x0, x1 = (0, N) if x_range is None else x_range
    c = (np.arange(N) + 0.5) / N
    out = np.empty((x1 - x0) * N * N, dtype=np.int32)
    yy, zz = np.meshgrid(c, c, indexing="ij")
    plane = np.stack([np.zeros(N * N), yy.ravel(), zz.ravel()], axis=1)
    for ix in range(x0, x1):
        plane[:, 0] = c[ix]
        _, idx = tree.query(plane, workers=-1)
        out[(ix - x0) * N * N:(ix - x0 + 1) * N * N] = idx
    return out


def grain_angles(ngrains: int = 1000, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed + 1)
    u = rng.random((ngrains, 3))
    psi = 360.0 * u[:, 0]
    theta = np.degrees(np.arccos(1.0 - 2.0 * u[:, 1]))
    phi = 360.0 * u[:, 2]
    return np.stack([psi, theta, phi], axis=1)


def polycrystal(N: int, ngrains: int = 1000, seed: int = SEED, nstep: int = 10, slip_type: int = 1,
                x_range=None, stress_bc: bool = False) -> Problem:
    gm = grain_map(N, ngrains, seed, x_range)
    ang = grain_angles(ngrains, seed)[gm]
    cry = Crystal(slip_type=slip_type, elastic_type=1, h_type=1, alter_mode=0, e=200000.0, nu=0.3,
                  mu=200000.0 / 2.6, harden_n=20.0, theta_0=100.0, voche_m=1.0, tau_v=100.0, tau_y=100.0)
    mat = Material(name="poly", type=10, crystal=1)
    FP = np.zeros(9)
    FP[0], FP[4], FP[8] = 0.01, -0.003, -0.003
    nbc = np.zeros(9, dtype=np.int32)
    if stress_bc:                 # uniaxial tension: F_xx driven, P_yy = P_zz = 0
        FP[4] = FP[8] = 0.0
        nbc[4] = nbc[8] = 1
    p = Problem(N=N, materials=[mat], crystals=[cry], matlist=np.ones(len(gm), dtype=np.int32), angles=ang,
                FP_max=FP, isNBC=nbc, mults=np.full(nstep, 1.0 / nstep),
                tolNR=1.0e-5, tolPCG=1.0e-10, maxIter=20, tstep=1.0)
    return p


def taylor_polycrystal(N: int, ncrystals: int = 4, ngrains: int = 1000, seed: int = SEED, nstep: int = 10,
                       mixed: bool = False) -> Problem:
    """Polycrystalline material points: every voxel carries ``ncrystals`` crystals whose stresses
    and tangents are Taylor-averaged (mm10 with n_crystals > 1, mm10_a.f:112-197).  Crystal 0 of
    a voxel has the orientation of its Voronoi grain, the others are further draws from the same
    orientation table.  ``mixed``: odd crystals are bcc48 (crystal_input file), even ones fcc."""
    p = polycrystal(N, ngrains, seed, nstep)
    gm = grain_map(N, ngrains, seed)
    table = grain_angles(ngrains + ncrystals, seed)
    ang = np.stack([table[(gm + c * 7) % len(table)] for c in range(ncrystals)], axis=1)   # (N3, nc, 3)
    p.angles = np.ascontiguousarray(ang)
    p.materials[0].n_crystals = ncrystals
    if mixed:
        import copy
        bcc = copy.copy(p.crystals[0]); bcc.slip_type = 8; bcc.tau_y = 120.0
        p.crystals = [p.crystals[0], bcc]
        p.materials[0].crystal_input = 2
        ids = np.ones((len(gm), ncrystals), dtype=np.int32)
        ids[:, 1::2] = 2
        p.crystal_ids = ids
    return p


def workload_variant(prob, variant, ngrains):
    """the benchmark polycrystal (possibly one rank's slab of it) with another material law: ``mts`` or
    ``taylorN`` = N crystals per material point.  For kernel measurements and the multi-GPU self-check,
    not the headline workload."""
    import dataclasses
    if variant == "bcc48":     # the 48-system bcc family of test_mm10.in (mod_crystals.f:812-1205) on every grain
        prob.crystals = [dataclasses.replace(prob.crystals[0], slip_type=8)]
        return prob
    if variant == "mts":       # `hardening mts` with the thresholds of tests/golden/decks/mts_mm10.in
        c = dataclasses.replace(prob.crystals[0], h_type=2, theta_0=1500.0, tau_a=20.0, tau_hat_y=180.0, g_0_y=0.4,
                                tau_hat_v=300.0, g_0_v=1.2, burgers=2.5e-7, mu_0=80000.0, D_0=3000.0, T_0=200.0)
        prob.crystals = [c]
        return prob
    nc = int(variant[-1])      # taylorN: N crystals per material point, further draws from the orientation table
    table = grain_angles(ngrains + nc)
    base = np.asarray(prob.angles)
    key = np.abs(base[:, 0] * 1000.0).astype(np.int64) % ngrains           # one key per grain (its first Kocks angle)
    ang = np.stack([base] + [table[(key + 7 * k) % len(table)] for k in range(1, nc)], axis=1)
    prob.angles = np.ascontiguousarray(ang)
    prob.materials[0].n_crystals = nc
    return prob
