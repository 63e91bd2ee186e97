#!/usr/bin/env python
"""Generate tests/golden/poly_*.npz: the benchmark polycrystal (cpfft_b200/polycrystal.py) at sizes the
CPU oracle needs minutes for, frozen as fixtures for the `-m gpu` parity tests.

    python tools/make_golden_poly.py            # all cases (about 15 min on 8 cores)
    python tools/make_golden_poly.py poly32_stress

Per case: Newton iterations per load step, CG iterations per solve, the macroscopic stress P_bar
per step, and -- at SAMPLE deterministic voxels (numpy default_rng(N)) -- F, P, the unrotated
stress and the whole mm10 history of the LAST step, plus max |P| and sum |F| over the grid.
These are ORACLE outputs (the reference cannot be built here: ifort + MKL), frozen so that the
GPU tests can hold the north-star tolerances on >= 6 plastic load steps at 32^3 / 64^3 without
re-running the oracle for ten minutes.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from oracle import Oracle  # noqa: E402
from cpfft_b200.polycrystal import polycrystal  # noqa: E402

SAMPLE = 1024
# name: (N, grains, stress_bc, load steps)
CASES = {
    "poly32_strain": (32, 64, False, 8),
    "poly32_stress": (32, 64, True, 8),
    "poly64_strain": (64, 200, False, 8),
}


def sample_voxels(N):
    return np.sort(np.random.default_rng(N).choice(N ** 3, size=min(4 * SAMPLE, N ** 3), replace=False))[::4]


def run(name):
    N, grains, sbc, nstep = CASES[name]
    p = polycrystal(N, ngrains=grains, stress_bc=sbc)
    o = Oracle(p, threads=os.cpu_count())
    o.drive_eps_sig(1, 0)
    t0 = time.time()
    r = o.FFT_nr3(nstep=nstep)
    assert r["rc"] == 0
    idx = sample_voxels(N)
    cg = np.full((nstep, 64), -1, dtype=np.int32)
    for s, row in enumerate(r["cg_iters"]):
        cg[s, :len(row)] = row
    out = dict(N=N, grains=grains, stress_bc=int(sbc), nstep=nstep, nr_iters=np.asarray(r["nr_iters"], dtype=np.int32),
               cg_iters=cg, Pbar=np.asarray(r["Pbar"]), idx=idx, F=o.Fn1[:, idx].copy(), P=o.Pn1[:, idx].copy(),
               urcs=o.urcs_n1[idx].copy(), hist=o.hist_n[idx].copy(), P_absmax=np.abs(o.Pn1).max(),
               F_abssum=np.abs(o.Fn1).sum(), failures=np.asarray(r["counters"][3:5]))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {time.time() - t0:.0f} s, newton {list(out['nr_iters'])}, wrote {path}", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run(n)
