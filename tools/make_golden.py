#!/usr/bin/env python
"""Generate tests/golden/deck_results.json: macroscopic stress-strain curves and iteration
counts of the reference's shipped decks (and the derived variants) computed by the CPU oracle.

The reference ships no golden vectors and cannot be built here (ifort + MKL), so these are
ORACLE outputs, frozen to (a) guard the oracle against drift and (b) give the GPU tests a
committed fixture.  Regenerate with `python tools/make_golden.py` after a deliberate change.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from helpers import deck, mm10_variant, stress_bc_variant  # noqa: E402
from oracle import Oracle  # noqa: E402


def run(prob, nstep=None):
    o = Oracle(prob, threads=1)       # one thread: the reductions are then bit-reproducible
    o.drive_eps_sig(1, 0)
    r = o.FFT_nr3(nstep=nstep)
    assert r["rc"] == 0
    return {"nr_iters": [int(v) for v in r["nr_iters"]], "cg_iters": [[int(v) for v in row] for row in r["cg_iters"]],
            "Pbar": [[float(v) for v in row] for row in r["Pbar"]],
            "P_absmax": float(np.abs(o.Pn1).max()), "F_checksum": float(np.abs(o.Fn1).sum())}


def main():
    out = {"generator": "tools/make_golden.py (CPU oracle, 1 thread)",
           "test_mm01.in": run(deck("test_mm01.in")),
           "test_mm10.in": run(deck("test_mm10.in")),
           "test_mm10.in+angle2.in": run(mm10_variant("angle2.in"), nstep=4),
           "test_mm01.in+P_yy=P_zz=0": run(stress_bc_variant(deck("test_mm01.in")), nstep=3),
           # derived deck: polycrystalline material points (n_crystals 2, fcc + bcc48 from a crystal file)
           "taylor_mm10.in": run(deck("taylor_mm10.in")),
           # derived deck: MTS hardening law
           "mts_mm10.in": run(deck("mts_mm10.in"))}
    path = os.path.join(ROOT, "tests", "golden", "deck_results.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
