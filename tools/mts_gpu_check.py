"""Development helper: the `--variant mts` bench workload (MTS hardening on the benchmark polycrystal) on the GPU against
the oracle at small sizes: Newton / CG counts, P, local failures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpfft_b200 import Solver
from cpfft_b200.polycrystal import polycrystal, workload_variant
from oracle import Oracle

for N, G, nstep in ((16, 20, 3), (32, 200, 3), (32, 1000, 3), (40, 1000, 3)):
    p = workload_variant(polycrystal(N, ngrains=G), "mts", G)
    o = Oracle(p); o.drive_eps_sig(1, 0)
    ro = o.FFT_nr3(nstep=nstep)
    s = Solver(workload_variant(polycrystal(N, ngrains=G), "mts", G)); s.drive_eps_sig(1, 0)
    try:
        rs = s.FFT_nr3(nstep=nstep)
        err = np.abs(s.download("PN1") - o.Pn1).max() / np.abs(o.Pn1).max()
        print(N, G, "oracle nr", [int(v) for v in ro["nr_iters"]], "gpu nr", [int(v) for v in rs["nr_iters"]],
              "cg max", [max(c) if len(c) else 0 for c in ro["cg_iters"]], [max(c) if len(c) else 0 for c in rs["cg_iters"]],
              "P err %.2e" % err, "fail", s.material_failures(), flush=True)
    except Exception as e:
        print(N, G, "oracle nr", [int(v) for v in ro["nr_iters"]], "GPU FAILED:", str(e)[:150], "fail", s.material_failures(), flush=True)
        # where do the fields differ after the failing run?
        P = s.download("PN1")
        print("   nan in P:", int(np.isnan(P).sum()), "max |P|", float(np.nanmax(np.abs(P))), "oracle max |P|", float(np.abs(o.Pn1).max()), flush=True)
