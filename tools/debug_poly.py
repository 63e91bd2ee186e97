"""Development helper: run the polycrystal on the GPU until a local mm10 failure, dump the
state of the failing voxels to gpurun_out/fail_dump.npz for analysis against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpfft_b200 import Solver
from cpfft_b200.polycrystal import polycrystal

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = polycrystal(N)
s = Solver(p)
s.drive_eps_sig(1, 0)
for step in range(1, 6):
    try:
        rs = s.FFT_nr3(nstep=1, first=step - 1)
        print("step", step, "nr", rs["nr_iters"], "cg", rs["cg_iters"], flush=True)
    except Exception as ex:
        print("GPU step", step, "failed:", ex)
        ff = s.fail_flags()
        li = s.local_iters()
        bad = np.nonzero(ff)[0]
        print("nfail", len(bad), "iters max", li.max(axis=0))
        bad = bad[:50]
        F1 = s.download("FN1")[:, bad]
        Fn = s.download("FN")[:, bad]
        H = s.download("HIST_N")[:, bad]
        U = s.download("URCS_N")[:, bad]
        E = s.download("EPS_N")[:, bad]
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/fail_dump.npz", bad=bad, Fn1=F1, Fn=Fn, hist_n=H, urcs_n=U, eps_n=E,
                 angles=p.angles[bad], step=step, msg=str(ex), liters=li[bad])
        print("dumped", len(bad))
        break
