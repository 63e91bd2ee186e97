"""Development helper: time the material sweep (k_update_mm10 + k_pk1_tangent) on a plastic
polycrystal state.  CPFFT_B200_LIB selects a library variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cpfft_b200 import Solver
from cpfft_b200.polycrystal import polycrystal

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
SLIP = sys.argv[2] if len(sys.argv) > 2 else "fcc"           # fcc | bcc48
p = polycrystal(N, ngrains=200, slip_type={"fcc": 1, "bcc48": 8}[SLIP])
s = Solver(p)
s.drive_eps_sig(1, 0)
s.FFT_nr3(nstep=3)
# a converged-size increment on top of the committed state: what a mid-Newton sweep sees
F = s.download("FN1")
F[0] += 1e-3; F[4] -= 3e-4; F[8] -= 3e-4
s.upload("FN1", F)
s.profile(True); s.profile_reset()
for it in range(5):
    s.drive_eps_sig(4, 1)
t = s.profile_table()
P = s.download("PN1")
print(os.path.basename(os.environ.get("CPFFT_B200_LIB", "default")), "std" if os.environ.get("CPFFT_MM10_LF") == "0" else "LF",
      "nouni" if os.environ.get("CPFFT_MM10_UNI") == "0" else "uni", SLIP, "N", N, {k: round(v[0] / max(v[1], 1), 3) for k, v in t.items() if v[1]},
      "checksum", float(np.abs(P).sum()), "fail", s.material_failures(), "iters", s.local_iters().mean(axis=0))
