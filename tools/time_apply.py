"""Development helper: per-kernel time of the spectral operator at one grid size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_oracle_spectral import _toy_problem
from cpfft_b200 import Solver
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = _toy_problem(N)
s = Solver(p)
s.drive_eps_sig(1, 0)
rng = np.random.default_rng(0)
x = rng.standard_normal((9, p.N3))
s.upload("DFM", x)
for _ in range(3):
    s.G_K_dF("DFM", "B", 1)
s.profile(True); s.profile_reset()
for _ in range(20):
    s.G_K_dF("DFM", "B", 1)
t = s.profile_table()
print(os.environ.get("CPFFT_B200_LIB", "default"), "N", N, {k: round(v[0] / max(v[1], 1), 4) for k, v in t.items() if v[1]},
      "checksum", float(np.abs(s.download("B")).sum()))
# the same kernels inside a CG solve (fused direction / solution updates in k_fz, p.Ap sums in k_iz_pipe)
F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
F += 0.002 * rng.standard_normal((9, p.N3))
s.upload("FN1", F); s.drive_eps_sig(1, 1)
s.upload("DFM", x); s.G_K_dF("DFM", "B", 1)
s.profile_reset()
it, rr = s.fftPcg("B", "DFM", 1e-6)
t = s.profile_table()
print("   CG", it, "iterations", {k: round(v[0] / max(v[1], 1), 4) for k, v in t.items() if v[1]},
      "ms/iteration", round(sum(v[0] for v in t.values()) / max(it, 1), 4))
