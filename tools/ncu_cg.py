"""Minimal CG workload for an `ncu --set full` capture of the kernels of one CG iteration at one
grid size (no torch import: starts in seconds).  See profiles/README.md for the command."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from test_oracle_spectral import _toy_problem  # noqa: E402
from cpfft_b200 import Solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
p = _toy_problem(N)
s = Solver(p)
s.drive_eps_sig(1, 0)
x = np.random.default_rng(0).standard_normal((9, p.N3))
s.upload("DFM", x)
s.G_K_dF("DFM", "B", 1)
print("cg iterations", s.fftPcg("B", "DFM", 1e-3))
