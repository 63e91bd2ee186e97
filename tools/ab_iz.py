"""Development helper: A/B timing of the kernel variants behind the development switches
(CPFFT_IZ_PIPE = 0 k_iz / 1 k_iz_pipe; CPFFT_CG_FUSE_X = 0 separate CG update pass / 1 solution
update fused into the next forward z pass) inside G_K_dF and inside a CG solve, one process, one
grid.  CUDA events on the launching stream (the library's kernel-class profiler).
    python tools/ab_iz.py [N=256] [repeats=20]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from test_oracle_spectral import _toy_problem  # noqa: E402
from cpfft_b200 import Solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p = _toy_problem(N)
rng = np.random.default_rng(0)
x = rng.standard_normal((9, p.N3))
out = {"grid": N, "repeats": REP, "modes": {}}
MODES = [("00", 8), ("10", 8), ("01", 8), ("11", 8), ("00", 8)]   # (iz_pipe, cg_fuse_x); baseline first and last: drift check
for mode, lpc in MODES:
    os.environ["CPFFT_IZ_PIPE"] = mode[0]
    os.environ["CPFFT_CG_FUSE_X"] = mode[1]
    os.environ["CPFFT_IZ_LPC"] = str(lpc)
    s = Solver(p)
    s.drive_eps_sig(1, 0)
    s.upload("DFM", x)
    for _ in range(3):
        s.G_K_dF("DFM", "B", 1)
    s.profile(True); s.profile_reset()
    for _ in range(REP):
        s.G_K_dF("DFM", "B", 1)
    t = s.profile_table()
    ent = {"apply_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in t.items() if v[1]}}
    chk = float(np.abs(s.download("B")).sum())
    s.profile_reset()
    it, rr = s.fftPcg("B", "DFM", 1e-6)       # CG: k_iz<DOT> with the fused p.Ap partial sums
    t = s.profile_table()
    ent["cg_ms"] = {k: round(v[0] / max(v[1], 1), 4) for k, v in t.items() if v[1]}
    ent["cg_iters"] = it
    ent["checksum"] = chk
    key = f"iz_pipe={mode[0]} cg_fuse_x={mode[1]}" + ("" if f"iz_pipe={mode[0]} cg_fuse_x={mode[1]}" not in out["modes"] else " (again)")
    ent["cg_ms_per_iteration"] = round(sum(v[0] for v in t.values()) / max(it, 1), 4)
    out["modes"][key] = ent
    print(key, json.dumps(ent), flush=True)
    del s
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"ab_iz_{N}.json"), "w") as f:
    json.dump(out, f, indent=1)
