"""Development helper: accuracy of the GPU spectral operator (fast and generic paths) against
the oracle, with the location of the largest error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_oracle_spectral import _toy_problem
from cpfft_b200 import Solver
from oracle import Oracle

for N in (16, 32, 64):
    p = _toy_problem(N)
    o = Oracle(p, threads=8)
    rng = np.random.default_rng(3)
    F = np.zeros((9, p.N3)); F[[0, 4, 8]] = 1.0
    F += 0.02 * rng.standard_normal((9, p.N3))
    o.Fn1[:] = F; o.drive_eps_sig(1, 1)
    x = rng.standard_normal((9, p.N3))
    for force in (False, True):
        if force:
            os.environ["CPFFT_GENERIC_FFT"] = "1"
        else:
            os.environ.pop("CPFFT_GENERIC_FFT", None)
        s = Solver(p)
        s.upload("FN1", F); s.drive_eps_sig(1, 1)
        k4err = np.abs(s.download("K4") - o.K4).max() / np.abs(o.K4).max()
        for flg in (0, 1):
            s.upload("DFM", x); s.G_K_dF("DFM", "B", flg)
            got = s.download("B"); ref = o.G_K_dF(x, flg)
            d = np.abs(got - ref)
            c, e = np.unravel_index(np.argmax(d), d.shape)
            # spectrum of the error: which frequencies carry it
            E = np.fft.fftn((got - ref)[c].reshape(N, N, N))
            kk = np.unravel_index(np.argmax(np.abs(E)), E.shape)
            print(f"N={N} generic={force} flgK={flg} relerr={d.max() / np.abs(ref).max():.3e} K4err={k4err:.2e} "
                  f"at comp {c} voxel {(e // (N * N), (e // N) % N, e % N)} err-spectrum peak {kk} "
                  f"rms={np.sqrt((d * d).mean()) / np.abs(ref).max():.3e}", flush=True)
        s.close()
