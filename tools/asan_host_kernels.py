"""Address / undefined-behaviour sanitizer run of the kernel source (host build, tests/native/material_host.cpp):
every crystal / grid variant of tests/test_host_kernels.py plus the decks and a mixed Taylor polycrystal, three load
steps each, the last one a 4 % increment that forces sub-stepping and local failures.  Usage:

    g++ -O1 -g -march=x86-64-v3 -std=c++17 -fopenmp -fPIC -shared -w -fsanitize=address,undefined \
        -fno-omit-frame-pointer -o /tmp/asan/libmaterial_host.so tests/native/material_host.cpp
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_host_kernels.py

Round 1: no report (all 23 cases, lattice-frame variant included).  The oracle was run the same way (its three
sources with the same flags, whole FFT_nr3 solves of the four decks, 8^3 / 9^3 polycrystals, a stress-BC run): no report,
and tests/native/fft_core_host_test.cpp (the FFT building blocks) likewise."""
import os
import sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(_ROOT, 'tests')); sys.path.insert(0, _ROOT)
import numpy as np
import host_kernels
host_kernels._OUT = os.environ.get('ASAN_HOST_LIB', '/tmp/asan/libmaterial_host.so')
host_kernels.build = lambda force=False: host_kernels._OUT
from host_kernels import HostKernels
from test_host_kernels import VARIANTS, _variant_problem
from helpers import deck
from cpfft_b200.polycrystal import taylor_polycrystal
probs = [(k, _variant_problem(k)) for k in VARIANTS] + [("mm01", deck("test_mm01.in")), ("mm10deck", deck("test_mm10.in")),
        ("taylor_mixed", taylor_polycrystal(4, ncrystals=3, ngrains=12, mixed=True))]
for name, p in probs:
    for lf in ((False, True) if name in ("voce_m_2", "diffusion") else (False,)):
        k = HostKernels(p, lattice_frame=lf)
        rng = np.random.default_rng(11)
        G = rng.standard_normal((9, p.N3))
        bar = np.zeros((9, 1)); bar[0] = 1.0; bar[4] = bar[8] = -0.45; bar[1] = 0.3
        I = np.zeros((9, p.N3)); I[[0, 4, 8]] = 1.0
        k.drive_eps_sig(1, 0)
        for step, amp in ((1, 0.003), (2, 0.006), (3, 0.04)):       # last one forces sub-stepping / failures
            for it in (0, 1):
                F = I + amp * (bar + 0.25 * G)
                k.Fn1[:] = F
                nf = k.drive_eps_sig(step, it)
            k.Fn[:] = F; k.update()
        print(name, lf, "ok, failures in last sweep:", nf, flush=True)
        del k
print("ASAN RUN COMPLETE")
