#!/usr/bin/env python
"""Development helper: build variants of libcpfft_b200.so with other compile-time knobs into
gpurun_variants/ (travels to the GPU box; git-ignored) for A/B timing with CPFFT_B200_LIB=...:

    python tools/build_variants.py t64c4:-DUPD_THREADS=64,-DMM10_MIN_CTAS=4 unroll2:-DMM10_SLIP_UNROLL=2
    python tools/build_variants.py iz5:tu=spectral_pow2_g1.cu,-DIZ_MINB=5   # another translation unit (256^3 lives in group 1)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpfft_b200.build import CSRC, NVCC_FLAGS  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_variants")


def build(name, defines):
    objdir = os.path.join(OUT, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    tus = [d[3:] for d in defines if d.startswith("tu=")] or ["material.cu", "material_taylor.cu", "material_mts.cu"]
    defines = [d for d in defines if not d.startswith("tu=")]
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + defines + ["-Xptxas", "-v"]
    objs = []
    for src in tus:
        obj = os.path.join(objdir, src[:-3] + ".o")
        log = subprocess.run(["nvcc"] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if log.returncode:
            raise SystemExit(log.stderr[-3000:])
        lines = log.stderr.splitlines()
        for i, l in enumerate(lines):
            if "Compiling entry function '_Z15k_update_mm10_u7UpdArgs'" in l or "Compiling entry function '_Z18k_update_mm10_lf_u7UpdArgs'" in l or ("Compiling" in l and any(k in l for k in ("k_iz_pipeILi256ELb1E", "k_fzILi320ELi3E", "k_fzILi400ELi3E", "k_fxILi400ELb0E", "k_fyfILi400ELb0E", "k_fyiILi400E"))):
                print(name, "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
        objs.append(obj)
    # the other translation units are the default build's objects
    base = os.path.join(ROOT, "cpfft_b200", "build")
    objs += [os.path.join(base, f[:-3] + ".o") for f in ("material.cu", "material_taylor.cu", "material_mts.cu", "spectral.cu", "spectral_pow2.cu",
                                                                 "spectral_pow2_g1.cu", "spectral_pow2_g2.cu", "spectral_pow2_g3.cu", "solver.cu") if f not in tus]
    lib = os.path.join(OUT, f"lib_{name}.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + ["-ldl"])
    return lib


if __name__ == "__main__":
    specs = [a.split(":", 1) for a in sys.argv[1:]]
    with ThreadPoolExecutor(max_workers=4) as pool:
        for lib in pool.map(lambda s: build(s[0], [d for d in s[1].split(",") if d]), specs):
            print("built", lib)
