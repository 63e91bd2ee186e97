"""The C-ABI boundary: libcpfft_b200.so loads here (no GPU), exports every symbol that
include/cpfft_b200.h declares, and the product fails loudly -- never falls back -- without a
CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cpfft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpfft_[a-zA-Z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from cpfft_b200.build import build
    build()
    import cpfft_b200
    return cpfft_b200.load_library()


def test_header_and_python_mirror_agree():
    import cpfft_b200
    assert header_symbols() == sorted(cpfft_b200.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_no_torch_or_oracle_dependency(lib):
    """plain C ABI: the shared object must not link libtorch, python or the oracle"""
    import subprocess
    import cpfft_b200
    out = subprocess.run(["ldd", cpfft_b200.library_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out and "oracle" not in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cpfft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".f90")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import deck
    from cpfft_b200 import Solver, CpfftError
    with pytest.raises(CpfftError) as ei:
        Solver(deck("test_mm01.in"))
    assert ei.value.code < 0          # CUDA error, no CPU fallback


def test_usage_errors_without_device(lib):
    assert lib.cpfft_hist_size(None) == 0
    assert lib.cpfft_local_voxels(None) == 0
    assert lib.cpfft_last_error(None) == b"null handle"
    assert lib.cpfft_profile_classes() == 11
    names = [lib.cpfft_profile_name(i).decode() for i in range(11)]
    assert "k_x_green" in names and "k_update_mm10" in names


def test_fortran_module_binds_every_symbol():
    """cpfft_b200/fortran/cpfft_iso_c.f90 (not compilable here: no Fortran compiler) must at
    least declare a bind(c) interface for every export and the same field ids as the header."""
    f90 = open(os.path.join(ROOT, "cpfft_b200", "fortran", "cpfft_iso_c.f90")).read()
    bound = sorted(set(re.findall(r"name='(cpfft_[A-Za-z0-9_]+)'", f90)))
    assert bound == header_symbols()
    hdr = open(os.path.join(ROOT, "include", "cpfft_b200.h")).read()
    enum = re.search(r"typedef enum \{(.*?)\} cpfft_field;", hdr, flags=re.S).group(1)
    enum = re.sub(r"/\*.*?\*/", "", enum, flags=re.S)
    names = [n.strip().split("=")[0].strip() for n in enum.split(",") if n.strip()]
    names = [n for n in names if n != "CPFFT_NUM_FIELDS"]
    for i, n in enumerate(names):
        assert re.search(rf"\b{n} = {i}\b", f90), (n, i)


def test_fortran_hooks_use_only_bound_symbols():
    """cpfft_b200/fortran/cpfft_hooks.f (the reference-side replacement bodies of INTEGRATION.md 4,
    fixed form): every cpfft_* routine it calls is bound by the interface module, and no line runs
    past column 72."""
    base = os.path.join(ROOT, "cpfft_b200", "fortran")
    hooks = open(os.path.join(base, "cpfft_hooks.f")).read()
    f90 = open(os.path.join(base, "cpfft_iso_c.f90")).read()
    bound = set(re.findall(r"name='(cpfft_[A-Za-z0-9_]+)'", f90))
    helpers = {"cpfft_check", "cpfft_h", "cpfft_model_to_gpu", "cpfft_download_results", "cpfft_iso_c",
               "cpfft_config", "cpfft_material", "cpfft_crystal", "cpfft_hooks"}
    code = "\n".join(l for l in hooks.splitlines() if l[:1] not in ("c", "C"))
    used = set(re.findall(r"\b(cpfft_[A-Za-z0-9_]+)", code))
    assert used - helpers <= bound, sorted(used - helpers - bound)
    assert {"cpfft_create", "cpfft_set_materials", "cpfft_set_voxels", "cpfft_drive_eps_sig", "cpfft_G_K_dF",
            "cpfft_FFT_nr3", "cpfft_step_log"} <= used
    assert all(len(l) <= 72 for l in hooks.splitlines())
    assert all(h in f90 or h in ("cpfft_model_to_gpu", "cpfft_download_results", "cpfft_hooks") for h in helpers)


def test_plain_c_host(lib, tmp_path):
    """a C99 program compiled against include/cpfft_b200.h and linked with the library (the view of a cgo /
    ISO_C_BINDING caller): the header is valid C, the struct sizes are those of the Python mirror, and without a
    GPU cpfft_create fails with a CUDA error instead of falling back to anything"""
    import ctypes
    import subprocess
    import torch
    from cpfft_b200.api import library_path, Config
    from cpfft_b200.problem import MaterialPOD, CrystalPOD
    exe = str(tmp_path / "c_host")
    libdir = os.path.dirname(library_path())
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "c_host.c"), "-o", exe, "-L", libdir, "-lcpfft_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0 and out.stdout.startswith("created N=8 voxels=512"), out.stdout
    else:
        assert out.returncode == 10, (out.returncode, out.stdout)
        sizes = dict(kv.split("=") for kv in out.stdout.split()[3:])
        assert int(sizes["sizeof(config)"]) == ctypes.sizeof(Config)
        assert int(sizes["sizeof(material)"]) == ctypes.sizeof(MaterialPOD)
        assert int(sizes["sizeof(crystal)"]) == ctypes.sizeof(CrystalPOD)
