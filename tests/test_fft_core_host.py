"""Host build of the CUDA FFT building blocks (cpfft_b200/csrc/fft_core.cuh): butterflies,
in-place digit-reversed stages and the real/half-complex packing against a naive DFT."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_core_on_host():
    src = os.path.join(ROOT, "tests", "native", "fft_core_host_test.cpp")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "fft_core_host_test")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
