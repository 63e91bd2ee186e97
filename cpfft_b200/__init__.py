"""cpfft_b200: B200-native hot path of CPFFT (finite-strain FFT homogenisation with bilinear
Mises and crystal-plasticity material updates), delivered as a C-ABI CUDA library
(``libcpfft_b200.so``, include/cpfft_b200.h) plus this thin host mirror."""
from .problem import Problem, Material, Crystal  # noqa: F401
from .deck import read_deck  # noqa: F401
from .api import Solver, CpfftError, load_library, library_path, EXPORTS  # noqa: F401
