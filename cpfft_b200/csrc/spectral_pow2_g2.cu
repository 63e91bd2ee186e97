// cpfft_b200: fast spectral path, instantiations for the grid sizes of group 2 (spectral_pow2_decl.cuh)
#include "spectral_pow2_impl.cuh"

CPF_POW2_SIZES_G2(CPF_POW2_INSTANTIATE)
