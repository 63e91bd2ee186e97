// cpfft_b200: mm10 sweep kernels for polycrystalline material points, n_crystals > 1 (Taylor average), Voce hardening.
// Same source as material.cu's kernels (upd_mm10_voxel<MULTI = true, ..>), a translation unit of its own for build time.
#include "material_kernels.cuh"

MM10_KERNEL(k_update_mm10_taylor, true, MM10_VOCE, false, false)
MM10_KERNEL(k_update_mm10_taylor_u, true, MM10_VOCE, false, true)
MM10_KERNEL(k_update_mm10_taylor_lf, true, MM10_VOCE, true, false)
MM10_KERNEL(k_update_mm10_taylor_lf_u, true, MM10_VOCE, true, true)
