// cpfft_b200: mm10 sweep kernels for the MTS hardening law (`hardening mts`), single crystals and Taylor points.
// Same source as material.cu's kernels (upd_mm10_voxel<.., HARD = MM10_MTS, ..>), a translation unit of its own for build time.
#include "material_kernels.cuh"

MM10_KERNEL(k_update_mm10_mts, false, MM10_MTS, false, false)
MM10_KERNEL(k_update_mm10_mts_u, false, MM10_MTS, false, true)
MM10_KERNEL(k_update_mm10_taylor_mts, true, MM10_MTS, false, false)
MM10_KERNEL(k_update_mm10_taylor_mts_u, true, MM10_MTS, false, true)
