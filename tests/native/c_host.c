/* A plain-C host of the C ABI (tests/test_abi.py::test_plain_c_host): what a cgo / ISO_C_BINDING /
 * ctypes caller sees.  Compiled as C99 against include/cpfft_b200.h and linked with the library.
 * On a machine without a GPU cpfft_create() must FAIL with a CUDA error and a message -- there is
 * no CPU fallback; with a GPU it creates and destroys a handle.  Prints one line for the test. */
#include <stdio.h>
#include <string.h>
#include "cpfft_b200.h"

int main(void) {
  cpfft_config cfg;
  cpfft_handle* h = NULL;
  int rc, nclass, i, found = 0;
  memset(&cfg, 0, sizeof cfg);
  cfg.N = 8; cfg.device = 0; cfg.rank = 0; cfg.world = 1; cfg.maxIter = 20;
  cfg.tolNR = 1e-5; cfg.tolPCG = 1e-10; cfg.tstep = 1.0;
  if (cpfft_hist_size(NULL) != 0 || cpfft_local_voxels(NULL) != 0) { puts("null-handle queries"); return 2; }
  if (strcmp(cpfft_last_error(NULL), "null handle") != 0) { puts("null-handle message"); return 2; }
  nclass = cpfft_profile_classes();
  for (i = 0; i < nclass; ++i) found += strcmp(cpfft_profile_name(i), "k_update_mm10") == 0;
  if (found != 1) { puts("profile classes"); return 2; }
  rc = cpfft_create(&cfg, &h);
  if (rc == 0) {
    printf("created N=%d voxels=%lld\n", cfg.N, (long long)cpfft_local_voxels(h));
    cpfft_destroy(h);
    return 0;
  }
  printf("create failed rc=%d sizeof(config)=%d sizeof(material)=%d sizeof(crystal)=%d\n", rc, (int)sizeof(cpfft_config),
         (int)sizeof(cpfft_material), (int)sizeof(cpfft_crystal));
  return rc < 0 ? 10 : 3;
}
