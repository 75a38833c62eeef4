/* Plain-C consumer of include/sageattn_b200.h (no CUDA, no torch): proves the boundary is a C ABI — the header compiles as C,
 * the library links with gcc alone, argument validation answers before any CUDA call.  Built and run by
 * tests/test_capi_cpu.py::test_plain_c_program_links_and_validates. */
#include <stdio.h>
#include <string.h>
#include "sageattn_b200.h"

int main(void) {
  if (sab_version() < 100) { printf("bad version %d\n", sab_version()); return 1; }
  /* null tensors: rejected with SAB_ERR_INVALID before the device is touched */
  int st = sab_qk_int8_sv_f8_attn(0, 0, 0, 0, 0, 0, 0, 0, 0, SAB_DTYPE_BF16, 1, 2, 2, 128, 128, 128, 0, 0, 0, 0, 0, 0, 128, 0, 0, 0,
                                  0, SAB_GRAN_PER_THREAD, SAB_GRAN_PER_THREAD, 1.0f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  if (st != SAB_ERR_INVALID || strstr(sab_last_error(), "null") == 0) { printf("unexpected: %d '%s'\n", st, sab_last_error()); return 2; }
  /* unsupported head dim */
  char x[16];
  st = sab_qk_int8_sv_f8_attn((const int8_t*)x, (const int8_t*)x, (const uint8_t*)x, x, 0, (const float*)x, (const float*)x, 0, 0,
                              SAB_DTYPE_FP16, 1, 2, 2, 128, 128, 96, 0, 0, 0, 0, 0, 0, 128, 0, 0, 0, 0, SAB_GRAN_PER_WARP,
                              SAB_GRAN_PER_WARP, 1.0f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  if (st != SAB_ERR_UNSUPPORTED) { printf("unexpected: %d '%s'\n", st, sab_last_error()); return 3; }
  printf("C_ABI_OK %d\n", sab_version());
  return 0;
}
