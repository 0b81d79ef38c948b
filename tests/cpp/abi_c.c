/* include/gsm.h must be valid C (not only C++): compile-only check, plus a trivial call sequence that a C host
 * program would make.  Built by tests/test_abi.py with gcc -std=c99. */
#include <stdio.h>
#include <string.h>
#include "gsm.h"

int main(void) {
  gsm_ctx* ctx = NULL;
  gsm_params p;
  memset(&p, 0, sizeof(p));
  p.mode = GSM_MODE_GF;
  p.radius = 9;
  p.num_disp = 64;
  printf("%s\n", gsm_version());
  if (gsm_create(&ctx, 0, 64, 64, 64, 1) != GSM_OK) {
    printf("no device: %s\n", gsm_last_error()); /* expected on a CPU-only box: loud failure, no fallback */
    return 0;
  }
  {
    unsigned char l[64 * 64], r[64 * 64], d[64 * 64];
    memset(l, 7, sizeof(l));
    memset(r, 7, sizeof(r));
    if (gsm_stereo_batch(ctx, &p, 1, l, r, d, NULL, 64, 64) != GSM_OK) { printf("error: %s\n", gsm_last_error()); return 1; }
    if (gsm_block_matching(ctx, l, r, d, 64, 64, 5, 16) != GSM_OK) { printf("error: %s\n", gsm_last_error()); return 1; }
    printf("ok %d launches\n", (int)gsm_launch_count(ctx));
  }
  gsm_destroy(ctx);
  return 0;
}
