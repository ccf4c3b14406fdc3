/* A host program written against the C library `futhark <backend> --library README.fut` generates (README.fut:9-13: entry
 * points init and step), compiled as C against include/flowamok_futhark.h and linked against libflowamok_b200.so instead --
 * the loop of INTEGRATION.md section 1.  Prints a checksum over all frames; tests/test_readme_main.py compares it with the
 * frames of the same program executed from the reference source (tests/golden/lys_traces.npz, readme/). */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include "flowamok_futhark.h"

#define CHECK(x)                                              \
  do {                                                        \
    if ((x) != 0) {                                           \
      char *m = futhark_context_get_error(ctx);               \
      fprintf(stderr, "%s failed: %s\n", #x, m ? m : "?");    \
      free(m);                                                \
      return 1;                                               \
    }                                                         \
  } while (0)

int main(int argc, char **argv) {
  const int n_frames = argc > 1 ? atoi(argv[1]) : 40;
  struct futhark_context_config *cfg = futhark_context_config_new();
  struct futhark_context *ctx = futhark_context_new(cfg);
  struct futhark_opaque_state *s = NULL, *s2 = NULL;
  CHECK(futhark_entry_init(ctx, &s, 30, 30, 10, 123));                 /* README.fut:15  init 30 30 10 123 */
  uint32_t *pixels = malloc(300 * 300 * sizeof(uint32_t));
  for (int k = 0; k < n_frames; k++) {                                 /* :video (step, ..., 400) */
    struct futhark_u32_2d *frame = NULL;
    CHECK(futhark_entry_step(ctx, &frame, &s2, s));
    const int64_t *shape = futhark_shape_u32_2d(ctx, frame);
    if (shape[0] != 300 || shape[1] != 300) return 2;
    CHECK(futhark_values_u32_2d(ctx, frame, pixels));
    CHECK(futhark_context_sync(ctx));
    CHECK(futhark_free_u32_2d(ctx, frame));
    CHECK(futhark_free_opaque_state(ctx, s));
    s = s2;
    uint64_t sum = 0;
    for (int i = 0; i < 300 * 300; i++) sum = sum * 1099511628211ull + pixels[i];
    printf("frame %d %016" PRIx64 "\n", k, sum);
  }
  free(pixels);
  CHECK(futhark_free_opaque_state(ctx, s));
  futhark_context_free(ctx);
  futhark_context_config_free(cfg);
  return 0;
}
