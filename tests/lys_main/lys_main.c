/* A lys-style C main without SDL: the calls lys' liblys.c makes on the generated library, in its order, against
 * include/flowamok_lys.h.  Built twice by tests/test_lys_main.py: against libflowamok_explorer.so and, with
 * -DFLOWAMOK_LYS_DESIGNER, against libflowamok_designer.so.  Prints a checksum of every frame. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include "flowamok_lys.h"

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
  const int n_frames = argc > 1 ? atoi(argv[1]) : 30;
  const int64_t height = 600, width = 800;
  struct futhark_context_config *cfg = futhark_context_config_new();
  futhark_context_config_set_device(cfg, "#0");
  struct futhark_context *ctx = futhark_context_new(cfg);
  struct futhark_opaque_state *state = NULL, *next = NULL;
  CHECK(futhark_entry_init(ctx, &state, 42u, height, width));

  struct futhark_u8_1d *fmt = NULL;
  CHECK(futhark_entry_text_format(ctx, &fmt));
  const int64_t fmt_len = futhark_shape_u8_1d(ctx, fmt)[0];
  char *fmt_s = calloc((size_t)fmt_len + 1, 1);
  CHECK(futhark_values_u8_1d(ctx, fmt, (uint8_t *)fmt_s));
  CHECK(futhark_free_u8_1d(ctx, fmt));
  bool grab = true;
  CHECK(futhark_entry_grab_mouse(ctx, &grab));

#ifndef FLOWAMOK_LYS_DESIGNER
  /* the explorer starts paused on scenario 0: go to "close crossing" and switch auto on */
  const int32_t keys[] = {0x4000004F, 0x4000004F, 0x4000004F, 0x4000004F, 'a'};
  for (unsigned k = 0; k < sizeof keys / sizeof keys[0]; k++) {
    CHECK(futhark_entry_key(ctx, &next, 0, keys[k], state));
    CHECK(futhark_free_opaque_state(ctx, state));
    state = next;
  }
#else
  /* the designer: draw a road down from the entry road */
  const int32_t strokes[][3] = {{1, 105, 255}, {1, 105, 455}, {0, 0, 0}};
  for (unsigned k = 0; k < 3; k++) {
    CHECK(futhark_entry_mouse(ctx, &next, strokes[k][0], strokes[k][1], strokes[k][2], state));
    CHECK(futhark_free_opaque_state(ctx, state));
    state = next;
  }
#endif

  uint32_t *pixels = malloc((size_t)(height * width) * 4);
  uint64_t sum = 0;
  for (int f = 0; f < n_frames; f++) {
    CHECK(futhark_entry_step(ctx, &next, 1.0f / 30.0f + 1e-4f, state));
    CHECK(futhark_free_opaque_state(ctx, state));
    state = next;
    struct futhark_u32_2d *frame = NULL;
    CHECK(futhark_entry_render(ctx, &frame, state));
    const int64_t *shape = futhark_shape_u32_2d(ctx, frame);
    if (shape[0] != height || shape[1] != width) return 2;
    CHECK(futhark_values_u32_2d(ctx, frame, pixels));
    CHECK(futhark_context_sync(ctx));
    CHECK(futhark_free_u32_2d(ctx, frame));
    for (int64_t i = 0; i < height * width; i++) sum = sum * 1099511628211ull + pixels[i];
  }
  uint32_t colour = 0;
  CHECK(futhark_entry_text_colour(ctx, &colour, state));
#ifndef FLOWAMOK_LYS_DESIGNER
  int32_t a, b, c, d, f;
  int64_t e, g, h;
  CHECK(futhark_entry_text_content(ctx, &a, &b, &c, &d, &e, &f, &g, &h, 30.0f, state));
  printf("explorer frames=%d checksum=%016" PRIx64 " text=(%d,%d,%d,%d,%" PRId64 ",%d,%" PRId64 ",%" PRId64 ") colour=%08x grab=%d fmt=%zu\n",
         n_frames, sum, a, b, c, d, e, f, g, h, colour, (int)grab, (size_t)fmt_len);
#else
  int32_t a, b, c;
  CHECK(futhark_entry_text_content(ctx, &a, &b, &c, 30.0f, state));
  printf("designer frames=%d checksum=%016" PRIx64 " text=(%d,%d,%d) colour=%08x grab=%d fmt=%zu\n", n_frames, sum, a, b, c, colour,
         (int)grab, (size_t)fmt_len);
#endif
  free(pixels);
  free(fmt_s);
  CHECK(futhark_free_opaque_state(ctx, state));
  futhark_context_free(ctx);
  futhark_context_config_free(cfg);
  return 0;
}
