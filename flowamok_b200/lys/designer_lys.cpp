// flowamok_b200 -- designer/designer.fut's `module lys` as the library a lys C main links (include/flowamok_lys.h).
//
// The state machine of designer.fut:8-136 restated over the operator-level C ABI: one grid of window size / scale, a car
// of random colour entering every tenth step (flowamok_designer_step: ticks on the device), roads drawn with the mouse
// through the point-edit entries.  Line numbers: /root/reference/designer/designer.fut.
#define FLOWAMOK_LYS_DESIGNER 1
#include "fa_lys.h"

using namespace falys;

struct futhark_opaque_state {          // :9-21
  int64_t h, w, gh, gw, scale;
  Rng rng;
  float time_unused;
  bool running;
  int32_t steps_per_second;
  GridRef grid;
  int64_t n_steps_total;
  int32_t n_leaks;
  bool cursor_some;                    // cursor_prev: option (direction i64)
  int64_t cursor_y, cursor_x;
};

#include "fa_lys_api.inc"

namespace {
typedef futhark_opaque_state State;

// init_grid gh gw rng (:32-36)
int init_grid(futhark_context *c, int64_t gh, int64_t gw, const Rng &rng, Rng *rng_out, GridRef *out) {
  GridRef g;
  if (create_grid(c, gh, gw, rng, rng_out, &g)) return 1;
  if (flowamok_add_line_horizontal(c->fc, g->g, 25, 0, 10, 1) || flowamok_add_line_horizontal(c->fc, g->g, 35, 0, 10, -1)) return 1;
  *out = g;
  return 0;
}

// `s.grid :> [gh][gw]cell` (:60, :135): after a resize the grid keeps its old size (:53-55) and the coercion fails at run time
int coerce(futhark_context *c, const State *s) {
  int64_t gh = 0, gw = 0;
  if (flowamok_grid_shape(c->fc, s->grid->g, &gh, &gw)) return 1;
  if (gh != s->gh || gw != s->gw)
    return raise(c, "Value of (core language) shape [%lld][%lld] cannot match shape of type `[%lld][%lld]cell`", (long long)gh,
                 (long long)gw, (long long)s->gh, (long long)s->gw);
  return 0;
}

// step s n_steps (:59-75)
int step_n(futhark_context *c, State *s, int64_t n_steps) {
  if (coerce(c, s)) return 1;
  GridRef g;
  if (clone(c, s->grid, &g)) return 1;                                       // :63 copy
  uint64_t st = s->rng.state, inc = s->rng.inc;
  int64_t total = s->n_steps_total;
  int32_t leaks = 0;
  if (flowamok_designer_step(c->fc, g->g, n_steps, &total, &st, &inc, &leaks)) return 1;
  s->rng = Rng{st, inc}; s->grid = g; s->n_steps_total = total; s->n_leaks = leaks;   // :72-75
  return 0;
}

int64_t sgn(int64_t v) { return (v > 0) - (v < 0); }

// add_lines (gy0, gx0) (gy1, gx1) grid (:97-110): an L from the previous cursor cell to this one, horizontal leg first
int add_lines(futhark_context *c, flowamok_grid *g, int64_t gy0, int64_t gx0, int64_t gy1, int64_t gx1) {
  const int64_t ylo = gy0 < gy1 ? gy0 : gy1, yhi = gy0 < gy1 ? gy1 : gy0;
  const int64_t xlo = gx0 < gx1 ? gx0 : gx1, xhi = gx0 < gx1 ? gx1 : gx0;
  if (gx0 != gx1 && flowamok_add_line_horizontal(c->fc, g, gy0, xlo, xhi, (int8_t)sgn(gx1 - gx0))) return 1;
  if (gy0 != gy1 && flowamok_add_line_vertical(c->fc, g, ylo, yhi, gx1, (int8_t)sgn(gy1 - gy0))) return 1;
  return 0;
}

// mouse buttons x y s (:112-125)
int mouse(futhark_context *c, State *s, int32_t buttons, int32_t x32, int32_t y32) {
  if (buttons != 1) { s->cursor_some = false; return 0; }                    // :125
  const int64_t y = y32, x = x32;
  if (!(y >= 0 && y < s->h && x >= 0 && x < s->w)) return 0;                 // :115, :124
  if (s->scale == 0) return raise(c, "division by zero");
  const int64_t gy = y / s->scale, gx = x / s->scale;                        // :116 (non-negative: floor == truncation)
  if (s->cursor_some) {                                                      // :117-121
    GridRef g;
    if (clone(c, s->grid, &g)) return 1;
    if (add_lines(c, g->g, s->cursor_y, s->cursor_x, gy, gx)) return 1;
    g->cycles_valid = false;
    s->grid = g;
  }
  s->cursor_some = true; s->cursor_y = gy; s->cursor_x = gx;                 // :122
  return 0;
}

template <class F>
int transition(futhark_context *ctx, futhark_opaque_state **out0, const futhark_opaque_state *s, F f) {
  if (!ctx || !out0 || !s) return 1;
  if (!ctx->fc) return raise(ctx, "no device context");
  std::unique_ptr<State> n(new State(*s));
  if (f(n.get())) return 1;
  *out0 = n.release();
  return 0;
}
}  // namespace

extern "C" {

int futhark_entry_init(struct futhark_context *ctx, struct futhark_opaque_state **out0, const uint32_t seed, const int64_t h,
                       const int64_t w) {                                    // :38-47
  if (!ctx || !out0) return 1;
  if (!ctx->fc) return raise(ctx, "no device context");
  std::unique_ptr<State> s(new State());
  s->scale = 10;
  s->h = h; s->w = w; s->gh = h / s->scale; s->gw = w / s->scale;
  if (h < 0 || w < 0) return raise(ctx, "init: negative window %lld x %lld", (long long)h, (long long)w);
  if (init_grid(ctx, s->gh, s->gw, rng_from_seed((int32_t)seed), &s->rng, &s->grid)) return 1;
  s->time_unused = 0.0f; s->running = true; s->steps_per_second = 30;
  s->n_steps_total = 0; s->n_leaks = 0; s->cursor_some = false; s->cursor_y = s->cursor_x = 0;
  *out0 = s.release();
  return 0;
}

int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int64_t h, const int64_t w,
                         const struct futhark_opaque_state *s) {             // :49-55: the grid itself is NOT resized (FIXME there)
  return transition(ctx, out0, s, [&](State *n) {
    if (h < 0 || w < 0) return raise(ctx, "resize: negative window %lld x %lld", (long long)h, (long long)w);
    n->h = h; n->w = w; n->gh = h / n->scale; n->gw = w / n->scale;
    return 0;
  });
}

int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t e, const int32_t key,
                      const struct futhark_opaque_state *s) {                // :88-95
  return transition(ctx, out0, s, [&](State *n) {
    if (e != 0) return 0;
    if (key == K_SPACE) { n->running = !n->running; return 0; }
    if (key == K_r) {
      Rng rng;
      GridRef g;
      if (init_grid(ctx, n->gh, n->gw, n->rng, &rng, &g)) return 1;
      n->rng = rng; n->grid = g;
    }
    return 0;
  });
}

int futhark_entry_mouse(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t buttons, const int32_t x,
                        const int32_t y, const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [&](State *n) { return mouse(ctx, n, buttons, x, y); });
}

int futhark_entry_wheel(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t, const int32_t,
                        const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [](State *) { return 0; });                // :131 `case _ -> s`
}

int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0, const float td,
                       const struct futhark_opaque_state *s) {               // :77-86
  return transition(ctx, out0, s, [&](State *n) {
    if (!n->running) return 0;
    float unused = n->time_unused;
    const int64_t steps_new = steps_due(td, &unused, n->steps_per_second);
    if (steps_new > 0 && step_n(ctx, n, steps_new)) return 1;
    n->time_unused = unused;
    return 0;
  });
}

int futhark_entry_render(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_opaque_state *s) {
  if (!ctx || !out0 || !s) return 1;
  if (!ctx->fc) return raise(ctx, "no device context");
  if (coerce(ctx, s)) return 1;                                              // :133-135
  return render_window(ctx, s->grid, s->h, s->w, s->scale, out0);
}

int futhark_entry_text_format(struct futhark_context *ctx, struct futhark_u8_1d **out0) {   // :25
  if (!ctx || !out0) return 1;
  *out0 = new_string("FPS: %d\nSteps per second: %d\nLeaks: %d");
  return 0;
}

int futhark_entry_text_content(struct futhark_context *ctx, int32_t *out0, int32_t *out1, int32_t *out2, const float fps,
                               const struct futhark_opaque_state *s) {       // :27-28
  if (!ctx || !s || !out0 || !out1 || !out2) return 1;
  *out0 = (int32_t)fps; *out1 = s->steps_per_second; *out2 = s->n_leaks;
  return 0;
}

struct flowamok_grid *flowamok_lys_state_grid(const struct futhark_opaque_state *s) { return s ? s->grid->g : nullptr; }
int flowamok_lys_state_fields(const struct futhark_opaque_state *s, uint64_t *f, int cap, float *time_unused) {
  if (!s || !f || cap < 14) return -1;
  f[0] = (uint64_t)s->h; f[1] = (uint64_t)s->w; f[2] = (uint64_t)s->gh; f[3] = (uint64_t)s->gw; f[4] = (uint64_t)s->scale;
  f[5] = s->running; f[6] = (uint64_t)(int64_t)s->steps_per_second; f[7] = (uint64_t)s->n_steps_total;
  f[8] = (uint64_t)(int64_t)s->n_leaks; f[9] = s->cursor_some; f[10] = (uint64_t)s->cursor_y; f[11] = (uint64_t)s->cursor_x;
  f[12] = s->rng.state; f[13] = s->rng.inc;
  if (time_unused) *time_unused = s->time_unused;
  return 14;
}

}  // extern "C"
