// flowamok_b200 -- what the two lys frontends share: the Futhark-style context and array values, value-semantic grid
// handles over the operator-level C ABI, the frontends' rng plumbing and the window render.
//
// Host code only (g++): everything that touches the device goes through include/flowamok.h.  There is no CPU path
// here either -- a context without a CUDA device fails in flowamok_context_new and every entry reports it.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/flowamok.h"
#include "../../include/flowamok_lys.h"
#include "../csrc/fa_bits.h"   // the pcg32 restatement (diku-dk/cpprandom 1.3.1 is not vendored: unpinned)

struct futhark_context_config { int device = -1; int debugging = 0, profiling = 0, logging = 0; };
struct futhark_context {
  flowamok_context *fc = nullptr;
  std::string err;   // errors raised by this layer; device-side errors stay in fc
};
struct futhark_u32_2d { std::vector<uint32_t> v; int64_t shape[2]; };
struct futhark_u8_1d { std::vector<uint8_t> v; int64_t shape[1]; };

namespace falys {
using fa::u32;
using fa::u64;

// SDL2 key codes (SDL_keycode.h), as lys hands them to `event`
enum : int32_t {
  K_SPACE = 32, K_1 = '1', K_2 = '2', K_a = 'a', K_p = 'p', K_r = 'r',
  K_RIGHT = 0x4000004F, K_LEFT = 0x40000050, K_DOWN = 0x40000051, K_UP = 0x40000052
};
const u32 ARGB_WHITE = 0xFFFFFFFFu, ARGB_BLACK = 0xFF000000u;

inline int raise(futhark_context *c, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return 1;
}

// rnge = pcg32 (src/random.fut:4)
struct Rng { u64 state = 0, inc = 0; };
inline Rng rng_from_seed(int32_t seed) {                 // rnge.rng_from_seed [seed]
  Rng r;
  fa::pcg32_from_seed(seed, r.state, r.inc);
  return r;
}
inline Rng split_one(const Rng &r, int64_t i) { return {fa::pcg32_split_state(r.state, r.inc, i), r.inc}; }   // (split_rng n r)[i]
inline Rng join(const Rng *rs, int64_t n) {               // rnge.join_rng: product of the states, OR of the incs
  Rng j{1, 0};
  for (int64_t i = 0; i < n; i++) { j.state *= rs[i].state; j.inc |= rs[i].inc; }
  return j;
}

// A grid VALUE: immutable once it is inside a state.  The planes go back to the library's pool with the last state
// that shows them (flowamok_grid_recycle), so `copy grids[sid]` of the next frame allocates nothing.
struct GridVal {
  flowamok_context *fc;
  flowamok_grid *g;
  bool cycles_valid = false;   // the ring list installed on g is find_cycles of g's own roads
  GridVal(flowamok_context *c, flowamok_grid *p) : fc(c), g(p) {}
  GridVal(const GridVal &) = delete;
  GridVal &operator=(const GridVal &) = delete;
  ~GridVal() { if (g) flowamok_grid_recycle(fc, g); }
};
typedef std::shared_ptr<GridVal> GridRef;

// create_grid gh gw () rng (src/utils.fut:14-17): cell i gets (split_rng (gh*gw) rng)[i]; the rng handed back is the join
inline int create_grid(futhark_context *c, int64_t gh, int64_t gw, const Rng &rng, Rng *joined, GridRef *out) {
  if (gh < 0 || gw < 0) return raise(c, "create_grid: negative size %lld x %lld", (long long)gh, (long long)gw);
  const int64_t n = gh * gw;
  std::vector<uint64_t> st((size_t)n);
  u64 prod = 1;
  for (int64_t i = 0; i < n; i++) { st[(size_t)i] = fa::pcg32_split_state(rng.state, rng.inc, i); prod *= st[(size_t)i]; }
  flowamok_grid *g = nullptr;
  if (flowamok_grid_new(c->fc, &g, gh, gw)) return 1;
  if (flowamok_grid_upload(c->fc, g, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st.data(), rng.inc)) {
    flowamok_grid_free(c->fc, g);
    return 1;
  }
  *joined = Rng{prod, n > 0 ? rng.inc : 0};
  *out = std::make_shared<GridVal>(c->fc, g);
  return 0;
}
inline int clone(futhark_context *c, const GridRef &src, GridRef *out) {   // `copy grid`
  flowamok_grid *g = nullptr;
  if (flowamok_grid_clone(c->fc, src->g, &g)) return 1;
  *out = std::make_shared<GridVal>(c->fc, g);
  (*out)->cycles_valid = src->cycles_valid;
  return 0;
}
// stepper_perfect.find_cycles grid (steppers.fut:233-300); a grid value's roads never change, so once is enough
inline int find_cycles(futhark_context *c, const GridRef &gv, int64_t *n) {
  if (!gv->cycles_valid) {
    if (flowamok_find_cycles(c->fc, gv->g, 0, nullptr)) return 1;
    gv->cycles_valid = true;
  }
  return flowamok_cycle_count(c->fc, gv->g, n, nullptr);
}

// The `#step td` arithmetic both programs share (explorer.fut:135-143, designer.fut:79-85), in f32 like the reference
inline int64_t steps_due(float td, float *time_unused, int32_t per_second) {
  volatile float time_total = td + *time_unused;
  volatile float steps_new = time_total * (float)per_second;
  const int64_t n = (int64_t)steps_new;                          // i64.f32 truncates
  volatile float q = (float)n / (float)per_second;
  *time_unused = time_total - q;
  return n;
}

// render h w gh gw scale grid (flowamok.fut:16-64): pixel (y, x) shows cell (y / scale, x / scale) when that is inside the
// grid, black otherwise -- i.e. the [gh*scale][gw*scale] image of flowamok_render cut or padded to the window.
inline int render_window(futhark_context *c, const GridRef &gv, int64_t h, int64_t w, int64_t scale, futhark_u32_2d **out) {
  int64_t gh = 0, gw = 0;
  if (flowamok_grid_shape(c->fc, gv->g, &gh, &gw)) return 1;
  if (scale < 1) return raise(c, "render: division by zero (scale %lld)", (long long)scale);
  if (h < 0 || w < 0) return raise(c, "render: negative window %lld x %lld", (long long)h, (long long)w);
  const int64_t ih = gh * scale, iw = gw * scale;
  std::vector<uint32_t> img((size_t)(ih * iw));
  if (ih * iw > 0 && flowamok_render(c->fc, gv->g, scale, img.data())) return 1;
  std::unique_ptr<futhark_u32_2d> a(new futhark_u32_2d());
  a->shape[0] = h; a->shape[1] = w;
  a->v.assign((size_t)(h * w), ARGB_BLACK);
  const int64_t ch = h < ih ? h : ih, cw = w < iw ? w : iw;
  for (int64_t y = 0; y < ch; y++) memcpy(&a->v[(size_t)(y * w)], &img[(size_t)(y * iw)], (size_t)cw * 4);
  *out = a.release();
  return 0;
}

inline futhark_u8_1d *new_string(const std::string &s) {
  futhark_u8_1d *a = new futhark_u8_1d();
  a->v.assign(s.begin(), s.end());
  a->shape[0] = (int64_t)s.size();
  return a;
}
}  // namespace falys
