// flowamok_b200 -- explorer/explorer.fut's `module lys` as the library a lys C main links (include/flowamok_lys.h).
//
// The state machine of explorer.fut:52-204 restated over the operator-level C ABI: fourteen 108 x 192 grids, one per
// scenario of explorer/scenarios.fut, the current one stepped by flowamok_explorer_step (ticks on the device, scenario
// hook and cycle recomputation inside the library).  Line numbers: /root/reference/explorer/explorer.fut.
#include <array>

#include "fa_lys.h"

using namespace falys;

namespace {
const int N_SCENARIOS = 14;            // :9-22, scenario.n_scenarios
const int64_t GH = 108, GW = 192;      // :48-49
const char *const NAMES[N_SCENARIOS] = {   // `name ()` of explorer/scenarios.fut:6,22,35,53,69,85,115,133,152,180,193,216,239,260
    "single fork", "tight queues", "basic cycle", "crossing", "close crossing", "crossroads", "small tight cycle",
    "independent tight cycles", "overlapping tight cycles", "hits an edge", "adding lines", "multi spill",
    "many cycles (n = 2)", "t cross"};
}  // namespace

struct futhark_opaque_state {          // :53-66
  int64_t h, w, scale;
  bool auto_;
  Rng rng;
  float time_unused;
  int32_t steps_auto_per_second;
  std::array<int64_t, N_SCENARIOS> steps;
  std::array<GridRef, N_SCENARIOS> grids;
  bool perfect_simulation;
  int64_t n_cycles;                    // length cycle_checks; the rings themselves live on the grid value they were found on
  int64_t n_cell_leaks;
  int64_t scenario_id;
};

#include "fa_lys_api.inc"

namespace {
typedef futhark_opaque_state State;

// init_grid gh gw rng scenario_id (:85-88)
int init_grid(futhark_context *c, const Rng &rng, int sid, Rng *rng_out, GridRef *out) {
  GridRef g;
  if (create_grid(c, GH, GW, rng, rng_out, &g)) return 1;
  if (flowamok_scenario_init(c->fc, g->g, sid)) return 1;
  *out = g;
  return 0;
}

// cycle_checks = find_cycles grids[scenario_id] wherever the reference recomputes it (:95-97, :156-158, :164-166, :172-175, :186-188)
int recompute_cycles(futhark_context *c, State *s) { return find_cycles(c, s->grids[(size_t)s->scenario_id], &s->n_cycles); }

// step s n_steps (:109-130)
int step_n(futhark_context *c, State *s, int64_t n_steps) {
  const size_t k = (size_t)s->scenario_id;
  if (s->perfect_simulation && !s->grids[k]->cycles_valid && recompute_cycles(c, s)) return 1;   // invariant; see fa_lys.h
  GridRef g;
  if (clone(c, s->grids[k], &g)) return 1;                                                       // :112, :114 copy
  int64_t step_cur = s->steps[k], leaks = 0;
  uint64_t st = s->rng.state, inc = s->rng.inc;
  if (flowamok_explorer_step(c->fc, g->g, (int)s->scenario_id, s->perfect_simulation ? 1 : 0, n_steps, &step_cur, &st, &inc, &leaks))
    return 1;
  // a scenario that edits roads asks for new cycle_checks; they are recomputed (inside the library) only in perfect mode (:123-125)
  g->cycles_valid = s->perfect_simulation;
  if (s->perfect_simulation && flowamok_cycle_count(c->fc, g->g, &s->n_cycles, nullptr)) return 1;
  s->rng = Rng{st, inc};                                                                         // :126
  s->steps[k] = step_cur;                                                                        // :127
  s->grids[k] = g;                                                                               // :128
  s->n_cell_leaks = leaks;                                                                       // :130
  return 0;
}

int floor_mod(int64_t a, int64_t n) { return (int)(((a % n) + n) % n); }   // Futhark's % rounds towards negative infinity

// event (#keydown {key}) (:145-198)
int keydown(futhark_context *c, State *s, int32_t key) {
  switch (key) {
    case K_SPACE:                                                           // :146-147
      if (step_n(c, s, 1)) return 1;
      s->auto_ = false;
      return 0;
    case K_LEFT:                                                            // :148-161
      s->scenario_id = floor_mod(s->scenario_id - 1, N_SCENARIOS);
      if (s->perfect_simulation && recompute_cycles(c, s)) return 1;
      s->n_cell_leaks = 0;
      return 0;
    case K_RIGHT:                                                           // :162-168 (n_cell_leaks stays)
      s->scenario_id = floor_mod(s->scenario_id + 1, N_SCENARIOS);
      if (s->perfect_simulation && recompute_cycles(c, s)) return 1;
      return 0;
    case K_a: s->auto_ = !s->auto_; return 0;                               // :169-170
    case K_p:                                                               // :171-175
      s->perfect_simulation = !s->perfect_simulation;
      return s->perfect_simulation ? recompute_cycles(c, s) : 0;
    case K_UP: s->steps_auto_per_second += 1; return 0;                     // :176-177
    case K_DOWN:                                                            // :178-179
      s->steps_auto_per_second = s->steps_auto_per_second - 1 > 1 ? s->steps_auto_per_second - 1 : 1;
      return 0;
    case K_r: {                                                             // :180-193
      Rng rng;
      GridRef g;
      if (init_grid(c, s->rng, (int)s->scenario_id, &rng, &g)) return 1;
      s->rng = rng;
      s->grids[(size_t)s->scenario_id] = g;
      if (s->perfect_simulation && recompute_cycles(c, s)) return 1;
      s->steps[(size_t)s->scenario_id] = 0;
      s->auto_ = false;
      return 0;
    }
    case K_1: s->scale += 1; return 0;                                      // :194-195
    case K_2: s->scale = s->scale - 1 > 1 ? s->scale - 1 : 1; return 0;     // :196-197
    default: return 0;                                                      // :198
  }
}

// every entry of the shape state -> state: copy the value, apply, hand the copy out
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
                       const int64_t w) {
  if (!ctx || !out0) return 1;
  if (!ctx->fc) return raise(ctx, "no device context");
  std::unique_ptr<State> s(new State());
  const Rng rng = rng_from_seed((int32_t)seed);                             // :90 rng_from_seed [i32.u32 seed]
  Rng outs[N_SCENARIOS];
  for (int k = 0; k < N_SCENARIOS; k++)                                     // :91-92 split_rng n_scenarios, map2 init_grid
    if (init_grid(ctx, split_one(rng, k), k, &outs[k], &s->grids[(size_t)k])) return 1;
  s->perfect_simulation = true;                                             // :93
  s->rng = join(outs, N_SCENARIOS);                                         // :94
  s->scenario_id = 0;                                                       // :95
  if (recompute_cycles(ctx, s.get())) return 1;                             // :96-98
  s->h = h; s->w = w; s->scale = 10; s->auto_ = false; s->time_unused = 0.0f;   // :99-101
  s->steps_auto_per_second = 30;
  s->steps.fill(0);
  s->n_cell_leaks = 0;
  *out0 = s.release();
  return 0;
}

int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int64_t h, const int64_t w,
                         const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [&](State *n) { n->h = h; n->w = w; return 0; });   // :105-107
}

int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t e, const int32_t key,
                      const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [&](State *n) { return e == 0 ? keydown(ctx, n, key) : 0; });   // #keyup: `case _ -> s`
}

int futhark_entry_mouse(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t, const int32_t, const int32_t,
                        const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [](State *) { return 0; });               // :199-200
}

int futhark_entry_wheel(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t, const int32_t,
                        const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [](State *) { return 0; });               // :199-200
}

int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0, const float td,
                       const struct futhark_opaque_state *s) {
  return transition(ctx, out0, s, [&](State *n) {                           // :134-144
    if (!n->auto_) return 0;
    float unused = n->time_unused;
    const int64_t steps_new = steps_due(td, &unused, n->steps_auto_per_second);
    if (steps_new > 0 && step_n(ctx, n, steps_new)) return 1;
    n->time_unused = unused;
    return 0;
  });
}

int futhark_entry_render(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_opaque_state *s) {
  if (!ctx || !out0 || !s) return 1;
  if (!ctx->fc) return raise(ctx, "no device context");
  return render_window(ctx, s->grids[(size_t)s->scenario_id], s->h, s->w, s->scale, out0);   // :202-203
}

int futhark_entry_text_format(struct futhark_context *ctx, struct futhark_u8_1d **out0) {     // :70-75
  if (!ctx || !out0) return 1;
  std::string names = NAMES[0];
  for (int i = 1; i < N_SCENARIOS; i++) names += std::string("|") + NAMES[i];
  *out0 = new_string("FPS: %d\nScenario: %[" + names +
                     "]\nAuto: %[no|yes]\nSteps per second: %d\nSteps: %ld\nPerfect: %[false|true]\nCycles detected: %ld\n"
                     "Cell leaks this step: %ld");
  return 0;
}

int futhark_entry_text_content(struct futhark_context *ctx, int32_t *out0, int32_t *out1, int32_t *out2, int32_t *out3,
                               int64_t *out4, int32_t *out5, int64_t *out6, int64_t *out7, const float fps,
                               const struct futhark_opaque_state *s) {                        // :77-81
  if (!ctx || !s || !out0 || !out1 || !out2 || !out3 || !out4 || !out5 || !out6 || !out7) return 1;
  *out0 = (int32_t)fps;                      // t32
  *out1 = (int32_t)s->scenario_id;
  *out2 = s->auto_ ? 1 : 0;
  *out3 = s->steps_auto_per_second;
  *out4 = s->steps[(size_t)s->scenario_id];
  *out5 = s->perfect_simulation ? 1 : 0;
  *out6 = s->n_cycles;
  *out7 = s->n_cell_leaks;
  return 0;
}

struct flowamok_grid *flowamok_lys_state_grid(const struct futhark_opaque_state *s) {
  return s ? s->grids[(size_t)s->scenario_id]->g : nullptr;
}
int flowamok_lys_state_fields(const struct futhark_opaque_state *s, uint64_t *f, int cap, float *time_unused) {
  if (!s || !f || cap < 12) return -1;
  f[0] = (uint64_t)s->h; f[1] = (uint64_t)s->w; f[2] = (uint64_t)s->scale; f[3] = s->auto_;
  f[4] = (uint64_t)(int64_t)s->steps_auto_per_second; f[5] = (uint64_t)s->scenario_id;
  f[6] = (uint64_t)s->steps[(size_t)s->scenario_id]; f[7] = s->perfect_simulation; f[8] = (uint64_t)s->n_cycles;
  f[9] = (uint64_t)s->n_cell_leaks; f[10] = s->rng.state; f[11] = s->rng.inc;
  if (time_unused) *time_unused = s->time_unused;
  return 12;
}

}  // extern "C"
