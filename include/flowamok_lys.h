/*
 * flowamok_b200 -- the C API of the two lys frontends (explorer/explorer.fut, designer/designer.fut), re-implemented
 * as host code over the operator-level C ABI (include/flowamok.h), so that the lys C `main` of either program links
 * against libflowamok_explorer.so / libflowamok_designer.so instead of the library `futhark <backend> --library`
 * generates from explorer_wrapper.fut / designer_wrapper.fut (explorer/Makefile, designer/Makefile: lys' common.mk).
 *
 * Both programs instantiate `module lys: lys with text_content = ...` (explorer.fut:52, designer.fut:8); the wrapper
 * that diku-dk/lys 0.1.24 (futhark.pkg:2, NOT vendored in the reference) generates turns that module into the entry
 * points below: init, resize, key, mouse, wheel, step, render, text_format, text_content, text_colour, grab_mouse.
 * The state machines behind them are pinned by explorer.fut / designer.fut themselves and restated line by line
 * (flowamok_b200/lys/); the NAMES AND ARGUMENT ORDER of the entry points follow lys' published wrapper and Futhark's C
 * API conventions and are [EXT-UNVERIFIED against lys 0.1.24: not in the image].  What each entry does to the state
 * is cited per function.
 *
 * Both libraries export the same symbol names (like the two generated libraries do); a program links ONE of them.
 * Compile the caller with -DFLOWAMOK_LYS_DESIGNER for the designer's text_content.
 *
 * Values are immutable like Futhark's: every entry returns a FRESH state and leaves its input valid; grids are shared
 * between states and cloned only by the entry that steps or edits them, so the lys loop (new state, free the old one)
 * allocates nothing per frame.  Errors: non-zero return, message from futhark_context_get_error (malloc'd).  Size
 * mismatches and out-of-bounds edits, which Futhark reports as run-time errors, are reported the same way.
 */
#ifndef FLOWAMOK_LYS_H
#define FLOWAMOK_LYS_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

struct futhark_context_config;
struct futhark_context;
struct futhark_opaque_state;   /* explorer.fut:53-66 / designer.fut:9-21 */
struct futhark_u32_2d;         /* [h][w]argb.colour */
struct futhark_u8_1d;          /* string */

struct futhark_context_config *futhark_context_config_new(void);
void futhark_context_config_free(struct futhark_context_config *cfg);
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s);   /* "#k" */
struct futhark_context *futhark_context_new(struct futhark_context_config *cfg);
void futhark_context_free(struct futhark_context *ctx);
int futhark_context_sync(struct futhark_context *ctx);
char *futhark_context_get_error(struct futhark_context *ctx);
/* the rest of the context API of a generated library (accepted; there is nothing to tune, cache or profile here) */
int futhark_context_config_set_tuning_param(struct futhark_context_config *cfg, const char *param_name, size_t new_value);
int futhark_get_tuning_param_count(void);
const char *futhark_get_tuning_param_name(int i);
const char *futhark_get_tuning_param_class(int i);
void futhark_context_config_set_cache_file(struct futhark_context_config *cfg, const char *f);
void futhark_context_set_logging_file(struct futhark_context *ctx, FILE *f);
void futhark_context_pause_profiling(struct futhark_context *ctx);
void futhark_context_unpause_profiling(struct futhark_context *ctx);
int futhark_context_clear_caches(struct futhark_context *ctx);
char *futhark_context_report(struct futhark_context *ctx);   /* malloc'd */

/* explorer.fut:89-101 / designer.fut:36-45   init (seed: u32) (h: i64) (w: i64) */
int futhark_entry_init(struct futhark_context *ctx, struct futhark_opaque_state **out0, const uint32_t seed, const int64_t h,
                       const int64_t w);
/* explorer.fut:105-107 / designer.fut:49-55 */
int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int64_t h, const int64_t w,
                         const struct futhark_opaque_state *s);
/* event (#keydown {key}) when e == 0, (#keyup {key}) otherwise: explorer.fut:145-198 / designer.fut:88-95.  SDL key codes. */
int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t e, const int32_t key,
                      const struct futhark_opaque_state *s);
/* event (#mouse {buttons, x, y}): designer.fut:112-125 (the explorer ignores it, explorer.fut:199-200) */
int futhark_entry_mouse(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t buttons, const int32_t x,
                        const int32_t y, const struct futhark_opaque_state *s);
/* event (#wheel {dx, dy}): ignored by both programs */
int futhark_entry_wheel(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t dx, const int32_t dy,
                        const struct futhark_opaque_state *s);
/* event (#step td): explorer.fut:134-144 (then :109-130) / designer.fut:77-86 (then :59-75) */
int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0, const float td,
                       const struct futhark_opaque_state *s);
/* explorer.fut:202-203 / designer.fut:133-135: flowamok.fut:16-64 `render s.h s.w gh gw s.scale grid` */
int futhark_entry_render(struct futhark_context *ctx, struct futhark_u32_2d **out0, const struct futhark_opaque_state *s);
/* explorer.fut:70-75 / designer.fut:25 */
int futhark_entry_text_format(struct futhark_context *ctx, struct futhark_u8_1d **out0);
#ifdef FLOWAMOK_LYS_DESIGNER
/* designer.fut:27-28: (fps, steps per second, leaks) */
int futhark_entry_text_content(struct futhark_context *ctx, int32_t *out0, int32_t *out1, int32_t *out2, const float fps,
                               const struct futhark_opaque_state *s);
#else
/* explorer.fut:77-81: (fps, scenario id, auto, steps per second, steps of this scenario, perfect, cycles detected, cell
 * leaks of the last step) */
int futhark_entry_text_content(struct futhark_context *ctx, int32_t *out0, int32_t *out1, int32_t *out2, int32_t *out3,
                               int64_t *out4, int32_t *out5, int64_t *out6, int64_t *out7, const float fps,
                               const struct futhark_opaque_state *s);
#endif
/* explorer.fut:83 / designer.fut:30: argb.white */
int futhark_entry_text_colour(struct futhark_context *ctx, uint32_t *out0, const struct futhark_opaque_state *s);
/* explorer.fut:103 / designer.fut:47: false */
int futhark_entry_grab_mouse(struct futhark_context *ctx, bool *out0);

int futhark_free_opaque_state(struct futhark_context *ctx, struct futhark_opaque_state *obj);
int futhark_free_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr);
int futhark_values_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr, uint32_t *data);
const int64_t *futhark_shape_u32_2d(struct futhark_context *ctx, struct futhark_u32_2d *arr);
int futhark_free_u8_1d(struct futhark_context *ctx, struct futhark_u8_1d *arr);
int futhark_values_u8_1d(struct futhark_context *ctx, struct futhark_u8_1d *arr, uint8_t *data);
const int64_t *futhark_shape_u8_1d(struct futhark_context *ctx, struct futhark_u8_1d *arr);

/* Extensions (not generated by Futhark), for tests and diagnostics: the operator-level context behind a lys context, the
 * grid a state shows (explorer: grids[scenario_id]; borrowed, valid while the state lives), and the scalar fields of a
 * state.  flowamok_lys_state_fields writes, explorer: {h, w, scale, auto, steps_auto_per_second, scenario_id,
 * steps[scenario_id], perfect_simulation, length cycle_checks, n_cell_leaks, rng.state, rng.inc};  designer: {h, w, gh, gw,
 * scale, running, steps_per_second, n_steps_total, n_leaks, cursor_prev is #some, cursor_prev.y, cursor_prev.x, rng.state,
 * rng.inc}; *time_unused receives the f32 field.  Returns the number of fields written. */
struct flowamok_context;
struct flowamok_grid;
struct flowamok_context *flowamok_lys_context(struct futhark_context *ctx);
struct flowamok_grid *flowamok_lys_state_grid(const struct futhark_opaque_state *s);
int flowamok_lys_state_fields(const struct futhark_opaque_state *s, uint64_t *fields, int cap, float *time_unused);

#ifdef __cplusplus
}
#endif
#endif
