/*
 * flowamok CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's per-tick grid update and of the few
 * helpers needed to build fixtures.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (flowamok_b200/csrc) never includes, links or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  The reference is pure Futhark and cannot be compiled in
 * this image (no futhark, no GHC), so there is no oracle/_ref build.
 *
 * Parity status:
 *   - movement semantics (U/R/M/Z, find_cycles, scenario 8): PINNED by the
 *     README GIF (400 frames, tests/golden/readme_gif_classes.npz).
 *   - RNG arithmetic (diku-dk/cpprandom 1.3.1, not vendored in the reference
 *     tree): "parity unpinned" -- restated from the library's published
 *     algorithm (PCG32 XSH-RR 64/32 + C++-style uniform_int_distribution); the
 *     398-bit known answer recovered from the GIF does NOT reproduce, see
 *     tests/test_oracle_rng.py (xfail) and DESIGN.md.
 *   - colour constants (athas/matte, not vendored): payload only, unpinned.
 *
 * Layout mirrors what the Futhark compiler gives `[gh][gw](cell aux)`: a
 * struct of arrays, one array per record field (~28 B/cell).
 */
#ifndef FLOWAMOK_ORACLE_H
#define FLOWAMOK_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* src/random.fut:4,7  rnge = pcg32, rng = {state:u64, inc:u64} (cpprandom) */
typedef struct { uint64_t state, inc; } fo_rng;

/* src/model.fut:5-24, struct-of-arrays */
typedef struct {
  int64_t gh, gw;
  uint64_t *rng_state, *rng_inc;   /* underlying.rng              */
  int8_t *uy, *ux;                 /* underlying.direction.{y,x}  */
  int8_t *dy, *dx;                 /* direction.{y,x}             */
  uint32_t *color;                 /* color (argb)                */
  uint8_t *calc, *my, *mx, *pref;  /* can_be_moved_to_from.{calculated,direction.y,direction.x,next_preference_flow_x} */
} fo_grid;

/* src/model.fut:26  cell_leak = (i64, i64, cell aux); the cell is the SOURCE cell before the move */
typedef struct {
  int64_t y_next, x_next;
  int64_t y_src, x_src;            /* not in the reference tuple; position of the copied cell, for tests */
  fo_rng rng; int8_t uy, ux, dy, dx; uint32_t color; uint8_t calc, my, mx, pref;
} fo_leak;

/* choose_direction is a parameter of both steppers (src/steppers.fut:180,305). */
enum { FO_CHOOSE_RANDOM = 0,   /* flowamok.fut:10-14 choose_direction_random */
       FO_CHOOSE_Y = 1,        /* always {y=u.y,x=0} */
       FO_CHOOSE_X = 2,        /* always {y=0,x=u.x} */
       FO_CHOOSE_TAPE = 3 };   /* choice = tape[tick % tape_len] for every draw of that tick */
typedef struct {
  int mode;
  const uint8_t *tape; int64_t tape_len;
  int64_t tick;                /* advanced by the caller once per step */
} fo_chooser;

/* ---- cpprandom restatement (EXT-UNVERIFIED) ---- */
uint32_t fo_pcg32_rand(fo_rng *r);
fo_rng   fo_rng_from_seed(const int32_t *xs, int64_t n);
void     fo_split_rng(int64_t n, fo_rng r, fo_rng *out);
fo_rng   fo_join_rng(const fo_rng *rs, int64_t n);
int8_t   fo_dist_i8_rand(int8_t lo, int8_t hi, fo_rng *r);
float    fo_dist_f32_rand(float lo, float hi, fo_rng *r);
uint32_t fo_random_color(fo_rng *r);                       /* src/random.fut:9-13 */

/* ---- grid construction: src/utils.fut ---- */
fo_grid *fo_grid_alloc(int64_t gh, int64_t gw);
void     fo_grid_free(fo_grid *g);
fo_grid *fo_grid_clone(const fo_grid *g);
fo_rng   fo_create_grid(fo_grid *g, fo_rng rng);           /* utils.fut:14-17 (fills g, returns join_rng) */
void     fo_add_line_vertical(fo_grid *g, int64_t y0, int64_t y1, int64_t x, int8_t flow);   /* utils.fut:19-25 */
void     fo_add_line_horizontal(fo_grid *g, int64_t y, int64_t x0, int64_t x1, int8_t flow); /* utils.fut:27-33 */
void     fo_add_cell(fo_grid *g, int64_t y, int64_t x, uint32_t color, int8_t dy, int8_t dx); /* utils.fut:35-38 */

/* ---- steppers: src/steppers.fut ---- */
void    fo_update_pass(fo_grid *g);                        /* :9-56 one Jacobi pass */
int64_t fo_move_until_fixpoint(fo_grid *g);                /* :147-164, returns the number of passes */
/* dense masks [n][gh][gw] of (y,x) int8 pairs, interleaved: mask[((k*gh+y)*gw+x)*2 + {0:y,1:x}] */
int     fo_resolve_cycles(fo_grid *g, const int8_t *masks, int64_t n);  /* :189-224; returns 0, or -1 on the OOB index of :207 */
/* returns number of leaks; *leaks (malloc'd, row-major order of the source cell) may be NULL to skip the list */
int64_t fo_move_and_reset_cells(fo_grid *g, fo_chooser *ch, fo_leak **leaks); /* :166-173 (58-145) */
int64_t fo_step_quick(fo_grid *g, fo_chooser *ch, fo_leak **leaks);           /* :177-185 */
int64_t fo_step_perfect(fo_grid *g, fo_chooser *ch, const int8_t *masks, int64_t n, fo_leak **leaks); /* :302-312 */
/* :233-300 literal worklist; returns n_cycles and malloc'd dense masks; max_paths bounds the worklist (returns -1 if exceeded) */
int64_t fo_find_cycles(const fo_grid *g, int8_t **masks_out, int64_t max_paths);

/* sparse restatement of resolve_cycles for grids where dense masks do not fit
 * (SURVEY 8a-R disjointness): ring k = cells [off[k], off[k+1]) of flat index
 * idx[], check direction (cy,cx) and the static grant (1 = my, 2 = mx).
 * Validated against fo_resolve_cycles in tests/test_oracle_cycles.py. */
void    fo_resolve_cycles_sparse(fo_grid *g, int64_t n, const int64_t *off, const int64_t *idx,
                                 const int8_t *cy, const int8_t *cx, const uint8_t *grant);
/* dense masks -> sparse lists (malloc'd); returns total cells */
int64_t fo_cycles_dense_to_sparse(const fo_grid *g, const int8_t *masks, int64_t n, int64_t **off,
                                  int64_t **idx, int8_t **cy, int8_t **cx, uint8_t **grant);
int64_t fo_step_perfect_sparse(fo_grid *g, fo_chooser *ch, int64_t n, const int64_t *off, const int64_t *idx,
                               const int8_t *cy, const int8_t *cx, const uint8_t *grant, fo_leak **leaks);

/* ---- scenarios: explorer/scenarios.fut, ids as explorer/explorer.fut:9-22 ---- */
#define FO_N_SCENARIOS 14
const char *fo_scenario_name(int sid);
void fo_scenario_init(int sid, fo_grid *g);
/* returns recompute_cycles; rng is the scenario-level rng (model.fut:33) */
int  fo_scenario_step(int sid, fo_grid *g, int64_t steps, fo_rng *rng);

/* ---- render: flowamok.fut:16-64 ---- */
void fo_render(const fo_grid *g, int64_t h, int64_t w, int64_t scale, uint32_t *pixels);

/* ---- synthetic road network of SURVEY.md App. E (the reference has no generator; this restates the
 * rule documented in DESIGN.md independently of the device generator) ----
 * fills g; returns the number of rings and malloc'd sparse ring arrays (as fo_cycles_dense_to_sparse). */
int64_t fo_generate_synthetic(fo_grid *g, uint64_t seed, uint32_t density16, uint32_t ring_full16,
                              int64_t **off, int64_t **idx, int8_t **cy, int8_t **cx, uint8_t **grant);

/* ---- misc ---- */
void fo_set_threads(int n);    /* OpenMP threads for the map-shaped passes (cpu baseline) */
int  fo_get_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
