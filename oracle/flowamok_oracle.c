/*
 * flowamok CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see flowamok_oracle.h).
 *
 * Literal restatement of /root/reference/src/{steppers,model,helpers,utils,random,option,reduce1}.fut,
 * flowamok.fut, explorer/scenarios.fut.  Written from the behaviour of those files; nothing is
 * copied (the reference is Futhark, this is C).  Citations are file:line of the reference.
 *
 * Pinned by: the README GIF (tests/golden/readme_gif_classes.npz), the scenario tables evaluated from
 * explorer/scenarios.fut (tests/golden/scenario_inits.npz) and -- cell for cell, tick for tick, leak lists and
 * find_cycles included -- vectors made by executing src/steppers.fut itself under a Futhark-subset interpreter
 * (tests/golden/stepper_vectors.npz, tests/test_stepper_vectors.py).  NOT pinned: the RNG arithmetic below.
 */
#include "flowamok_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * RNG: diku-dk/cpprandom 1.3.1 (futhark.pkg:3), NOT in the reference tree.  Restated from the
 * library's published algorithm as recalled.  "parity unpinned".
 * ---------------------------------------------------------------------------------------- */

/* pcg32.rand: PCG-XSH-RR 64/32, output from the old state */
uint32_t fo_pcg32_rand(fo_rng *r) {
  uint64_t old = r->state;
  r->state = old * 6364136223846793005ULL + (r->inc | 1ULL);
  uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  return (xorshifted >> rot) | (xorshifted << ((-rot) & 31u));
}

/* pcg32.rng_from_seed (call sites: animation.fut:15, explorer.fut:90) */
fo_rng fo_rng_from_seed(const int32_t *xs, int64_t n) {
  fo_rng r;
  r.state = 0;
  r.inc = (0xda3e39cb94b95bdbULL << 1) | 1ULL;
  (void)fo_pcg32_rand(&r);
  for (int64_t i = 0; i < n; i++) r.state += (uint64_t)(int64_t)xs[i]; /* u64.i32 sign-extends */
  (void)fo_pcg32_rand(&r);
  return r;
}

static uint32_t fo_hash_u32(uint32_t x) {
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  x = (x >> 16) ^ x;
  return x;
}

/* pcg32.split_rng (call sites: utils.fut:15, explorer.fut:91): per-index state mix, one advance, inc shared */
void fo_split_rng(int64_t n, fo_rng r, fo_rng *out) {
  for (int64_t i = 0; i < n; i++) {
    fo_rng c;
    c.state = r.state ^ (uint64_t)(int64_t)(int32_t)fo_hash_u32((uint32_t)(int32_t)i);
    c.inc = r.inc;
    (void)fo_pcg32_rand(&c);
    out[i] = c;
  }
}

/* pcg32.join_rng (call sites: utils.fut:17, explorer.fut:94) */
fo_rng fo_join_rng(const fo_rng *rs, int64_t n) {
  fo_rng r = {1ULL, 0ULL};
  for (int64_t i = 0; i < n; i++) { r.state *= rs[i].state; r.inc |= rs[i].inc; }
  return r;
}

/* uniform_int_distribution i8 pcg32 .rand (random.fut:5; call sites flowamok.fut:11, scenarios.fut:95-98,208) */
int8_t fo_dist_i8_rand(int8_t lo, int8_t hi, fo_rng *r) {
  uint32_t emin = (uint32_t)(int64_t)lo, emax = (uint32_t)(int64_t)hi;
  uint32_t range = emax - emin + 1u;
  if ((int32_t)range <= 0) return 0; /* degenerate: library returns E.min */
  uint32_t secure_max = 0xFFFFFFFFu - 0xFFFFFFFFu % range;
  uint32_t x = fo_pcg32_rand(r);
  while (x >= secure_max) x = fo_pcg32_rand(r);
  return (int8_t)(emin + x / (secure_max / range));
}

/* uniform_real_distribution f32 pcg32 .rand (random.fut:6) */
float fo_dist_f32_rand(float lo, float hi, fo_rng *r) {
  uint32_t x = fo_pcg32_rand(r);
  float xf = ((float)(uint64_t)x - 0.0f) / ((float)(uint64_t)0xFFFFFFFFu - 0.0f);
  return lo + xf * (hi - lo);
}

/* athas/matte argb.from_rgba (not vendored): clamp to [0,1], scale by 255, truncate */
static uint32_t fo_from_rgba(float r, float g, float b, float a) {
#define CH(v) ((uint32_t)(fminf(1.0f, fmaxf(0.0f, (v))) * 255.0f))
  return (CH(a) << 24) | (CH(r) << 16) | (CH(g) << 8) | CH(b);
#undef CH
}

/* src/random.fut:9-13 */
uint32_t fo_random_color(fo_rng *r) {
  float cr = fo_dist_f32_rand(0.2f, 0.8f, r);
  float cg = fo_dist_f32_rand(0.2f, 0.8f, r);
  float cb = fo_dist_f32_rand(0.2f, 0.8f, r);
  return fo_from_rgba(cr, cg, cb, 1.0f);
}

/* matte named colours used by explorer/scenarios.fut (payload only; unpinned) */
#define FO_WHITE   0xFFFFFFFFu
#define FO_BLACK   0xFF000000u
#define FO_RED     0xFFFF0000u
#define FO_GREEN   0xFF00FF00u
#define FO_BLUE    0xFF0000FFu
#define FO_YELLOW  0xFFFFFF00u
#define FO_ORANGE  0xFFFFFF00u /* matte: orange = add yellow red, saturates to yellow (GIF: (246,247,6)) */
#define FO_MAGENTA 0xFFFF00FFu
#define FO_VIOLET  0xFFFF00FFu /* matte: violet = add magenta blue, saturates to magenta (GIF) */
#define FO_BROWN   0xFF7C301Cu /* matte: from_rgba 0.49 0.19 0.11 1 (GIF mean ~ (117,49,26)) */

/* ------------------------------------------------------------------------------------------
 * grid helpers
 * ---------------------------------------------------------------------------------------- */
void fo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int fo_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

fo_grid *fo_grid_alloc(int64_t gh, int64_t gw) {
  fo_grid *g = (fo_grid *)calloc(1, sizeof(fo_grid));
  size_t n = (size_t)(gh * gw);
  if (n == 0) n = 1;
  g->gh = gh; g->gw = gw;
  g->rng_state = (uint64_t *)calloc(n, 8); g->rng_inc = (uint64_t *)calloc(n, 8);
  g->uy = (int8_t *)calloc(n, 1); g->ux = (int8_t *)calloc(n, 1);
  g->dy = (int8_t *)calloc(n, 1); g->dx = (int8_t *)calloc(n, 1);
  g->color = (uint32_t *)calloc(n, 4);
  g->calc = (uint8_t *)calloc(n, 1); g->my = (uint8_t *)calloc(n, 1);
  g->mx = (uint8_t *)calloc(n, 1); g->pref = (uint8_t *)calloc(n, 1);
  return g;
}
void fo_grid_free(fo_grid *g) {
  if (!g) return;
  free(g->rng_state); free(g->rng_inc); free(g->uy); free(g->ux); free(g->dy); free(g->dx);
  free(g->color); free(g->calc); free(g->my); free(g->mx); free(g->pref); free(g);
}
fo_grid *fo_grid_clone(const fo_grid *s) {
  fo_grid *g = fo_grid_alloc(s->gh, s->gw);
  size_t n = (size_t)(s->gh * s->gw);
  memcpy(g->rng_state, s->rng_state, n * 8); memcpy(g->rng_inc, s->rng_inc, n * 8);
  memcpy(g->uy, s->uy, n); memcpy(g->ux, s->ux, n); memcpy(g->dy, s->dy, n); memcpy(g->dx, s->dx, n);
  memcpy(g->color, s->color, n * 4);
  memcpy(g->calc, s->calc, n); memcpy(g->my, s->my, n); memcpy(g->mx, s->mx, n); memcpy(g->pref, s->pref, n);
  return g;
}

/* helpers.fut:27-28 */
static inline int in_bounds(int64_t h, int64_t w, int64_t y, int64_t x) { return y >= 0 && y < h && x >= 0 && x < w; }

/* utils.fut:14-17 (create_cell :7-12: white, empty direction, no road, cmtf all false) */
fo_rng fo_create_grid(fo_grid *g, fo_rng rng) {
  int64_t n = g->gh * g->gw;
  fo_rng *rngs = (fo_rng *)malloc(sizeof(fo_rng) * (size_t)(n > 0 ? n : 1));
  fo_split_rng(n, rng, rngs);
  for (int64_t i = 0; i < n; i++) {
    g->rng_state[i] = rngs[i].state; g->rng_inc[i] = rngs[i].inc;
    g->uy[i] = g->ux[i] = g->dy[i] = g->dx[i] = 0;
    g->color[i] = FO_WHITE;
    g->calc[i] = g->my[i] = g->mx[i] = g->pref[i] = 0;
  }
  fo_rng joined = fo_join_rng(rngs, n);
  free(rngs);
  return joined;
}

/* utils.fut:19-25 */
void fo_add_line_vertical(fo_grid *g, int64_t y0, int64_t y1, int64_t x, int8_t flow) {
  for (int64_t y = y0; y <= y1; y++) g->uy[y * g->gw + x] = flow;
}
/* utils.fut:27-33 */
void fo_add_line_horizontal(fo_grid *g, int64_t y, int64_t x0, int64_t x1, int8_t flow) {
  for (int64_t x = x0; x <= x1; x++) g->ux[y * g->gw + x] = flow;
}
/* utils.fut:35-38 */
void fo_add_cell(fo_grid *g, int64_t y, int64_t x, uint32_t color, int8_t dy, int8_t dx) {
  int64_t i = y * g->gw + x;
  g->color[i] = color; g->dy[i] = dy; g->dx[i] = dx;
}

/* ------------------------------------------------------------------------------------------
 * steppers.fut:9-56  update_can_be_moved_to_from -- ONE Jacobi pass (tabulate_2d reads the old grid)
 * ---------------------------------------------------------------------------------------- */
void fo_update_pass(fo_grid *g) {
  const int64_t gh = g->gh, gw = g->gw;
  size_t n = (size_t)(gh * gw);
  uint8_t *ncalc = (uint8_t *)malloc(n ? n : 1), *nmy = (uint8_t *)malloc(n ? n : 1);
  uint8_t *nmx = (uint8_t *)malloc(n ? n : 1), *npref = (uint8_t *)malloc(n ? n : 1);
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < gh; y++) {
    for (int64_t x = 0; x < gw; x++) {
      int64_t i = y * gw + x;
      uint8_t calc = g->calc[i], my = g->my[i], mx = g->mx[i], pref = g->pref[i];
      if (!calc) {                                                           /* :15-16 */
        int ready, base;
        if (g->dy[i] != 0 || g->dx[i] != 0) {                                /* is_occupied :18 */
          int64_t yn = y + g->dy[i], xn = x + g->dx[i];                      /* :19 */
          if (in_bounds(gh, gw, yn, xn)) {                                   /* :20 */
            int64_t j = yn * gw + xn;
            if (g->uy[j] != 0 || g->ux[j] != 0) {                            /* has_underlying :22 */
              ready = g->calc[j];                                            /* :23 */
              base = (g->dy[i] != 0 && g->my[j]) || (g->dx[i] != 0 && g->mx[j]); /* :24-25 */
            } else { ready = 1; base = 1; }                                  /* :26 */
          } else { ready = 1; base = 1; }                                    /* :27 */
        } else { ready = 1; base = 1; }                                      /* :28 */
        if (ready) {                                                         /* :29 */
          calc = 1;                                                          /* :30 */
          if (base) {                                                        /* :31 */
            if (g->uy[i] != 0 || g->ux[i] != 0) {                            /* :32 */
              int64_t yp = y - g->uy[i], xp = x - g->ux[i];                  /* :33-34 */
              int py = g->uy[i] != 0 && in_bounds(gh, gw, yp, x) && g->dy[yp * gw + x] != 0; /* :35-37 */
              int px = g->ux[i] != 0 && in_bounds(gh, gw, y, xp) && g->dx[y * gw + xp] != 0; /* :38-40 */
              /* match :41-47, first matching case wins */
              if (px && pref)        { my = 0; mx = 1; pref = 0; }            /* (_, true, true)      :43 */
              else if (py && !pref)  { my = 1; mx = 0; pref = 1; }            /* (true, _, false)     :44 */
              else if (py && !px && pref) { my = 1; mx = 0; pref = 1; }       /* (true, false, true)  :45 */
              else if (!py && px && !pref) { my = 0; mx = 1; pref = 0; }      /* (false, true, false) :46 */
              else                   { my = 0; mx = 0; /* pref = b */ }       /* (false, false, b)    :47 */
            } else { my = 1; mx = 1; }                                       /* :51-53 accept leaks */
          }
        }
      }
      ncalc[i] = calc; nmy[i] = my; nmx[i] = mx; npref[i] = pref;
    }
  }
  free(g->calc); free(g->my); free(g->mx); free(g->pref);
  g->calc = ncalc; g->my = nmy; g->mx = nmx; g->pref = npref;
}

/* steppers.fut:147-164: repeat until the number of calculated cells stops changing (initial previous = 0) */
int64_t fo_move_until_fixpoint(fo_grid *g) {
  int64_t n = g->gh * g->gw, prev = 0, passes = 0;
  int running = 1;
  while (running) {
    fo_update_pass(g);
    passes++;
    int64_t cnt = 0;
#pragma omp parallel for reduction(+ : cnt) schedule(static)
    for (int64_t i = 0; i < n; i++) cnt += g->calc[i];
    running = cnt != prev;
    prev = cnt;
  }
  return passes;
}

/* ------------------------------------------------------------------------------------------
 * steppers.fut:189-224  resolve_cycles (dense masks), reduce1 merge = first grid whose cell is calculated
 * ---------------------------------------------------------------------------------------- */
int fo_resolve_cycles(fo_grid *g, const int8_t *masks, int64_t n) {
  const int64_t gh = g->gh, gw = g->gw, ncell = gh * gw;
  if (n == 0) return 0;                                                      /* :224 #neutral */
  /* all tests (:197-200) look at the INPUT grid; keep its calc plane */
  uint8_t *calc0 = (uint8_t *)malloc((size_t)(ncell ? ncell : 1));
  memcpy(calc0, g->calc, (size_t)ncell);
  int err = 0;
  for (int64_t k = 0; k < n; k++) {
    const int8_t *m = masks + (size_t)k * (size_t)ncell * 2;
    int pass = 1;
    for (int64_t i = 0; i < ncell && pass; i++) {
      int8_t cy = m[2 * i], cx = m[2 * i + 1];
      if (!((cy == 0 && cx == 0) || (!calc0[i] && g->dy[i] == cy && g->dx[i] == cx))) pass = 0;
    }
    if (!pass) continue;                                                     /* else cells :215 */
    for (int64_t y = 0; y < gh; y++)
      for (int64_t x = 0; x < gw; x++) {
        int64_t i = y * gw + x;
        int8_t cy = m[2 * i], cx = m[2 * i + 1];
        if (cy == 0 && cx == 0) continue;                                    /* :202,:212 */
        /* merge (:216-221): an earlier grid's calculated version wins */
        if (g->calc[i]) continue;
        uint8_t my = g->my[i], mx = g->mx[i];
        if (g->uy[i] != 0 && g->ux[i] != 0) {                                /* :206 corner */
          int64_t yp = y - g->uy[i];
          if (!in_bounds(gh, gw, yp, x)) { err = -1; continue; }             /* :207 unchecked index in the reference */
          const int8_t *q = m + 2 * (yp * gw + x);
          if (q[0] == g->uy[i] && q[1] == 0) my = 1; else mx = 1;            /* :207-209 */
        } else { my = g->uy[i] != 0; mx = g->ux[i] != 0; }                   /* :210-211 */
        g->calc[i] = 1; g->my[i] = my; g->mx[i] = mx;                        /* :205 ; pref untouched */
      }
  }
  free(calc0);
  return err;
}

/* ------------------------------------------------------------------------------------------
 * flowamok.fut:10-14 choose_direction_random, plus the deterministic choosers used by tests
 * returns 0 -> {y=u.y, x=0}, 1 -> {y=0, x=u.x}; may advance the cell's own rng
 * ---------------------------------------------------------------------------------------- */
static int fo_choose(fo_chooser *ch, fo_rng *rng) {
  switch (ch ? ch->mode : FO_CHOOSE_RANDOM) {
  case FO_CHOOSE_Y: return 0;
  case FO_CHOOSE_X: return 1;
  case FO_CHOOSE_TAPE: return ch->tape[ch->tick % ch->tape_len] != 0;
  default: return fo_dist_i8_rand(0, 1, rng) != 0;
  }
}

/* ------------------------------------------------------------------------------------------
 * steppers.fut:58-145  move_cell / move_movable_cells / reset_cells, fused as in :166-173
 * ---------------------------------------------------------------------------------------- */
int64_t fo_move_and_reset_cells(fo_grid *g, fo_chooser *ch, fo_leak **leaks_out) {
  const int64_t gh = g->gh, gw = g->gw, n = gh * gw;
  size_t sn = (size_t)(n ? n : 1);
  int8_t *ndy = (int8_t *)malloc(sn), *ndx = (int8_t *)malloc(sn);
  uint32_t *ncolor = (uint32_t *)malloc(sn * 4);
  uint8_t *leakflag = (uint8_t *)calloc(sn, 1);
  /* cells that leak AND draw in the same tick: the leak tuple carries the PRE-draw rng (:122 copies the input cell) */
  int64_t nld = 0, capld = 0, *ld_idx = NULL; uint64_t *ld_state = NULL;
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < gh; y++) {
    for (int64_t x = 0; x < gw; x++) {
      int64_t i = y * gw + x;
      const int8_t uy = g->uy[i], ux = g->ux[i], dy = g->dy[i], dx = g->dx[i];
      const int64_t yn = y + dy, xn = x + dx;                                /* :66 */
      int8_t rdy = dy, rdx = dx; uint32_t rcolor = g->color[i];
      const uint64_t rng0 = g->rng_state[i]; int drew = 0;
      int vacate = 0;                                                        /* run handle_cell_moved_to_cell_next :67-74 */
      if (g->my[i] || g->mx[i]) {                                            /* :76 */
        int64_t yp, xp;
        if (g->my[i]) { yp = y - uy; xp = x; } else { yp = y; xp = x - ux; } /* :77-79 */
        if (in_bounds(gh, gw, yp, xp)) {                                     /* :80 */
          int64_t p = yp * gw + xp;
          int y_ok = g->my[i] && g->uy[p] != 0 && g->uy[p] == uy;            /* :82 */
          int x_ok = g->mx[i] && g->ux[p] != 0 && g->ux[p] == ux;            /* :83 */
          if (y_ok || x_ok) {                                                /* :84 */
            int fy = 0, fyi = 0, fx = 0, fxi = 0;
            if (uy != 0 && in_bounds(gh, gw, y + uy, x)) {                   /* :85-92 */
              int64_t q = (y + uy) * gw + x;
              fy = g->uy[q] == uy; fyi = fy || (g->uy[q] == 0 && g->ux[q] == 0);
            }
            if (ux != 0 && in_bounds(gh, gw, y, x + ux)) {                   /* :93-100 */
              int64_t q = y * gw + (x + ux);
              fx = g->ux[q] == ux; fxi = fx || (g->uy[q] == 0 && g->ux[q] == 0);
            }
            if (fy && fx) {                                                  /* :102-103 */
              fo_rng r = {g->rng_state[i], g->rng_inc[i]};
              int c = fo_choose(ch, &r); drew = 1;
              g->rng_state[i] = r.state; g->rng_inc[i] = r.inc;              /* :112 (own cell only: race-free) */
              if (c == 0) { rdy = uy; rdx = 0; } else { rdy = 0; rdx = ux; }
            } else if (!fx && fyi) { rdy = uy; rdx = 0; }                    /* :104-105 */
            else if (!fy && fxi) { rdy = 0; rdx = ux; }                      /* :106-107 */
            else { rdy = g->dy[p]; rdx = g->dx[p]; }                         /* :108 */
            rcolor = g->color[p];                                            /* :109 (aux = () :110) */
          } else vacate = 1;                                                 /* :113 */
        } /* else cell unchanged :114 */
      } else vacate = 1;                                                     /* :115 */
      if (vacate) {
        if (in_bounds(gh, gw, yn, xn)) {                                     /* :68 */
          int64_t j = yn * gw + xn;
          if ((g->my[j] && uy != 0) || (g->mx[j] && ux != 0)) { rdy = 0; rdx = 0; } /* :70-72 */
        } else { rdy = 0; rdx = 0; }                                         /* :74 */
      }
      if ((dy != 0 || dx != 0) && in_bounds(gh, gw, yn, xn)) {               /* :117 */
        int64_t j = yn * gw + xn;
        if (g->uy[j] == 0 && g->ux[j] == 0 &&                                /* :119 */
            ((g->my[j] && dy != 0) || (g->mx[j] && dx != 0))) leakflag[i] = 1; /* :120-121 */
      }
      if (leakflag[i] && drew && g->rng_state[i] != rng0) {
#pragma omp critical
        {
          if (nld == capld) { capld = capld ? capld * 2 : 16;
            ld_idx = (int64_t *)realloc(ld_idx, sizeof(int64_t) * (size_t)capld);
            ld_state = (uint64_t *)realloc(ld_state, 8 * (size_t)capld); }
          ld_idx[nld] = i; ld_state[nld] = rng0; nld++;
        }
      }
      ndy[i] = rdy; ndx[i] = rdx; ncolor[i] = rcolor;
    }
  }
  /* all_somes (option.fut:3-12): order-preserving, row-major in the source cell */
  int64_t nleaks = 0;
  for (int64_t i = 0; i < n; i++) nleaks += leakflag[i];
  if (leaks_out) {
    fo_leak *L = (fo_leak *)malloc(sizeof(fo_leak) * (size_t)(nleaks ? nleaks : 1));
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++)
      if (leakflag[i]) {
        fo_leak *l = &L[k++];
        l->y_src = i / gw; l->x_src = i % gw;
        l->y_next = l->y_src + g->dy[i]; l->x_next = l->x_src + g->dx[i];
        /* the tuple carries the source cell as it was BEFORE the move (:122) */
        l->rng.state = g->rng_state[i]; l->rng.inc = g->rng_inc[i];
        for (int64_t q = 0; q < nld; q++) if (ld_idx[q] == i) l->rng.state = ld_state[q];
        l->uy = g->uy[i]; l->ux = g->ux[i]; l->dy = g->dy[i]; l->dx = g->dx[i];
        l->color = g->color[i]; l->calc = g->calc[i]; l->my = g->my[i]; l->mx = g->mx[i]; l->pref = g->pref[i];
      }
    *leaks_out = L;
  }
  free(leakflag); free(ld_idx); free(ld_state);
  free(g->dy); free(g->dx); free(g->color);
  g->dy = ndy; g->dx = ndx; g->color = ncolor;
  /* reset_cells :140-145 (pref kept) */
  memset(g->calc, 0, (size_t)n); memset(g->my, 0, (size_t)n); memset(g->mx, 0, (size_t)n);
  return nleaks;
}

/* steppers.fut:177-185 */
int64_t fo_step_quick(fo_grid *g, fo_chooser *ch, fo_leak **leaks) {
  fo_move_until_fixpoint(g);
  return fo_move_and_reset_cells(g, ch, leaks);
}
/* steppers.fut:302-312 */
int64_t fo_step_perfect(fo_grid *g, fo_chooser *ch, const int8_t *masks, int64_t n, fo_leak **leaks) {
  fo_move_until_fixpoint(g);
  fo_resolve_cycles(g, masks, n);
  return fo_move_and_reset_cells(g, ch, leaks);
}

/* ------------------------------------------------------------------------------------------
 * sparse ring form of resolve_cycles (test-infrastructure extension for grids whose dense masks do
 * not fit; equality with the dense form is itself tested)
 * ---------------------------------------------------------------------------------------- */
void fo_resolve_cycles_sparse(fo_grid *g, int64_t n, const int64_t *off, const int64_t *idx,
                              const int8_t *cy, const int8_t *cx, const uint8_t *grant) {
  /* pass tests use the input calc; passing rings are pairwise disjoint, so marking in
     place cannot influence another ring's test except through shared cells, which we guard like the merge */
  uint8_t *passk = (uint8_t *)malloc((size_t)(n ? n : 1));
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t k = 0; k < n; k++) {
    int pass = 1;
    for (int64_t t = off[k]; t < off[k + 1] && pass; t++) {
      int64_t i = idx[t];
      if (g->calc[i] || g->dy[i] != cy[t] || g->dx[i] != cx[t]) pass = 0;
    }
    passk[k] = (uint8_t)pass;
  }
  for (int64_t k = 0; k < n; k++) {
    if (!passk[k]) continue;
    for (int64_t t = off[k]; t < off[k + 1]; t++) {
      int64_t i = idx[t];
      if (g->calc[i]) continue;
      g->calc[i] = 1;
      if (grant[t] == 1) g->my[i] = 1; else g->mx[i] = 1;
    }
  }
  free(passk);
}

int64_t fo_cycles_dense_to_sparse(const fo_grid *g, const int8_t *masks, int64_t n, int64_t **off_o,
                                  int64_t **idx_o, int8_t **cy_o, int8_t **cx_o, uint8_t **grant_o) {
  const int64_t gh = g->gh, gw = g->gw, ncell = gh * gw;
  int64_t total = 0;
  for (int64_t k = 0; k < n; k++)
    for (int64_t i = 0; i < ncell; i++) {
      const int8_t *m = masks + ((size_t)k * (size_t)ncell + (size_t)i) * 2;
      total += (m[0] != 0 || m[1] != 0);
    }
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(total ? total : 1));
  int8_t *cy = (int8_t *)malloc((size_t)(total ? total : 1)), *cx = (int8_t *)malloc((size_t)(total ? total : 1));
  uint8_t *grant = (uint8_t *)malloc((size_t)(total ? total : 1));
  int64_t t = 0;
  for (int64_t k = 0; k < n; k++) {
    off[k] = t;
    const int8_t *m = masks + (size_t)k * (size_t)ncell * 2;
    for (int64_t y = 0; y < gh; y++)
      for (int64_t x = 0; x < gw; x++) {
        int64_t i = y * gw + x;
        if (m[2 * i] == 0 && m[2 * i + 1] == 0) continue;
        idx[t] = i; cy[t] = m[2 * i]; cx[t] = m[2 * i + 1];
        if (g->uy[i] != 0 && g->ux[i] != 0) {                                /* steppers.fut:206-209 */
          int64_t yp = y - g->uy[i];
          int isy = in_bounds(gh, gw, yp, x) && m[2 * (yp * gw + x)] == g->uy[i] && m[2 * (yp * gw + x) + 1] == 0;
          grant[t] = isy ? 1 : 2;
        } else grant[t] = (g->uy[i] != 0) ? 1 : 2;                           /* :210-211 (exactly one is set on a ring cell) */
        t++;
      }
  }
  off[n] = t;
  *off_o = off; *idx_o = idx; *cy_o = cy; *cx_o = cx; *grant_o = grant;
  return total;
}

int64_t fo_step_perfect_sparse(fo_grid *g, fo_chooser *ch, int64_t n, const int64_t *off, const int64_t *idx,
                               const int8_t *cy, const int8_t *cx, const uint8_t *grant, fo_leak **leaks) {
  fo_move_until_fixpoint(g);
  fo_resolve_cycles_sparse(g, n, off, idx, cy, cx, grant);
  return fo_move_and_reset_cells(g, ch, leaks);
}

/* ------------------------------------------------------------------------------------------
 * steppers.fut:233-300  find_cycles, literal worklist with dense path masks
 * ---------------------------------------------------------------------------------------- */
typedef struct { int8_t *mask; int64_t y, x, ys, xs; } fo_path;

int64_t fo_find_cycles(const fo_grid *g, int8_t **masks_out, int64_t max_paths) {
  const int64_t gh = g->gh, gw = g->gw, ncell = gh * gw;
  size_t msz = (size_t)ncell * 2;
  int64_t cap = 16, np = 0;
  fo_path *ps = (fo_path *)malloc(sizeof(fo_path) * (size_t)cap);
#define PUSH(M, Y, X, YS, XS) do { if (np == cap) { cap *= 2; ps = (fo_path *)realloc(ps, sizeof(fo_path) * (size_t)cap); } \
    ps[np].mask = (M); ps[np].y = (Y); ps[np].x = (X); ps[np].ys = (YS); ps[np].xs = (XS); np++; } while (0)
  /* starts :235-259 (corners in row-major order; y start before x start) */
  for (int64_t y = 0; y < gh; y++)
    for (int64_t x = 0; x < gw; x++) {
      int64_t i = y * gw + x;
      int8_t cy = g->uy[i], cx = g->ux[i];
      if (!(cy != 0 && cx != 0)) continue;
      int64_t yn = y + cy, xn = x + cx;
      int y_ok = in_bounds(gh, gw, yn, x) && g->uy[yn * gw + x] == cy;       /* :244 */
      int x_ok = in_bounds(gh, gw, y, xn) && g->ux[y * gw + xn] == cx;       /* :245 */
      if (y_ok) { int8_t *m = (int8_t *)calloc(msz, 1); m[2 * i] = cy; m[2 * i + 1] = 0; PUSH(m, yn, x, y, x); }
      if (x_ok) { int8_t *m = (int8_t *)calloc(msz, 1); m[2 * i] = 0; m[2 * i + 1] = cx; PUSH(m, y, xn, y, x); }
    }
  /* worklist :260-296 */
  int64_t cur = 0;
  int overflow = 0;
  while (cur < np) {
    fo_path *p = &ps[cur];
    int64_t y = p->y, x = p->x;
    if (y == p->ys && x == p->xs) { cur++; continue; }                       /* cycle :264-265 */
    /* NB: the reference indexes grid[y, x] at :266 before any bounds test; positions handed on are
       always in bounds except through the plain-road step at :292-293, which can walk off the grid
       (a Futhark index error).  We discard such a path like an off-road one. */
    if (!in_bounds(gh, gw, y, x)) { p->y = -1; p->x = -1; cur++; continue; }
    int64_t i = y * gw + x;
    if (p->mask[2 * i] != 0 || p->mask[2 * i + 1] != 0) { p->y = -1; p->x = -1; cur++; continue; } /* :267-269 */
    int8_t cy = g->uy[i], cx = g->ux[i];
    if (cy != 0 && cx != 0) {                                                /* :272 */
      int64_t yn = y + cy, xn = x + cx;
      int y_ok = in_bounds(gh, gw, yn, x) && g->uy[yn * gw + x] == cy;
      int x_ok = in_bounds(gh, gw, y, xn) && g->ux[y * gw + xn] == cx;
      if (y_ok && x_ok) {                                                    /* :281-285 */
        if (max_paths > 0 && np >= max_paths) { overflow = 1; break; }
        int8_t *mb = (int8_t *)malloc(msz); memcpy(mb, p->mask, msz);
        int64_t ys = p->ys, xs = p->xs;
        p->mask[2 * i] = cy; p->mask[2 * i + 1] = 0; p->y = yn; p->x = x;
        mb[2 * i] = 0; mb[2 * i + 1] = cx;
        PUSH(mb, y, xn, ys, xs);
      } else if (y_ok) { p->mask[2 * i] = cy; p->mask[2 * i + 1] = 0; p->y = yn; p->x = x; }   /* :286-288 */
      else { p->mask[2 * i] = 0; p->mask[2 * i + 1] = cx; p->y = y; p->x = xn; }               /* :289-290 (x_grid even if !x_ok) */
    } else if (cy != 0 || cx != 0) {                                         /* :291-294 */
      p->mask[2 * i] = cy; p->mask[2 * i + 1] = cx; p->y = y + cy; p->x = x + cx;
    } else { p->y = -1; p->x = -1; cur++; }                                  /* :295-296 */
  }
  int64_t nout = 0;
  int8_t *out = NULL;
  if (!overflow) {
    /* :298 keep successful; :299 nub_2d keeps the first of equal masks (helpers.fut:30-33) */
    int64_t *keep = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np ? np : 1));
    for (int64_t a = 0; a < np; a++) {
      if (ps[a].y == -1) continue;
      int dup = 0;
      for (int64_t b = 0; b < nout && !dup; b++) dup = memcmp(ps[keep[b]].mask, ps[a].mask, msz) == 0;
      if (!dup) keep[nout++] = a;
    }
    out = (int8_t *)malloc(msz * (size_t)(nout ? nout : 1));
    for (int64_t b = 0; b < nout; b++) memcpy(out + msz * (size_t)b, ps[keep[b]].mask, msz);
    free(keep);
  }
  for (int64_t a = 0; a < np; a++) free(ps[a].mask);
  free(ps);
#undef PUSH
  if (overflow) { *masks_out = NULL; return -1; }
  *masks_out = out;
  return nout;
}

/* ------------------------------------------------------------------------------------------
 * explorer/scenarios.fut:5-274, ids per explorer/explorer.fut:9-22
 * ---------------------------------------------------------------------------------------- */
static const char *fo_names[FO_N_SCENARIOS] = {
  "single fork", "tight queues", "basic cycle", "crossing", "close crossing", "crossroads",
  "small tight cycle", "independent tight cycles", "overlapping tight cycles", "hits an edge",
  "adding lines", "multi spill", "many cycles (n = 2)", "t cross"};
const char *fo_scenario_name(int sid) { return (sid >= 0 && sid < FO_N_SCENARIOS) ? fo_names[sid] : ""; }

#define LH(y, x0, x1, f) fo_add_line_horizontal(g, (y), (x0), (x1), (int8_t)(f))
#define LV(y0, y1, x, f) fo_add_line_vertical(g, (y0), (y1), (x), (int8_t)(f))
#define AC(y, x, c, dy, dx) fo_add_cell(g, (y), (x), (c), (int8_t)(dy), (int8_t)(dx))

static void fo_small_tight_cycle_init(fo_grid *g) {                          /* :117-126 */
  LH(11, 11, 13, 1); LV(11, 13, 13, 1); LH(13, 11, 13, -1); LV(11, 13, 11, -1);
  for (int i = 0; i < 2; i++) AC(11, 11 + i, FO_RED, 0, 1);
  for (int i = 0; i < 2; i++) AC(11 + i, 13, FO_BLUE, 1, 0);
  for (int i = 0; i < 2; i++) AC(13, 12 + i, FO_ORANGE, 0, -1);
  for (int i = 0; i < 2; i++) AC(12 + i, 11, FO_MAGENTA, -1, 0);
}

void fo_scenario_init(int sid, fo_grid *g) {
  switch (sid) {
  case 0: LH(10, 10, 20, 1); LV(10, 40, 20, 1); LH(30, 5, 25, -1); break;    /* :8-12 */
  case 1: LH(40, 10, 20, 1); LV(10, 40, 20, -1); break;                      /* :24-27 */
  case 2:                                                                    /* :37-46 */
    LH(10, 10, 30, 1); LV(10, 30, 30, 1); LH(30, 10, 30, -1); LV(10, 30, 10, -1);
    AC(10, 20, FO_RED, 0, 1); AC(20, 30, FO_BLUE, 1, 0); AC(30, 20, FO_BROWN, 0, -1); AC(20, 10, FO_MAGENTA, -1, 0);
    break;
  case 3: case 4: LH(20, 10, 30, 1); LV(10, 30, 20, 1); break;               /* :55-58, :71-74 */
  case 5: LH(29, 1, 58, -1); LH(30, 1, 58, 1); LV(1, 58, 29, 1); LV(1, 58, 30, -1); break; /* :87-92 */
  case 6: fo_small_tight_cycle_init(g); break;
  case 7:                                                                    /* :135-145 */
    fo_small_tight_cycle_init(g);
    LH(10, 10, 14, -1); LV(10, 14, 14, -1); LH(14, 10, 14, 1); LV(10, 14, 10, 1);
    for (int i = 0; i < 4; i++) AC(10, 11 + i, FO_RED, 0, -1);
    for (int i = 0; i < 4; i++) AC(11 + i, 14, FO_BLUE, -1, 0);
    for (int i = 0; i < 4; i++) AC(14, 10 + i, FO_ORANGE, 0, 1);
    for (int i = 0; i < 4; i++) AC(10 + i, 10, FO_MAGENTA, 1, 0);
    break;
  case 8:                                                                    /* :154-173 */
    LH(10, 10, 14, -1); LV(10, 24, 14, -1); LH(24, 10, 14, 1); LV(10, 24, 10, 1);
    LH(14, 14, 24, 1); LV(14, 20, 24, 1); LH(20, 14, 24, -1);
    for (int i = 0; i < 4; i++) AC(10, 11 + i, FO_RED, 0, -1);
    for (int i = 0; i < 14; i++) AC(11 + i, 14, FO_BLUE, -1, 0);
    for (int i = 0; i < 4; i++) AC(24, 10 + i, FO_ORANGE, 0, 1);
    for (int i = 0; i < 14; i++) AC(10 + i, 10, FO_MAGENTA, 1, 0);
    for (int i = 0; i < 9; i++) AC(14, 15 + i, FO_BROWN, 0, 1);
    for (int i = 0; i < 6; i++) AC(14 + i, 24, FO_VIOLET, 1, 0);
    for (int i = 0; i < 10; i++) AC(20, 15 + i, FO_BROWN, 0, -1);
    break;
  case 9: LH(10, 5, 15, 1); LV(5, 15, 16, -1); AC(10, 10, FO_BROWN, 0, 1); break; /* :182-186 */
  case 10: LH(1, 1, 6, 1); break;                                            /* :195-197 */
  case 11: LV(1, 9, 10, 1); LH(10, 1, 9, 1); LV(11, 19, 10, -1); LH(10, 11, 19, -1); break; /* :218-223 */
  case 12: {                                                                 /* :241-254, n = 2 */
    const int64_t n = 2, yo = 19, xo = 10;
    for (int64_t y = 0; y < n; y++)
      for (int64_t x = 0; x < n; x++) {
        int64_t y1 = y + 1, x1 = x + 1;
        LH(yo + 10 * y, xo + 10 * x, xo + 10 * x1 - (x == n - 1), 1);
        LV(yo + 10 * y, yo + 10 * y1 - (y == n - 1), xo + 10 * x1 - 1, 1);
        LH(yo + 10 * y1 - 1, xo + 10 * x - 1 + (x == 0), xo + 10 * x1 - 1, -1);
        LV(yo + 10 * y - 1 + (y == 0), yo + 10 * y1 - 1, xo + 10 * x, -1);
      }
    AC(yo, xo, FO_GREEN, 0, 1);
    break; }
  case 13: {                                                                 /* :262-270 */
    const int64_t y = 25;
    LH(y, 0, 10, 1); LH(y - 1, 10, 30, 1); LH(y + 1, 10, 30, 1); LV(y, y + 1, 10, 1); LV(y - 1, y, 10, -1);
    AC(y, 0, FO_MAGENTA, 0, 1);
    break; }
  default: break;
  }
}

/* Futhark `%` on i64 is floored modulo; all operands below are non-negative. */
int fo_scenario_step(int sid, fo_grid *g, int64_t steps, fo_rng *rng) {
  switch (sid) {
  case 0:                                                                    /* :14-18 */
    if (steps % 5 == 0) { uint32_t c = fo_random_color(rng); AC(10, 10, c, 0, 1); }
    return 0;
  case 1: { uint32_t c = fo_random_color(rng); AC(40, 10, c, 0, 1); return 0; } /* :29-31 */
  case 3:                                                                    /* :60-65 */
    if (steps % 5 == 0) { AC(20, 10, FO_RED, 0, 1); AC(10, 20, FO_BLUE, 1, 0); }
    return 0;
  case 4: {                                                                  /* :76-81 */
    uint32_t c = fo_random_color(rng); AC(20, 10, c, 0, 1);
    c = fo_random_color(rng); AC(10, 20, c, 1, 0);
    return 0; }
  case 5: {                                                                  /* :94-111 */
    int8_t east = fo_dist_i8_rand(0, 1, rng), south = fo_dist_i8_rand(0, 4, rng);
    int8_t north = fo_dist_i8_rand(0, 9, rng), west = fo_dist_i8_rand(0, 19, rng);
    if (north == 0) AC(1, 29, FO_BROWN, 1, 0);
    if (east == 0) AC(29, 58, FO_GREEN, 0, -1);
    if (south == 0) AC(58, 30, FO_VIOLET, -1, 0);
    if (west == 0) AC(30, 1, FO_YELLOW, 0, 1);
    return 0; }
  case 10: {                                                                 /* :199-212 */
    int64_t s = steps + 10;
    int recompute = 0;
    if (s % 5 == 0 && s / 5 < 20) {
      int64_t k = (s / 10) * 5 + 1;
      if ((s / 5) % 2 == 1) LH(k, k, k + 5, 1); else LV(k - 5, k, k, 1);
      recompute = 1;
    }
    int8_t nc = fo_dist_i8_rand(0, 3, rng);
    if (nc == 0) { uint32_t c = fo_random_color(rng); AC(1, 1, c, 0, 1); }
    return recompute; }
  case 11:                                                                   /* :225-232 */
    if (steps % 5 == 0) {
      AC(1, 10, FO_RED, 1, 0); AC(10, 1, FO_BLUE, 0, 1); AC(19, 10, FO_GREEN, -1, 0); AC(10, 19, FO_YELLOW, 0, -1);
    }
    return 0;
  default: return 0;                                                         /* sids 2,6,7,8,9,12,13: identity */
  }
}
#undef LH
#undef LV
#undef AC

/* ------------------------------------------------------------------------------------------
 * flowamok.fut:16-64 render (matte gray/mix restated from memory: unpinned)
 * ---------------------------------------------------------------------------------------- */
static int64_t fdiv(int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }
static int64_t fmod_(int64_t a, int64_t b) { return a - fdiv(a, b) * b; }

static uint32_t fo_mix(float m1, uint32_t c1, float m2, uint32_t c2) {
  float r1 = ((c1 >> 16) & 255) / 255.0f, g1 = ((c1 >> 8) & 255) / 255.0f, b1 = (c1 & 255) / 255.0f, a1 = (c1 >> 24) / 255.0f;
  float r2 = ((c2 >> 16) & 255) / 255.0f, g2 = ((c2 >> 8) & 255) / 255.0f, b2 = (c2 & 255) / 255.0f, a2 = (c2 >> 24) / 255.0f;
  float m12 = m1 + m2, w1 = m1 / m12, w2 = m2 / m12;
  float rs = w1 * r1 * r1 + w2 * r2 * r2, gs = w1 * g1 * g1 + w2 * g2 * g2, bs = w1 * b1 * b1 + w2 * b2 * b2;
  return fo_from_rgba(sqrtf(rs), sqrtf(gs), sqrtf(bs), (m1 * a1 + m2 * a2) / m12);
}

static int fo_flow_draw_at(int64_t scale, int8_t flow_x, int8_t flow_y, int64_t y, int64_t x) { /* :34-47 */
  int64_t d2 = fdiv(scale, 2), m2 = fmod_(scale, 2);
  y -= d2; x -= d2;
  if (m2 == 0) { y += (y >= 0); x += (x >= 0); }
  int64_t d = d2 - m2;
  int64_t yq = (int64_t)flow_y * y, xq = (int64_t)flow_x * x;
  int64_t ya = y < 0 ? -y : y, xa = x < 0 ? -x : x;
  int y_draw = (yq == d - xa) && xa <= d, x_draw = (xq == d - ya) && ya <= d;
  return (flow_x == 0 && y_draw) || (flow_y == 0 && x_draw) || (flow_y != 0 && flow_x != 0 && (y_draw || x_draw));
}

void fo_render(const fo_grid *g, int64_t h, int64_t w, int64_t scale, uint32_t *pixels) {
  const int64_t gh = g->gh, gw = g->gw;
  const uint32_t grey = fo_from_rgba(0.3f, 0.3f, 0.3f, 1.0f);               /* argb.gray 0.3 :56 */
  for (int64_t y = 0; y < h; y++)
    for (int64_t x = 0; x < w; x++) {
      int64_t yg = fdiv(y, scale), xg = fdiv(x, scale);                      /* :50 */
      uint32_t px = FO_BLACK;
      if (in_bounds(gh, gw, yg, xg)) {
        int64_t i = yg * gw + xg;
        int occupied = g->dy[i] != 0 || g->dx[i] != 0;
        int has_flow = g->ux[i] != 0 || g->uy[i] != 0;
        uint32_t base = occupied ? g->color[i] : (has_flow ? FO_WHITE : FO_BLACK); /* :18-29 */
        if (has_flow) {                                                      /* :53 */
          if (fo_flow_draw_at(scale, g->ux[i], g->uy[i], fmod_(y, scale), fmod_(x, scale)))
            px = occupied ? fo_mix(0.5f, grey, 0.5f, base) : grey;           /* :56-59 */
          else px = base;                                                    /* :60 */
        } else px = FO_BLACK;                                                /* :61 */
      }
      pixels[y * w + x] = px;
    }
}

/* ------------------------------------------------------------------------------------------
 * synthetic road network (rule documented in DESIGN.md):
 *  - avenues: rows y%16==0 carry u.x=+1, columns x%16==8 carry u.y=+1 (monotone: no cycles)
 *  - ring tiles: block (by,bx) owns rows 16by+2..16by+14, cols 16bx+10..16bx+22; if hash bit 0 is set it
 *    holds a one-way 13x13 rectangular ring built like scenarios.fut:117-120 (bit 1: clockwise)
 *  - cars: probability density16/65536 per road cell; a ring is completely full with prob ring_full16/16
 * ---------------------------------------------------------------------------------------- */
static uint64_t sm64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
static uint64_t shash(uint64_t seed, uint64_t tag, uint64_t a, uint64_t b) {
  uint64_t k = sm64(seed ^ (tag * 0xD6E8FEB86659FD93ULL));
  k = sm64(k ^ a);
  return sm64(k ^ b);
}
static void ring_dir(int cw, int ly, int lx, int *hy, int *hx) {
  *hy = 0; *hx = 0;
  if (cw) {
    if (ly == 0 && lx < 12) *hx = 1; else if (lx == 12 && ly < 12) *hy = 1;
    else if (ly == 12 && lx > 0) *hx = -1; else *hy = -1;
  } else {
    if (ly == 0 && lx > 0) *hx = -1; else if (lx == 0 && ly < 12) *hy = 1;
    else if (ly == 12 && lx < 12) *hx = 1; else *hy = -1;
  }
}
int64_t fo_generate_synthetic(fo_grid *g, uint64_t seed, uint32_t density16, uint32_t ring_full16,
                              int64_t **off_o, int64_t **idx_o, int8_t **cy_o, int8_t **cx_o, uint8_t **grant_o) {
  const int64_t gh = g->gh, gw = g->gw;
  const uint64_t inc = (0xda3e39cb94b95bdbULL << 1) | 1ULL;
#pragma omp parallel for schedule(static)
  for (int64_t y = 0; y < gh; y++)
    for (int64_t x = 0; x < gw; x++) {
      int64_t i = y * gw + x;
      uint64_t hc = shash(seed, 2, (uint64_t)y, (uint64_t)x);
      int car = (uint32_t)(hc & 0xFFFF) < density16;
      int arow = (y % 16) == 0, acol = (x % 16) == 8;
      g->uy[i] = acol ? 1 : 0; g->ux[i] = arow ? 1 : 0; g->dy[i] = 0; g->dx[i] = 0;
      if ((arow || acol) && car) {
        int alongx = arow && (!acol || ((hc >> 16) & 1));
        if (alongx) g->dx[i] = 1; else g->dy[i] = 1;
      }
      g->pref[i] = 0; g->calc[i] = g->my[i] = g->mx[i] = 0;
      g->rng_state[i] = shash(seed, 4, (uint64_t)y, (uint64_t)x); g->rng_inc[i] = inc;
      g->color[i] = 0xFFFFFFFFu;
    }
  int64_t nby = gh >= 15 ? (gh - 15) / 16 + 1 : 0, nbx = gw >= 23 ? (gw - 23) / 16 + 1 : 0;
  int64_t nr = 0;
  for (int64_t by = 0; by < nby; by++)
    for (int64_t bx = 0; bx < nbx; bx++) nr += (int64_t)(shash(seed, 1, (uint64_t)by, (uint64_t)bx) & 1);
  int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nr + 1));
  int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nr * 48 + 1));
  int8_t *cy = (int8_t *)malloc((size_t)(nr * 48 + 1)), *cx = (int8_t *)malloc((size_t)(nr * 48 + 1));
  uint8_t *grant = (uint8_t *)malloc((size_t)(nr * 48 + 1));
  int64_t k = 0, t = 0;
  for (int64_t by = 0; by < nby; by++)
    for (int64_t bx = 0; bx < nbx; bx++) {
      uint64_t hb = shash(seed, 1, (uint64_t)by, (uint64_t)bx);
      if (!(hb & 1)) continue;
      int cw = (int)((hb >> 1) & 1), full = (uint32_t)((hb >> 8) & 15) < ring_full16;
      int s = cw ? 1 : -1;
      off[k++] = t;
      for (int ly = 0; ly <= 12; ly++)
        for (int lx = 0; lx <= 12; lx++) {
          if (!(ly == 0 || ly == 12 || lx == 0 || lx == 12)) continue;
          int64_t y = 16 * by + 2 + ly, x = 16 * bx + 10 + lx, i = y * gw + x;
          if (ly == 0) g->ux[i] = (int8_t)s;
          if (ly == 12) g->ux[i] = (int8_t)-s;
          if (lx == 12) g->uy[i] = (int8_t)s;
          if (lx == 0) g->uy[i] = (int8_t)-s;
          int hy, hx;
          ring_dir(cw, ly, lx, &hy, &hx);
          uint64_t hc = shash(seed, 2, (uint64_t)y, (uint64_t)x);
          if (full || (uint32_t)(hc & 0xFFFF) < density16) { g->dy[i] = (int8_t)hy; g->dx[i] = (int8_t)hx; }
          idx[t] = i; cy[t] = (int8_t)hy; cx[t] = (int8_t)hx;
          int corner = (ly == 0 || ly == 12) && (lx == 0 || lx == 12);
          /* steppers.fut:206-211: a corner is granted along the axis it is ENTERED on = not its heading */
          grant[t] = (uint8_t)(corner ? (hy != 0 ? 2 : 1) : (hx != 0 ? 2 : 1));
          t++;
        }
    }
  off[k] = t;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < gh * gw; i++)
    if (g->dy[i] != 0 || g->dx[i] != 0)
      g->color[i] = 0xFF000000u | (uint32_t)(shash(seed, 3, (uint64_t)(i / gw), (uint64_t)(i % gw)) & 0xFFFFFFu);
  *off_o = off; *idx_o = idx; *cy_o = cy; *cx_o = cx; *grant_o = grant;
  return nr;
}
