/*
 * flowamok_b200 -- C ABI of the B200-native per-tick grid update of nqpz/flowamok.
 *
 * Two layers, both plain C (pointers and sizes only, no torch / C++ types):
 *
 *  1. flowamok_*  : the OPERATOR surface.  One entry point per reference operator on the hot path
 *                   (SURVEY.md 8b "inner" boundary); citations are /root/reference/<file>:<line>.
 *  2. futhark_*   : the surface the Futhark compiler emits for README.fut's two entry points
 *                   (`init`, `step`, README.fut:12-13), so a frontend written against the generated
 *                   library links against this one unchanged (SURVEY.md 8b "outer" boundary).
 *                   Declared in include/flowamok_futhark.h.
 *
 * Conventions (same as a Futhark-generated library): every function returns 0 on success and
 * non-zero on failure; the message is available from flowamok_context_get_error (malloc'd, caller
 * frees).  Handles are opaque; grids live in device memory (B200 HBM) between calls, exactly like
 * Futhark's opaque values on its GPU backends.  A context is single-threaded.  Work is enqueued on
 * the context's CUDA stream; results are valid after flowamok_context_sync or any call documented
 * as synchronous.  There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef FLOWAMOK_H
#define FLOWAMOK_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct flowamok_context flowamok_context;
typedef struct flowamok_grid flowamok_grid;

/* choose_direction parameter of both steppers (src/steppers.fut:180, :305) */
enum flowamok_chooser {
  FLOWAMOK_CHOOSE_RANDOM = 0, /* flowamok.fut:10-14 choose_direction_random (per-cell pcg32, src/random.fut:4-5) */
  FLOWAMOK_CHOOSE_Y = 1,      /* always {y = u.y, x = 0} */
  FLOWAMOK_CHOOSE_X = 2,      /* always {y = 0, x = u.x} */
  FLOWAMOK_CHOOSE_TAPE = 3    /* every draw of tick t returns tape[t % tape_len] (README GIF replay) */
};

/* ---- context (replaces futhark_context_config_new / futhark_context_new / _free / _sync / _get_error) ---- */
/* device < 0: current device.  stream: a cudaStream_t to enqueue on, or NULL for a private stream. */
int flowamok_context_new(flowamok_context **out, int device, void *stream);
int flowamok_context_free(flowamok_context *ctx);
int flowamok_context_sync(flowamok_context *ctx);
char *flowamok_context_get_error(flowamok_context *ctx);
int flowamok_context_device(flowamok_context *ctx);

/* ---- grid values: `[gh][gw](cell ())` of src/model.fut:19-24 as an opaque device value ---- */
int flowamok_grid_new(flowamok_context *ctx, flowamok_grid **out, int64_t gh, int64_t gw);
int flowamok_grid_free(flowamok_context *ctx, flowamok_grid *g);
int flowamok_grid_shape(flowamok_context *ctx, const flowamok_grid *g, int64_t *gh, int64_t *gw);

/* create_grid (src/utils.fut:14-17): every cell white, empty, no road, rng_i = split_rng(gh*gw, rng_from_seed [seed])[i].
 * joined_state/joined_inc (may be NULL) receive join_rng of the per-cell rngs (utils.fut:17).  Synchronous. */
int flowamok_create_grid(flowamok_context *ctx, flowamok_grid *g, int32_t seed, uint64_t *joined_state, uint64_t *joined_inc);

/* Host arrays in the reference's record-of-arrays layout, row-major [gh][gw].  Any pointer may be
 * NULL: on upload the field is zeroed (colour: white, rng: state 0); on download it is skipped.
 * The per-tick fields calculated/direction.{y,x} are always false between ticks (steppers.fut:140-145)
 * and are therefore not part of the value.  rng_inc is shared by all cells (utils.fut:15).  Synchronous. */
int flowamok_grid_upload(flowamok_context *ctx, flowamok_grid *g, const int8_t *uy, const int8_t *ux,
                         const int8_t *dy, const int8_t *dx, const uint32_t *color, const uint8_t *pref,
                         const uint64_t *rng_state, uint64_t rng_inc);
int flowamok_grid_download(flowamok_context *ctx, const flowamok_grid *g, int8_t *uy, int8_t *ux, int8_t *dy,
                           int8_t *dx, uint32_t *color, uint8_t *pref, uint64_t *rng_state, uint64_t *rng_inc);

/* Sparse access to underlying.rng.  choose_direction (flowamok.fut:10-14) is only called for a cell whose road has both axes
 * (steppers.fut:101-103), so only those cells' states ever change: a host that keeps the grid in the reference's layout can
 * pass rng_state = NULL to upload / download (the resident states stay) and move (global flat index, state) of the corner
 * cells instead of 8 bytes per cell.  corners: row-major, up to cap, *n = how many there are (cap = 0 sizes the arrays). */
int flowamok_grid_rng_corners(flowamok_context *ctx, flowamok_grid *g, int64_t cap, int64_t *idx, uint64_t *state, int64_t *n);
int flowamok_grid_rng_scatter(flowamok_context *ctx, flowamok_grid *g, int64_t n, const int64_t *idx, const uint64_t *state);

/* point edits between ticks: src/utils.fut:19-25, :27-33, :35-38 */
int flowamok_add_line_vertical(flowamok_context *ctx, flowamok_grid *g, int64_t y_start, int64_t y_end, int64_t x, int8_t flow);
int flowamok_add_line_horizontal(flowamok_context *ctx, flowamok_grid *g, int64_t y, int64_t x_start, int64_t x_end, int8_t flow);
int flowamok_add_cell(flowamok_context *ctx, flowamok_grid *g, int64_t y, int64_t x, uint32_t color, int8_t dy, int8_t dx);

/* chooser: tape (host bytes, copied) only for FLOWAMOK_CHOOSE_TAPE */
int flowamok_set_chooser(flowamok_context *ctx, flowamok_grid *g, int mode, const uint8_t *tape, int64_t tape_len);

/* cycle_checks of stepper_perfect.step (src/steppers.fut:305) in sparse form: ring k is the cells
 * [off[k], off[k+1]) with position (y,x), check direction (cy,cx) and grant (1: direction.y, 2: direction.x,
 * i.e. the outcome of steppers.fut:206-211 for that cell).  Host arrays, copied.  n = 0 clears.  Every cycle must be closed --
 * each cell's check direction leads to another cell of the same cycle, as find_cycles produces them (steppers.fut:247-293) --
 * or the call fails: ring resolution relies on it (a ring whose cells all head along it is never touched by the fixpoint).
 * On a row band (after flowamok_band_set_cuts) pass the GLOBAL list: cycles may straddle the cuts. */
int flowamok_set_cycles(flowamok_context *ctx, flowamok_grid *g, int64_t n, const int64_t *off, const int64_t *y,
                        const int64_t *x, const int8_t *cy, const int8_t *cx, const uint8_t *grant);
/* stepper_perfect.find_cycles (src/steppers.fut:233-300) on the grid's roads; stores the result as
 * the grid's cycle_checks and returns how many were found.  max_paths bounds the worklist (0: default). */
int flowamok_find_cycles(flowamok_context *ctx, flowamok_grid *g, int64_t max_paths, int64_t *n_cycles);
int flowamok_cycle_count(flowamok_context *ctx, const flowamok_grid *g, int64_t *n_cycles, int64_t *n_cells);
/* dense masks [n][gh][gw][2] (y,x) as the reference returns them; n from flowamok_cycle_count */
int flowamok_get_cycles_dense(flowamok_context *ctx, const flowamok_grid *g, int8_t *masks);

/* stepper_quick.step (src/steppers.fut:177-185) / stepper_perfect.step (:302-312), n_ticks times, the
 * grid being consumed and replaced like the reference's unique `*[gh][gw]` argument.  Enqueues only;
 * if n_leaks is non-NULL the call synchronises and stores the summed number of cell leaks
 * (explorer/explorer.fut:125). */
int flowamok_step_quick(flowamok_context *ctx, flowamok_grid *g, int64_t n_ticks, int64_t *n_leaks);
int flowamok_step_perfect(flowamok_context *ctx, flowamok_grid *g, int64_t n_ticks, int64_t *n_leaks);

/* Like flowamok_step_*, but brackets every kernel of every tick with CUDA events on the context's stream
 * and returns the summed device time in milliseconds of {k_u_local, k_u_iter, k_rings, k_move}.  Synchronous;
 * n_ticks <= 4096.  Measurement aid for bench.py's roofline object, not a reference operator. */
int flowamok_step_profile(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, double *ms);

/* The production tick on a time line (it forks after k_u_local: the global fixpoint on one stream, k_move on the snapshot
 * beside it, then the fix-up), summed over n_ticks <= 1024, in ms: {k_u_local (+ k_rings), fork -> end of k_u_iter,
 * fork -> end of the speculative k_move, join -> end of k_move_fix, whole tick}.  Measurement aid. */
int flowamok_step_timeline(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, double *ms5);

/* Ordered cell_leak list of ONE tick (src/model.fut:26, option.fut:3-12): runs one tick and returns up
 * to cap leaks in row-major order of the source cell as 8 int64 per leak:
 * {y_next, x_next, y_src, x_src, dy, dx, color, pref}.  Synchronous. */
int flowamok_step_leaks(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t cap, int64_t *leaks, int64_t *n_leaks);
/* The same with the WHOLE source cell of the reference's tuple (y_next, x_next, cell) (src/model.fut:26, steppers.fut:122), as
 * move_cell saw it: 14 int64 per leak {y_next, x_next, y_src, x_src, underlying.direction.y, .x, direction.y, .x, color,
 * can_be_moved_to_from.calculated, .direction.y, .direction.x, .next_preference_flow_x, underlying.rng.state (bit pattern; inc is
 * the grid's)}.  aux = () has no storage. */
int flowamok_step_leaks_cells(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t cap, int64_t *leaks, int64_t *n_leaks);

/* render (flowamok.fut:16-64) into a host buffer [gh*scale][gw*scale] of argb.  Synchronous. */
int flowamok_render(flowamok_context *ctx, const flowamok_grid *g, int64_t scale, uint32_t *pixels);

/* like flowamok_render but leaves the pixels in device memory (cudaMalloc'd, caller cudaFree's): what
 * animation.fut:22 hands back as the first component of `step`'s result. */
int flowamok_render_device(flowamok_context *ctx, const flowamok_grid *g, int64_t scale, void **pixels_dev);

/* `copy s.grid` (animation.fut:23): a fresh grid value with the same cells, cycle_checks and chooser. */
int flowamok_grid_clone(flowamok_context *ctx, const flowamok_grid *src, flowamok_grid **out);

/* flowamok_grid_free that keeps the planes for the next flowamok_grid_clone of a grid of the same shape (at most two are kept) */
int flowamok_grid_recycle(flowamok_context *ctx, flowamok_grid *g);

/* Tuning knob.  Normally the global fixpoint runs as a small kernel BESIDE the move; on a jammed network -- recent
 * fixpoint iterations visited more than `visits_per_iteration` words each (default 40000) -- a tick runs it at full size before the
 * move instead.  The results are identical either way. */
int flowamok_set_jam_threshold(flowamok_context *ctx, double visits_per_iteration);
int flowamok_jam_ticks(flowamok_context *ctx, const flowamok_grid *g, int64_t *n);   /* ticks this grid stepped in jam order */

/* diagnostics of the ticks since the last call: ticks, fixpoint iterations, worklist entries relaxed */
int flowamok_stats(flowamok_context *ctx, flowamok_grid *g, int64_t *ticks, int64_t *passes, int64_t *sweeps);

/* ticks by number of frontier iterations of the U fixpoint ("iterations-per-tick histogram so the jam depth behind
 * a number is visible"): hist32[k], k = 0..30, hist32[31] = 31 and more; counted since the grid was created */
int flowamok_stats_hist(flowamok_context *ctx, flowamok_grid *g, int64_t *hist32);
/* P2P transport diagnostics: convergence rounds so far, nanoseconds the band spent waiting for its neighbours' rows */
int flowamok_band_p2p_stats(flowamok_context *ctx, flowamok_grid *g, int64_t *rounds, int64_t *wait_ns);

/* Digest of the owned rows, computed on the device: sum and xor (mod 2^64) over the cells of a hash of (global cell index,
 * heading, preference, colour, rng state) -- see digest_cell in csrc/fa_kernels.cuh.  Sums / xors of bands combine to
 * the whole grid's: the size-independent equality check of BASELINE config 5 (65536^2 does not fit the host).  Synchronous. */
int flowamok_grid_digest(flowamok_context *ctx, const flowamok_grid *g, uint64_t *sum, uint64_t *xor_);

/* Synthetic road network of SURVEY.md App. E (monotone avenue lattice + ring tiles), generated on the
 * device; density16 = car probability per road cell in 1/65536, ring_full16 = share of rings that start
 * completely full in 1/16.  Also installs the analytic cycle list.  Synchronous. */
int flowamok_generate_synthetic(flowamok_context *ctx, flowamok_grid *g, uint64_t seed, uint32_t density16, uint32_t ring_full16);

/* ---- row bands for multi-GPU runs (the reference is single device) ----
 * A band holds rows [row_begin, row_end) of a gh_global x gw grid plus two halo rows per side.  upload/download use
 * WHOLE-grid host arrays (upload reads the rows the band holds, download writes the rows it owns); edits and cycle
 * lists take global coordinates.  One tick of a band, driven by the host (flowamok_b200/bands.py):
 *     pack_state -> send/recv with both neighbours -> unpack_state          (heading, preference, colour rows)
 *     u_init; pack_u -> send/recv -> merge_u(wake=0); u_iter                 (local fixpoint against the neighbours' post-sweep rows)
 *     repeat: pack_u -> send/recv -> merge_u(wake=1); all-reduce(max) of the woken count; u_resume   until 0
 *     [perfect: rings; pack_u -> send/recv -> merge_u(wake=0)]
 *     move
 * side 0 = neighbour with smaller rows, 1 = larger rows; buffers are device pointers of the sizes msg_bytes reports. */
int flowamok_band_new(flowamok_context *ctx, flowamok_grid **out, int64_t gh_global, int64_t gw, int64_t row_begin, int64_t row_end);
int flowamok_band_msg_bytes(flowamok_context *ctx, const flowamok_grid *g, int64_t *state_bytes, int64_t *u_bytes);
int flowamok_band_pack_state(flowamok_context *ctx, flowamok_grid *g, int side, void *dev_buf);
int flowamok_band_unpack_state(flowamok_context *ctx, flowamok_grid *g, int side, const void *dev_buf);
int flowamok_band_u_begin(flowamok_context *ctx, flowamok_grid *g);   /* = u_init then u_iter */
int flowamok_band_u_init(flowamok_context *ctx, flowamok_grid *g);
int flowamok_band_u_iter(flowamok_context *ctx, flowamok_grid *g);
int flowamok_band_pack_u(flowamok_context *ctx, flowamok_grid *g, int side, void *dev_buf);
int flowamok_band_merge_u(flowamok_context *ctx, flowamok_grid *g, int side, const void *dev_buf, int wake, void *dev_woken);
int flowamok_band_woken(flowamok_context *ctx, flowamok_grid *g, void *dev_woken);
int flowamok_band_u_resume(flowamok_context *ctx, flowamok_grid *g);
int flowamok_band_rings(flowamok_context *ctx, flowamok_grid *g);
int flowamok_band_move(flowamok_context *ctx, flowamok_grid *g);
int flowamok_band_leaks(flowamok_context *ctx, flowamok_grid *g, int64_t *n_leaks);
/* The same protocol driven natively: NCCL send/recv + one all-reduced word on the context's stream, no host code
 * per phase.  comm_id (rank 0) produces the 128-byte NCCL unique id to broadcast; comm_init joins the communicator;
 * band_step runs n_ticks (rounds: total convergence rounds so far, may be NULL). */
int flowamok_band_comm_id(void *id128);
int flowamok_band_comm_init(flowamok_context *ctx, const void *id128, int rank, int world);
int flowamok_band_step(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, int64_t *rounds);
/* ---- P2P transport (NVLink / NVSwitch peer memory; csrc/fa_band.cuh): the protocol above without NCCL and without the
 * host.  Every band owns a mailbox the other bands map; senders store rows and message numbers into the receiver's
 * mailbox, receivers poll their own memory on the device; the convergence rounds of the U phase and their termination
 * vote run inside one persistent kernel.  Set-up, on every band:
 *     set_cuts(world, cuts)            row boundaries of ALL bands; then install the cycles (flowamok_set_cycles with the
 *                                      GLOBAL list, or flowamok_generate_synthetic): rings may straddle cuts
 *     p2p_open(rank, world, h, &ptr)   allocates + exports the mailbox (64-byte CUDA IPC handle / raw device pointer)
 *     p2p_connect(handles, ptrs, k)    maps the peers (handles gathered from all ranks, or pointers inside one process)
 * afterwards flowamok_band_step runs whole ticks with no host synchronisation (rounds == NULL keeps it asynchronous). */
int flowamok_band_set_cuts(flowamok_context *ctx, flowamok_grid *g, int world, const int64_t *cuts);
int flowamok_band_p2p_open(flowamok_context *ctx, flowamok_grid *g, int rank, int world, void *handle64, void **local_ptr);
int flowamok_band_p2p_connect(flowamok_context *ctx, flowamok_grid *g, const void *handles, void *const *local_ptrs, int bands_on_device);
/* flowamok_step_profile for a band on the P2P transport: summed device ms of {k_band_xstate, k_u_local, k_u_iter with the
 * convergence rounds, ring resolution with its swaps, k_move} over n_ticks; every band must call it */
int flowamok_band_step_profile(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, double *ms5);
/* unmap the peers (every band; then synchronise the processes before any grid is freed) */
int flowamok_band_p2p_close(flowamok_context *ctx, flowamok_grid *g);
/* rings that straddle a cut under the host-driven protocol: verdict words of this band (after flowamok_band_rings),
 * to be AND-ed over all bands and applied */
int flowamok_band_ring_words(flowamok_context *ctx, const flowamok_grid *g, int64_t *n_words);
int flowamok_band_ring_verdicts(flowamok_context *ctx, flowamok_grid *g, void *dev_buf);
int flowamok_band_rings_apply(flowamok_context *ctx, flowamok_grid *g, const void *dev_pass);
/* car state only (heading, colour, preference, rng) of the owned rows from whole-grid host arrays; roads, cycles
 * and halo rows are left alone.  Synchronous. */
int flowamok_grid_upload_cars(flowamok_context *ctx, flowamok_grid *g, const int8_t *dy, const int8_t *dx, const uint32_t *color,
                              const uint8_t *pref, const uint64_t *rng_state);

/* ---- the frontends' `step s n` loops (explorer/explorer.fut:109-130, designer/designer.fut:59-75) and scenario.init, so that a
 * lys C main (the lys wrapper itself is generated from diku-dk/lys 0.1.24, not vendored) steps through this library:
 * explorer_step: n_steps x { stepper_{perfect,quick}.step; scenario.step id grid step_cur rng; find_cycles if asked }; in/out
 * *step_cur and the explorer's rng (state, inc); out the cell leaks of this call.  designer_step: n_steps x { a car of random
 * colour at (25, 0) every 10th step; stepper_quick.step }.  The scenario rng arithmetic is cpprandom's (unpinned). */
int flowamok_scenario_init(flowamok_context *ctx, flowamok_grid *g, int scenario_id);
int flowamok_explorer_step(flowamok_context *ctx, flowamok_grid *g, int scenario_id, int perfect, int64_t n_steps, int64_t *step_cur,
                           uint64_t *rng_state, uint64_t *rng_inc, int64_t *n_cell_leaks);
int flowamok_designer_step(flowamok_context *ctx, flowamok_grid *g, int64_t n_steps, int64_t *n_steps_total, uint64_t *rng_state,
                           uint64_t *rng_inc, int32_t *n_leaks);

/* library identification */
const char *flowamok_version(void);

#ifdef __cplusplus
}
#endif
#endif
