// flowamok_b200 -- C ABI (include/flowamok.h) over the CUDA kernels.  Host-side state management,
// packing, the sparse find_cycles and the per-tick launch sequence.  No CPU compute fallback exists:
// every stepping call needs a CUDA device.
#include "../../include/flowamok.h"

#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_set>
#include <vector>

#include "fa_kernels.cuh"

using namespace fa;

// NCCL is reached through dlopen: the process (torch) has already loaded its own libnccl.so.2, and the library
// must keep loading on machines without NCCL.
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
#define SYM(f, name) f = (decltype(f))dlsym(lib, name); if (!f) return false
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce"); SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};
static NcclApi g_nccl;

struct flowamok_context {
  ncclComm_t comm = nullptr;   // row-band communicator (flowamok_band_comm_init)
  int rank = 0, world = 1;
  int device;
  cudaStream_t stream;
  bool own_stream;
  // The tick forks after k_u_local: the global fixpoint (k_u_iter: latency-bound, one small CTA per SM) runs on `hi` (highest
  // priority: its cooperative grid is placed as soon as k_move's CTAs make room), k_move runs beside it on the main stream,
  // and the streams join before the fix-up (fa_stream.h, fix_word).
  cudaStream_t hi = nullptr;
  cudaEvent_t ev_local = nullptr, ev_fix = nullptr;
  bool split = true;
  int num_sms;
  int iter_blocks; // co-resident CTAs of k_u_iter
  int iter_blocks_p2p = 0;   // the same for a band with mapped peers (smaller when several bands share one device)
  int iter_blocks_big = 0;   // every CTA of k_u_iter an SM can hold: jammed networks (the fixpoint alone on the device)
  double jam_visits = 40000.0;   // visits per fixpoint iteration above which a tick runs in jam order (flowamok_set_jam_threshold)
  size_t smem_optin = 0;   // largest dynamic shared memory a CTA may ask for
  std::string err;
  std::vector<struct flowamok_grid *> grids;   // live grids: flowamok_context_sync reads their watchdog flags
  std::vector<struct flowamok_grid *> pool;    // recycled grids (flowamok_grid_recycle): flowamok_grid_clone reuses their planes
};

struct flowamok_grid {
  flowamok_context *ctx;
  int gh, gw, pitch, wvalid;   // gh = rows held locally (band + halo rows)
  int gh_global, row0, own0, own1, halo;
  size_t nwords, ncells_p;  // plane words, pitched cells
  U4 *road, *head[2];
  u32 *pref[2], *calc, *color[2], *leak_plane;
  unsigned char *dirty[2];   // per plane word: colour sectors that received a car in the last tick (lazy double buffer)
  u64 *mymx;
  u64 *rng;
  u64 rng_inc;
  int cur;
  int chooser;
  std::vector<uint8_t> tape;
  long long tick;
  // cycles (sparse)
  RingWord *ring_words;             // per ring: one entry per plane word it touches
  unsigned long long *ring_off;     // CSR offsets into ring_words, or nullptr with a uniform ring_stride
  long long n_rings, n_ring_cells, n_ring_words;
  int ring_stride;  // >0: uniform rings, ring_off == nullptr
  std::vector<int64_t> h_off, h_y, h_x;  // host copy for flowamok_get_cycles_dense
  std::vector<int8_t> h_cy, h_cx;
  // per-tick machinery
  IterState *is;
  Diag *diag;
  Diag *h_diag;  // pinned
  UEntry *table;          // worklist entries of the tick, indexed by plane word
  unsigned int *qlist[2]; // frontier lists (entry indices), ping-pong; k_u_local appends the first one to qlist[0]
  u32 *ustate;            // per word: U_NONE or entry index (| U_QUEUED)
  int nsx, nsy, rs;       // strips across / down, rows per strip (k_u_local)
  int mtx, mty, rs_m;     // k_move tiles (M_WARPS strips wide, rs_m rows)
  // speculative move (fa_stream.h, fix_word): the fixpoint's delta plane, the words that have a delta, fix-up claims
  u64 *mymx_B = nullptr;
  u32 *chg_list = nullptr, *fix_claim = nullptr;
  u32 fix_epoch = 0;                     // tag of the current tick's claims (never reset, unlike `tick`)
  Diag base;  // diag values already reported
  // native band stepping: message buffers [side] and the all-reduced convergence word
  char *msg_s_send[2] = {nullptr, nullptr}, *msg_s_recv[2] = {nullptr, nullptr};
  char *msg_u_send[2] = {nullptr, nullptr}, *msg_u_recv[2] = {nullptr, nullptr};
  unsigned int *d_woken = nullptr, *h_woken = nullptr;
  long long band_rounds = 0;
  int ring_touch = 1;   // some ring cell lies on a boundary row of this band (1 = unknown/assume yes)
  int ring_swap = -1;   // collective decision (max over ranks of ring_touch); -1 = not agreed yet
  // rings that straddle a cut (judged by every band they touch, AND-ed, then marked)
  std::vector<int64_t> cuts;        // row boundaries of all bands (world + 1 entries), empty: unknown
  int *ring_sid = nullptr;          // per listed ring: straddle id or -1
  u32 *ring_partial = nullptr, *ring_pass = nullptr;   // this band's verdicts / the AND over all bands (bit = straddle id)
  long long n_straddle = 0;         // straddle ids in use (the same number in every band)
  int sbit_words = 0;               // 32-bit words of ring_partial / ring_pass
  // P2P transport (fa_band.cuh): my mailbox, the peers' mailboxes, device-side protocol state
  bool p2p = false;
  BandLink link;
  char *mailbox = nullptr;
  size_t mailbox_bytes = 0;
  int sbit_words_cap = 0;
  BandDev *band_dev = nullptr;
  void *ipc_opened[BAND_MAXW] = {};
  unsigned int s_seq = 0;           // state messages sent so far (= ticks stepped over P2P)
  int p2p_blocks = 0;               // cooperative grid of k_u_iter<true>
  // Jam detection: k_u_iter stores its cumulative {iterations, visits} into pinned host memory when it ends (posted stores, no
  // copy, no synchronisation: an asynchronous 16-byte copy on the fixpoint's stream cost 34 us per tick); the host reads
  // whatever has arrived when it enqueues a tick -- a few ticks late is fine.  A frontier of tens of thousands of words per iteration
  // needs the whole device: the tick then runs k_u_iter at full size before k_move instead of a small one beside it.
  unsigned long long *h_fix = nullptr, fix_seen[2] = {0, 0};
  bool jam = false;
  long long jam_ticks = 0;          // ticks stepped in jam order
  int8_t *stage = nullptr;          // staging buffer of upload / download (grows, never shrinks; no cudaMalloc per call)
  size_t stage_bytes = 0;
  const char *poisoned = nullptr;   // a device-side watchdog tripped: the planes are not trustworthy any more
  char poison_text[160];
};

static int fail(flowamok_context *ctx, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return 1;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// scratch device memory / events of one call: released on every return path
struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T> T *as() const { return (T *)p; }
};
struct EventSet {
  std::vector<cudaEvent_t> ev;
  ~EventSet() { for (auto e : ev) if (e) cudaEventDestroy(e); }
  cudaError_t create(size_t n) {
    ev.assign(n, nullptr);
    for (auto &e : ev) { cudaError_t r = cudaEventCreate(&e); if (r != cudaSuccess) return r; }
    return cudaSuccess;
  }
};

static inline unsigned int nblk(long long n, int b) { return (unsigned int)((n + b - 1) / b); }
// rows per k_u_local strip: long strips amortise the 2K+3 pipeline-fill rows every strip pays, but there must be
// enough strips (warps) to occupy every SM.  FLOWAMOK_STRIP_ROWS overrides (tuning).
static int strip_rows(const flowamok_context *ctx, int owned_rows, int pitch) {
  if (const char *e = getenv("FLOWAMOK_STRIP_ROWS")) { int v = atoi(e); if (v >= 1) return v; }
  const long long nsx = (pitch + STRIP_W - 1) / STRIP_W;
  const long long want = (long long)ctx->num_sms * 12;   // about the number of resident warps per SM
  int rs = 128;
  while (rs > 16 && nsx * ((owned_rows + rs - 1) / rs) < want) rs >>= 1;
  return rs;
}
struct flowamok_grid;
static int sync_color_buffers(flowamok_context *ctx, flowamok_grid *g);
static int ensure_stage(flowamok_context *ctx, flowamok_grid *g, size_t bytes);
extern "C" { static int check_device_errors(flowamok_context *ctx, flowamok_grid *g); }

// FLOWAMOK_DEBUG_SYNC=1: synchronise after every kernel so that a fault is reported at its launch site
static bool debug_sync() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("FLOWAMOK_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
#define KCHECK(name)                                                                              \
  do {                                                                                            \
    CU(cudaGetLastError());                                                                       \
    if (debug_sync()) {                                                                           \
      cudaError_t e2_ = cudaStreamSynchronize(ctx->stream);                                       \
      if (e2_ != cudaSuccess) return fail(ctx, "kernel %s faulted: %s", name, cudaGetErrorString(e2_)); \
    }                                                                                             \
  } while (0)

extern "C" {

const char *flowamok_version(void) { return "flowamok_b200 0.1 (sm_100a)"; }

int flowamok_context_new(flowamok_context **out, int device, void *stream) {
  if (!out) return 1;
  flowamok_context *ctx = new flowamok_context();
  *out = ctx;
  ctx->own_stream = false;
  ctx->stream = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    ctx->device = -1;
    return fail(ctx, "no CUDA device: flowamok_b200 has no CPU path (%s)", cudaGetErrorString(e));
  }
  if (device < 0) CU(cudaGetDevice(&device));
  ctx->device = device;
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  ctx->num_sms = prop.multiProcessorCount;
  if (stream) ctx->stream = (cudaStream_t)stream;
  else { CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
  {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo: least priority (numerically largest), hi: greatest
    CU(cudaStreamCreateWithPriority(&ctx->hi, cudaStreamNonBlocking, hi));
    CU(cudaEventCreateWithFlags(&ctx->ev_local, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_fix, cudaEventDisableTiming));
    const char *e = getenv("FLOWAMOK_SPLIT");
    ctx->split = !(e && e[0] == '0');
  }
  CU(cudaFuncSetAttribute(k_move<M_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
  ctx->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_u_iter<false>, UIT_THREADS, 0));
  if (per_sm < 1) return fail(ctx, "k_u_iter does not fit on an SM");
  ctx->iter_blocks = ctx->num_sms;   // ONE small CTA per SM: k_move runs beside the fixpoint
  ctx->iter_blocks_big = std::min(per_sm, 4) * ctx->num_sms;   // 4 x 128 threads per SM: more CTAs only make the grid barrier dearer
  {   // Load every kernel of the stepping path NOW.  With CUDA's lazy module loading the first launch of a kernel may have to
      // wait for running kernels to finish -- and the band kernels wait for each other on the device, so a load in the middle
      // of a tick could wait forever for a kernel whose neighbour has not been enqueued yet.
    const void *fns[] = {(const void *)k_u_local<false>, (const void *)k_u_local<true>, (const void *)k_u_iter<false>,
                         (const void *)k_u_iter<true>, (const void *)k_rings, (const void *)k_rings_mark, (const void *)k_move<M_WARPS>,
                         (const void *)k_band_xstate, (const void *)k_band_xu, (const void *)k_band_xring, (const void *)k_band_state,
                         (const void *)k_band_pack_u, (const void *)k_band_merge, (const void *)k_digest, (const void *)k_pack, (const void *)k_move_fix,
                         (const void *)k_unpack};
    for (const void *f : fns) {
      cudaFuncAttributes at;
      CU(cudaFuncGetAttributes(&at, f));
    }
  }
  int per_sm_p2p = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_p2p, k_u_iter<true>, UIT_THREADS, 0));
  if (per_sm_p2p < 1) return fail(ctx, "k_u_iter<p2p> does not fit on an SM");
  ctx->iter_blocks_p2p = ctx->num_sms;
  return 0;
}

int flowamok_context_free(flowamok_context *ctx) {
  if (!ctx) return 0;
  while (!ctx->pool.empty()) { flowamok_grid *g = ctx->pool.back(); ctx->pool.pop_back(); flowamok_grid_free(ctx, g); }
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  if (ctx->device >= 0) {
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->hi) { cudaStreamSynchronize(ctx->hi); cudaStreamDestroy(ctx->hi); }
    if (ctx->ev_local) cudaEventDestroy(ctx->ev_local);
    if (ctx->ev_fix) cudaEventDestroy(ctx->ev_fix);
  }
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}
int flowamok_context_sync(flowamok_context *ctx) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  CU(cudaStreamSynchronize(ctx->stream));
  for (flowamok_grid *g : ctx->grids)
    if (check_device_errors(ctx, g)) return 1;
  return 0;
}
char *flowamok_context_get_error(flowamok_context *ctx) {
  if (!ctx || ctx->err.empty()) return nullptr;
  char *s = (char *)malloc(ctx->err.size() + 1);
  memcpy(s, ctx->err.c_str(), ctx->err.size() + 1);
  ctx->err.clear();
  return s;
}
int flowamok_context_device(flowamok_context *ctx) { return ctx ? ctx->device : -1; }
/* tuning knob: a tick whose recent fixpoint iterations visited more than this many words each runs k_u_iter at full size before
 * k_move instead of a small one beside a speculative k_move (default 40000; results are identical either way) */
int flowamok_jam_ticks(flowamok_context *ctx, const flowamok_grid *g, int64_t *n) {
  if (!g || !n) return fail(ctx, "null argument");
  *n = g->jam_ticks;
  return 0;
}
int flowamok_set_jam_threshold(flowamok_context *ctx, double visits_per_iteration) {
  if (!ctx || !(visits_per_iteration > 0.0)) return 1;
  ctx->jam_visits = visits_per_iteration;
  return 0;
}

// ---------------------------------------------------------------------------------------------
}  // extern "C"

static int grid_alloc(flowamok_context *ctx, flowamok_grid **out, int64_t gh_global, int64_t gw, int64_t row_begin, int64_t row_end, int halo);
// a grid that could not be allocated completely is freed again: *out is either a whole grid or null
static int grid_new_impl(flowamok_context *ctx, flowamok_grid **out, int64_t gh_global, int64_t gw, int64_t row_begin,
                         int64_t row_end, int halo) {
  *out = nullptr;
  flowamok_grid *g = nullptr;
  if (grid_alloc(ctx, &g, gh_global, gw, row_begin, row_end, halo)) {
    if (g) {
      std::string keep = ctx->err;
      flowamok_grid_free(ctx, g);
      ctx->err = keep;
    }
    return 1;
  }
  *out = g;
  return 0;
}
static int grid_alloc(flowamok_context *ctx, flowamok_grid **out, int64_t gh_global, int64_t gw, int64_t row_begin,
                      int64_t row_end, int halo) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (gh_global < 1 || gw < 1 || gh_global > (1 << 20) || gw > (1 << 20)) return fail(ctx, "grid shape out of range");
  if (row_begin < 0 || row_end > gh_global || row_end <= row_begin) return fail(ctx, "bad row band");
  CU(cudaSetDevice(ctx->device));
  flowamok_grid *g = new flowamok_grid();
  g->ctx = ctx;
  g->gh_global = (int)gh_global; g->halo = halo;
  g->row0 = (int)row_begin - halo; g->own0 = halo; g->own1 = halo + (int)(row_end - row_begin);
  g->gh = g->own1 + halo; g->gw = (int)gw;
  g->wvalid = (int)((gw + 31) / 32);
  g->pitch = ((g->wvalid + 3) / 4) * 4;
  g->nwords = (size_t)g->gh * g->pitch;
  g->ncells_p = g->nwords * 32;
  if (g->nwords >= (1ull << 32)) { delete g; return fail(ctx, "grid too large for 32-bit word indices"); }
  g->cur = 0; g->chooser = CHOOSE_RANDOM; g->tick = 0;
  g->rng_inc = (0xda3e39cb94b95bdbULL << 1) | 1ULL;
  g->ring_words = nullptr; g->ring_off = nullptr; g->n_rings = g->n_ring_cells = g->n_ring_words = 0; g->ring_stride = 0;
  g->leak_plane = nullptr;
  memset(&g->link, 0, sizeof g->link);
  *out = g;
  CU(cudaMalloc(&g->road, g->nwords * sizeof(U4)));
  for (int k = 0; k < 2; k++) {
    CU(cudaMalloc(&g->head[k], g->nwords * sizeof(U4)));
    CU(cudaMalloc(&g->pref[k], g->nwords * 4));
    CU(cudaMalloc(&g->color[k], g->ncells_p * 4));
    CU(cudaMalloc(&g->dirty[k], g->nwords));
  }
  CU(cudaMalloc(&g->calc, g->nwords * 4));
  CU(cudaMalloc(&g->mymx, g->nwords * 8));
  CU(cudaMalloc(&g->rng, g->ncells_p * 8));
  const int owned = g->own1 - g->own0;
  g->rs = strip_rows(ctx, owned, g->pitch);
  g->nsx = (g->pitch + STRIP_W - 1) / STRIP_W; g->nsy = (owned + g->rs - 1) / g->rs;
  g->rs_m = 8;    // k_move tile rows: 31 KB of shared memory per CTA (descriptors + lists)
  if (const char *e = getenv("FLOWAMOK_MOVE_ROWS")) { int v = atoi(e); if (v >= 1 && v <= 128) g->rs_m = v; }
  while (g->rs_m > 1 && MTL::Smem::bytes(g->rs_m) > ctx->smem_optin) g->rs_m--;
  g->mtx = (g->nsx + M_WARPS - 1) / M_WARPS; g->mty = (owned + g->rs_m - 1) / g->rs_m;
  CU(cudaMalloc(&g->mymx_B, g->nwords * 8));
  CU(cudaMalloc(&g->chg_list, g->nwords * 4));
  CU(cudaMalloc(&g->fix_claim, g->nwords * 4));
  CU(cudaMemsetAsync(g->mymx_B, 0, g->nwords * 8, ctx->stream));
  CU(cudaMemsetAsync(g->fix_claim, 0, g->nwords * 4, ctx->stream));
  CU(cudaMalloc(&g->table, g->nwords * sizeof(UEntry)));
  {   // frontier lists; qlist[1] doubles as the per-strip staging area of k_u_local (STRIP_W * rs slots per strip)
    const size_t nslots = std::max(g->nwords, (size_t)g->nsx * g->nsy * STRIP_W * g->rs);
    CU(cudaMalloc(&g->qlist[0], nslots * 4));
    CU(cudaMalloc(&g->qlist[1], nslots * 4));
  }
  CU(cudaMalloc(&g->ustate, g->nwords * 4));
  CU(cudaMalloc(&g->is, sizeof(IterState)));
  CU(cudaMalloc(&g->diag, sizeof(Diag)));
  CU(cudaMallocHost(&g->h_diag, sizeof(Diag)));
  CU(cudaMallocHost(&g->h_fix, 16));
  g->h_fix[0] = g->h_fix[1] = 0;
  memset(&g->base, 0, sizeof(Diag));
  cudaStream_t st = ctx->stream;
  CU(cudaMemsetAsync(g->road, 0, g->nwords * sizeof(U4), st));
  for (int k = 0; k < 2; k++) {
    CU(cudaMemsetAsync(g->head[k], 0, g->nwords * sizeof(U4), st));
    CU(cudaMemsetAsync(g->pref[k], 0, g->nwords * 4, st));
    CU(cudaMemsetAsync(g->color[k], 0, g->ncells_p * 4, st));
    CU(cudaMemsetAsync(g->dirty[k], 0, g->nwords, st));
  }
  CU(cudaMemsetAsync(g->calc, 0, g->nwords * 4, st));
  CU(cudaMemsetAsync(g->mymx, 0, g->nwords * 8, st));
  CU(cudaMemsetAsync(g->rng, 0, g->ncells_p * 8, st));
  CU(cudaMemsetAsync(g->is, 0, sizeof(IterState), st));
  CU(cudaMemsetAsync(g->ustate, 0xFF, g->nwords * 4, st));   // U_NONE: halo rows never get entries
  CU(cudaMemsetAsync(g->diag, 0, sizeof(Diag), st));
  CU(cudaStreamSynchronize(st));
  ctx->grids.push_back(g);
  return 0;
}

static int ensure_stage(flowamok_context *ctx, flowamok_grid *g, size_t bytes) {
  if (bytes <= g->stage_bytes) return 0;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaFree(g->stage));
  g->stage = nullptr; g->stage_bytes = 0;
  CU(cudaMalloc(&g->stage, bytes));
  g->stage_bytes = bytes;
  return 0;
}

// After the colour plane of the current buffer was (re)written from outside, make the other buffer identical and
// forget the dirty sectors: the lazy double buffer's invariant (k_move only rewrites the 32-byte sectors that received a
// car in this tick or the previous one).
static int sync_color_buffers(flowamok_context *ctx, flowamok_grid *g) {
  CU(cudaMemcpyAsync(g->color[g->cur ^ 1], g->color[g->cur], g->ncells_p * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemsetAsync(g->dirty[0], 0, g->nwords, ctx->stream));
  CU(cudaMemsetAsync(g->dirty[1], 0, g->nwords, ctx->stream));
  return 0;
}

extern "C" {

int flowamok_grid_new(flowamok_context *ctx, flowamok_grid **out, int64_t gh, int64_t gw) {
  return grid_new_impl(ctx, out, gh, gw, 0, gh, 0);
}
/* a band of a larger grid: rows [row_begin, row_end) plus two halo rows on each side */
int flowamok_band_new(flowamok_context *ctx, flowamok_grid **out, int64_t gh_global, int64_t gw, int64_t row_begin,
                      int64_t row_end) {
  return grid_new_impl(ctx, out, gh_global, gw, row_begin, row_end, 2);
}

int flowamok_grid_free(flowamok_context *ctx, flowamok_grid *g) {
  if (!g) return 0;
  if (ctx) ctx->grids.erase(std::remove(ctx->grids.begin(), ctx->grids.end(), g), ctx->grids.end());
  if (ctx && ctx->device >= 0) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(g->road);
  for (int k = 0; k < 2; k++) { cudaFree(g->head[k]); cudaFree(g->pref[k]); cudaFree(g->color[k]); cudaFree(g->dirty[k]); }
  cudaFree(g->calc); cudaFree(g->mymx); cudaFree(g->rng); cudaFree(g->is); cudaFree(g->diag);
  for (int k = 0; k < 2; k++) { cudaFree(g->msg_s_send[k]); cudaFree(g->msg_s_recv[k]); cudaFree(g->msg_u_send[k]); cudaFree(g->msg_u_recv[k]); }
  cudaFree(g->d_woken); if (g->h_woken) cudaFreeHost(g->h_woken);
  for (int r = 0; r < BAND_MAXW; r++) if (g->ipc_opened[r]) cudaIpcCloseMemHandle(g->ipc_opened[r]);
  cudaFree(g->mailbox); cudaFree(g->band_dev); cudaFree(g->ring_sid); cudaFree(g->ring_partial); cudaFree(g->ring_pass);
  cudaFree(g->table); cudaFree(g->qlist[0]); cudaFree(g->qlist[1]); cudaFree(g->ustate); cudaFree(g->mymx_B); cudaFree(g->chg_list); cudaFree(g->fix_claim);
  cudaFree(g->ring_words); cudaFree(g->ring_off); cudaFree(g->leak_plane); cudaFree(g->stage);
  cudaFreeHost(g->h_diag);
  if (g->h_fix) cudaFreeHost(g->h_fix);
  delete g;
  return 0;
}

/* Like flowamok_grid_free, but the planes are kept for the next flowamok_grid_clone of a grid of the same shape (animation.fut:23
 * copies the grid every frame: the generated-library surface would otherwise allocate and clear 19 bytes per cell per frame). */
int flowamok_grid_recycle(flowamok_context *ctx, flowamok_grid *g) {
  if (!g) return 0;
  if (!ctx || ctx->device < 0 || g->poisoned || g->mailbox || ctx->pool.size() >= 2) return flowamok_grid_free(ctx, g);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->grids.erase(std::remove(ctx->grids.begin(), ctx->grids.end(), g), ctx->grids.end());
  ctx->pool.push_back(g);
  return 0;
}

int flowamok_grid_shape(flowamok_context *ctx, const flowamok_grid *g, int64_t *gh, int64_t *gw) {
  if (!g) return fail(ctx, "null grid");
  if (gh) *gh = g->gh_global;
  if (gw) *gw = g->gw;
  return 0;
}

int flowamok_create_grid(flowamok_context *ctx, flowamok_grid *g, int32_t seed, uint64_t *joined_state, uint64_t *joined_inc) {
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  u64 s0, inc;
  pcg32_from_seed(seed, s0, inc);
  g->rng_inc = inc;
  CU(cudaMemsetAsync(g->road, 0, g->nwords * sizeof(U4), st));
  CU(cudaMemsetAsync(g->head[g->cur], 0, g->nwords * sizeof(U4), st));
  CU(cudaMemsetAsync(g->pref[g->cur], 0, g->nwords * 4, st));
  CU(cudaMemsetAsync(g->color[g->cur], 0, g->ncells_p * 4, st));
  CU(cudaMemsetAsync(g->rng, 0, g->ncells_p * 8, st));
  long long n = (long long)g->gh * g->gw;
  k_create<<<nblk(n, 256), 256, 0, st>>>(g->gh, g->gw, g->pitch, g->row0, g->gh_global, g->color[g->cur], g->rng, s0, inc);
  CU(cudaGetLastError());
  if (sync_color_buffers(ctx, g)) return 1;
  CU(cudaStreamSynchronize(st));
  if (joined_state || joined_inc) {  // join_rng (utils.fut:17): product of states, OR of incs
    u64 js = 1;
    for (long long i = 0; i < (long long)g->gh_global * g->gw; i++) js *= pcg32_split_state(s0, inc, i);
    if (joined_state) *joined_state = js;
    if (joined_inc) *joined_inc = inc;
  }
  g->tick = 0;
  return 0;
}

int flowamok_grid_upload(flowamok_context *ctx, flowamok_grid *g, const int8_t *uy, const int8_t *ux,
                         const int8_t *dy, const int8_t *dx, const uint32_t *color, const uint8_t *pref,
                         const uint64_t *rng_state, uint64_t rng_inc) {
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // host arrays are [gh_global][gw]; this device takes the rows its planes hold (band + halo rows)
  const int lo = std::max(0, g->row0), hi = std::min(g->gh_global, g->row0 + g->gh);   // global rows present
  const size_t nrow = (size_t)(hi - lo), n = nrow * g->gw;
  if (ensure_stage(ctx, g, n * 5)) return 1;
  int8_t *tmp = g->stage;
  const void *src[5] = {uy, ux, dy, dx, pref};
  int8_t *dptr[5];
  for (int k = 0; k < 5; k++) {
    dptr[k] = src[k] ? tmp + n * k : nullptr;
    if (src[k]) CU(cudaMemcpyAsync(dptr[k], (const int8_t *)src[k] + (size_t)lo * g->gw, n, cudaMemcpyHostToDevice, st));
  }
  // the staging arrays start at global row lo: present them to k_pack as a grid whose row 0 is global row lo
  k_pack<<<nblk((long long)g->nwords * 32, 256), 256, 0, st>>>(g->gh, g->gw, g->pitch, g->row0 - lo, (int)nrow, dptr[0], dptr[1],
                                                              dptr[2], dptr[3], (const uint8_t *)dptr[4], g->road,
                                                              g->head[g->cur], g->pref[g->cur]);
  CU(cudaGetLastError());
  size_t cp = (size_t)g->pitch * 32;
  const int l0 = lo - g->row0;   // local row of global row lo
  CU(cudaMemsetAsync(g->color[g->cur], color ? 0 : 0xFF, g->ncells_p * 4, st));
  if (rng_state) CU(cudaMemsetAsync(g->rng, 0, g->ncells_p * 8, st));   // NULL: the resident states stay (see flowamok_grid_rng_*)
  if (color)
    CU(cudaMemcpy2DAsync(g->color[g->cur] + (size_t)l0 * cp, cp * 4, color + (size_t)lo * g->gw, (size_t)g->gw * 4,
                         (size_t)g->gw * 4, nrow, cudaMemcpyHostToDevice, st));
  if (rng_state)
    CU(cudaMemcpy2DAsync(g->rng + (size_t)l0 * cp, cp * 8, rng_state + (size_t)lo * g->gw, (size_t)g->gw * 8,
                         (size_t)g->gw * 8, nrow, cudaMemcpyHostToDevice, st));
  g->rng_inc = rng_inc;
  if (sync_color_buffers(ctx, g)) return 1;
  CU(cudaStreamSynchronize(st));
  return 0;
}

int flowamok_grid_download(flowamok_context *ctx, const flowamok_grid *g, int8_t *uy, int8_t *ux, int8_t *dy,
                           int8_t *dx, uint32_t *color, uint8_t *pref, uint64_t *rng_state, uint64_t *rng_inc) {
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // host arrays are [gh_global][gw]; only the OWNED rows are written
  const int owned = g->own1 - g->own0, grow = g->row0 + g->own0;
  size_t n = (size_t)owned * g->gw;
  void *dst[5] = {uy, ux, dy, dx, pref};
  bool any = false;
  for (int k = 0; k < 5; k++) any |= dst[k] != nullptr;
  int8_t *tmp = nullptr;
  if (any) {
    if (ensure_stage(ctx, const_cast<flowamok_grid *>(g), n * 5)) return 1;
    tmp = g->stage;
    k_unpack<<<nblk((long long)n, 256), 256, 0, st>>>(g->own0, g->own1, g->gw, g->pitch, g->road, g->head[g->cur], g->pref[g->cur],
                                                      uy ? tmp : nullptr, ux ? tmp + n : nullptr, dy ? tmp + 2 * n : nullptr,
                                                      dx ? tmp + 3 * n : nullptr, pref ? (uint8_t *)(tmp + 4 * n) : nullptr);
    CU(cudaGetLastError());
    for (int k = 0; k < 5; k++)
      if (dst[k]) CU(cudaMemcpyAsync((int8_t *)dst[k] + (size_t)grow * g->gw, tmp + n * k, n, cudaMemcpyDeviceToHost, st));
  }
  size_t cp = (size_t)g->pitch * 32;
  if (color)
    CU(cudaMemcpy2DAsync(color + (size_t)grow * g->gw, (size_t)g->gw * 4, g->color[g->cur] + (size_t)g->own0 * cp, cp * 4,
                         (size_t)g->gw * 4, owned, cudaMemcpyDeviceToHost, st));
  if (rng_state)
    CU(cudaMemcpy2DAsync(rng_state + (size_t)grow * g->gw, (size_t)g->gw * 8, g->rng + (size_t)g->own0 * cp, cp * 8,
                         (size_t)g->gw * 8, owned, cudaMemcpyDeviceToHost, st));
  if (rng_inc) *rng_inc = g->rng_inc;
  CU(cudaStreamSynchronize(st));
  return check_device_errors(ctx, const_cast<flowamok_grid *>(g));   // a synchronisation point: watchdog flags are fatal
}

int flowamok_add_line_vertical(flowamok_context *ctx, flowamok_grid *g, int64_t y0, int64_t y1, int64_t x, int8_t flow) {
  if (!g) return fail(ctx, "null grid");
  if (y1 < y0) return 0;  // n <= 0: the reference loop body never runs (utils.fut:21-23)
  if (y0 < 0 || y1 >= g->gh_global || x < 0 || x >= g->gw) return fail(ctx, "add_line_vertical: index out of bounds");
  CU(cudaSetDevice(ctx->device));
  int64_t a = std::max<int64_t>(y0, g->row0), b = std::min<int64_t>(y1, (int64_t)g->row0 + g->gh - 1);   // rows held here
  if (b < a) return 0;
  int n = (int)(b - a + 1);
  k_line_v<<<nblk(n, 128), 128, 0, ctx->stream>>>(g->pitch, g->road, (int)(a - g->row0), n, (int)x, flow);
  CU(cudaGetLastError());
  return 0;
}
int flowamok_add_line_horizontal(flowamok_context *ctx, flowamok_grid *g, int64_t y, int64_t x0, int64_t x1, int8_t flow) {
  if (!g) return fail(ctx, "null grid");
  if (x1 < x0) return 0;
  if (x0 < 0 || x1 >= g->gw || y < 0 || y >= g->gh_global) return fail(ctx, "add_line_horizontal: index out of bounds");
  CU(cudaSetDevice(ctx->device));
  if (y < g->row0 || y >= (int64_t)g->row0 + g->gh) return 0;   // not held by this band
  int nw = (int)((x1 >> 5) - (x0 >> 5) + 1);
  k_line_h<<<nblk(nw, 128), 128, 0, ctx->stream>>>(g->pitch, g->road, (int)(y - g->row0), (int)x0, (int)x1, flow);
  CU(cudaGetLastError());
  return 0;
}
int flowamok_add_cell(flowamok_context *ctx, flowamok_grid *g, int64_t y, int64_t x, uint32_t color, int8_t dy, int8_t dx) {
  if (!g) return fail(ctx, "null grid");
  if (y < 0 || y >= g->gh_global || x < 0 || x >= g->gw) return fail(ctx, "add_cell: index out of bounds");
  CU(cudaSetDevice(ctx->device));
  if (y < g->row0 || y >= (int64_t)g->row0 + g->gh) return 0;   // not held by this band
  k_add_cell<<<1, 1, 0, ctx->stream>>>(g->pitch, g->head[g->cur], g->color[g->cur], g->color[g->cur ^ 1], (int)(y - g->row0), (int)x, color, dy, dx);
  CU(cudaGetLastError());
  return 0;
}

int flowamok_set_chooser(flowamok_context *ctx, flowamok_grid *g, int mode, const uint8_t *tape, int64_t tape_len) {
  if (!g) return fail(ctx, "null grid");
  if (mode < 0 || mode > 3) return fail(ctx, "unknown chooser mode %d", mode);
  if (mode == CHOOSE_TAPE && (!tape || tape_len <= 0)) return fail(ctx, "tape chooser needs a tape");
  g->chooser = mode;
  g->tape.clear();
  if (mode == CHOOSE_TAPE) g->tape.assign(tape, tape + tape_len);
  g->tick = 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// cycles
// ---------------------------------------------------------------------------------------------
static int band_of_row(const std::vector<int64_t> &cuts, int64_t y) {
  int r = 0;
  while (r + 2 < (int)cuts.size() && y >= cuts[(size_t)r + 1]) r++;
  return r;
}
// (re)allocate the verdict bits of the straddling rings: all ones = "no objection"
static int alloc_straddle_bits(flowamok_context *ctx, flowamok_grid *g, long long n_straddle, const std::vector<int> &sids) {
  CU(cudaFree(g->ring_sid)); CU(cudaFree(g->ring_partial)); CU(cudaFree(g->ring_pass));
  g->ring_sid = nullptr; g->ring_partial = nullptr; g->ring_pass = nullptr;
  g->n_straddle = n_straddle; g->sbit_words = (int)((n_straddle + 31) / 32);
  if (n_straddle <= 0) return 0;
  if (g->p2p && g->sbit_words > g->sbit_words_cap)
    return fail(ctx, "%lld rings straddle band cuts, more than the connected mailboxes hold: install the cycles before flowamok_band_p2p_open",
                n_straddle);
  if (!sids.empty()) {
    CU(cudaMalloc(&g->ring_sid, sids.size() * sizeof(int)));
    CU(cudaMemcpy(g->ring_sid, sids.data(), sids.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  CU(cudaMalloc(&g->ring_partial, (size_t)g->sbit_words * 4));
  CU(cudaMalloc(&g->ring_pass, (size_t)g->sbit_words * 4));
  CU(cudaMemset(g->ring_partial, 0xFF, (size_t)g->sbit_words * 4));
  CU(cudaMemset(g->ring_pass, 0, (size_t)g->sbit_words * 4));
  g->link.ring_words = g->sbit_words;
  return 0;
}

// The cycle list is GLOBAL (the same list on every band).  A band keeps the rings that have a cell in its rows, with
// the words of its own rows only; a ring whose rows lie in more than one band gets a straddle id -- its position among the
// straddling rings of the global list, hence the same number in every band (needs the cuts: flowamok_band_set_cuts).
static int install_rings(flowamok_context *ctx, flowamok_grid *g, int64_t n, const int64_t *off, const int64_t *y,
                         const int64_t *x, const int8_t *cy, const int8_t *cx, const uint8_t *grant) {
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaFree(g->ring_words)); CU(cudaFree(g->ring_off));
  g->ring_words = nullptr; g->ring_off = nullptr; g->n_rings = 0; g->n_ring_cells = 0; g->n_ring_words = 0; g->ring_stride = 0;
  g->h_off.clear(); g->h_y.clear(); g->h_x.clear(); g->h_cy.clear(); g->h_cx.clear();
  std::vector<int> sids;
  if (n <= 0) { g->ring_touch = 0; g->ring_swap = g->cuts.empty() ? -1 : 0; return alloc_straddle_bits(ctx, g, 0, sids); }
  const int64_t tot = off[n], lo = g->row0 + g->own0, hi = g->row0 + g->own1;
  const bool band = g->halo != 0, have_cuts = !g->cuts.empty();
  std::vector<RingWord> words;
  std::vector<unsigned long long> offs;
  long long n_straddle = 0, kept = 0;
  int touch = 0, touch_global = 0;
  for (int64_t k = 0; k < n; k++) {
    bool in = false, out = false;
    int64_t ymin = INT64_MAX, ymax = INT64_MIN;
    for (int64_t t = off[k]; t < off[k + 1]; t++) {
      if (y[t] < 0 || y[t] >= g->gh_global || x[t] < 0 || x[t] >= g->gw) return fail(ctx, "cycle cell out of bounds");
      const bool single = (cy[t] != 0) != (cx[t] != 0);
      if (!single) return fail(ctx, "cycle check direction must be single-axis (steppers.fut:239-294)");
      if (grant[t] != 1 && grant[t] != 2) return fail(ctx, "grant must be 1 (y) or 2 (x)");
      (y[t] >= lo && y[t] < hi ? in : out) = true;
      ymin = std::min(ymin, y[t]); ymax = std::max(ymax, y[t]);
      if (have_cuts)
        for (size_t c = 1; c + 1 < g->cuts.size(); c++) touch_global |= (y[t] == g->cuts[c] - 1 || y[t] == g->cuts[c]);
    }
    {   // A cycle must be CLOSED: every cell's check direction leads to another cell of the same cycle (what find_cycles
        // produces, steppers.fut:247-293).  k_rings relies on it: such a ring's cells are never calculated by the fixpoint
        // when they all head along it, so `calculated` need not be read (ring_word_ok).
      std::unordered_set<int64_t> cells;
      for (int64_t t = off[k]; t < off[k + 1]; t++) cells.insert(y[t] * (int64_t)g->gw + x[t]);
      for (int64_t t = off[k]; t < off[k + 1]; t++)
        if (!cells.count((y[t] + cy[t]) * (int64_t)g->gw + (x[t] + cx[t])) || x[t] + cx[t] < 0 || x[t] + cx[t] >= g->gw)
          return fail(ctx, "cycle %lld is not closed: the check direction of cell (%lld, %lld) leaves it", (long long)k, (long long)y[t],
                      (long long)x[t]);
    }
    int sid = -1;
    if (band && have_cuts) {
      if (ymin <= ymax && band_of_row(g->cuts, ymin) != band_of_row(g->cuts, ymax)) sid = (int)n_straddle++;
    } else if (band && in && out) {
      return fail(ctx, "a cycle straddles this row band: give every band the global cycle list after flowamok_band_set_cuts");
    }
    if (band && !in) continue;   // another band's ring
    offs.push_back(words.size());
    sids.push_back(sid);
    kept++;
    const size_t first = words.size();
    for (int64_t t = off[k]; t < off[k + 1]; t++) {
      if (y[t] < lo || y[t] >= hi) continue;   // the other band's part of a straddling ring
      const u32 wi = (u32)((size_t)(y[t] - g->row0) * g->pitch + (x[t] >> 5)), bit = 1u << (x[t] & 31);
      size_t j = first;
      while (j < words.size() && words[j].wi != wi) j++;
      if (j == words.size()) { RingWord e = {wi, 0, 0, 0, 0, 0, 0, 0}; words.push_back(e); }
      RingWord &e = words[j];
      if (cy[t] < 0) e.mN |= bit; else if (cy[t] > 0) e.mS |= bit; else if (cx[t] < 0) e.mW |= bit; else e.mE |= bit;
      if (grant[t] == 2) e.gX |= bit; else e.gY |= bit;
      if (y[t] == lo || y[t] == hi - 1) touch = 1;
    }
  }
  offs.push_back(words.size());
  if (kept > 0) {
    CU(cudaMalloc(&g->ring_words, std::max<size_t>(1, words.size()) * sizeof(RingWord)));
    CU(cudaMalloc(&g->ring_off, ((size_t)kept + 1) * 8));
    CU(cudaMemcpy(g->ring_words, words.data(), words.size() * sizeof(RingWord), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g->ring_off, offs.data(), ((size_t)kept + 1) * 8, cudaMemcpyHostToDevice));
  }
  g->n_rings = kept; g->n_ring_cells = tot; g->n_ring_words = (long long)words.size();
  g->ring_touch = touch;
  g->ring_swap = have_cuts ? ((touch_global || n_straddle > 0) ? 1 : 0) : -1;   // with the cuts every band decides alike
  g->h_off.assign(off, off + n + 1); g->h_y.assign(y, y + tot); g->h_x.assign(x, x + tot);
  g->h_cy.assign(cy, cy + tot); g->h_cx.assign(cx, cx + tot);
  bool any_sid = false;
  for (int d : sids) any_sid |= d >= 0;
  if (!any_sid) sids.clear();
  return alloc_straddle_bits(ctx, g, n_straddle, sids);
}

int flowamok_set_cycles(flowamok_context *ctx, flowamok_grid *g, int64_t n, const int64_t *off, const int64_t *y,
                        const int64_t *x, const int8_t *cy, const int8_t *cx, const uint8_t *grant) {
  if (!g) return fail(ctx, "null grid");
  return install_rings(ctx, g, n, off, y, x, cy, cx, grant);
}

// stepper_perfect.find_cycles (steppers.fut:233-300) with sparse path sets instead of dense masks.
// Same worklist discipline (in-place continuation on the y branch, x branch appended, :281-285), so
// the surviving cycles come out in the reference's order; nub_2d (:299) keeps the first of equals.
namespace {
struct Path {
  std::vector<int64_t> idx;   // cells in visiting order
  std::vector<int8_t> cy, cx;
  std::unordered_set<int64_t> seen;
  int64_t y, x, ys, xs;
  void add(int64_t i, int8_t a, int8_t b) { idx.push_back(i); cy.push_back(a); cx.push_back(b); seen.insert(i); }
};
}  // namespace

int flowamok_find_cycles(flowamok_context *ctx, flowamok_grid *g, int64_t max_paths, int64_t *n_cycles) {
  if (!g) return fail(ctx, "null grid");
  if (max_paths <= 0) max_paths = 1 << 16;
  if (g->halo != 0) return fail(ctx, "find_cycles needs the whole grid (not available on a row band)");
  const int64_t gh = g->gh, gw = g->gw;
  std::vector<int8_t> uy((size_t)gh * gw), ux((size_t)gh * gw);
  if (flowamok_grid_download(ctx, g, uy.data(), ux.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)) return 1;
  auto inb = [&](int64_t y, int64_t x) { return y >= 0 && y < gh && x >= 0 && x < gw; };
  std::vector<Path> ps;
  for (int64_t y = 0; y < gh; y++)
    for (int64_t x = 0; x < gw; x++) {
      int64_t i = y * gw + x;
      int8_t cy = uy[i], cx = ux[i];
      if (!(cy != 0 && cx != 0)) continue;
      int64_t yn = y + cy, xn = x + cx;
      bool y_ok = inb(yn, x) && uy[yn * gw + x] == cy, x_ok = inb(y, xn) && ux[y * gw + xn] == cx;
      if (y_ok) { Path p; p.add(i, cy, 0); p.y = yn; p.x = x; p.ys = y; p.xs = x; ps.push_back(std::move(p)); }
      if (x_ok) { Path p; p.add(i, 0, cx); p.y = y; p.x = xn; p.ys = y; p.xs = x; ps.push_back(std::move(p)); }
      if ((int64_t)ps.size() > max_paths) return fail(ctx, "find_cycles: more than %lld paths", (long long)max_paths);
    }
  size_t cur = 0;
  while (cur < ps.size()) {
    Path &p = ps[cur];
    int64_t y = p.y, x = p.x;
    if (y == p.ys && x == p.xs) { cur++; continue; }
    if (!inb(y, x)) { p.y = p.x = -1; cur++; continue; }  // the reference would fault indexing here (:266)
    int64_t i = y * gw + x;
    if (p.seen.count(i)) { p.y = p.x = -1; cur++; continue; }
    int8_t cy = uy[i], cx = ux[i];
    if (cy != 0 && cx != 0) {
      int64_t yn = y + cy, xn = x + cx;
      bool y_ok = inb(yn, x) && uy[yn * gw + x] == cy, x_ok = inb(y, xn) && ux[y * gw + xn] == cx;
      if (y_ok && x_ok) {
        if ((int64_t)ps.size() >= max_paths) return fail(ctx, "find_cycles: more than %lld paths", (long long)max_paths);
        Path b = p;
        p.add(i, cy, 0); p.y = yn; p.x = x;
        b.add(i, 0, cx); b.y = y; b.x = xn;
        ps.push_back(std::move(b));  // may reallocate: p is not used afterwards
      } else if (y_ok) { p.add(i, cy, 0); p.y = yn; p.x = x; }
      else { p.add(i, 0, cx); p.y = y; p.x = xn; }
    } else if (cy != 0 || cx != 0) {
      p.add(i, cy, cx); p.y = y + cy; p.x = x + cx;
    } else { p.y = p.x = -1; cur++; }
  }
  // keep cycles, drop duplicates (same cell set with the same directions), first occurrence wins
  struct Canon { std::vector<std::pair<int64_t, int>> v; };
  std::vector<Canon> canon;
  std::vector<size_t> keep;
  for (size_t a = 0; a < ps.size(); a++) {
    if (ps[a].y == -1) continue;
    Canon c;
    for (size_t t = 0; t < ps[a].idx.size(); t++) c.v.push_back({ps[a].idx[t], (ps[a].cy[t] + 1) * 3 + (ps[a].cx[t] + 1)});
    std::sort(c.v.begin(), c.v.end());
    bool dup = false;
    for (auto &o : canon) if (o.v == c.v) { dup = true; break; }
    if (!dup) { canon.push_back(std::move(c)); keep.push_back(a); }
  }
  std::vector<int64_t> off{0}, ry, rx;
  std::vector<int8_t> rcy, rcx;
  std::vector<uint8_t> grant;
  for (size_t a : keep) {
    const Path &p = ps[a];
    for (size_t t = 0; t < p.idx.size(); t++) {
      int64_t i = p.idx[t], y = i / gw, x = i % gw;
      ry.push_back(y); rx.push_back(x); rcy.push_back(p.cy[t]); rcx.push_back(p.cx[t]);
      uint8_t gr;
      if (uy[i] != 0 && ux[i] != 0) {  // steppers.fut:206-209: is the ring cell at y - u.y heading {u.y, 0}?
        int64_t yp = y - uy[i];
        bool isy = false;
        if (inb(yp, x)) {
          int64_t j = yp * gw + x;
          for (size_t q = 0; q < p.idx.size(); q++)
            if (p.idx[q] == j) { isy = p.cy[q] == uy[i] && p.cx[q] == 0; break; }
        }
        gr = isy ? 1 : 2;
      } else gr = uy[i] != 0 ? 1 : 2;
      grant.push_back(gr);
    }
    off.push_back((int64_t)ry.size());
  }
  int64_t n = (int64_t)keep.size();
  if (n_cycles) *n_cycles = n;
  return install_rings(ctx, g, n, off.data(), ry.data(), rx.data(), rcy.data(), rcx.data(), grant.data());
}

int flowamok_cycle_count(flowamok_context *ctx, const flowamok_grid *g, int64_t *n_cycles, int64_t *n_cells) {
  if (!g) return fail(ctx, "null grid");
  if (n_cycles) *n_cycles = g->n_rings;
  if (n_cells) *n_cells = g->n_ring_cells;
  return 0;
}
int flowamok_get_cycles_dense(flowamok_context *ctx, const flowamok_grid *g, int8_t *masks) {
  if (!g) return fail(ctx, "null grid");
  if (g->n_rings > 0 && g->h_off.empty()) return fail(ctx, "cycle list was generated on the device; dense form unavailable");
  size_t per = (size_t)g->gh_global * g->gw * 2;
  const int64_t ng = g->h_off.empty() ? 0 : (int64_t)g->h_off.size() - 1;   // the list as it was installed (global on a band)
  memset(masks, 0, per * (size_t)ng);
  for (int64_t k = 0; k < ng; k++)
    for (int64_t t = g->h_off[k]; t < g->h_off[k + 1]; t++) {
      size_t o = per * (size_t)k + ((size_t)g->h_y[t] * g->gw + g->h_x[t]) * 2;
      masks[o] = g->h_cy[t]; masks[o + 1] = g->h_cx[t];
    }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// stepping
// ---------------------------------------------------------------------------------------------
static GridView make_view(flowamok_grid *g) {
  GridView v;
  v.gh = g->gh; v.gw = g->gw; v.pitch = g->pitch; v.wvalid = g->wvalid;
  v.gh_global = g->gh_global; v.row0 = g->row0; v.own0 = g->own0; v.own1 = g->own1;
  v.road = g->road;
  v.head_in = g->head[g->cur]; v.head_out = g->head[g->cur ^ 1];
  v.pref_in = g->pref[g->cur]; v.pref_out = g->pref[g->cur ^ 1];
  v.calc = g->calc; v.mymx = g->mymx; v.ustate = g->ustate;
  v.mymx_B = nullptr; v.chg_list = nullptr; v.chg_n = nullptr;
  v.color_in = g->color[g->cur]; v.color_out = g->color[g->cur ^ 1];
  v.dirty_in = g->dirty[g->cur]; v.dirty_out = g->dirty[g->cur ^ 1];
  v.rng = g->rng; v.rng_inc = g->rng_inc;
  v.chooser = g->chooser;
  v.tape_bit = (g->chooser == CHOOSE_TAPE) ? (g->tape[(size_t)(g->tick % (long long)g->tape.size())] != 0) : 0u;
  return v;
}

static int launch_u_init(flowamok_context *ctx, flowamok_grid *g, const GridView &v) {
  const unsigned int nb = nblk((long long)g->nsx * g->nsy, STRIP_THREADS / 32);
  if (g->halo != 0)   // a row band: ghost rows, boundary rows always on the frontier
    k_u_local<true><<<nb, STRIP_THREADS, 0, ctx->stream>>>(v, g->nsx, g->nsy, g->rs, g->table, g->qlist[0], g->qlist[1], g->is);
  else
    k_u_local<false><<<nb, STRIP_THREADS, 0, ctx->stream>>>(v, g->nsx, g->nsy, g->rs, g->table, g->qlist[0], g->qlist[1], g->is);
  KCHECK("k_u_local");
  return 0;
}
// first_it = 0: start from the frontier k_u_local left in qlist[0]; 1: resume from qlist[1] (band merge)
static int launch_u_iter(flowamok_context *ctx, flowamok_grid *g, GridView &v, unsigned int first_it, bool p2p = false,
                         cudaStream_t st = nullptr, int blocks = 0) {
  if (!st) st = ctx->stream;
  if (!blocks) blocks = ctx->iter_blocks;
  void *args[] = {(void *)&v, (void *)&g->table, (void *)&g->qlist[0], (void *)&g->qlist[1],
                  (void *)&first_it, (void *)&g->is, (void *)&g->diag, (void *)&g->link, (void *)&g->h_fix};
  if (p2p) CU(cudaLaunchCooperativeKernel((void *)k_u_iter<true>, dim3(g->p2p_blocks), dim3(UIT_THREADS), args, 0, st));
  else CU(cudaLaunchCooperativeKernel((void *)k_u_iter<false>, dim3(blocks), dim3(UIT_THREADS), args, 0, st));
  KCHECK("k_u_iter");
  return 0;
}
static int launch_rings(flowamok_context *ctx, flowamok_grid *g, const GridView &v, cudaStream_t st = nullptr) {
  if (!st) st = ctx->stream;
  if (g->n_rings > 0) {
    k_rings<<<nblk(g->n_rings, RINGS_PER_BLOCK), RINGS_PER_BLOCK, 0, st>>>(v, g->ring_words, g->ring_off, g->n_rings, g->ring_stride,
                                                                           g->ring_sid, g->ring_partial);
    KCHECK("k_rings");
  }
  return 0;
}
static size_t move_smem_bytes(const flowamok_grid *g) {
#ifdef FA_B1_TMA   // A/B variant: staging buffer + mbarrier behind the tile's lists
  return ((MTL::Smem::bytes(g->rs_m) + 127) & ~(size_t)127) + B1_TMA_CHUNK * 32 + 16;
#else
  return MTL::Smem::bytes(g->rs_m);
#endif
}
static int launch_move(flowamok_context *ctx, flowamok_grid *g, const GridView &v, u32 *leak_plane) {
  k_move<M_WARPS><<<g->mtx * g->mty, M_WARPS * 32, move_smem_bytes(g), ctx->stream>>>(v, g->nsx, g->mtx, g->rs_m, g->diag, leak_plane);
  KCHECK("k_move");
  g->cur ^= 1;
  g->tick++;
  return 0;
}

// The view of the global fixpoint (and of the fix-up) when the tick moves speculatively: grants go into the delta plane
static GridView spec_view(flowamok_grid *g) {
  GridView v = make_view(g);
  v.mymx_B = g->mymx_B; v.chg_list = g->chg_list; v.chg_n = &g->is->chg_n;
  return v;
}
// Fork: k_u_local (ring resolution, a band's first boundary swap) are enqueued on the main stream.  The global fixpoint goes
// to the high-priority stream and keeps its grants in the delta plane; k_move follows k_u_local at once on the main stream,
// gathering from the plane as published; join; fix-up of the words around the ones the fixpoint changed.
static int fork_fix_and_move(flowamok_context *ctx, flowamok_grid *g, bool p2p, cudaEvent_t *tev = nullptr) {
  cudaStream_t st = ctx->stream;
  if (tev) CU(cudaEventRecord(tev[1], st));
  CU(cudaEventRecord(ctx->ev_local, st));
  CU(cudaStreamWaitEvent(ctx->hi, ctx->ev_local, 0));
  GridView vf = spec_view(g);
  if (launch_u_iter(ctx, g, vf, 0u, p2p, ctx->hi)) return 1;
  if (tev) CU(cudaEventRecord(tev[2], ctx->hi));
  CU(cudaEventRecord(ctx->ev_fix, ctx->hi));
  GridView vs = make_view(g);
  k_move<M_WARPS><<<g->mtx * g->mty, M_WARPS * 32, move_smem_bytes(g), st>>>(vs, g->nsx, g->mtx, g->rs_m, g->diag, nullptr);
  KCHECK("k_move");
  if (tev) CU(cudaEventRecord(tev[3], st));
  CU(cudaStreamWaitEvent(st, ctx->ev_fix, 0));
  if (++g->fix_epoch == 0u) {   // the tag wrapped (after 2^32 ticks): forget the old claims
    CU(cudaMemsetAsync(g->fix_claim, 0, g->nwords * 4, st));
    g->fix_epoch = 1u;
  }
  k_move_fix<<<ctx->num_sms * 2, 256, 0, st>>>(vf, g->fix_claim, g->fix_epoch, g->is, g->diag);
  KCHECK("k_move_fix");
  if (tev) CU(cudaEventRecord(tev[4], st));
  g->cur ^= 1;
  g->tick++;
  return 0;
}

// ev != nullptr (flowamok_step_profile) or a leak plane: the kernels one after the other on the main stream
static int enqueue_tick(flowamok_context *ctx, flowamok_grid *g, bool perfect, u32 *leak_plane, cudaEvent_t *ev = nullptr,
                        cudaEvent_t *tev = nullptr) {
  if (g->halo != 0) return fail(ctx, "a row band is stepped with the flowamok_band_* calls");
  cudaStream_t st = ctx->stream;
  if (ctx->split && !ev && !leak_plane) {
    {   // jam detection on what the last ticks reported: visits per fixpoint iteration, with hysteresis
      const unsigned long long p = ((volatile unsigned long long *)g->h_fix)[0], w = ((volatile unsigned long long *)g->h_fix)[1];
      if (p > g->fix_seen[0] && w >= g->fix_seen[1]) {
        const double per_it = (double)(w - g->fix_seen[1]) / (double)(p - g->fix_seen[0]);
        if (per_it > ctx->jam_visits) g->jam = true;
        else if (per_it < 0.4 * ctx->jam_visits) g->jam = false;
        g->fix_seen[0] = p; g->fix_seen[1] = w;
      }
    }
    if (g->jam && !tev) {   // the frontier is large: the fixpoint gets the whole device, then the move (no speculation)
      g->jam_ticks++;
      GridView v = make_view(g);
      if (launch_u_init(ctx, g, v)) return 1;
      if (launch_u_iter(ctx, g, v, 0u, false, st, ctx->iter_blocks_big)) return 1;
      if (perfect && launch_rings(ctx, g, v)) return 1;
      return launch_move(ctx, g, v, nullptr);
    }
    if (tev) CU(cudaEventRecord(tev[0], st));
    GridView v = make_view(g);
    if (launch_u_init(ctx, g, v)) return 1;
    if (perfect && launch_rings(ctx, g, v)) return 1;   // ring resolution does not depend on the fixpoint (ring_grant)
    return fork_fix_and_move(ctx, g, false, tev);
  }
  GridView v = make_view(g);
  if (ev) CU(cudaEventRecord(ev[0], st));
  if (launch_u_init(ctx, g, v)) return 1;
  if (ev) CU(cudaEventRecord(ev[1], st));
  if (launch_u_iter(ctx, g, v, 0u)) return 1;
  if (ev) CU(cudaEventRecord(ev[2], st));
  if (perfect && launch_rings(ctx, g, v)) return 1;
  if (ev) CU(cudaEventRecord(ev[3], st));
  if (launch_move(ctx, g, v, leak_plane)) return 1;
  if (ev) CU(cudaEventRecord(ev[4], st));
  return 0;
}

// The device-side watchdogs (grid barrier of k_u_iter, waits of the band protocol) leave the planes half-relaxed when
// they trip: the grid is poisoned, and every later stepping or reading call fails.
static int check_device_errors(flowamok_context *ctx, flowamok_grid *g) {
  if (g->poisoned) return fail(ctx, "grid is poisoned: %s", g->poisoned);
  unsigned int err = 0;
  BandDev bd;
  memset(&bd, 0, sizeof bd);
  CU(cudaMemcpyAsync(&err, &g->is->error, 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (g->band_dev) CU(cudaMemcpyAsync(&bd, g->band_dev, sizeof bd, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (err) g->poisoned = "the grid barrier watchdog of k_u_iter tripped (fixpoint left half-relaxed)";
  else if (bd.error) {
    static const char *what[] = {"?", "state rows", "u rows", "termination vote", "ring verdicts"};
    snprintf(g->poison_text, sizeof g->poison_text, "band %d: a wait of the row-band protocol timed out (%s: message %u never came, last seen %u)",
             g->link.rank, what[(bd.error >> 8) < 5 ? (bd.error >> 8) : 0], bd.err_want, bd.err_got);
    g->poisoned = g->poison_text;
  }
  if (g->poisoned) return fail(ctx, "%s", g->poisoned);
  return 0;
}
static int read_diag(flowamok_context *ctx, flowamok_grid *g, Diag *d) {
  CU(cudaMemcpyAsync(g->h_diag, g->diag, sizeof(Diag), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *d = *g->h_diag;
  return check_device_errors(ctx, g);
}

static int step_n(flowamok_context *ctx, flowamok_grid *g, bool perfect, int64_t n_ticks, int64_t *n_leaks) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device: flowamok_b200 has no CPU path");
  if (!g) return fail(ctx, "null grid");
  if (g->poisoned) return fail(ctx, "grid is poisoned: %s", g->poisoned);
  CU(cudaSetDevice(ctx->device));
  unsigned long long before = 0;
  if (n_leaks) {
    Diag d;
    if (read_diag(ctx, g, &d)) return 1;
    before = d.leaks;
  }
  for (int64_t t = 0; t < n_ticks; t++)
    if (enqueue_tick(ctx, g, perfect, nullptr)) return 1;
  if (n_leaks) {
    Diag d;
    if (read_diag(ctx, g, &d)) return 1;
    *n_leaks = (int64_t)(d.leaks - before);
  }
  return 0;
}

int flowamok_step_quick(flowamok_context *ctx, flowamok_grid *g, int64_t n_ticks, int64_t *n_leaks) {
  return step_n(ctx, g, false, n_ticks, n_leaks);
}
int flowamok_step_perfect(flowamok_context *ctx, flowamok_grid *g, int64_t n_ticks, int64_t *n_leaks) {
  return step_n(ctx, g, true, n_ticks, n_leaks);
}

int flowamok_step_profile(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, double *ms) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || !ms || n_ticks < 1 || n_ticks > 4096) return fail(ctx, "bad profile arguments");
  CU(cudaSetDevice(ctx->device));
  EventSet es;
  CU(es.create((size_t)n_ticks * 5));
  std::vector<cudaEvent_t> &ev = es.ev;
  for (int64_t t = 0; t < n_ticks; t++)
    if (enqueue_tick(ctx, g, perfect != 0, nullptr, &ev[(size_t)t * 5])) return 1;
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 4; k++) ms[k] = 0.0;
  for (int64_t t = 0; t < n_ticks; t++)
    for (int k = 0; k < 4; k++) {
      float f = 0.f;
      CU(cudaEventElapsedTime(&f, ev[(size_t)t * 5 + k], ev[(size_t)t * 5 + k + 1]));
      ms[k] += f;
    }
  return check_device_errors(ctx, g);
}

/* The production tick (speculative move, three streams) on a time line, summed over n_ticks <= 1024, in ms:
 * ms[0] k_u_local (+ k_rings) on the main stream; ms[1] fork -> end of k_u_iter (high-priority stream); ms[2] fork -> end of
 * the speculative k_move (side stream); ms[3] join -> end of k_move_fix; ms[4] the whole tick. */
int flowamok_step_timeline(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, double *ms5) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || !ms5 || n_ticks < 1 || n_ticks > 1024 || !ctx->split) return fail(ctx, "bad timeline arguments");
  CU(cudaSetDevice(ctx->device));
  EventSet es;
  CU(es.create((size_t)n_ticks * 5));
  std::vector<cudaEvent_t> &ev = es.ev;
  int rc = 0;
  for (int64_t t = 0; t < n_ticks && !rc; t++) rc = enqueue_tick(ctx, g, perfect != 0, nullptr, nullptr, &ev[(size_t)t * 5]);
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  for (int k = 0; k < 5; k++) ms5[k] = 0.0;
  if (!rc && se == cudaSuccess)
    for (int64_t t = 0; t < n_ticks; t++) {
      cudaEvent_t *e = &ev[(size_t)t * 5];
      float a = 0, b = 0, c = 0, d1 = 0, d2 = 0, tot = 0;
      cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]); cudaEventElapsedTime(&c, e[1], e[3]);
      cudaEventElapsedTime(&d1, e[2], e[4]); cudaEventElapsedTime(&d2, e[3], e[4]); cudaEventElapsedTime(&tot, e[0], e[4]);
      ms5[0] += a; ms5[1] += b; ms5[2] += c; ms5[3] += d1 < d2 ? d1 : d2; ms5[4] += tot;
    }
  if (se != cudaSuccess) return fail(ctx, "timeline: %s", cudaGetErrorString(se));
  return rc ? 1 : check_device_errors(ctx, g);
}

static int step_leaks_impl(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t cap, int64_t *leaks, int64_t *n_leaks, int nf) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g) return fail(ctx, "null grid");
  if (g->poisoned) return fail(ctx, "grid is poisoned: %s", g->poisoned);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!g->leak_plane) CU(cudaMalloc(&g->leak_plane, g->nwords * 8));   // leak mask plane + "drew in this tick" plane
  int old = g->cur;
  if (enqueue_tick(ctx, g, perfect != 0, g->leak_plane)) return 1;
  DevBuf cbuf, obuf, tbuf, dbuf;
  CU(cbuf.alloc((g->nwords + 1) * 4));
  CU(obuf.alloc((g->nwords + 1) * 4));
  unsigned int *counts = cbuf.as<unsigned int>(), *offs = obuf.as<unsigned int>();
  CU(cudaMemsetAsync(counts, 0, (g->nwords + 1) * 4, st));
  k_leak_counts<<<nblk((long long)g->nwords, 256), 256, 0, st>>>((long long)g->nwords, g->leak_plane, counts);
  CU(cudaGetLastError());
  size_t tmpb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmpb, counts, offs, (int)(g->nwords + 1), st));
  CU(tbuf.alloc(tmpb));
  CU(cub::DeviceScan::ExclusiveSum(tbuf.p, tmpb, counts, offs, (int)(g->nwords + 1), st));
  unsigned int total = 0;
  CU(cudaMemcpyAsync(&total, offs + g->nwords, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  long long emit = std::min<long long>(total, cap);
  if (emit > 0 && leaks) {
    CU(dbuf.alloc((size_t)emit * 8 * nf));
    long long *dout = dbuf.as<long long>();
    // the tuple copies the cell as move_cell saw it (steppers.fut:122): heading/colour of the old state,
    // next_preference_flow_x already updated by the fixpoint (written by U into the new buffer)
    k_leak_emit<<<nblk((long long)g->nwords, 256), 256, 0, st>>>(g->gh, g->gw, g->pitch, g->row0, g->leak_plane, offs, g->head[old],
                                                                g->pref[old ^ 1], g->color[old], g->road, g->calc, g->mymx, g->rng,
                                                                g->rng_inc, emit, nf, dout);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(leaks, dout, (size_t)emit * 8 * nf, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  if (n_leaks) *n_leaks = total;
  return check_device_errors(ctx, g);
}

int flowamok_step_leaks(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t cap, int64_t *leaks, int64_t *n_leaks) {
  return step_leaks_impl(ctx, g, perfect, cap, leaks, n_leaks, 8);
}
int flowamok_step_leaks_cells(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t cap, int64_t *leaks, int64_t *n_leaks) {
  return step_leaks_impl(ctx, g, perfect, cap, leaks, n_leaks, 14);
}

int flowamok_render(flowamok_context *ctx, const flowamok_grid *g, int64_t scale, uint32_t *pixels) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || scale < 1) return fail(ctx, "bad render arguments");
  if (g->halo != 0) return fail(ctx, "render needs the whole grid (not available on a row band)");
  CU(cudaSetDevice(ctx->device));
  long long n = (long long)g->gh * scale * g->gw * scale;
  DevBuf buf;
  CU(buf.alloc((size_t)n * 4));
  u32 *d = buf.as<u32>();
  k_render<<<nblk(n, 256), 256, 0, ctx->stream>>>(g->gh, g->gw, g->pitch, scale, g->road, g->head[g->cur], g->color[g->cur], d);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(pixels, d, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return check_device_errors(ctx, const_cast<flowamok_grid *>(g));
}

int flowamok_render_device(flowamok_context *ctx, const flowamok_grid *g, int64_t scale, void **pixels_dev) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || scale < 1 || !pixels_dev) return fail(ctx, "bad render arguments");
  CU(cudaSetDevice(ctx->device));
  long long n = (long long)g->gh * scale * g->gw * scale;
  DevBuf buf;
  CU(buf.alloc((size_t)n * 4));
  k_render<<<nblk(n, 256), 256, 0, ctx->stream>>>(g->gh, g->gw, g->pitch, scale, g->road, g->head[g->cur], g->color[g->cur], buf.as<u32>());
  CU(cudaGetLastError());
  *pixels_dev = buf.p;
  buf.p = nullptr;   // the caller owns the pixels now
  return 0;
}

int flowamok_grid_clone(flowamok_context *ctx, const flowamok_grid *src, flowamok_grid **out) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!src || !out) return fail(ctx, "null grid");
  flowamok_grid *g = nullptr;
  for (size_t i = 0; i < ctx->pool.size() && !g; i++) {   // a recycled grid of the same shape: no allocation, no memsets
    flowamok_grid *q = ctx->pool[i];
    if (q->gh_global == src->gh_global && q->gw == src->gw && q->row0 == src->row0 && q->own0 == src->own0 && q->own1 == src->own1 &&
        q->halo == src->halo) {
      g = q;
      ctx->pool.erase(ctx->pool.begin() + (long)i);
      ctx->grids.push_back(g);
      cudaFree(g->ring_words); cudaFree(g->ring_off); cudaFree(g->ring_sid); cudaFree(g->ring_partial); cudaFree(g->ring_pass);
      g->ring_words = nullptr; g->ring_off = nullptr; g->ring_sid = nullptr; g->ring_partial = nullptr; g->ring_pass = nullptr;
      g->n_rings = g->n_ring_cells = g->n_ring_words = 0; g->ring_stride = 0; g->n_straddle = 0; g->sbit_words = 0;
      g->h_off.clear(); g->h_y.clear(); g->h_x.clear(); g->h_cy.clear(); g->h_cx.clear();
    }
  }
  if (!g && grid_new_impl(ctx, &g, src->gh_global, src->gw, src->row0 + src->own0, src->row0 + src->own1, src->halo)) return 1;
  cudaStream_t st = ctx->stream;
  CU(cudaMemcpyAsync(g->road, src->road, src->nwords * sizeof(U4), cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(g->head[0], src->head[src->cur], src->nwords * sizeof(U4), cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(g->pref[0], src->pref[src->cur], src->nwords * 4, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(g->color[0], src->color[src->cur], src->ncells_p * 4, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(g->rng, src->rng, src->ncells_p * 8, cudaMemcpyDeviceToDevice, st));
  g->cur = 0;
  if (sync_color_buffers(ctx, g)) return 1; g->rng_inc = src->rng_inc; g->chooser = src->chooser; g->tape = src->tape; g->tick = src->tick;
  if (src->n_rings > 0) {
    CU(cudaMalloc(&g->ring_words, std::max<size_t>(1, (size_t)src->n_ring_words) * sizeof(RingWord)));
    CU(cudaMemcpyAsync(g->ring_words, src->ring_words, (size_t)src->n_ring_words * sizeof(RingWord), cudaMemcpyDeviceToDevice, st));
    if (src->ring_off) {
      CU(cudaMalloc(&g->ring_off, ((size_t)src->n_rings + 1) * 8));
      CU(cudaMemcpyAsync(g->ring_off, src->ring_off, ((size_t)src->n_rings + 1) * 8, cudaMemcpyDeviceToDevice, st));
    }
    g->n_rings = src->n_rings; g->n_ring_cells = src->n_ring_cells; g->n_ring_words = src->n_ring_words; g->ring_stride = src->ring_stride;
    g->ring_touch = src->ring_touch; g->ring_swap = src->ring_swap; g->cuts = src->cuts;
    if (src->n_straddle > 0) {
      std::vector<int> none;
      if (alloc_straddle_bits(ctx, g, src->n_straddle, none)) return 1;
      if (src->ring_sid) {
        CU(cudaMalloc(&g->ring_sid, (size_t)src->n_rings * sizeof(int)));
        CU(cudaMemcpyAsync(g->ring_sid, src->ring_sid, (size_t)src->n_rings * sizeof(int), cudaMemcpyDeviceToDevice, st));
      }
    }
    g->h_off = src->h_off; g->h_y = src->h_y; g->h_x = src->h_x; g->h_cy = src->h_cy; g->h_cx = src->h_cx;
  }
  *out = g;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Row bands.  side 0 = towards smaller rows, 1 = towards larger rows.
//   state message: 2 heading rows + 2 preference rows + 1 colour row (what k_u_local / k_move read across the cut)
//   u message    : 1 row of calc + 1 row of my|mx (what the fixpoint reads across the cut)
// ---------------------------------------------------------------------------------------------
int flowamok_band_msg_bytes(flowamok_context *ctx, const flowamok_grid *g, int64_t *state_bytes, int64_t *u_bytes) {
  if (!g) return fail(ctx, "null grid");
  if (state_bytes) *state_bytes = (int64_t)g->pitch * (2 * 16 + 2 * 4 + 128);
  if (u_bytes) *u_bytes = (int64_t)g->pitch * (4 + 8);
  return 0;
}
static int band_check(flowamok_context *ctx, const flowamok_grid *g, int side) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || g->halo < 2) return fail(ctx, "not a row band");
  if (side != 0 && side != 1) return fail(ctx, "side must be 0 or 1");
  if (g->own1 - g->own0 < 2) return fail(ctx, "a band needs at least 2 rows");
  return 0;
}
static int band_state(flowamok_context *ctx, flowamok_grid *g, void *b0, void *b1, int pack) {
  BandBufs bufs = {{(char *)b0, (char *)b1}};
  k_band_state<<<dim3(nblk(8ll * g->pitch, 256), 2), 256, 0, ctx->stream>>>(g->pitch, g->own0, g->own1, g->head[g->cur], g->pref[g->cur],
                                                                             g->color[g->cur], bufs, pack);
  KCHECK("k_band_state");
  return 0;
}
int flowamok_band_pack_state(flowamok_context *ctx, flowamok_grid *g, int side, void *dev_buf) {
  if (band_check(ctx, g, side)) return 1;
  CU(cudaSetDevice(ctx->device));
  return band_state(ctx, g, side == 0 ? dev_buf : nullptr, side == 1 ? dev_buf : nullptr, 1);
}
int flowamok_band_unpack_state(flowamok_context *ctx, flowamok_grid *g, int side, const void *dev_buf) {
  if (band_check(ctx, g, side)) return 1;
  CU(cudaSetDevice(ctx->device));
  return band_state(ctx, g, side == 0 ? (void *)dev_buf : nullptr, side == 1 ? (void *)dev_buf : nullptr, 0);
}
/* u_init: k_u_local only (statics, local levels, ghost rows seeded with this band's view of the neighbour's
 * boundary row); u_iter: the frontier iterations.  u_begin = both.  Swapping the boundary rows BETWEEN the two lets
 * the frontier run against the neighbour's post-sweep rows, so the first convergence check usually passes. */
int flowamok_band_u_init(flowamok_context *ctx, flowamok_grid *g) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  // a tick always starts with empty frontier counters, whatever the host did with the previous round
  CU(cudaMemsetAsync(g->is->qn, 0, sizeof(g->is->qn), ctx->stream));
  GridView v = make_view(g);
  return launch_u_init(ctx, g, v);
}
int flowamok_band_u_iter(flowamok_context *ctx, flowamok_grid *g) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  GridView v = make_view(g);
  return launch_u_iter(ctx, g, v, 0u);
}
int flowamok_band_u_begin(flowamok_context *ctx, flowamok_grid *g) {
  return flowamok_band_u_init(ctx, g) || flowamok_band_u_iter(ctx, g);
}
static int band_pack_u(flowamok_context *ctx, flowamok_grid *g, void *b0, void *b1) {
  BandBufs bufs = {{(char *)b0, (char *)b1}};
  k_band_pack_u<<<dim3(nblk(g->pitch, 128), 2), 128, 0, ctx->stream>>>(g->pitch, g->own0, g->own1, g->calc, g->mymx, bufs);
  KCHECK("k_band_pack_u");
  return 0;
}
static int band_merge_u(flowamok_context *ctx, flowamok_grid *g, const void *b0, const void *b1, int wake) {
  BandBufs bufs = {{(char *)b0, (char *)b1}};
  GridView v = make_view(g);
  k_band_merge<<<dim3(nblk(g->pitch, 128), 2), 128, 0, ctx->stream>>>(v, bufs, wake, g->table, g->qlist[1], g->is);
  KCHECK("k_band_merge");
  return 0;
}
int flowamok_band_pack_u(flowamok_context *ctx, flowamok_grid *g, int side, void *dev_buf) {
  if (band_check(ctx, g, side)) return 1;
  CU(cudaSetDevice(ctx->device));
  return band_pack_u(ctx, g, side == 0 ? dev_buf : nullptr, side == 1 ? dev_buf : nullptr);
}
/* OR the neighbour's boundary row into my ghost row; wake=1 queues the owned words next to changed ghost words.
 * dev_woken (device u32, may be NULL) receives the number of entries queued so far this round. */
int flowamok_band_merge_u(flowamok_context *ctx, flowamok_grid *g, int side, const void *dev_buf, int wake, void *dev_woken) {
  if (band_check(ctx, g, side)) return 1;
  CU(cudaSetDevice(ctx->device));
  if (band_merge_u(ctx, g, side == 0 ? dev_buf : nullptr, side == 1 ? dev_buf : nullptr, wake)) return 1;
  if (dev_woken) CU(cudaMemcpyAsync(dev_woken, &g->is->qn[1], 4, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
int flowamok_band_woken(flowamok_context *ctx, flowamok_grid *g, void *dev_woken) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(dev_woken, &g->is->qn[1], 4, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
int flowamok_band_u_resume(flowamok_context *ctx, flowamok_grid *g) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  GridView v = make_view(g);
  return launch_u_iter(ctx, g, v, 1u);
}
int flowamok_band_rings(flowamok_context *ctx, flowamok_grid *g) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  GridView v = make_view(g);
  return launch_rings(ctx, g, v);
}
/* Rings that straddle a cut, host-driven protocol: flowamok_band_rings judges them (and marks the local ones);
 * ring_verdicts copies this band's verdict words (n_words of them, 1 = no objection) to dev_buf and resets them;
 * the caller ANDs the buffers of all bands; rings_apply marks the rings whose bit survived. */
int flowamok_band_ring_words(flowamok_context *ctx, const flowamok_grid *g, int64_t *n_words) {
  if (!g) return fail(ctx, "null grid");
  if (n_words) *n_words = g->sbit_words;
  return 0;
}
int flowamok_band_ring_verdicts(flowamok_context *ctx, flowamok_grid *g, void *dev_buf) {
  if (band_check(ctx, g, 0)) return 1;
  if (g->sbit_words == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(dev_buf, g->ring_partial, (size_t)g->sbit_words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemsetAsync(g->ring_partial, 0xFF, (size_t)g->sbit_words * 4, ctx->stream));
  return 0;
}
int flowamok_band_rings_apply(flowamok_context *ctx, flowamok_grid *g, const void *dev_pass) {
  if (band_check(ctx, g, 0)) return 1;
  if (g->sbit_words == 0 || g->n_rings == 0 || !g->ring_sid) return 0;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(g->ring_pass, dev_pass, (size_t)g->sbit_words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  GridView v = make_view(g);
  k_rings_mark<<<nblk(g->n_rings, 128), 128, 0, ctx->stream>>>(v, g->ring_words, g->ring_off, g->n_rings, g->ring_stride, g->ring_sid, g->ring_pass);
  KCHECK("k_rings_mark");
  return 0;
}
int flowamok_band_move(flowamok_context *ctx, flowamok_grid *g) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  GridView v = make_view(g);
  return launch_move(ctx, g, v, nullptr);
}
#define NC(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) return fail(ctx, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
  } while (0)

/* NCCL bootstrap: rank 0 calls comm_id and broadcasts the 128 bytes (e.g. with torch.distributed); every rank then
 * calls comm_init on its own context. */
int flowamok_band_comm_id(void *id128) {
  if (!g_nccl.load()) return 1;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return 1;
  memcpy(id128, &id, sizeof id);
  return 0;
}
int flowamok_band_comm_init(flowamok_context *ctx, const void *id128, int rank, int world) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g_nccl.load()) return fail(ctx, "libnccl.so.2 could not be loaded: %s", dlerror());
  CU(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NC(g_nccl.CommInitRank(&ctx->comm, world, id, rank));
  ctx->rank = rank; ctx->world = world;
  return 0;
}

/* Row boundaries of ALL bands (world + 1 entries, cuts[0] = 0, cuts[world] = gh_global), the same on every band.  Needed
 * before cycles are installed on a band whose rings may straddle a cut, and before flowamok_band_p2p_open. */
int flowamok_band_set_cuts(flowamok_context *ctx, flowamok_grid *g, int world, const int64_t *cuts) {
  if (band_check(ctx, g, 0)) return 1;
  if (world < 1 || world > BAND_MAXW || !cuts) return fail(ctx, "bad cuts (1 <= world <= %d)", BAND_MAXW);
  if (cuts[0] != 0 || cuts[world] != g->gh_global) return fail(ctx, "cuts must run from 0 to the grid height");
  bool mine = false;
  for (int r = 0; r < world; r++) {
    if (cuts[r + 1] <= cuts[r]) return fail(ctx, "cuts must increase");
    mine |= cuts[r] == g->row0 + g->own0 && cuts[r + 1] == g->row0 + g->own1;
  }
  if (!mine) return fail(ctx, "this band's rows are not one of the intervals of the cuts");
  g->cuts.assign(cuts, cuts + world + 1);
  return 0;
}

static size_t mailbox_layout(const flowamok_grid *g, int world, int ring_words_cap, BandLink *L) {
  const size_t sb = (size_t)g->pitch * (2 * 16 + 2 * 4 + 128), ub = (size_t)g->pitch * (4 + 8);
  size_t o = sizeof(MailHdr);
  L->s_bytes = sb; L->u_bytes = ub;
  L->off_s = o; o += 4 * sb;
  L->off_u = o; o += 4 * ub;
  L->off_ring = o; o += (size_t)2 * world * ring_words_cap * 4;
  return (o + 255) & ~(size_t)255;
}

/* P2P transport, step 1 of 2: allocate this band's mailbox and export it.  handle64 receives the 64-byte CUDA IPC handle
 * (for bands in other processes), *local_ptr the device pointer (for bands in this process).  All bands must have the same
 * pitch and world, and the cycles should already be installed (the mailbox is sized for the straddling rings). */
int flowamok_band_p2p_open(flowamok_context *ctx, flowamok_grid *g, int rank, int world, void *handle64, void **local_ptr) {
  if (band_check(ctx, g, 0)) return 1;
  if (world < 1 || world > BAND_MAXW || rank < 0 || rank >= world) return fail(ctx, "bad rank / world (world <= %d)", BAND_MAXW);
  if ((int)g->cuts.size() != world + 1) return fail(ctx, "flowamok_band_set_cuts must be called first (with the same world)");
  if (g->cuts[(size_t)rank] != g->row0 + g->own0) return fail(ctx, "rank %d does not own the rows this band holds", rank);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  if (g->mailbox) return fail(ctx, "the band's mailbox is open already");
  g->sbit_words_cap = std::max(1024, g->sbit_words);
  BandLink L;
  memset(&L, 0, sizeof L);
  g->mailbox_bytes = mailbox_layout(g, world, g->sbit_words_cap, &L);
  CU(cudaMalloc(&g->mailbox, g->mailbox_bytes));
  CU(cudaMemset(g->mailbox, 0, g->mailbox_bytes));
  CU(cudaMalloc(&g->band_dev, sizeof(BandDev)));
  CU(cudaMemset(g->band_dev, 0, sizeof(BandDev)));
  L.rank = rank; L.world = world; L.self = g->mailbox; L.dev = g->band_dev; L.ring_words = g->sbit_words;
  g->link = L;
  ctx->rank = rank; ctx->world = world;
  if (handle64) {
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaIpcGetMemHandle(&h, g->mailbox));
    memcpy(handle64, &h, 64);
  }
  if (local_ptr) *local_ptr = g->mailbox;
  return 0;
}

/* step 2 of 2: map the other bands' mailboxes.  local_ptrs (world entries, may be NULL): device pointers of bands that
 * live in THIS process (taken as they are); otherwise handles (world x 64 bytes, by rank) are opened with CUDA IPC.
 * bands_on_device: how many bands share this band's GPU (1 in production; > 1 only when several bands are stepped
 * concurrently on one device for tests: their persistent kernels must all be resident at once). */
int flowamok_band_p2p_connect(flowamok_context *ctx, flowamok_grid *g, const void *handles, void *const *local_ptrs, int bands_on_device) {
  if (band_check(ctx, g, 0)) return 1;
  if (!g->mailbox) return fail(ctx, "flowamok_band_p2p_open has not been called");
  CU(cudaSetDevice(ctx->device));
  BandLink &L = g->link;
  for (int r = 0; r < L.world; r++) {
    if (r == L.rank) { L.all[r] = g->mailbox; continue; }
    if (local_ptrs && local_ptrs[r]) { L.all[r] = (char *)local_ptrs[r]; continue; }
    if (!handles) return fail(ctx, "no handle for band %d", r);
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)r * 64, 64);
    void *ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g->ipc_opened[r] = ptr;
    L.all[r] = (char *)ptr;
  }
  L.nb[0] = L.rank > 0 ? L.all[L.rank - 1] : nullptr;
  L.nb[1] = L.rank + 1 < L.world ? L.all[L.rank + 1] : nullptr;
  if (bands_on_device < 1) bands_on_device = 1;
  g->p2p_blocks = bands_on_device == 1 ? ctx->iter_blocks_p2p : std::max(1, ctx->iter_blocks_p2p / (2 * bands_on_device));
  g->p2p = true;
  return 0;
}

/* Unmap the peers' mailboxes (call on every band, then synchronise the processes, then free the grids: a mailbox must
 * not be freed while another process still maps it). */
int flowamok_band_p2p_close(flowamok_context *ctx, flowamok_grid *g) {
  if (band_check(ctx, g, 0)) return 1;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  g->p2p = false;
  for (int r = 0; r < BAND_MAXW; r++) {
    if (g->ipc_opened[r]) CU(cudaIpcCloseMemHandle(g->ipc_opened[r]));
    g->ipc_opened[r] = nullptr;
    g->link.all[r] = nullptr;
  }
  g->link.nb[0] = g->link.nb[1] = nullptr;
  return check_device_errors(ctx, g);
}

// swap `bytes` with both neighbours: send[side] goes to the neighbour on that side, recv[side] comes from it
static int band_swap(flowamok_context *ctx, char *const send[2], char *const recv[2], size_t bytes) {
  NC(g_nccl.GroupStart());
  for (int side = 0; side < 2; side++) {
    int peer = side == 0 ? ctx->rank - 1 : ctx->rank + 1;
    if (peer < 0 || peer >= ctx->world) continue;
    NC(g_nccl.Send(send[side], bytes, ncclUint8, peer, ctx->comm, ctx->stream));
    NC(g_nccl.Recv(recv[side], bytes, ncclUint8, peer, ctx->comm, ctx->stream));
  }
  NC(g_nccl.GroupEnd());
  return 0;
}

// The per-tick protocol over the P2P transport (fa_band.cuh): five to eight launches per tick on the context's stream,
// no NCCL call and no host synchronisation at all -- the convergence rounds run inside k_u_iter<true>.
//   k_band_xstate            heading / preference / colour rows across the cuts
//   k_u_local<band>          local levels
//   k_u_iter<p2p>            pre-swap, local fixpoint, { swap + wake + vote, local fixpoint }*
//   [perfect] k_rings        local rings are marked, straddling rings judged;  k_band_xring + k_rings_mark if any ring
//                            straddles a cut;  k_band_xu if any ring touches a boundary row
//   k_move
static int band_step_p2p(flowamok_context *ctx, flowamok_grid *g, bool perfect, int64_t n_ticks, int64_t *rounds, cudaEvent_t *ev = nullptr) {
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const bool multi = g->link.world > 1;
  const int xs_blocks = std::max(1, std::min(64, (int)nblk(8ll * g->pitch, 256)));
  for (int64_t t = 0; t < n_ticks; t++) {
    cudaEvent_t *e = ev ? ev + t * 6 : nullptr;   // brackets: xstate | u_local | rings + verdicts + boundary swap | u_iter | move
    if (e) CU(cudaEventRecord(e[0], st));
    if (multi) {
      k_band_xstate<<<xs_blocks, 256, 0, st>>>(g->pitch, g->own0, g->own1, g->head[g->cur], g->pref[g->cur], g->color[g->cur], g->link,
                                               ++g->s_seq);
      KCHECK("k_band_xstate");
    }
    if (e) CU(cudaEventRecord(e[1], st));
    const bool spec = ctx->split && !e && g->p2p_blocks == ctx->iter_blocks_p2p;   // several bands on one device: one after the other
    CU(cudaMemsetAsync(g->is->qn, 0, sizeof(g->is->qn), st));   // a tick starts with empty frontier counters
    GridView v = make_view(g);
    if (launch_u_init(ctx, g, v)) return 1;
    if (e) CU(cudaEventRecord(e[2], st));
    if (perfect) {   // ring resolution does not depend on the fixpoint (ring_grant): it runs first, so its marks travel with
                     // the boundary swap below and belong to the snapshot the speculative k_move reads
      if (launch_rings(ctx, g, v)) return 1;
      if (multi && g->n_straddle > 0) {
        k_band_xring<<<1, 512, 0, st>>>(g->ring_partial, g->ring_pass, g->link);
        KCHECK("k_band_xring");
        if (g->n_rings > 0 && g->ring_sid) {
          k_rings_mark<<<nblk(g->n_rings, 128), 128, 0, st>>>(v, g->ring_words, g->ring_off, g->n_rings, g->ring_stride, g->ring_sid, g->ring_pass);
          KCHECK("k_rings_mark");
        }
      }
    }
    if (multi) {   // ghost rows := the neighbours' boundary rows after THEIR local levels (and ring marks)
      k_band_xu<<<1, 1024, 0, st>>>(v, g->link);
      KCHECK("k_band_xu");
    }
    if (e) CU(cudaEventRecord(e[3], st));
    if (spec) {
      if (fork_fix_and_move(ctx, g, true)) return 1;
      continue;
    }
    if (launch_u_iter(ctx, g, v, 0u, true)) return 1;
    if (e) CU(cudaEventRecord(e[4], st));
    if (launch_move(ctx, g, v, nullptr)) return 1;
    if (e) CU(cudaEventRecord(e[5], st));
  }
  if (rounds) {   // convergence rounds so far: a device counter (this read synchronises; pass NULL to stay asynchronous)
    unsigned long long r = 0;
    CU(cudaMemcpyAsync(&r, &g->band_dev->rounds, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *rounds = (int64_t)r;
    if (check_device_errors(ctx, g)) return 1;
  }
  return 0;
}

/* flowamok_step_profile for a band on the P2P transport: summed device milliseconds of {k_band_xstate, k_u_local, ring
 * resolution + boundary swap, k_u_iter (with the convergence rounds), k_move}, the kernels one after the other, over
 * n_ticks <= 1024.  Every band must call it. */
int flowamok_band_step_profile(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, double *ms5) {
  if (band_check(ctx, g, 0)) return 1;
  if (!g->p2p || !ms5 || n_ticks < 1 || n_ticks > 1024) return fail(ctx, "bad band profile arguments (P2P transport only)");
  CU(cudaSetDevice(ctx->device));
  EventSet es;
  CU(es.create((size_t)n_ticks * 6));
  std::vector<cudaEvent_t> &ev = es.ev;
  int rc = band_step_p2p(ctx, g, perfect != 0, n_ticks, nullptr, ev.data());
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  for (int k = 0; k < 5; k++) ms5[k] = 0.0;
  if (!rc && se == cudaSuccess)
    for (int64_t t = 0; t < n_ticks; t++)
      for (int k = 0; k < 5; k++) {
        float f = 0.f;
        cudaEventElapsedTime(&f, ev[(size_t)t * 6 + k], ev[(size_t)t * 6 + k + 1]);
        ms5[k] += f;
      }
  if (se != cudaSuccess) return fail(ctx, "band profile: %s", cudaGetErrorString(se));
  return rc ? 1 : check_device_errors(ctx, g);
}

/* n_ticks of stepper_quick / stepper_perfect on this band, the whole per-tick protocol (see above) driven from here
 * with NCCL send/recv on the context's stream: one host synchronisation per convergence round, no Python in the loop. */
int flowamok_band_step(flowamok_context *ctx, flowamok_grid *g, int perfect, int64_t n_ticks, int64_t *rounds) {
  if (band_check(ctx, g, 0)) return 1;
  if (g->poisoned) return fail(ctx, "grid is poisoned: %s", g->poisoned);
  if (g->p2p) return band_step_p2p(ctx, g, perfect != 0, n_ticks, rounds);
  const bool multi = ctx->world > 1;
  if (multi && !ctx->comm) return fail(ctx, "neither flowamok_band_p2p_connect nor flowamok_band_comm_init has been called");
  if (multi && perfect && g->n_straddle > 0)
    return fail(ctx, "rings straddle a band cut: the NCCL transport has no bitwise reduction, use flowamok_band_p2p_*");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t sb = (size_t)g->pitch * (2 * 16 + 2 * 4 + 128), ub = (size_t)g->pitch * (4 + 8);
  if (!g->d_woken) {
    for (int k = 0; k < 2; k++) {
      CU(cudaMalloc(&g->msg_s_send[k], sb)); CU(cudaMalloc(&g->msg_s_recv[k], sb));
      CU(cudaMalloc(&g->msg_u_send[k], ub)); CU(cudaMalloc(&g->msg_u_recv[k], ub));
    }
    CU(cudaMalloc(&g->d_woken, 4));
    CU(cudaMallocHost(&g->h_woken, 4));
  }
  auto has = [&](int side) { int peer = side == 0 ? ctx->rank - 1 : ctx->rank + 1; return multi && peer >= 0 && peer < ctx->world; };
  if (multi && g->ring_swap < 0) {   // all bands agree once whether the post-ring swap is needed at all
    unsigned int f = g->n_rings > 0 ? (unsigned)g->ring_touch : 0u;
    CU(cudaMemcpyAsync(g->d_woken, &f, 4, cudaMemcpyHostToDevice, st));
    NC(g_nccl.AllReduce(g->d_woken, g->d_woken, 1, ncclUint32, ncclMax, ctx->comm, st));
    CU(cudaMemcpyAsync(g->h_woken, g->d_woken, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    g->ring_swap = *g->h_woken ? 1 : 0;
  }
  void *ss[2] = {has(0) ? g->msg_s_send[0] : nullptr, has(1) ? g->msg_s_send[1] : nullptr};
  void *sr[2] = {has(0) ? g->msg_s_recv[0] : nullptr, has(1) ? g->msg_s_recv[1] : nullptr};
  void *us[2] = {has(0) ? g->msg_u_send[0] : nullptr, has(1) ? g->msg_u_send[1] : nullptr};
  void *ur[2] = {has(0) ? g->msg_u_recv[0] : nullptr, has(1) ? g->msg_u_recv[1] : nullptr};
  auto swap_u = [&](int wake) -> int {
    if (band_pack_u(ctx, g, us[0], us[1])) return 1;
    if (band_swap(ctx, g->msg_u_send, g->msg_u_recv, ub)) return 1;
    return band_merge_u(ctx, g, ur[0], ur[1], wake);
  };
  // One convergence round = swap the boundary rows, wake the waiting cars that read a changed ghost cell, resume the
  // frontier (a no-op kernel when nothing was woken anywhere).  SPEC rounds are enqueued back to back without asking
  // whether they are needed; only then do the bands agree (one all-reduced word, ONE host synchronisation) whether the
  // last merge still woke something, and keep going in that case.  A second round is needed whenever, at any cut, a car
  // waits on a cell of the other band that itself resolved late: measured in 29 % of the ticks with one cut at 65536
  // columns, i.e. almost always with seven -- so 4 and more bands run two rounds unasked.
  static const int spec_env = [] { const char *e = getenv("FLOWAMOK_BAND_SPEC"); return e ? atoi(e) : 0; }();
  const int SPEC = spec_env >= 1 ? spec_env : (ctx->world >= 4 ? 2 : 1);
  static const bool PRESWAP = [] { const char *e = getenv("FLOWAMOK_BAND_PRESWAP"); return !e || e[0] != '0'; }();
  for (int64_t t = 0; t < n_ticks; t++) {
    if (multi) {
      if (band_state(ctx, g, ss[0], ss[1], 1)) return 1;
      if (band_swap(ctx, g->msg_s_send, g->msg_s_recv, sb)) return 1;
      if (band_state(ctx, g, sr[0], sr[1], 0)) return 1;
    }
    if (flowamok_band_u_init(ctx, g)) return 1;
    if (multi && PRESWAP && swap_u(0)) return 1;   // ghost rows := the neighbour's rows after ITS local levels
    if (flowamok_band_u_iter(ctx, g)) return 1;
    for (int round = 0; multi; round++) {
      if (swap_u(1)) return 1;
      if (round + 1 < SPEC) { if (flowamok_band_u_resume(ctx, g)) return 1; g->band_rounds++; continue; }
      CU(cudaMemcpyAsync(g->d_woken, &g->is->qn[1], 4, cudaMemcpyDeviceToDevice, st));
      NC(g_nccl.AllReduce(g->d_woken, g->d_woken, 1, ncclUint32, ncclMax, ctx->comm, st));
      CU(cudaMemcpyAsync(g->h_woken, g->d_woken, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));            // the host round trip of the tick: one 4-byte word
      g->band_rounds++;
      if (*g->h_woken == 0) break;
      if (flowamok_band_u_resume(ctx, g)) return 1;
    }
    if (perfect) {
      if (flowamok_band_rings(ctx, g)) return 1;
      if (multi && g->ring_swap && swap_u(0)) return 1;   // ring marks on a boundary row must reach the neighbour
    }
    if (flowamok_band_move(ctx, g)) return 1;
  }
  if (rounds) *rounds = g->band_rounds;
  return 0;
}

/* like flowamok_grid_upload for the car state only: heading, colour, preference and rng of the OWNED rows from
 * whole-grid host arrays; roads, cycles and halo rows are left alone (halo rows are refreshed by the next tick's swap) */
int flowamok_grid_upload_cars(flowamok_context *ctx, flowamok_grid *g, const int8_t *dy, const int8_t *dx, const uint32_t *color,
                              const uint8_t *pref, const uint64_t *rng_state) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || !dy || !dx || !color || !pref) return fail(ctx, "upload_cars needs heading, colour and preference");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int lo = g->row0 + g->own0, owned = g->own1 - g->own0;
  const size_t n = (size_t)owned * g->gw, cp = (size_t)g->pitch * 32;
  if (ensure_stage(ctx, g, n * 3)) return 1;
  int8_t *tmp = g->stage;
  const void *src[3] = {dy, dx, pref};
  for (int k = 0; k < 3; k++) CU(cudaMemcpyAsync(tmp + n * k, (const int8_t *)src[k] + (size_t)lo * g->gw, n, cudaMemcpyHostToDevice, st));
  k_pack<<<nblk((long long)g->nwords * 32, 256), 256, 0, st>>>(g->gh, g->gw, g->pitch, -g->own0, owned, nullptr, nullptr, tmp, tmp + n,
                                                              (const uint8_t *)(tmp + 2 * n), nullptr, g->head[g->cur], g->pref[g->cur]);
  CU(cudaGetLastError());
  CU(cudaMemcpy2DAsync(g->color[g->cur] + (size_t)g->own0 * cp, cp * 4, color + (size_t)lo * g->gw, (size_t)g->gw * 4,
                       (size_t)g->gw * 4, owned, cudaMemcpyHostToDevice, st));
  if (rng_state)   // NULL: the resident states stay
    CU(cudaMemcpy2DAsync(g->rng + (size_t)g->own0 * cp, cp * 8, rng_state + (size_t)lo * g->gw, (size_t)g->gw * 8,
                         (size_t)g->gw * 8, owned, cudaMemcpyHostToDevice, st));
  if (sync_color_buffers(ctx, g)) return 1;
  CU(cudaStreamSynchronize(st));
  return 0;
}

/* The cells whose RNG state can ever change are the ones with a road on both axes (k_corner_counts).  corners: (global
 * flat index, state) of every such OWNED cell, in row-major order, up to cap; *n receives their number (call with cap = 0
 * to size the arrays).  rng_scatter: write states back (cells outside this band's rows are skipped).  Host arrays. */
int flowamok_grid_rng_corners(flowamok_context *ctx, flowamok_grid *g, int64_t cap, int64_t *idx, uint64_t *state, int64_t *n) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || !n) return fail(ctx, "null argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const long long nw = (long long)(g->own1 - g->own0) * g->pitch;
  DevBuf counts, offs, tmp, dout;
  CU(counts.alloc((size_t)(nw + 1) * 4)); CU(offs.alloc((size_t)(nw + 1) * 4));
  CU(cudaMemsetAsync(counts.p, 0, (size_t)(nw + 1) * 4, st));
  k_corner_counts<<<nblk(nw, 256), 256, 0, st>>>(g->own0, g->own1, g->pitch, g->road, counts.as<unsigned int>());
  CU(cudaGetLastError());
  size_t tmpb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmpb, counts.as<unsigned int>(), offs.as<unsigned int>(), (int)(nw + 1), st));
  CU(tmp.alloc(tmpb));
  CU(cub::DeviceScan::ExclusiveSum(tmp.p, tmpb, counts.as<unsigned int>(), offs.as<unsigned int>(), (int)(nw + 1), st));
  unsigned int total = 0;
  CU(cudaMemcpyAsync(&total, offs.as<unsigned int>() + nw, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *n = total;
  const long long emit = std::min<long long>(total, cap);
  if (emit > 0 && idx && state) {
    CU(dout.alloc((size_t)emit * 16));
    long long *di = dout.as<long long>();
    u64 *ds = (u64 *)(di + emit);
    k_corner_emit<<<nblk(nw, 256), 256, 0, st>>>(g->own0, g->own1, g->gw, g->pitch, g->row0, g->road, offs.as<unsigned int>(), g->rng, emit, di, ds);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(idx, di, (size_t)emit * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(state, ds, (size_t)emit * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return check_device_errors(ctx, g);
}
int flowamok_grid_rng_scatter(flowamok_context *ctx, flowamok_grid *g, int64_t n, const int64_t *idx, const uint64_t *state) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g || n < 0 || (n > 0 && (!idx || !state))) return fail(ctx, "bad rng_scatter arguments");
  if (n == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (ensure_stage(ctx, g, (size_t)n * 16)) return 1;
  long long *di = (long long *)g->stage;
  u64 *ds = (u64 *)(di + n);
  CU(cudaMemcpyAsync(di, idx, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ds, state, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  k_rng_scatter<<<nblk(n, 256), 256, 0, st>>>(n, di, ds, g->gw, g->pitch, g->row0, g->gh, g->rng);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  return 0;
}

int flowamok_band_leaks(flowamok_context *ctx, flowamok_grid *g, int64_t *n_leaks) {
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  Diag d;
  if (read_diag(ctx, g, &d)) return 1;
  if (n_leaks) *n_leaks = (int64_t)d.leaks;
  return 0;
}

int flowamok_stats(flowamok_context *ctx, flowamok_grid *g, int64_t *ticks, int64_t *passes, int64_t *sweeps) {
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  Diag d;
  if (read_diag(ctx, g, &d)) return 1;
  if (ticks) *ticks = (int64_t)(d.ticks - g->base.ticks);
  if (passes) *passes = (int64_t)(d.passes - g->base.passes);
  if (sweeps) *sweeps = (int64_t)(d.sweeps - g->base.sweeps);
  g->base = d;
  return 0;
}

/* ticks by number of k_u_iter iterations since the grid was created: hist[k] for k = 0..30, hist[31] = 31 and more */
int flowamok_stats_hist(flowamok_context *ctx, flowamok_grid *g, int64_t *hist32) {
  if (!g || !hist32) return fail(ctx, "null argument");
  CU(cudaSetDevice(ctx->device));
  Diag d;
  if (read_diag(ctx, g, &d)) return 1;
  for (int k = 0; k < 32; k++) hist32[k] = (int64_t)d.hist[k];
  return 0;
}
/* row bands over the P2P transport: convergence rounds so far and the time CTA 0 spent waiting for neighbours (ns) */
int flowamok_band_p2p_stats(flowamok_context *ctx, flowamok_grid *g, int64_t *rounds, int64_t *wait_ns) {
  if (!g || !g->band_dev) return fail(ctx, "not a band with an open mailbox");
  CU(cudaSetDevice(ctx->device));
  BandDev bd;
  CU(cudaMemcpyAsync(&bd, g->band_dev, sizeof bd, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (rounds) *rounds = (int64_t)bd.rounds;
  if (wait_ns) *wait_ns = (int64_t)bd.waits_ns;
  return check_device_errors(ctx, g);
}

int flowamok_grid_digest(flowamok_context *ctx, const flowamok_grid *g, uint64_t *sum, uint64_t *xor_) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  unsigned long long *d = nullptr, h[2] = {0, 0};
  CU(cudaMalloc(&d, 16));
  cudaError_t e = cudaMemsetAsync(d, 0, 16, ctx->stream);
  if (e == cudaSuccess) {
    k_digest<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(g->own0, g->own1, g->gw, g->pitch, g->row0, g->head[g->cur], g->pref[g->cur],
                                                        g->color[g->cur], g->rng, d);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, "flowamok_grid_digest: %s", cudaGetErrorString(e));
  if (sum) *sum = h[0];
  if (xor_) *xor_ = h[1];
  return check_device_errors(ctx, const_cast<flowamok_grid *>(g));
}

int flowamok_generate_synthetic(flowamok_context *ctx, flowamok_grid *g, uint64_t seed, uint32_t density16, uint32_t ring_full16) {
  if (!ctx || ctx->device < 0) return fail(ctx, "no CUDA device");
  if (!g) return fail(ctx, "null grid");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CU(cudaStreamSynchronize(st));
  k_synth<<<nblk((long long)g->nwords * 32, 256), 256, 0, st>>>(g->gh, g->gw, g->pitch, g->row0, g->gh_global, seed, density16,
                                                               ring_full16, g->road, g->head[g->cur], g->pref[g->cur],
                                                               g->color[g->cur], g->rng);
  CU(cudaGetLastError());
  g->rng_inc = (0xda3e39cb94b95bdbULL << 1) | 1ULL;
  if (sync_color_buffers(ctx, g)) return 1;
  CU(cudaFree(g->ring_words)); CU(cudaFree(g->ring_off));
  g->ring_words = nullptr; g->ring_off = nullptr; g->n_rings = g->n_ring_cells = g->n_ring_words = 0; g->ring_stride = 0;
  g->h_off.clear(); g->h_y.clear(); g->h_x.clear(); g->h_cy.clear(); g->h_cx.clear();
  const int ghg = g->gh_global;
  int nby = ghg >= 15 ? (ghg - 15) / 16 + 1 : 0, nbx = g->gw >= 23 ? (g->gw - 23) / 16 + 1 : 0;
  std::vector<int> no_sids;
  if (alloc_straddle_bits(ctx, g, 0, no_sids)) return 1;
  g->ring_touch = 0; g->ring_swap = g->cuts.empty() ? -1 : 0;
  if (nby > 0 && nbx > 0) {
    const int lo = g->row0 + g->own0, hi = g->row0 + g->own1;
    // ring rows of block row by: 16*by + 2 .. 16*by + 14.  A cut at row c (rows < c above, >= c below) runs THROUGH the
    // rings of block row c / 16 iff c % 16 is in 3..14, and a boundary row (c - 1 or c) holds ring cells iff c % 16 is in 2..15.
    SynthCuts cutrows;
    cutrows.n = 0;
    bool touch = false;
    std::vector<int64_t> cs = g->cuts;
    if (cs.empty() && g->halo != 0) cs = {0, lo, hi, ghg};   // cuts unknown: only this band's own boundaries can be judged
    for (size_t i = 1; i + 1 < cs.size(); i++) {
      const int64_t c = cs[i];
      if (c <= 0 || c >= ghg) continue;
      const int m = (int)(c % 16), by = (int)(c / 16);
      if (by >= nby) continue;
      touch |= m >= 2;
      if (m >= 3 && m <= 14) {
        if (g->cuts.empty()) return fail(ctx, "synthetic rings straddle this row band: call flowamok_band_set_cuts first (or cut at multiples of 16)");
        bool dup = false;
        for (int q = 0; q < cutrows.n; q++) dup |= cutrows.by[q] == by;
        if (!dup) cutrows.by[cutrows.n++] = by;
      }
    }
    DevBuf cntbuf;
    CU(cntbuf.alloc(8));
    unsigned int *cnt = cntbuf.as<unsigned int>();
    CU(cudaMemsetAsync(cnt, 0, 8, st));
    const long long my_blocks = ((long long)(hi - lo) / 16 + 3) * nbx;   // ring blocks that can intersect this band
    CU(cudaMalloc(&g->ring_words, (size_t)my_blocks * SYNTH_RING_WORDS * sizeof(RingWord)));
    int *sid = nullptr;
    if (cutrows.n > 0) CU(cudaMalloc(&sid, (size_t)my_blocks * sizeof(int)));
    k_synth_rings<<<nblk((long long)nby * nbx, 128), 128, 0, st>>>(ghg, g->gw, g->pitch, g->row0, lo, hi, seed, nby, nbx,
                                                                   g->ring_words, sid, cnt, cutrows);
    CU(cudaGetLastError());
    unsigned int n[2] = {0, 0};
    CU(cudaMemcpyAsync(n, cnt, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    g->n_rings = n[0]; g->n_ring_cells = (long long)n[0] * 48; g->n_ring_words = (long long)n[0] * SYNTH_RING_WORDS;
    g->ring_stride = SYNTH_RING_WORDS;
    if (cutrows.n > 0) {
      if (alloc_straddle_bits(ctx, g, (long long)cutrows.n * nbx, no_sids)) return 1;
      g->ring_sid = sid;
    }
    // the same decision on every band (it depends on the cuts only); without the cuts: this band's own view, agreed later
    g->ring_touch = touch ? 1 : 0;
    g->ring_swap = g->cuts.empty() ? -1 : (touch ? 1 : 0);
  }
  CU(cudaStreamSynchronize(st));
  g->tick = 0;
  return 0;
}

}  // extern "C"

#include "fa_futhark.inc"
