// flowamok_b200 -- CUDA kernels (sm_100a) of the per-tick grid update.
//
// One tick = k_u_local, [k_rings], then -- side by side -- k_u_iter on a high-priority stream and k_move, then k_move_fix:
//   k_u_local  : streaming, one warp per strip of 30 plane words x RS rows (fa_stream.h, UStrip): statics and the first
//                FA_U_LEVELS (3) Jacobi levels of update_can_be_moved_to_from (steppers.fut:9-56) in registers, worklist
//                entries for the words that still hold an unresolved waiting car, frontier.
//   k_rings    : resolve_cycles (steppers.fut:189-224) on the sparse ring list.  It does not depend on the fixpoint
//                (ring_word_ok), so it runs before it and its grants are part of what k_move reads.
//   k_u_iter   : cooperative persistent kernel, one small CTA per SM; relaxes the frontier under a grid-wide barrier and
//                device-side list lengths until the global fixpoint (steppers.fut:147-164) -- no host round trip per
//                iteration, each iteration touches only unresolved words.  Its grants go into a delta plane (mymx_B), so
//                that k_move can gather from the planes k_u_local published at the same time.  On a row band with mapped
//                peers (fa_band.cuh) the convergence rounds with the neighbour bands run inside this launch too.
//   k_move     : streaming like k_u_local (MTile): move_cell + reset_cells (steppers.fut:58-145): new car planes, colour
//                payload through a lazy double buffer (only the 32-byte sectors that received a car in this tick or the
//                previous one are rewritten), per-cell pcg32 draws at forks, leak count.
//   k_move_fix : recomputes the words around every word whose my|mx the fixpoint changed after k_u_local (fa_stream.h,
//                fix_word); a few thousand words per tick.
// plus packing / unpacking / generator / render helpers.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "fa_band.cuh"
#include "fa_stream.h"
#include "fa_tiles.h"

namespace fa {

constexpr int STRIP_W = 30;                // owned plane words per strip (lanes 0 and 31 carry the neighbours)
constexpr int STRIP_THREADS = 128;         // k_u_local / k_move block: 4 warps = 4 strips side by side
#ifndef FA_U_LEVELS
#define FA_U_LEVELS 3                      // Jacobi levels k_u_local resolves in registers (queues of <= 3 cars)
#endif
#ifndef FA_U_MINB
#define FA_U_MINB 4                        // resident CTAs per SM the register allocation of k_u_local aims for
#endif
#ifndef FA_M_MINB
#define FA_M_MINB 5                        // the same for k_move
#endif
#ifndef FA_UIT_THREADS
#define FA_UIT_THREADS 128
#endif
#ifndef FA_UIT_MINB
#define FA_UIT_MINB 8                      // register budget of k_u_iter: 64 K / (threads * this) per thread
#endif
constexpr int UIT_THREADS = FA_UIT_THREADS;   // k_u_iter block; ONE small CTA per SM: k_move runs beside the fixpoint

struct IterState {
  unsigned int qn[3];       // frontier lengths: iteration it >= 1 reads slot it % 3 and writes (it + 1) % 3
  unsigned int bar[2];      // grid barrier: arrivals, generation
  unsigned int error;       // barrier watchdog tripped
  unsigned int chg_n;       // words listed for the fix-up of the speculative move (k_u_iter wipes the previous tick's first)
  unsigned int pad[1];
  unsigned long long cum_it, cum_visits;   // iterations / visited words of all k_u_iter launches so far (jam detection)
};

struct Diag {
  unsigned long long leaks;    // summed cell leaks
  unsigned long long passes;   // fixpoint iterations of k_u_iter
  unsigned long long sweeps;   // worklist entries relaxed
  unsigned long long ticks;
  unsigned long long hist[32]; // ticks by number of k_u_iter iterations (last bucket: 31 and more)
};

// ---------------------------------------------------------------------------------------------
// k_u_local / k_move: warp `wg` of the grid takes strip (wg / nsx, wg % nsx): rows [own0 + sy*rs, +rs), words
// [sx*STRIP_W, +STRIP_W).  The four warps of a CTA are four strips side by side, so the halo words they share hit L1.
// ---------------------------------------------------------------------------------------------
template <bool BAND>
__global__ void __launch_bounds__(STRIP_THREADS, FA_U_MINB)
k_u_local(const __grid_constant__ GridView g, int nsx, int nsy, int rs, UEntry *table, unsigned int *qlist0,
          unsigned int *qstage, IterState *is) {
  const int wg = (int)(blockIdx.x * (STRIP_THREADS / 32) + (threadIdx.x >> 5));
  if (wg >= nsx * nsy) return;   // whole warp
  const int sy = wg / nsx, sx = wg - sy * nsx;
  const int ya = g.own0 + sy * rs, yb = min(g.own1, ya + rs);
  // staging segment of this strip: qstage is the second frontier list, idle until k_u_iter starts
  UStrip<STRIP_W, FA_U_LEVELS, BAND>::run(g, sx, ya, yb, (int)(threadIdx.x & 31), table, qstage + (size_t)wg * STRIP_W * rs, qlist0,
                                    &is->qn[0]);
}

// Sense-reversing grid barrier; co-residency is guaranteed by the cooperative launch.  A watchdog
// turns a would-be hang into an error flag.
__device__ __forceinline__ bool grid_barrier(IterState *is, unsigned int nblocks) {
  __shared__ int ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    ok = 1;
    volatile unsigned int *vgen = &is->bar[1];
    unsigned int gen = *vgen;
    __threadfence();
    unsigned int prev = atomicAdd(&is->bar[0], 1u);
    if (prev == nblocks - 1) {
      atomicExch(&is->bar[0], 0u);
      __threadfence();
      atomicAdd(&is->bar[1], 1u);
    } else {
      long long spins = 0;
      while (*vgen == gen) {
        if (++spins > (1LL << 28)) { atomicExch(&is->error, 1u); ok = 0; break; }   // about half a minute
        __nanosleep(100);
      }
    }
    __threadfence();
  }
  __syncthreads();
  return ok != 0;
}

// ---------------------------------------------------------------------------------------------
// k_u_iter: frontier iterations.  Iteration `it` visits the entry indices in qlist[it & 1] (iteration 0: the
// frontier k_u_local appended to qlist0) and appends the words it wakes to the other list, warp-aggregated with one
// global atomic per warp and push slot.  Ends when an iteration produces an empty list.
// ---------------------------------------------------------------------------------------------
#ifndef FA_UIT_TAIL
#define FA_UIT_TAIL 128
#endif
constexpr unsigned int UIT_TAIL = FA_UIT_TAIL;   // frontier length below which CTA 0 finishes the fixpoint alone
#ifndef FA_UIT_CHASE
#define FA_UIT_CHASE 0
#endif
constexpr int UIT_CHASE = FA_UIT_CHASE;   // links of a waiting queue one visit follows before it hands over to the list

// first_it = 0: start from the frontier k_u_local left in qlist0 / qn[0]; 1: resume from qlist1 / qn[1] (band merge)
//
// P2P = true (a row band whose neighbours' mailboxes are mapped, fa_band.cuh): the convergence rounds of the band protocol
// run inside this launch.  (k_band_xu has swapped the boundary rows once before it: the ghost rows hold the neighbours' rows
// after THEIR local levels.)  After every local fixpoint CTA 0 swaps the boundary rows again; the words its merge wakes
// become the next frontier, and all bands vote -- device to device -- whether anybody woke anything.  No NCCL call, no
// kernel boundary and no host round trip per convergence round; the other CTAs wait at the grid barrier.
template <bool P2P>
__global__ void __launch_bounds__(UIT_THREADS, FA_UIT_MINB)   // <= 64 registers: 8 K of an SM's 64 K, so that k_move keeps 4 of its 5 CTAs beside it
k_u_iter(const __grid_constant__ GridView g, const UEntry *table, unsigned int *qlist0, unsigned int *qlist1,
         unsigned int first_it, IterState *is, Diag *diag, const __grid_constant__ BandLink L, unsigned long long *host_stats) {
  const int tid = threadIdx.x, nt = blockDim.x, b = blockIdx.x, nb = gridDim.x, lane = tid & 31;
  volatile IterState *vis = is;
  unsigned int *ql[2] = {qlist0, qlist1};
  unsigned long long relaxed = 0;

  // Visit entry `idx` (if `has`) and CHASE the chain of waiting cars behind it: of the words the visit wakes, the
  // thread keeps the first for itself and visits it at once (it has claimed it: QUEUED is set), the others are
  // appended to dst, warp-aggregated with one global atomic per warp and push slot.  A queue of cars thus resolves
  // in one iteration, link by link, instead of one grid-wide iteration per car.  No block barrier anywhere.
  auto visit_and_flush = [&](auto cta_scope, bool has, unsigned int idx, unsigned int *dst, unsigned int *out_cnt) {
    for (int hop = 0; __any_sync(0xFFFFFFFFu, has); hop++) {
      unsigned int push[9];
      int np = 0;
      if (has) {
        const UEntry e = table[idx];
        np = u_iter_entry<decltype(cta_scope)::value>(g, e, idx, push);
        relaxed++;
      }
      has = np > 0 && hop < UIT_CHASE;
      if (has) {   // keep the first woken word: shift the rest down
        idx = push[0];
#pragma unroll
        for (int j = 0; j < 8; j++) push[j] = push[j + 1];
        np--;
      }
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const bool p = j < np;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, p);
        if (m) {
          const int leader = __ffs(m) - 1;
          unsigned bs = 0;
          if (lane == leader) bs = atomicAdd(out_cnt, (unsigned)__popc(m));
          bs = __shfl_sync(0xFFFFFFFFu, bs, leader);
          if (p) dst[bs + __popc(m & ((1u << lane) - 1u))] = push[j];
        }
      }
    }
  };

  // iteration `it` visits the entry indices in qlist[it & 1] (length qn[it % 3]) and appends to qlist[(it + 1) & 1]
  unsigned int it = first_it, its_done = 0;
  unsigned long long visits = 0;   // sum of the frontier lengths (the same number in every CTA that takes part)
  unsigned int useq = P2P ? *(volatile unsigned int *)&L.dev->u_seq : 0u;   // number of the last u message sent (k_band_xu counts too)
  if (g.mymx_B) {   // speculative move: the previous tick's deltas (its fix-up is done) are wiped, the list starts empty
    const unsigned int n = *(volatile unsigned int *)g.chg_n;
    for (unsigned int k = (unsigned)b * nt + tid; k < n; k += (unsigned)nb * nt) g.mymx_B[g.chg_list[k]] = 0ull;
    if (!grid_barrier(is, nb)) return;
    if (b == 0 && tid == 0) *g.chg_n = 0u;
    if (!grid_barrier(is, nb)) return;
  }
  for (;;) {   // convergence rounds (one, unless P2P; the boundary rows were swapped once before this launch: k_band_xu)
    bool tail = false;   // once the frontier is short, CTA 0 finishes alone: no more grid-wide barriers
    for (;; it++) {
      const unsigned int n = vis->qn[it % 3u];
      if (n == 0) break;                       // same value in every CTA: uniform exit
      if (!tail && n <= UIT_TAIL) {            // uniform decision: every CTA read the same n after the barrier
        if (b != 0) break;
        tail = true;
      }
      const unsigned int *src = ql[it & 1];
      unsigned int *dst = ql[(it + 1) & 1];
      unsigned int *out_cnt = &is->qn[(it + 1) % 3u];
      if (b == 0 && tid == 0) is->qn[(it + 2) % 3u] = 0;   // output slot of iteration it+1; idle since barrier(it-1)
      const unsigned int stride = tail ? (unsigned)nt : (unsigned)nb * nt;
      for (unsigned int k0 = tail ? 0u : (unsigned)b * nt; k0 < n; k0 += stride) {
        const unsigned int k = k0 + tid;
        if (tail) visit_and_flush(std::true_type(), k < n, k < n ? src[k] : 0u, dst, out_cnt);
        else visit_and_flush(std::false_type(), k < n, k < n ? src[k] : 0u, dst, out_cnt);
      }
      its_done++;
      visits += n;
      if (tail) { __threadfence(); __syncthreads(); }
      else if (!grid_barrier(is, nb)) return;
    }
    if (!P2P) break;
    // One convergence round, the whole grid sharing the row work.  CTA 0 stands at the iteration whose list is empty (slot
    // it % 3 == 0, slot (it + 1) % 3 was cleared one iteration ago): the merge appends the woken words to ql[it & 1] /
    // qn[it % 3], which is where the next iteration reads.  The other CTAs may have left the loop early (tail): they take `it`
    // from CTA 0.
    if (b == 0 && tid == 0) L.dev->round_it = it;
    if (!grid_barrier(is, nb)) return;                      // the local fixpoint is over: the boundary rows are final for this round
    it = *(volatile unsigned int *)&L.dev->round_it;
    useq++;
    band_push_u(g, L, useq, b * nt + tid, nb * nt);
    __threadfence_system();
    if (!grid_barrier(is, nb)) return;                      // every slice is pushed and fenced
    if (b == 0) { band_flags_u(L, useq); __syncthreads(); }
    if (!grid_barrier(is, nb)) return;                      // the neighbours' rows have arrived
    band_merge_u(g, L, table, useq, true, ql[it & 1], &is->qn[it % 3u], b * nt + tid, nb * nt);
    if (!grid_barrier(is, nb)) return;                      // the woken list is complete
    if (b == 0) {
      const bool any = band_vote(L, vis->qn[it % 3u] != 0u);
      if (tid == 0) {
        L.dev->go = any ? 1u : 0u; L.dev->rounds++; L.dev->u_seq = useq;
        is->qn[(it + 1) % 3u] = 0;
      }
    }
    if (!grid_barrier(is, nb)) return;
    if (*(volatile unsigned int *)&L.dev->go == 0u) break;
  }
  {   // visits of this CTA: warp sums, then one shared and one global atomic
    __shared__ unsigned long long cta_relaxed;
    if (tid == 0) cta_relaxed = 0ull;
    __syncthreads();
    unsigned int r32 = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)relaxed);
    if (lane == 0 && r32) atomicAdd(&cta_relaxed, (unsigned long long)r32);
    __syncthreads();
    relaxed = cta_relaxed;
  }
  if (tid == 0) {
    if (relaxed) atomicAdd(&diag->sweeps, relaxed);
    if (b == 0) {
      is->qn[0] = 0; is->qn[1] = 0; is->qn[2] = 0;   // the slot everyone just read is 0 already
      atomicAdd(&diag->passes, (unsigned long long)its_done);
      atomicAdd(&diag->hist[its_done < 31u ? its_done : 31u], 1ull);
      is->cum_it += its_done; is->cum_visits += visits;
      if (host_stats) {   // pinned host memory: posted stores, nobody waits (the host's jam detection reads them whenever)
        host_stats[0] = is->cum_it; host_stats[1] = is->cum_visits;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row bands.  One launch serves both sides of the band (blockIdx.y = side; a side without a neighbour has a null
// buffer): the per-tick protocol is latency-bound, so every launch saved shortens the tick.
//   state message: 2 heading rows + 2 preference rows + 1 colour row     u message: 1 row of calc + 1 row of my|mx
// ---------------------------------------------------------------------------------------------
struct BandBufs {
  char *buf[2];
};
// pack = true: owned boundary rows -> message; false: message -> halo rows
__global__ void k_band_state(int pitch, int own0, int own1, U4 *head, u32 *pref, u32 *color, BandBufs bufs, int pack) {
  const int side = blockIdx.y;
  char *b = bufs.buf[side];
  if (!b) return;
  const size_t P = (size_t)pitch;
  const int r2 = pack ? (side == 0 ? own0 : own1 - 2) : (side == 0 ? own0 - 2 : own1);
  const int rc = pack ? (side == 0 ? own0 : own1 - 1) : (side == 0 ? own0 - 1 : own1);
  U4 *mh = reinterpret_cast<U4 *>(b);                 // [2P] headings
  u32 *mp = reinterpret_cast<u32 *>(b + 2 * P * 16);  // [2P] preferences
  U4 *mc = reinterpret_cast<U4 *>(b + 2 * P * 20);    // [8P] colours of one row, 16 bytes at a time
  U4 *gh = head + (size_t)r2 * P;
  u32 *gp = pref + (size_t)r2 * P;
  U4 *gc = reinterpret_cast<U4 *>(color + (size_t)rc * P * 32);
  const size_t n = 8 * P;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (pack) {
      mc[i] = gc[i];
      if (i < 2 * P) { mh[i] = gh[i]; mp[i] = gp[i]; }
    } else {
      gc[i] = mc[i];
      if (i < 2 * P) { gh[i] = mh[i]; gp[i] = mp[i]; }
    }
  }
}
__global__ void k_band_pack_u(int pitch, int own0, int own1, const u32 *calc, const u64 *mymx, BandBufs bufs) {
  const int side = blockIdx.y;
  char *b = bufs.buf[side];
  if (!b) return;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= pitch) return;
  const size_t o = (size_t)(side == 0 ? own0 : own1 - 1) * pitch + w;
  reinterpret_cast<u32 *>(b)[w] = calc[o];
  reinterpret_cast<u64 *>(b + (size_t)pitch * 4)[w] = mymx[o];
}
// OR the neighbour's boundary row (calc, my|mx) into the ghost row; wake the owned words that hold a waiting car which
// reads a changed cell.  One thread per word and side.
__global__ void k_band_merge(const __grid_constant__ GridView g, BandBufs bufs, int wake, const UEntry *table, unsigned int *qlist1,
                             IterState *is) {
  const int side = blockIdx.y;
  const char *b = bufs.buf[side];
  if (!b) return;
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= g.pitch) return;
  const int ghost_row = side == 0 ? g.own0 - 1 : g.own1, own_row = side == 0 ? g.own0 : g.own1 - 1;
  const u32 *recv_calc = reinterpret_cast<const u32 *>(b);
  const u64 *recv_mymx = reinterpret_cast<const u64 *>(b + (size_t)g.pitch * 4);
  size_t o = (size_t)ghost_row * g.pitch + w;
  u32 oc = g.calc[o], nc = oc | recv_calc[w];
  u64 om = g.mymx[o], nm = om | recv_mymx[w];
  if (nc == oc && nm == om) return;
  g.mymx[o] = nm;
  g.calc[o] = nc;
  if (!wake) return;
  const u32 chg = (oc ^ nc) | (u32)(om ^ nm) | (u32)((om ^ nm) >> 32);
  for (int dx = -1; dx <= 1; dx++) {
    if (!band_word_reads_ghost(g, table, side, own_row, w, dx, chg)) continue;
    u32 q = u_try_queue(g, own_row, w + dx);
    if (q != U_NONE) qlist1[atomicAdd(&is->qn[1], 1u)] = q & ~U_QUEUED;
  }
}

// ---------------------------------------------------------------------------------------------
// resolve_cycles on sparse rings (steppers.fut:189-224; disjointness argument SURVEY 8a-R).
// A ring is stored per plane word it touches: the ring's cells in that word by check direction, and which
// of them are granted along y / x if the whole ring passes (the outcome of steppers.fut:206-211).
// ---------------------------------------------------------------------------------------------
struct alignas(32) RingWord {
  u32 wi;                 // plane word index
  u32 mN, mS, mW, mE;     // ring cells of this word whose check direction is N / S / W / E
  u32 gY, gX;             // cells granted direction.y / direction.x when the ring passes
  u32 pad;
};
// The grants of a passing ring go into my|mx.  (A ring whose cells all head along it passes whatever the rest of the fixpoint
// does: every ring cell waits for the next one, so none is ever calculated, steppers.fut:197-200.  Ring resolution can
// therefore run BEFORE k_u_iter -- the production order: its grants are then part of the plane the speculative k_move reads.
// The visits never take a ring's grants for the fixpoint's: they only count grants of calculated cells.)
__device__ __forceinline__ void ring_grant(const GridView &g, const RingWord &w) {
  if (w.gY | w.gX) atomicOr((unsigned long long *)&g.mymx[w.wi], (unsigned long long)w.gY | ((unsigned long long)w.gX << 32));
}
// Every ring cell of the word holds a car heading exactly the check direction (:197-200).  `calculated` is not read: if ALL
// cells of a ring head along it, each waits for the next (find_cycles steps to pos + check, :247-293), so none of them is ever
// calculated by the fixpoint -- and if one does not, the ring fails here anyway.  (This is also why ring resolution may run
// before k_u_iter: see ring_grant.)
__device__ __forceinline__ bool ring_word_ok(const GridView &g, const RingWord &e) {
  const u32 m = e.mN | e.mS | e.mW | e.mE;
  if (m == 0u) return true;   // padding entry
  const U4 h = g.head_in[e.wi];
  const u32 n = h.x & h.y & ~h.z, s_ = h.x & ~h.y & ~h.z, w = h.z & h.w & ~h.x, ea = h.z & ~h.w & ~h.x;
  return (n & e.mN) == e.mN && (s_ & e.mS) == e.mS && (w & e.mW) == e.mW && (ea & e.mE) == e.mE;
}
constexpr int RINGS_PER_BLOCK = 256;
constexpr int RING_ITEMS = 4;   // (survivor, word) items a thread tests at once in the uniform-stride path
// Phase 1: one thread per ring tests the ring's first word (almost every ring that is not completely full fails
// here); survivors are compacted in shared memory.  Phase 2 tests all words of the survivors and, if a ring passes,
// ORs its grants into my|mx (one 64-bit atomic per word).  Rings of one uniform length (off == nullptr: the synthetic
// network) are handled as a flat list of (survivor, word) items spread over all threads, every thread holding several
// independent loads in flight -- two memory latencies per CTA; CSR rings take one warp per survivor.
//
// Row bands: a ring that straddles a cut is listed by every band it touches, each with the words of its own rows
// only, and carries its straddle id (sid[k] >= 0, the same number in every band).  Such a ring is only JUDGED here:
// a band that finds an objection clears bit sid of `partial`; k_band_xring ANDs the bands' verdicts and k_rings_mark
// applies the grants.  sid == nullptr: no ring straddles.
__global__ void __launch_bounds__(RINGS_PER_BLOCK)
k_rings(const __grid_constant__ GridView g, const RingWord *words, const unsigned long long *off, long long n, int stride,
        const int *sid, u32 *partial) {
  __shared__ unsigned int surv[RINGS_PER_BLOCK];
  __shared__ unsigned int okf[RINGS_PER_BLOCK];
  __shared__ unsigned int wbase[RINGS_PER_BLOCK / 32 + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long k = (long long)blockIdx.x * RINGS_PER_BLOCK + tid;
  bool pass0 = false;
  int my_sid = -1;
  if (k < n) {
    const unsigned long long b = off ? off[k] : (unsigned long long)k * stride;
    const unsigned long long e = off ? off[k + 1] : b + stride;
    pass0 = e > b && ring_word_ok(g, words[b]);
    if (sid) my_sid = sid[k];
    if (my_sid >= 0 && !pass0) atomicAnd(&partial[my_sid >> 5], ~(1u << (my_sid & 31)));
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, pass0);
  if (lane == 0) wbase[warp + 1] = __popc(m);
  okf[tid] = 1u;
  __syncthreads();
  if (tid == 0) { wbase[0] = 0; for (int i = 1; i <= RINGS_PER_BLOCK / 32; i++) wbase[i] += wbase[i - 1]; }
  __syncthreads();
  if (pass0) surv[wbase[warp] + __popc(m & ((1u << lane) - 1u))] = (unsigned int)tid;
  __syncthreads();
  const int nsurv = (int)wbase[RINGS_PER_BLOCK / 32];
  const long long ring0 = (long long)blockIdx.x * RINGS_PER_BLOCK;
  if (!off) {   // uniform ring length: flat (survivor, word) items
    const int nitems = nsurv * stride;
    for (int i0 = tid; i0 < nitems; i0 += RINGS_PER_BLOCK * RING_ITEMS) {   // test: RING_ITEMS independent loads in flight
      RingWord w[RING_ITEMS];
      int si[RING_ITEMS];
#pragma unroll
      for (int q = 0; q < RING_ITEMS; q++) {
        const int i = i0 + q * RINGS_PER_BLOCK;
        si[q] = -1;
        if (i < nitems) {
          si[q] = i / stride;
          w[q] = words[(unsigned long long)(ring0 + surv[si[q]]) * stride + (unsigned)(i - si[q] * stride)];
        }
      }
      bool ok[RING_ITEMS];
#pragma unroll
      for (int q = 0; q < RING_ITEMS; q++) ok[q] = si[q] < 0 || ring_word_ok(g, w[q]);
#pragma unroll
      for (int q = 0; q < RING_ITEMS; q++)
        if (!ok[q]) okf[si[q]] = 0u;   // benign race: everybody writes 0
    }
    __syncthreads();
    if (sid)   // straddling survivors: verdict only
      for (int s_ = tid; s_ < nsurv; s_ += RINGS_PER_BLOCK) {
        const int d = sid[ring0 + surv[s_]];
        if (d >= 0) {
          if (!okf[s_]) atomicAnd(&partial[d >> 5], ~(1u << (d & 31)));
          okf[s_] = 0u;   // not marked here
        }
      }
    if (sid) __syncthreads();
    for (int i = tid; i < nitems; i += RINGS_PER_BLOCK) {   // mark the rings that passed
      const int s_ = i / stride;
      if (!okf[s_]) continue;
      const RingWord w = words[(unsigned long long)(ring0 + surv[s_]) * stride + (unsigned)(i - s_ * stride)];
      ring_grant(g, w);
    }
    return;
  }
  for (int si = warp; si < nsurv; si += RINGS_PER_BLOCK / 32) {
    const long long r = ring0 + surv[si];
    const unsigned long long b = off[r], e = off[r + 1];
    const int d = sid ? sid[r] : -1;
    bool ok = true;
    for (unsigned long long t = b + lane; t < e; t += 32) ok = ok && ring_word_ok(g, words[t]);
    ok = __all_sync(0xFFFFFFFFu, ok);
    if (d >= 0) {
      if (!ok && lane == 0) atomicAnd(&partial[d >> 5], ~(1u << (d & 31)));
      continue;
    }
    if (!ok) continue;
    for (unsigned long long t = b + lane; t < e; t += 32) {
      const RingWord w = words[t];
      ring_grant(g, w);
    }
  }
}
// the grants of the straddling rings that passed in EVERY band (pass: AND of the bands' verdicts); a thread per ring
__global__ void k_rings_mark(const __grid_constant__ GridView g, const RingWord *words, const unsigned long long *off, long long n, int stride,
                             const int *sid, const u32 *pass) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int d = sid[k];
  if (d < 0 || !((pass[d >> 5] >> (d & 31)) & 1u)) return;
  const unsigned long long b = off ? off[k] : (unsigned long long)k * stride, e = off ? off[k + 1] : b + stride;
  for (unsigned long long t = b; t < e; t++) {
    const RingWord w = words[t];
    ring_grant(g, w);
  }
}

#ifdef FA_B1_TMA
// ---------------------------------------------------------------------------------------------
// A/B VARIANT, not the production path (tools/tune_variants.py build tma=-DFA_B1_TMA): phase B1 of k_move -- the copy of the
// listed 32-byte colour sectors -- through the bulk-async copy engine (cp.async.bulk global -> shared with an mbarrier,
// shared -> global with a bulk group) instead of LDG.128 / STG.128 through registers.  Measured and rejected: DESIGN 6b.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
}
constexpr int B1_TMA_CHUNK = 256;   // sectors staged per round: 8 KB of shared memory
// all threads of the CTA; stage: B1_TMA_CHUNK * 32 bytes, 128-byte aligned; bar: one mbarrier (initialised, count 1)
template <class MT>
__device__ __forceinline__ void phaseB1_tma(const GridView &g, typename MT::Smem sm, int tsx, int ya, int tid, int nt, char *stage, uint64_t *bar) {
  const unsigned short *sect = sm.sect();
  const int ns = (int)sm.cnt()[0];
  uint32_t parity = 0;
  for (int c0 = 0; c0 < ns; c0 += B1_TMA_CHUNK) {
    const int n = min(B1_TMA_CHUNK, ns - c0);
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)n * 32u);
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
      const int id = sect[c0 + k], row = id / (MT::TW * 4), sc = id - row * (MT::TW * 4);
      const size_t off = ((size_t)(ya + row) * g.pitch + (size_t)tsx * STRIP_W) * 32 + (size_t)sc * 8;
      bulk_g2s(stage + (size_t)k * 32, g.color_in + off, 32u, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int k = tid; k < n; k += nt) {
      const int id = sect[c0 + k], row = id / (MT::TW * 4), sc = id - row * (MT::TW * 4);
      const size_t off = ((size_t)(ya + row) * g.pitch + (size_t)tsx * STRIP_W) * 32 + (size_t)sc * 8;
      bulk_s2g(g.color_out + off, stage + (size_t)k * 32, 32u);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging buffer may be overwritten
    __syncthreads();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // the writes are done before the arrival patches
}
#endif

// ---------------------------------------------------------------------------------------------
// move + reset
// ---------------------------------------------------------------------------------------------
// k_move CTA: NW strips side by side (4 measured best: 2 and 3 lose memory parallelism in phase B, 6 and 8 lose
// resident CTAs to overlap phase A of one tile with phase B of another)
constexpr int M_WARPS = 4;
typedef MTile<STRIP_W, M_WARPS> MTL;
template <int NW>
__global__ void __launch_bounds__(NW * 32, FA_M_MINB)
k_move(const __grid_constant__ GridView g, int nsx, int ntx, int rs, Diag *diag, unsigned int *leak_plane) {
  typedef MTile<STRIP_W, NW> MTL;   // shadows the production typedef: the kernel body is width-generic
  extern __shared__ __align__(16) u32 smem_raw[];
  const typename MTL::Smem sm = {smem_raw, rs};
  const int tid = threadIdx.x, nt = NW * 32, wi = tid >> 5, lane = tid & 31;
  const int ty = blockIdx.x / ntx, tx = blockIdx.x - ty * ntx;
  const int ya = g.own0 + ty * rs, yb = min(g.own1, ya + rs);
  const int sx = tx * NW + wi;
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&diag->ticks, 1ull);
  MTL::phaseB_clear(sm, tid);
  const unsigned int nleak = MTL::phaseA(g, sx < nsx, sx, wi, ya, yb, lane, leak_plane, sm);
  if (nleak) atomicAdd(&diag->leaks, (unsigned long long)nleak);
  __syncthreads();
  MTL::phaseB0(sm, yb - ya, tid, nt);
  __syncthreads();
#ifdef FA_B1_TMA
  {
    const size_t lists = ((size_t)rs * MTL::TW * 6 + MTL::ARR_CAP + 4) * 4;   // MTile::Smem::bytes(rs)
    char *stage = reinterpret_cast<char *>(smem_raw) + ((lists + 127) & ~(size_t)127);
    uint64_t *bar = reinterpret_cast<uint64_t *>(stage + B1_TMA_CHUNK * 32);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    phaseB1_tma<MTL>(g, sm, tx * NW, ya, tid, nt, stage, bar);
  }
#else
  MTL::phaseB1(g, sm, tx * NW, ya, tid, nt);
#endif
  __syncthreads();   // the sector copies are ordered before the arrival patches
  MTL::phaseB2(g, sm, tx * NW, ya, yb - ya, tid, nt);
}

// The fix-up of the speculative move (fa_stream.h, fix_word): the words around every word whose my|mx the fixpoint changed
// after the snapshot.  The list length lives on the device (is->chg_n): a fixed grid strides over it.
__global__ void __launch_bounds__(256)
k_move_fix(const __grid_constant__ GridView g, u32 *claim, u32 tag, IterState *is, Diag *diag) {
  const unsigned int n = *(volatile unsigned int *)&is->chg_n * 9u;
  unsigned int leaks = 0;
  for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) leaks += fix_item(g, claim, tag, t);
  leaks = __reduce_add_sync(0xFFFFFFFFu, leaks);
  if ((threadIdx.x & 31) == 0 && leaks) atomicAdd(&diag->leaks, (unsigned long long)leaks);
}

// ---------------------------------------------------------------------------------------------
// packing helpers: one warp builds one plane word with __ballot_sync
// ---------------------------------------------------------------------------------------------
// host arrays are [gh_global][gw]; local row y holds global row y + row0 (rows outside the grid stay zero)
__global__ void k_pack(int gh, int gw, int pitch, int row0, int gh_global, const int8_t *uy, const int8_t *ux,
                       const int8_t *dy, const int8_t *dx, const uint8_t *pref, U4 *road, U4 *head, u32 *prefp) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long long nwords = (long long)gh * pitch;
  if (warp >= nwords) return;
  int y = (int)(warp / pitch), w = (int)(warp % pitch);
  int x = w * 32 + lane;
  int yg = y + row0;
  bool in = x < gw && yg >= 0 && yg < gh_global;
  size_t i = (size_t)(in ? yg : 0) * gw + (in ? x : 0);
  int vuy = (in && uy) ? uy[i] : 0, vux = (in && ux) ? ux[i] : 0;
  int vdy = (in && dy) ? dy[i] : 0, vdx = (in && dx) ? dx[i] : 0;
  int vp = (in && pref) ? pref[i] : 0;
  U4 r, h;
  r.x = __ballot_sync(0xFFFFFFFFu, vuy != 0); r.y = __ballot_sync(0xFFFFFFFFu, vuy < 0);
  r.z = __ballot_sync(0xFFFFFFFFu, vux != 0); r.w = __ballot_sync(0xFFFFFFFFu, vux < 0);
  h.x = __ballot_sync(0xFFFFFFFFu, vdy != 0); h.y = __ballot_sync(0xFFFFFFFFu, vdy < 0);
  h.z = __ballot_sync(0xFFFFFFFFu, vdx != 0); h.w = __ballot_sync(0xFFFFFFFFu, vdx < 0);
  u32 p = __ballot_sync(0xFFFFFFFFu, vp != 0);
  if (lane == 0) {
    if (road) road[warp] = r;
    if (head) head[warp] = h;
    if (prefp) prefp[warp] = p;
  }
}

// unpacks the owned rows [own0, own1) into arrays of shape [own1-own0][gw]
__global__ void k_unpack(int own0, int own1, int gw, int pitch, const U4 *road, const U4 *head, const u32 *prefp,
                         int8_t *uy, int8_t *ux, int8_t *dy, int8_t *dx, uint8_t *pref) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(own1 - own0) * gw) return;
  int y = own0 + (int)(i / gw), x = (int)(i % gw);
  size_t o = (size_t)y * pitch + (x >> 5);
  u32 bit = 1u << (x & 31);
  U4 r = road[o], h = head[o];
  if (uy) uy[i] = (r.x & bit) ? ((r.y & bit) ? -1 : 1) : 0;
  if (ux) ux[i] = (r.z & bit) ? ((r.w & bit) ? -1 : 1) : 0;
  if (dy) dy[i] = (h.x & bit) ? ((h.y & bit) ? -1 : 1) : 0;
  if (dx) dx[i] = (h.z & bit) ? ((h.w & bit) ? -1 : 1) : 0;
  if (pref) pref[i] = (prefp[o] & bit) ? 1 : 0;
}

// create_grid (utils.fut:7-17): white, empty; rng_i = split_rng element i
__global__ void k_create(int gh, int gw, int pitch, int row0, int gh_global, u32 *color, u64 *rng, u64 state0, u64 inc) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)gh * gw) return;
  int y = (int)(i / gw), x = (int)(i % gw);
  long long yg = (long long)y + row0;
  if (yg < 0 || yg >= gh_global) return;
  size_t c = (size_t)y * pitch * 32 + x;
  color[c] = 0xFFFFFFFFu;
  rng[c] = pcg32_split_state(state0, inc, yg * gw + x);   // split_rng index = flat index in the WHOLE grid
}

// point edits (utils.fut:19-38)
__global__ void k_line_v(int pitch, U4 *road, int y0, int n, int x, int flow) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  size_t o = (size_t)(y0 + k) * pitch + (x >> 5);
  u32 bit = 1u << (x & 31);
  // distinct rows -> distinct words, but other columns of the word may be edited concurrently: atomics
  if (flow != 0) atomicOr(&road[o].x, bit); else atomicAnd(&road[o].x, ~bit);
  if (flow < 0) atomicOr(&road[o].y, bit); else atomicAnd(&road[o].y, ~bit);
}
__global__ void k_line_h(int pitch, U4 *road, int y, int x0, int x1, int flow) {
  int w = (x0 >> 5) + blockIdx.x * blockDim.x + threadIdx.x;
  if (w > (x1 >> 5)) return;
  int lo = max(x0, w * 32) - w * 32, hi = min(x1, w * 32 + 31) - w * 32;
  u32 m = (hi == 31 ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
  size_t o = (size_t)y * pitch + w;
  if (flow != 0) road[o].z |= m; else road[o].z &= ~m;
  if (flow < 0) road[o].w |= m; else road[o].w &= ~m;
}
__global__ void k_add_cell(int pitch, U4 *head, u32 *color, u32 *color_other, int y, int x, u32 col, int dy, int dx) {
  size_t o = (size_t)y * pitch + (x >> 5);
  u32 bit = 1u << (x & 31);
  U4 h = head[o];
  h.x = dy != 0 ? (h.x | bit) : (h.x & ~bit); h.y = dy < 0 ? (h.y | bit) : (h.y & ~bit);
  h.z = dx != 0 ? (h.z | bit) : (h.z & ~bit); h.w = dx < 0 ? (h.w | bit) : (h.w & ~bit);
  head[o] = h;
  color[(size_t)y * pitch * 32 + x] = col;
  color_other[(size_t)y * pitch * 32 + x] = col;   // keep the lazy double buffer's invariant
}

// ---------------------------------------------------------------------------------------------
// synthetic road network: monotone avenue lattice + ring tiles.  One warp per word.
// The same rule is restated independently in tests/synth_ref.py.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline u64 synth_hash(u64 seed, u64 tag, u64 a, u64 b) {
  u64 k = splitmix64(seed ^ (tag * 0xD6E8FEB86659FD93ULL));
  k = splitmix64(k ^ a);
  return splitmix64(k ^ b);
}
struct SynthCell {
  int uy, ux, dy, dx;
  u32 color;
  u64 rng;
};
__host__ __device__ inline SynthCell synth_cell(int gh, int gw, int y, int x, u64 seed, u32 density16, u32 ring_full16) {
  SynthCell c = {0, 0, 0, 0, 0xFFFFFFFFu, 0};
  u64 hc = synth_hash(seed, 2, (u64)y, (u64)x);
  bool car_draw = (u32)(hc & 0xFFFFu) < density16;
  bool arow = (y % 16) == 0, acol = (x % 16) == 8;
  if (arow) c.ux = 1;
  if (acol) c.uy = 1;
  if (arow || acol) {
    if (car_draw) {
      bool alongx = arow && (!acol || ((hc >> 16) & 1));
      if (alongx) c.dx = 1; else c.dy = 1;
    }
  } else if (y >= 2 && x >= 10) {
    int by = (y - 2) / 16, bx = (x - 10) / 16;
    int ly = (y - 2) - by * 16, lx = (x - 10) - bx * 16;
    bool fits = (16 * by + 14 <= gh - 1) && (16 * bx + 22 <= gw - 1);
    if (fits && ly <= 12 && lx <= 12 && (ly == 0 || ly == 12 || lx == 0 || lx == 12)) {
      u64 hb = synth_hash(seed, 1, (u64)by, (u64)bx);
      if (hb & 1) {
        bool cw = (hb >> 1) & 1;
        bool full = (u32)((hb >> 8) & 15u) < ring_full16;
        int s = cw ? 1 : -1;  // cw: top row +x, right col +y, bottom row -x, left col -y
        if (ly == 0) c.ux = s;
        if (ly == 12) c.ux = -s;
        if (lx == 12) c.uy = s;
        if (lx == 0) c.uy = -s;
        if (full || car_draw) {
          // heading = outgoing travel direction (as explorer/scenarios.fut:121-125 places them)
          int hy = 0, hx = 0;
          if (cw) {
            if (ly == 0 && lx < 12) hx = 1; else if (lx == 12 && ly < 12) hy = 1;
            else if (ly == 12 && lx > 0) hx = -1; else hy = -1;
          } else {
            if (ly == 0 && lx > 0) hx = -1; else if (lx == 0 && ly < 12) hy = 1;
            else if (ly == 12 && lx < 12) hx = 1; else hy = -1;
          }
          c.dy = hy; c.dx = hx;
        }
      }
    }
  }
  if (c.dy != 0 || c.dx != 0) c.color = 0xFF000000u | (u32)(synth_hash(seed, 3, (u64)y, (u64)x) & 0xFFFFFFu);
  c.rng = synth_hash(seed, 4, (u64)y, (u64)x);
  return c;
}
__global__ void k_synth(int gh, int gw, int pitch, int row0, int gh_global, u64 seed, u32 density16, u32 ring_full16,
                        U4 *road, U4 *head, u32 *prefp, u32 *color, u64 *rng) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= (long long)gh * pitch) return;
  int y = (int)(warp / pitch), w = (int)(warp % pitch), x = w * 32 + lane;
  int yg = y + row0;
  SynthCell c = {0, 0, 0, 0, 0, 0};
  if (x < gw && yg >= 0 && yg < gh_global) c = synth_cell(gh_global, gw, yg, x, seed, density16, ring_full16);
  U4 r, h;
  r.x = __ballot_sync(0xFFFFFFFFu, c.uy != 0); r.y = __ballot_sync(0xFFFFFFFFu, c.uy < 0);
  r.z = __ballot_sync(0xFFFFFFFFu, c.ux != 0); r.w = __ballot_sync(0xFFFFFFFFu, c.ux < 0);
  h.x = __ballot_sync(0xFFFFFFFFu, c.dy != 0); h.y = __ballot_sync(0xFFFFFFFFu, c.dy < 0);
  h.z = __ballot_sync(0xFFFFFFFFu, c.dx != 0); h.w = __ballot_sync(0xFFFFFFFFu, c.dx < 0);
  if (lane == 0) { road[warp] = r; head[warp] = h; prefp[warp] = 0; }
  size_t ci = (size_t)y * pitch * 32 + x;
  color[ci] = c.color;   // pad cells get 0
  rng[ci] = c.rng;
}
// analytic ring list of the generator: 48 cells per ring, stored as SYNTH_RING_WORDS word entries (unused ones empty)
constexpr int SYNTH_RING_WORDS = 26;
// rows [lo, hi) (global) are the band this device owns: a ring is listed by every device that owns at least one of its
// rows, with the words of the owned rows only (the other entries stay empty).  A ring cut by a band boundary gets the
// straddle id idx(by) * nbx + bx, idx = position of its block row in `cutrows` (the block rows some cut runs through,
// the same list on every device).
struct SynthCuts {
  int n;
  int by[BAND_MAXW];
};
__global__ void k_synth_rings(int gh, int gw, int pitch, int row0, int lo, int hi, u64 seed, int nby, int nbx,
                              RingWord *words, int *sid, unsigned int *nrings, const SynthCuts cutrows) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nby * nbx) return;
  int by = b / nbx, bx = b % nbx;
  if (!((16 * by + 14 <= gh - 1) && (16 * bx + 22 <= gw - 1))) return;
  u64 hb = synth_hash(seed, 1, (u64)by, (u64)bx);
  if (!(hb & 1)) return;
  const int r_lo = 16 * by + 2, r_hi = 16 * by + 14;
  if (r_hi < lo || r_lo >= hi) return;                       // another band's ring
  int d = -1;
  for (int i = 0; i < cutrows.n; i++)
    if (cutrows.by[i] == by) d = i * nbx + bx;
  bool cw = (hb >> 1) & 1;
  unsigned int slot = atomicAdd(nrings, 1u);
  if (sid) sid[slot] = d;
  RingWord *out = words + (size_t)slot * SYNTH_RING_WORDS;
  const int x0 = 16 * bx + 10, wa = x0 >> 5, wb = (x0 + 12) >> 5;
  for (int i = 0; i < SYNTH_RING_WORDS; i++) {
    // entries 2*ly and 2*ly+1: the part of ring row ly lying in word wa / in word wb (empty if wa == wb)
    int ly = i >> 1, w = (i & 1) ? wb : wa;
    RingWord e = {0, 0, 0, 0, 0, 0, 0, 0};
    int y = 16 * by + 2 + ly;
    const bool mine = y >= lo && y < hi;
    e.wi = mine ? (u32)((size_t)(y - row0) * pitch + w) : 0u;
    if (mine && !((i & 1) && wa == wb))
      for (int lx = 0; lx <= 12; lx++) {
        if (!(ly == 0 || ly == 12 || lx == 0 || lx == 12)) continue;
        int x = x0 + lx;
        if ((x >> 5) != w) continue;
        int hy = 0, hx = 0;
        if (cw) {
          if (ly == 0 && lx < 12) hx = 1; else if (lx == 12 && ly < 12) hy = 1;
          else if (ly == 12 && lx > 0) hx = -1; else hy = -1;
        } else {
          if (ly == 0 && lx > 0) hx = -1; else if (lx == 0 && ly < 12) hy = 1;
          else if (ly == 12 && lx < 12) hx = 1; else hy = -1;
        }
        u32 bit = 1u << (x & 31);
        if (hy < 0) e.mN |= bit; else if (hy > 0) e.mS |= bit; else if (hx < 0) e.mW |= bit; else e.mE |= bit;
        // grant (steppers.fut:206-211): a corner is entered along the axis it does NOT leave on
        bool corner = (ly == 0 || ly == 12) && (lx == 0 || lx == 12);
        bool grant_x = corner ? (hy != 0) : (hx != 0);
        if (grant_x) e.gX |= bit; else e.gY |= bit;
      }
    out[i] = e;
  }
}

// ---------------------------------------------------------------------------------------------
// render (flowamok.fut:16-64); matte gray/mix restated from memory (unpinned)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 d_from_rgba(float r, float g, float b, float a) {
  u32 R = (u32)__fmul_rn(fminf(1.0f, fmaxf(0.0f, r)), 255.0f), G = (u32)__fmul_rn(fminf(1.0f, fmaxf(0.0f, g)), 255.0f);
  u32 B = (u32)__fmul_rn(fminf(1.0f, fmaxf(0.0f, b)), 255.0f), A = (u32)__fmul_rn(fminf(1.0f, fmaxf(0.0f, a)), 255.0f);
  return (A << 24) | (R << 16) | (G << 8) | B;
}
__device__ __forceinline__ float d_mix_ch(float w1, float c1, float w2, float c2) {
  return sqrtf(__fadd_rn(__fmul_rn(__fmul_rn(w1, c1), c1), __fmul_rn(__fmul_rn(w2, c2), c2)));
}
__device__ __forceinline__ u32 d_mix(float m1, u32 c1, float m2, u32 c2) {
  float r1 = __fdiv_rn((float)((c1 >> 16) & 255u), 255.0f), g1 = __fdiv_rn((float)((c1 >> 8) & 255u), 255.0f);
  float b1 = __fdiv_rn((float)(c1 & 255u), 255.0f), a1 = __fdiv_rn((float)(c1 >> 24), 255.0f);
  float r2 = __fdiv_rn((float)((c2 >> 16) & 255u), 255.0f), g2 = __fdiv_rn((float)((c2 >> 8) & 255u), 255.0f);
  float b2 = __fdiv_rn((float)(c2 & 255u), 255.0f), a2 = __fdiv_rn((float)(c2 >> 24), 255.0f);
  float m12 = __fadd_rn(m1, m2), w1 = __fdiv_rn(m1, m12), w2 = __fdiv_rn(m2, m12);
  float a = __fdiv_rn(__fadd_rn(__fmul_rn(m1, a1), __fmul_rn(m2, a2)), m12);
  return d_from_rgba(d_mix_ch(w1, r1, w2, r2), d_mix_ch(w1, g1, w2, g2), d_mix_ch(w1, b1, w2, b2), a);
}
__global__ void k_render(int gh, int gw, int pitch, long long scale, const U4 *road, const U4 *head,
                         const u32 *color, u32 *pixels) {
  long long h = gh * scale, w = gw * scale;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  long long y = i / w, x = i % w;
  int yg = (int)(y / scale), xg = (int)(x / scale);
  size_t o = (size_t)yg * pitch + (xg >> 5);
  u32 bit = 1u << (xg & 31);
  U4 r = road[o], hd = head[o];
  int fy = (r.x & bit) ? ((r.y & bit) ? -1 : 1) : 0, fx = (r.z & bit) ? ((r.w & bit) ? -1 : 1) : 0;
  bool occupied = ((hd.x | hd.z) & bit) != 0;
  u32 px = 0xFF000000u;
  if (fy != 0 || fx != 0) {
    u32 base = occupied ? color[(size_t)yg * pitch * 32 + xg] : 0xFFFFFFFFu;
    long long yi = y % scale, xi = x % scale;
    long long d2 = scale / 2, m2 = scale % 2;
    long long yy = yi - d2, xx = xi - d2;
    if (m2 == 0) { yy += (yy >= 0); xx += (xx >= 0); }
    long long d = d2 - m2;
    long long ya = yy < 0 ? -yy : yy, xa = xx < 0 ? -xx : xx;
    bool y_draw = (fy * yy == d - xa) && xa <= d, x_draw = (fx * xx == d - ya) && ya <= d;
    bool draw = (fx == 0 && y_draw) || (fy == 0 && x_draw) || (fy != 0 && fx != 0 && (y_draw || x_draw));
    const u32 grey = d_from_rgba(0.3f, 0.3f, 0.3f, 1.0f);
    px = draw ? (occupied ? d_mix(0.5f, grey, 0.5f, base) : grey) : base;
  }
  pixels[i] = px;
}

// Position-dependent digest of the owned rows: out[0] += sum, out[1] ^= xor over the owned cells of
// H(global cell index, heading, preference, colour, rng state).  Sums of bands add up to the whole grid's, so a banded run
// is compared with the single-grid run -- or with numpy on a downloaded state -- without moving the planes to the host.
__device__ __host__ inline u64 digest_cell(u64 idx, u32 bits, u32 color, u64 rng) {
  return splitmix64(splitmix64(idx * 0x9E3779B97F4A7C15ULL + (((u64)bits << 32) | color)) ^ rng);
}
__global__ void k_digest(int own0, int own1, int gw, int pitch, int row0, const U4 *head, const u32 *prefp, const u32 *color,
                         const u64 *rng, unsigned long long *out) {
  const long long n = (long long)(own1 - own0) * gw;
  u64 sum = 0, x = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = own0 + (int)(i / gw), xx = (int)(i % gw);
    const size_t o = (size_t)y * pitch + (xx >> 5), c = (size_t)y * pitch * 32 + xx;
    const u32 bit = 1u << (xx & 31);
    const U4 h = head[o];
    const u32 bits = ((h.x & bit) ? 1u : 0u) | ((h.y & bit) ? 2u : 0u) | ((h.z & bit) ? 4u : 0u) | ((h.w & bit) ? 8u : 0u) |
                     ((prefp[o] & bit) ? 16u : 0u);
    const u64 v = digest_cell((u64)(y + row0) * (u64)gw + (u64)xx, bits, color[c], rng[c]);
    sum += v; x ^= v;
  }
  for (int d = 16; d > 0; d >>= 1) {
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    x ^= __shfl_xor_sync(0xFFFFFFFFu, x, d);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&out[0], sum); atomicXor(&out[1], x); }
}

// Sparse view of the per-cell RNG plane.  choose_direction is only ever called for a cell whose road has BOTH axes
// (steppers.fut:101-103 needs flow_y_ok && flow_x_ok), so only those "corner" cells' states can change: hosts that keep the grid
// in the reference's layout need not move 8 bytes per cell in and out, only (flat index, state) of the corners.
__global__ void k_corner_counts(int own0, int own1, int pitch, const U4 *road, unsigned int *counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(own1 - own0) * pitch) return;
  const U4 r = road[(size_t)own0 * pitch + i];
  counts[i] = __popc(r.x & r.z);   // pad cells have no road
}
__global__ void k_corner_emit(int own0, int own1, int gw, int pitch, int row0, const U4 *road, const unsigned int *offs, const u64 *rng,
                              long long cap, long long *idx, u64 *state) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(own1 - own0) * pitch) return;
  const int y = own0 + (int)(i / pitch), w = (int)(i % pitch);
  const U4 r = road[(size_t)y * pitch + w];
  u32 m = r.x & r.z;
  long long k = offs[i];
  while (m && k < cap) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    idx[k] = (long long)(y + row0) * gw + (w * 32 + b);
    state[k] = rng[((size_t)y * pitch + w) * 32 + b];
    k++;
  }
}
__global__ void k_rng_scatter(long long n, const long long *idx, const u64 *state, int gw, int pitch, int row0, int gh, u64 *rng) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long y = idx[i] / gw - row0, x = idx[i] % gw;
  if (y < 0 || y >= gh) return;   // a row this band does not hold
  rng[(size_t)y * pitch * 32 + x] = state[i];
}

// leak list compaction helpers (flowamok_step_leaks)
__global__ void k_leak_counts(long long nwords, const u32 *leak_plane, unsigned int *counts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nwords) counts[i] = __popc(leak_plane[i]);
}
// nf = 8 (flowamok_step_leaks) or 14 (flowamok_step_leaks_cells) int64 per leak; the record is the SOURCE cell as move_cell saw it
// (steppers.fut:122 copies `cell` of the old grid): old heading and colour, calculated / direction / preference as the fixpoint
// left them, and the cell's rng state BEFORE this tick -- if the cell itself drew in this tick (a car entered it while its own
// left: fork plane), the draw is undone (pcg32's state update is invertible; a redrawn rejection, probability 2^-31 per draw, is
// undone too).
__device__ __forceinline__ u64 pcg32_undo_draw(u64 s, u64 inc) {
  const u64 MINV = 13877824140714322085ULL;   // 6364136223846793005^-1 mod 2^64
  for (int k = 0; k < 4; k++) {
    s = (s - (inc | 1ULL)) * MINV;
    const u64 before = (s - (inc | 1ULL)) * MINV;   // was the draw before this one a rejected one of the same call?
    u64 t = before;
    if (pcg32_next(t, inc) < 0xFFFFFFFEu) break;
  }
  return s;
}
__global__ void k_leak_emit(int gh, int gw, int pitch, int row0, const u32 *leak_plane, const unsigned int *offs,
                            const U4 *head_old, const u32 *pref_new, const u32 *color_old, const U4 *road, const u32 *calc, const u64 *mymx,
                            const u64 *rng, u64 rng_inc, long long cap, int nf, long long *out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)gh * pitch) return;
  u32 m = leak_plane[i];
  if (!m) return;
  long long k = offs[i];
  int y = (int)(i / pitch), w = (int)(i % pitch);
  U4 h = head_old[i];
  u32 pr = pref_new[i];
  const u32 drew = leak_plane[i + (size_t)gh * pitch];
  while (m && k < cap) {
    int b = __ffs(m) - 1;
    m &= m - 1;
    u32 bit = 1u << b;
    int x = w * 32 + b;
    int dy = (h.x & bit) ? ((h.y & bit) ? -1 : 1) : 0, dx = (h.z & bit) ? ((h.w & bit) ? -1 : 1) : 0;
    long long *o = out + k * nf;
    const size_t c = (size_t)y * pitch * 32 + x;
    if (nf == 8) {
      o[0] = y + row0 + dy; o[1] = x + dx; o[2] = y + row0; o[3] = x; o[4] = dy; o[5] = dx;
      o[6] = color_old[c]; o[7] = (pr & bit) ? 1 : 0;
    } else {
      const U4 r = road[i];
      const u64 mm = mymx[i];
      u64 st = rng[c];
      if (drew & bit) st = pcg32_undo_draw(st, rng_inc);
      o[0] = y + row0 + dy; o[1] = x + dx; o[2] = y + row0; o[3] = x;
      o[4] = (r.x & bit) ? ((r.y & bit) ? -1 : 1) : 0; o[5] = (r.z & bit) ? ((r.w & bit) ? -1 : 1) : 0;
      o[6] = dy; o[7] = dx; o[8] = color_old[c];
      o[9] = (calc[i] & bit) ? 1 : 0; o[10] = ((u32)mm & bit) ? 1 : 0; o[11] = ((u32)(mm >> 32) & bit) ? 1 : 0;
      o[12] = (pr & bit) ? 1 : 0; o[13] = (long long)st;
    }
    k++;
  }
}

}  // namespace fa
