// flowamok_b200 -- row bands over NVLink without NCCL and without the host: peer-mapped mailboxes, remote stores, and
// sequence flags polled on the device ("direct P2P stores into the peer's halo buffer + a flag").
//
// Every band owns ONE mailbox allocation, mapped by the other bands (CUDA IPC across processes, plain pointers inside one
// process).  A band never reads remote memory: a sender stores the payload into the RECEIVER's mailbox, fences at
// system scope, and then stores the message number into the receiver's header (st.release.sys); the receiver polls its
// own header (ld.acquire.sys, local memory) and reads the payload with L1-bypassing loads.
//
//   header     s_seq[side], u_seq[side]  : number of the last state / u message that arrived from the neighbour on `side`
//              vote[parity][rank]        : termination votes of the fixpoint rounds, (epoch << 1) | woke-something
//              ring_seq[rank]            : number of the last ring-verdict row that arrived from `rank`
//   payload    state inbox [side][parity] : 2 heading rows + 2 preference rows + 1 colour row  (k_band_xstate)
//              u inbox     [side][parity] : 1 row of calc + 1 row of my|mx                    (band_exchange_u)
//              ring rows   [parity][rank] : partial pass bits of the rings that straddle a cut  (k_band_xring)
//
// Two slots per channel suffice: every exchange step is "send message k, then wait for the neighbour's message k", so a
// band is never more than one message ahead of the band it talks to (when it writes k + 1 it has consumed the
// neighbour's k, which the neighbour sent after it had consumed k - 1).
//
// Every wait has a wall-clock watchdog (globaltimer): a protocol error becomes an error flag the host reads at its next
// synchronisation point, never a hung GPU.
#pragma once
#include <cuda_runtime.h>

#include "fa_tiles.h"

namespace fa {

constexpr int BAND_MAXW = 16;   // bands per grid the header has room for
#ifndef FA_BAND_TIMEOUT_NS
#define FA_BAND_TIMEOUT_NS 20000000000ULL   // 20 s
#endif

struct alignas(256) MailHdr {
  unsigned int s_seq[2];
  unsigned int u_seq[2];
  unsigned int vote[2][BAND_MAXW];
  unsigned int ring_seq[BAND_MAXW];
};
static_assert(sizeof(MailHdr) == 256, "mailbox header layout");

// device-local protocol state of one band (never written by a peer)
struct BandDev {
  unsigned int u_seq, epoch, ring_seq, pad0;    // numbers of the last messages SENT (state messages: the host counts)
  unsigned int xstate_arrived;                  // CTA arrival counter of k_band_xstate (monotone)
  unsigned int error;                           // 2 | tag << 8: a wait timed out (tag: 1 state, 2 u rows, 3 vote, 4 ring verdicts)
  unsigned int err_want, err_got;               // the message number that never came / the last one seen
  unsigned int go, round_it;                    // CTA 0 -> grid hand-over inside k_u_iter
  unsigned long long rounds;                    // convergence rounds so far
  unsigned long long waits_ns;                  // time CTA 0 spent waiting for neighbours (diagnostic)
};

struct BandLink {
  int rank, world;
  char *self;                  // my mailbox
  char *nb[2];                 // mailbox of the neighbour on side 0 (smaller rows) / 1 (larger rows); null at the grid's edge
  char *all[BAND_MAXW];        // every band's mailbox by rank (all[rank] == self)
  BandDev *dev;
  unsigned long long s_bytes, u_bytes;            // message sizes
  unsigned long long off_s, off_u, off_ring;      // payload offsets inside a mailbox
  int ring_words;                                 // 32-bit words of straddling-ring pass bits (0: none)
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long band_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until the message number at *p has reached `want` (numbers wrap; compare as a signed difference); shift: the
// vote words carry the epoch above bit 0.  Returns the word read.  On a timeout the band's error flag is set and every
// later wait returns at once, so the queued ticks drain quickly and the host reports the failure.
__device__ __forceinline__ unsigned int band_wait(const unsigned int *p, unsigned int want, int shift, BandDev *dev, unsigned int tag) {
  unsigned long long t0 = 0;
  for (unsigned int spins = 0;; spins++) {
    const unsigned int v = ld_acquire_sys(p);
    if ((int)((v >> shift) - want) >= 0) {
      if (t0) atomicAdd(&dev->waits_ns, band_now() - t0);
      return v;
    }
    if (*(volatile unsigned int *)&dev->error) return v;
    if (spins == 32) t0 = band_now();
    if (spins > 32) {
      __nanosleep(100);
      if ((spins & 1023u) == 0 && band_now() - t0 > FA_BAND_TIMEOUT_NS) {
        if (atomicCAS(&dev->error, 0u, 2u | (tag << 8)) == 0u) { dev->err_want = want; dev->err_got = v >> shift; }
        return v;
      }
    }
  }
}

__device__ __forceinline__ U4 ldcg_u4(const U4 *p) {   // 128-bit load that bypasses L1 (the peer wrote the data)
  const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
  U4 r = {v.x, v.y, v.z, v.w};
  return r;
}
__device__ __forceinline__ char *band_inbox_u(char *box, const BandLink &L, int side, unsigned int seq) {
  return box + L.off_u + ((size_t)side * 2 + (seq & 1u)) * L.u_bytes;
}
__device__ __forceinline__ char *band_inbox_s(char *box, const BandLink &L, int side, unsigned int seq) {
  return box + L.off_s + ((size_t)side * 2 + (seq & 1u)) * L.s_bytes;
}

// ---------------------------------------------------------------------------------------------
// u exchange: push this band's boundary rows of calc / my|mx into the neighbours' inboxes, wait for theirs, OR them into the
// ghost rows.  In three steps, so that a whole grid can share the work (k_u_iter: a slice per thread, grid barriers between
// the steps) as well as a single CTA (k_band_xu):
//   band_push_u   thread gtid of gthreads copies its words of both boundary rows (L1-bypassing loads: other CTAs wrote them);
//                 the caller fences at system scope and synchronises all pushing threads before
//   band_flags_u  (threads 0 and 1 of ONE CTA) publishes message number `seq` to the neighbours and waits for theirs;
//   band_merge_u  ORs the received rows into the ghost rows.  wake: the owned words that hold a waiting car which reads a changed
//                 ghost cell go on `list` (length *cnt) -- the same rule as k_band_merge.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void band_push_u(const GridView &g, const BandLink &L, unsigned int seq, int gtid, int gthreads) {
  const int P = g.pitch;
  for (int side = 0; side < 2; side++) {
    if (!L.nb[side]) continue;
    char *slot = band_inbox_u(L.nb[side], L, 1 - side, seq);   // my side-0 neighbour files me under ITS side 1
    u32 *dc = reinterpret_cast<u32 *>(slot);
    u64 *dm = reinterpret_cast<u64 *>(slot + (size_t)P * 4);
    const size_t o = (size_t)(side == 0 ? g.own0 : g.own1 - 1) * P;
    for (int w = gtid; w < P; w += gthreads) {
      dc[w] = __ldcg(&g.calc[o + w]);
      dm[w] = __ldcg(&g.mymx[o + w]) | (g.mymx_B ? __ldcg(&g.mymx_B[o + w]) : 0ull);
    }
  }
}
__device__ __forceinline__ void band_flags_u(const BandLink &L, unsigned int seq) {
  const int tid = threadIdx.x;
  if (tid < 2 && L.nb[tid]) {
    st_release_sys(&reinterpret_cast<MailHdr *>(L.nb[tid])->u_seq[1 - tid], seq);
    band_wait(&reinterpret_cast<MailHdr *>(L.self)->u_seq[tid], seq, 0, L.dev, 2u);
  }
}
__device__ __forceinline__ void band_merge_u(const GridView &g, const BandLink &L, const UEntry *table, unsigned int seq, bool wake,
                                             unsigned int *list, unsigned int *cnt, int gtid, int gthreads) {
  const int P = g.pitch;
  for (int side = 0; side < 2; side++) {
    if (!L.nb[side]) continue;
    const char *slot = band_inbox_u(L.self, L, side, seq);
    const u32 *rc = reinterpret_cast<const u32 *>(slot);
    const u64 *rm = reinterpret_cast<const u64 *>(slot + (size_t)P * 4);
    const int ghost_row = side == 0 ? g.own0 - 1 : g.own1, own_row = side == 0 ? g.own0 : g.own1 - 1;
    for (int w = gtid; w < P; w += gthreads) {
      const size_t o = (size_t)ghost_row * P + w;
      const u32 oc = __ldcg(&g.calc[o]), nc = oc | __ldcg(&rc[w]);
      const u64 om = __ldcg(&g.mymx[o]) | (g.mymx_B ? __ldcg(&g.mymx_B[o]) : 0ull), nm = om | __ldcg(&rm[w]);
      if (nc == oc && nm == om) continue;
      u_grant(g, o, om, nm);   // into the plane, or (speculative move under way) into the delta plane + fix-up list
      g.calc[o] = nc;
      if (!wake) continue;
      const u32 chg = (oc ^ nc) | (u32)(om ^ nm) | (u32)((om ^ nm) >> 32);
      for (int dx = -1; dx <= 1; dx++) {
        if (!band_word_reads_ghost(g, table, side, own_row, w, dx, chg)) continue;
        const u32 q = u_try_queue(g, own_row, w + dx);
        if (q != U_NONE) list[atomicAdd(cnt, 1u)] = q & ~U_QUEUED;
      }
    }
  }
}
// the three steps by ALL threads of ONE CTA
__device__ __forceinline__ void band_exchange_u(const GridView &g, const BandLink &L, const UEntry *table, bool wake, unsigned int *list,
                                                unsigned int *cnt) {
  __shared__ unsigned int seq_sh;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) seq_sh = ++L.dev->u_seq;
  __syncthreads();
  const unsigned int seq = seq_sh;
  band_push_u(g, L, seq, tid, nt);
  __threadfence_system();
  __syncthreads();
  band_flags_u(L, seq);
  __syncthreads();
  band_merge_u(g, L, table, seq, wake, list, cnt, tid, nt);
  __threadfence();
  __syncthreads();
}

// Termination vote of one convergence round, by all threads of ONE CTA: every band tells every band whether its merge
// woke something; returns the OR.  An all-to-all of world 4-byte stores over NVSwitch instead of an all-reduce + host read.
__device__ __forceinline__ bool band_vote(const BandLink &L, bool any) {
  __shared__ unsigned int ep_sh, any_sh;
  const int tid = threadIdx.x;
  if (tid == 0) { ep_sh = ++L.dev->epoch; any_sh = 0u; }
  __syncthreads();
  const unsigned int ep = ep_sh;
  if (tid < L.world) {
    st_release_sys(&reinterpret_cast<MailHdr *>(L.all[tid])->vote[ep & 1u][L.rank], (ep << 1) | (any ? 1u : 0u));
    const unsigned int v = band_wait(&reinterpret_cast<MailHdr *>(L.self)->vote[ep & 1u][tid], ep, 1, L.dev, 3u);
    if (v & 1u) atomicOr(&any_sh, 1u);
  }
  __syncthreads();
  return any_sh != 0u;
}

// ---------------------------------------------------------------------------------------------
// k_band_xstate: the per-tick state exchange (what k_u_local / k_move read across the cut): 2 heading rows + 2
// preference rows + 1 colour row per side, pushed straight into the neighbours' inboxes; then the neighbours' rows are
// copied from this band's inbox into its halo rows.  At most one CTA per SM, so every CTA is resident and the push of
// this band never waits for anything (the pull does: for the neighbours' pushes, which do not depend on this band).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_band_xstate(int pitch, int own0, int own1, U4 *head, u32 *pref, u32 *color, const __grid_constant__ BandLink L, unsigned int seq) {
  // seq: number of this state message = ticks stepped so far + 1 (counted by the host; the grid size never changes)
  const int tid = threadIdx.x;
  const size_t P = (size_t)pitch, n = 8 * P, stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + tid;
  for (int side = 0; side < 2; side++) {
    if (!L.nb[side]) continue;
    char *b = band_inbox_s(L.nb[side], L, 1 - side, seq);
    U4 *mh = reinterpret_cast<U4 *>(b);
    u32 *mp = reinterpret_cast<u32 *>(b + 2 * P * 16);
    U4 *mc = reinterpret_cast<U4 *>(b + 2 * P * 20);
    const int r2 = side == 0 ? own0 : own1 - 2, rc = side == 0 ? own0 : own1 - 1;
    const U4 *gh = head + (size_t)r2 * P;
    const u32 *gp = pref + (size_t)r2 * P;
    const U4 *gc = reinterpret_cast<const U4 *>(color + (size_t)rc * P * 32);
    for (size_t i = i0; i < n; i += stride) {
      mc[i] = gc[i];
      if (i < 2 * P) { mh[i] = gh[i]; mp[i] = gp[i]; }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned int arrived = atomicAdd(&L.dev->xstate_arrived, 1u) + 1u;
    if (arrived == seq * gridDim.x) {   // the last CTA of this launch: all pushes are fenced
      __threadfence_system();
      for (int side = 0; side < 2; side++)
        if (L.nb[side]) st_release_sys(&reinterpret_cast<MailHdr *>(L.nb[side])->s_seq[1 - side], seq);
    }
    for (int side = 0; side < 2; side++)
      if (L.nb[side]) band_wait(&reinterpret_cast<MailHdr *>(L.self)->s_seq[side], seq, 0, L.dev, 1u);
  }
  __syncthreads();
  for (int side = 0; side < 2; side++) {
    if (!L.nb[side]) continue;
    const char *b = band_inbox_s(L.self, L, side, seq);
    const U4 *mh = reinterpret_cast<const U4 *>(b);
    const u32 *mp = reinterpret_cast<const u32 *>(b + 2 * P * 16);
    const U4 *mc = reinterpret_cast<const U4 *>(b + 2 * P * 20);
    const int r2 = side == 0 ? own0 - 2 : own1, rc = side == 0 ? own0 - 1 : own1;
    U4 *gh = head + (size_t)r2 * P;
    u32 *gp = pref + (size_t)r2 * P;
    U4 *gc = reinterpret_cast<U4 *>(color + (size_t)rc * P * 32);
    for (size_t i = i0; i < n; i += stride) {
      gc[i] = ldcg_u4(&mc[i]);
      if (i < 2 * P) { gh[i] = ldcg_u4(&mh[i]); gp[i] = __ldcg(&mp[i]); }
    }
  }
}
// k_band_xu: one more u exchange without waking anybody (after ring resolution: ring marks on a boundary row must reach
// the neighbour's ghost row before k_move).  One CTA.
__global__ void __launch_bounds__(1024)
k_band_xu(const __grid_constant__ GridView g, const __grid_constant__ BandLink L) {
  band_exchange_u(g, L, nullptr, false, nullptr, nullptr);
}

// k_band_xring: all-to-all of the partial pass bits of the rings that straddle a cut (steppers.fut:197-200 needs ALL cells
// of a ring): `partial` (this band's verdict on its part of every straddling ring, 1 = no objection) goes to every band;
// pass = AND over all bands.  One CTA.  partial is reset to all-ones for the next tick.
__global__ void __launch_bounds__(512)
k_band_xring(u32 *partial, u32 *pass, const __grid_constant__ BandLink L) {
  __shared__ unsigned int seq_sh;
  const int tid = threadIdx.x, nt = blockDim.x, nw = L.ring_words;
  if (tid == 0) seq_sh = ++L.dev->ring_seq;
  __syncthreads();
  const unsigned int seq = seq_sh;
  for (int r = 0; r < L.world; r++) {
    u32 *row = reinterpret_cast<u32 *>(L.all[r] + L.off_ring) + ((size_t)(seq & 1u) * L.world + L.rank) * nw;
    for (int w = tid; w < nw; w += nt) row[w] = __ldcg(&partial[w]);
  }
  __threadfence_system();
  __syncthreads();
  if (tid < L.world) {
    st_release_sys(&reinterpret_cast<MailHdr *>(L.all[tid])->ring_seq[L.rank], seq);
    band_wait(&reinterpret_cast<MailHdr *>(L.self)->ring_seq[tid], seq, 0, L.dev, 4u);
  }
  __syncthreads();
  const u32 *rows = reinterpret_cast<const u32 *>(L.self + L.off_ring) + (size_t)(seq & 1u) * L.world * nw;
  for (int w = tid; w < nw; w += nt) {
    u32 v = 0xFFFFFFFFu;
    for (int r = 0; r < L.world; r++) v &= __ldcg(&rows[(size_t)r * nw + w]);
    pass[w] = v;
    partial[w] = 0xFFFFFFFFu;
  }
}

}  // namespace fa
