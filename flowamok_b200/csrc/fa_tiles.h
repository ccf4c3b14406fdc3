// flowamok_b200 -- tile programs of the per-tick update, written once for device and host.
//
// Each program is a sequence of phases; inside a phase every "thread" tid in [0, nt) works on a
// strided set of words and phases are separated by a block barrier.  The CUDA kernels
// (fa_kernels.cu) call phase(threadIdx.x, blockDim.x) with __syncthreads() between phases; the CPU
// emulation used by tests/emul runs the same phases with a loop over tid.  So the bit logic, the
// indexing, the halo/outer-ring bookkeeping and the pending/changed protocol are all unit-tested
// on the CPU against the oracle before they ever reach a GPU.
//
// Geometry of the U (fixpoint) tile, in words (32 cells) and rows:
//
//        c = 0        1 .. TW          TW+1
//      +--------+-------------------+--------+   r = 0            outer ring row (not computed)
//      | halo   |                   | halo   |   r = 1 .. H       computed halo rows
//      | word   |   tile interior   | word   |   r = H+1 .. H+TR  tile rows (owned, written back)
//      | 31 bits|                   | 31 bits|   r = H+TR+1 ..    computed halo rows
//      +--------+-------------------+--------+   r = SR-1         outer ring row
//   The outermost bit of each halo word and the two outer rows form the "outer ring": their
//   CALC/MY/MX come from the global planes (previous pass; zero in pass 1) and are never updated.
//   Everything else is recomputed locally to its least fixpoint (steppers.fut:147-164 is confluent,
//   SURVEY 8a-U), so a chain of waiting cars is resolved inside one tile visit unless it runs more
//   than H rows / 31 columns past the tile, in which case the tile is queued for another pass.
#pragma once
#include "fa_bits.h"

namespace fa {

struct alignas(16) U4 {
  u32 x, y, z, w;
};

// chooser modes of the C ABI (flowamok.h); choose_direction is a parameter of the reference steppers
enum { CHOOSE_RANDOM = 0, CHOOSE_Y = 1, CHOOSE_X = 2, CHOOSE_TAPE = 3 };

// Device-resident grid in bit-plane form.
struct GridView {
  int gh, gw;        // cells
  int pitch;         // words per plane row (multiple of 4 -> 16-byte aligned rows)
  int wvalid;        // ceil(gw / 32): words that hold cells
  const U4 *road;    // {UYN, UYS, UXN, UXS} per word                       [gh][pitch]
  const U4 *head_in; // {DYN, DYS, DXN, DXS} per word, state before the tick [gh][pitch]
  U4 *head_out;      // state after the tick
  const u32 *pref_in;
  u32 *pref_out;
  u32 *calc, *my, *mx;     // per-tick planes (steppers.fut:140-145 resets them: here they are scratch)
  const u32 *color_in;     // [gh][pitch*32] argb payload
  u32 *color_out;
  u64 *rng;                // [gh][pitch*32] pcg32 state of each cell (inc is shared, utils.fut:15)
  u64 rng_inc;
  int chooser;             // CHOOSE_*
  u32 tape_bit;            // CHOOSE_TAPE: this tick's choice
};

struct TickCounters {
  unsigned long long leaks;      // cell leaks of this tick (explorer.fut:125 only needs the count)
  unsigned int n_pending;        // tiles queued for the next fixpoint pass
  unsigned int changed;          // a queued tile made progress in this pass
  unsigned int u_iters;          // total local relaxation sweeps (diagnostics)
  unsigned int passes;           // fixpoint passes executed this tick (diagnostics)
};

#if defined(__CUDA_ARCH__)
#define FA_ATOMIC_ADD_U32(p, v) atomicAdd((p), (v))
#define FA_ATOMIC_ADD_U64(p, v) atomicAdd((p), (v))
#define FA_ATOMIC_OR_U32(p, v) atomicOr((p), (v))
#else
template <class T> inline T fa_host_add(T *p, T v) { T o = *p; *p = o + v; return o; }
#define FA_ATOMIC_ADD_U32(p, v) fa_host_add<unsigned int>((p), (v))
#define FA_ATOMIC_ADD_U64(p, v) fa_host_add<unsigned long long>((p), (v))
#define FA_ATOMIC_OR_U32(p, v) (*(p) |= (v))
#endif

FA_HD int popc32(u32 v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
FA_HD int ffs32(u32 v) {  // index of lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return __ffs(v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

// =============================================================================================
// U tile: local least fixpoint of update_can_be_moved_to_from (steppers.fut:9-56, :147-164)
//
// Only words that hold a waiting car ("link" cells: occupied, successor in bounds and on a road) can
// ever change after initialisation, so they are compacted into a worklist in shared memory and the
// relaxation sweeps touch nothing else.  A sweep is compute -> barrier -> commit -> barrier, which
// makes every committed word a consistent (R, MY, MX) triple: a cell with R = 1 has final MY/MX.
// =============================================================================================
FA_HD void fa_fence_block() {
#if defined(__CUDA_ARCH__)
  __threadfence_block();
#endif
}
FA_HD void fa_fence_device() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#endif
}
// append v to list if pred (warp-aggregated on the device)
FA_HD void wl_push(unsigned short *list, unsigned int *cnt, bool pred, unsigned short v) {
#if defined(__CUDA_ARCH__)
  unsigned act = __activemask();
  unsigned m = __ballot_sync(act, pred);
  if (m) {
    int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(cnt, (unsigned)__popc(m));
    base = __shfl_sync(act, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = v;
  }
#else
  if (pred) list[(*cnt)++] = v;
#endif
}

template <int TR_, int TW_, int H_>
struct UTile {
  static constexpr int TR = TR_, TW = TW_, H = H_;
  static constexpr int SR = TR + 2 * H + 2;
  static constexpr int SW = TW + 2;
  static constexpr int NW = SR * SW;
  struct Smem {
    u32 DYN[NW], DYS[NW], DXN[NW], DXS[NW];
    u32 T[NW], GY[NW], GX[NW];
    u32 R[NW], MY[NW], MX[NW];   // R doubles as ROAD staging during p0/p1; MY as the taint plane in p5
    unsigned short wl[2][NW];    // worklists of active words (ping-pong)
    unsigned int wn[2];
  };
  struct Pend {  // a thread's computed-but-uncommitted word of the current sweep
    int i;
    u32 R, MY, MX;
    bool changed;
  };

  const GridView &g;
  Smem &s;
  int y0, w0;      // first tile row / word
  bool first_pass; // outer ring unknown (all zero) instead of read from the global planes

  FA_HD UTile(const GridView &g_, Smem &s_, int tile_y, int tile_w, bool first)
      : g(g_), s(s_), y0(tile_y * TR), w0(tile_w * TW), first_pass(first) {}

  FA_HD static int idx(int r, int c) { return r * SW + c; }
  FA_HD int grow(int r) const { return y0 - H - 1 + r; }
  FA_HD int gword(int c) const { return w0 - 1 + c; }
  FA_HD bool in_grid(int r, int c) const {
    int y = grow(r), w = gword(c);
    return y >= 0 && y < g.gh && w >= 0 && w < g.pitch;
  }
  // bits of word (r,c) that this tile recomputes
  FA_HD static u32 cmask(int r, int c) {
    if (r == 0 || r == SR - 1) return 0u;
    if (c == 0) return 0xFFFFFFFEu;
    if (c == SW - 1) return 0x7FFFFFFFu;
    return 0xFFFFFFFFu;
  }
  FA_HD W3 w3(const u32 *p, int r, int c) const {
    W3 v;
    v.c = p[idx(r, c)];
    v.l = c > 0 ? p[idx(r, c - 1)] : 0u;
    v.r = c < SW - 1 ? p[idx(r, c + 1)] : 0u;
    return v;
  }

  // p0: stage car planes and the road-presence plane of the whole smem region
  FA_HD void p0(int tid, int nt) {
    if (tid == 0) { s.wn[0] = 0; s.wn[1] = 0; }
    for (int i = tid; i < NW; i += nt) {
      int r = i / SW, c = i - r * SW;
      U4 hd = {0, 0, 0, 0};
      u32 road = 0;
      if (in_grid(r, c)) {
        size_t o = (size_t)grow(r) * g.pitch + gword(c);
        hd = g.head_in[o];
        U4 rd = g.road[o];
        road = rd.x | rd.z;
      }
      s.DYN[i] = hd.x; s.DYS[i] = hd.y; s.DXN[i] = hd.z; s.DXS[i] = hd.w;
      s.R[i] = road;
    }
  }
  // p1: statics of every computed word
  FA_HD void p1(int tid, int nt) {
    for (int i = tid; i < NW; i += nt) {
      int r = i / SW, c = i - r * SW;
      u32 T = 0, GY = 0, GX = 0;
      if (cmask(r, c) != 0u && in_grid(r, c)) {
        int y = grow(r), w = gword(c);
        size_t o = (size_t)y * g.pitch + w;
        U4 rd = g.road[o];
        u32 pref = g.pref_in[o];
        Edge e = edge_masks(y, w, g.gh, g.gw);
        UStatic st = u_statics(rd.x, rd.y, rd.z, rd.w, s.DYN[i], s.DYS[i], s.DXN[i], s.DXS[i], pref,
                               w3(s.R, r - 1, c), w3(s.R, r, c), w3(s.R, r + 1, c), s.DYN[idx(r - 1, c)],
                               s.DYN[idx(r + 1, c)], w3(s.DXN, r, c), e);
        T = st.T; GY = st.GYr; GX = st.GXr;
      }
      s.T[i] = T; s.GY[i] = GY; s.GX[i] = GX;
    }
  }
  // p2: initial state (computed bits start from the terminals, outer bits come from the global planes)
  //     and the first worklist: every word holding a link cell
  FA_HD void p2(int tid, int nt) {
    for (int i0 = 0; i0 < NW; i0 += nt) {
      int i = i0 + tid;
      bool act = false;
      if (i < NW) {
        int r = i / SW, c = i - r * SW;
        u32 cm = cmask(r, c);
        u32 oR = 0, oMY = 0, oMX = 0;
        if (!first_pass && cm != 0xFFFFFFFFu && in_grid(r, c)) {
          size_t o = (size_t)grow(r) * g.pitch + gword(c);
          oR = g.calc[o];       // calc BEFORE my/mx: the writer publishes my/mx first (p4)
          fa_fence_device();
          oMY = g.my[o]; oMX = g.mx[o];
        }
        u32 T = s.T[i];
        s.R[i] = (cm & T) | (~cm & oR);
        s.MY[i] = (cm & T & s.GY[i]) | (~cm & oMY);
        s.MX[i] = (cm & T & s.GX[i]) | (~cm & oMX);
        act = (~T & cm) != 0u;
      }
      wl_push(s.wl[0], &s.wn[0], act, (unsigned short)i);
    }
  }
  FA_HD UStatic links(int i) const {
    UStatic st;
    Heading h = heading(s.DYN[i], s.DYS[i], s.DXN[i], s.DXS[i]);
    u32 link = ~s.T[i];
    st.T = s.T[i];
    st.LN = link & h.cN; st.LS = link & h.cS; st.LW = link & h.cW; st.LE = link & h.cE;
    st.LNW = link & h.dNW; st.LNE = link & h.dNE; st.LSW = link & h.dSW; st.LSE = link & h.dSE;
    st.GYr = s.GY[i]; st.GXr = s.GX[i];
    return st;
  }
  // p3a: compute the new state of worklist entry (base + tid) of list `cur` (no smem writes)
  FA_HD void p3a(int tid, int cur, int base, Pend &pd) const {
    pd.i = -1; pd.changed = false;
    int k = base + tid;
    if (k >= (int)s.wn[cur]) return;
    int i = s.wl[cur][k];
    int r = i / SW, c = i - r * SW;
    u32 cm = cmask(r, c);
    UStatic st = links(i);
    UState n = u_relax(st, w3(s.R, r - 1, c), w3(s.R, r, c), w3(s.R, r + 1, c), w3(s.MY, r - 1, c),
                       w3(s.MY, r + 1, c), w3(s.MX, r, c), w3(s.MX, r - 1, c), w3(s.MX, r + 1, c));
    pd.i = i;
    pd.R = (cm & n.R) | (~cm & s.R[i]);
    pd.MY = (cm & n.MY) | (~cm & s.MY[i]);
    pd.MX = (cm & n.MX) | (~cm & s.MX[i]);
    pd.changed = pd.R != s.R[i] || pd.MY != s.MY[i] || pd.MX != s.MX[i];
  }
  // p3b: commit, and keep the word on the next list while it still holds an uncalculated link cell
  FA_HD bool p3b(int cur, const Pend &pd) {
    bool keep = false;
    if (pd.i >= 0) {
      if (pd.changed) { s.R[pd.i] = pd.R; s.MY[pd.i] = pd.MY; s.MX[pd.i] = pd.MX; }
      int r = pd.i / SW, c = pd.i - r * SW;
      keep = (~pd.R & ~s.T[pd.i] & cmask(r, c)) != 0u;
    }
    wl_push(s.wl[cur ^ 1], &s.wn[cur ^ 1], keep, (unsigned short)(pd.i >= 0 ? pd.i : 0));
    return pd.changed;
  }
  // p4: write the tile's own words; returns bit0 = some owned cell is still uncalculated,
  //     bit1 = the written planes differ from what the previous pass left in global memory
  FA_HD int p4(int tid, int nt) {
    int res = 0;
    for (int i = tid; i < TR * TW; i += nt) {
      int tr = i / TW, tc = i - tr * TW;
      int r = tr + H + 1, c = tc + 1;
      int y = grow(r), w = gword(c);
      if (y >= g.gh || w >= g.pitch) continue;
      int si = idx(r, c);
      size_t o = (size_t)y * g.pitch + w;
      Edge e = edge_masks(y, w, g.gh, g.gw);
      u32 R = s.R[si] & e.valid, MY = s.MY[si] & e.valid, MX = s.MX[si] & e.valid;
      if (!first_pass && (g.calc[o] != R || g.my[o] != MY || g.mx[o] != MX)) res |= 2;
      g.my[o] = MY; g.mx[o] = MX;
      if (!first_pass) fa_fence_device();   // concurrent readers (p2 of a neighbour) must never see calc without its my/mx
      g.calc[o] = R;
      U4 rd = g.road[o];
      UState f = {R, MY, MX};
      UStatic dummy = {};
      g.pref_out[o] = u_pref_next(rd.x | rd.z, g.pref_in[o], f, dummy) & e.valid;
      if ((~R & e.valid) != 0u) res |= 1;
    }
    return res;
  }
  // p5: does an uncalculated owned cell wait (transitively) on an uncalculated OUTER cell?  Only words
  // still on the worklist (list `cur`) can be tainted.  The taint plane reuses MY (dead after p4).
  FA_HD void p5a(int tid, int nt) {
    for (int i = tid; i < NW; i += nt) {
      int r = i / SW, c = i - r * SW;
      s.MY[i] = ~cmask(r, c) & ~s.R[i];
    }
  }
  FA_HD u32 p5b_compute(int tid, int cur, int base, int &iw) const {
    iw = -1;
    int k = base + tid;
    if (k >= (int)s.wn[cur]) return 0u;
    int i = s.wl[cur][k];
    int r = i / SW, c = i - r * SW;
    UStatic st = links(i);
    W3 Xu = w3(s.MY, r - 1, c), Xc = w3(s.MY, r, c), Xd = w3(s.MY, r + 1, c);
    u32 x = (st.LN & Xu.c) | (st.LS & Xd.c) | (st.LW & west(Xc)) | (st.LE & east(Xc)) |
            (st.LNW & west(Xu)) | (st.LNE & east(Xu)) | (st.LSW & west(Xd)) | (st.LSE & east(Xd));
    iw = i;
    return s.MY[i] | (cmask(r, c) & x);
  }
  FA_HD bool p5b_commit(int iw, u32 x) {
    if (iw < 0 || x == s.MY[iw]) return false;
    s.MY[iw] = x;
    return true;
  }
  FA_HD bool p5c(int tid, int nt, int cur) const {
    bool pending = false;
    for (int k = tid; k < (int)s.wn[cur]; k += nt) {
      int i = s.wl[cur][k];
      int r = i / SW, c = i - r * SW;
      if (r < H + 1 || r >= H + 1 + TR || c < 1 || c > TW) continue;  // not an owned word
      Edge e = edge_masks(grow(r), gword(c), g.gh, g.gw);
      if ((s.MY[i] & e.valid) != 0u) pending = true;
    }
    return pending;
  }
};

// =============================================================================================
// M tile: move_cell + reset_cells (steppers.fut:58-145), colour/aux payload gather, per-cell RNG
// =============================================================================================
template <int TR_, int TW_>
struct MTile {
  static constexpr int TR = TR_, TW = TW_;
  static constexpr int SR = TR + 2, SW = TW + 2, NW = SR * SW;
  struct Smem {
    u32 UYN[NW], UYS[NW], UXN[NW], UXS[NW], DYN[NW], DYS[NW], DXN[NW], DXS[NW], MY[NW], MX[NW];
    u32 AY[TR * TW], AX[TR * TW];
    unsigned int leaks;
  };
  const GridView &g;
  Smem &s;
  int y0, w0;
  FA_HD MTile(const GridView &g_, Smem &s_, int tile_y, int tile_w) : g(g_), s(s_), y0(tile_y * TR), w0(tile_w * TW) {}
  FA_HD static int idx(int r, int c) { return r * SW + c; }

  FA_HD void p0(int tid, int nt) {
    if (tid == 0) s.leaks = 0;
    for (int i = tid; i < NW; i += nt) {
      int r = i / SW, c = i - r * SW;
      int y = y0 - 1 + r, w = w0 - 1 + c;
      U4 rd = {0, 0, 0, 0}, hd = {0, 0, 0, 0};
      u32 my = 0, mx = 0;
      if (y >= 0 && y < g.gh && w >= 0 && w < g.pitch) {
        size_t o = (size_t)y * g.pitch + w;
        rd = g.road[o]; hd = g.head_in[o]; my = g.my[o]; mx = g.mx[o];
      }
      s.UYN[i] = rd.x; s.UYS[i] = rd.y; s.UXN[i] = rd.z; s.UXS[i] = rd.w;
      s.DYN[i] = hd.x; s.DYS[i] = hd.y; s.DXN[i] = hd.z; s.DXS[i] = hd.w;
      s.MY[i] = my; s.MX[i] = mx;
    }
  }
  FA_HD W3 w3(const u32 *p, int r, int c) const {
    W3 v = {p[idx(r, c - 1)], p[idx(r, c)], p[idx(r, c + 1)]};
    return v;
  }
  // centre row: x-neighbours of every field that is looked at sideways (UYS is not)
  FA_HD MRow mrow_center(int r, int c) const {
    MRow m;
    m.UYN = w3(s.UYN, r, c); m.UXN = w3(s.UXN, r, c); m.UXS = w3(s.UXS, r, c);
    m.UYS.c = s.UYS[idx(r, c)]; m.UYS.l = m.UYS.r = 0u;
    m.DYN = w3(s.DYN, r, c); m.DYS = w3(s.DYS, r, c); m.DXN = w3(s.DXN, r, c); m.DXS = w3(s.DXS, r, c);
    m.MY = w3(s.MY, r, c); m.MX = w3(s.MX, r, c);
    return m;
  }
  // rows above/below: only the same word, unless a diagonal heading needs the corner cells
  FA_HD MRow mrow_side(int r, int c, bool diag) const {
    MRow m;
    int i = idx(r, c);
    m.UYN.c = s.UYN[i]; m.UYS.c = s.UYS[i]; m.UXN.c = s.UXN[i]; m.UXS.c = s.UXS[i];
    m.DYN.c = s.DYN[i]; m.DYS.c = s.DYS[i]; m.DXN.c = s.DXN[i]; m.DXS.c = s.DXS[i];
    m.MY.c = s.MY[i]; m.MX.c = s.MX[i];
    m.UYN.l = m.UYN.r = m.UYS.l = m.UYS.r = m.UXN.l = m.UXN.r = m.UXS.l = m.UXS.r = 0u;
    m.DYN.l = m.DYN.r = m.DYS.l = m.DYS.r = m.DXN.l = m.DXN.r = m.DXS.l = m.DXS.r = 0u;
    m.MY.l = m.MY.r = m.MX.l = m.MX.r = 0u;
    if (diag) {
      m.UYN = w3(s.UYN, r, c); m.UXN = w3(s.UXN, r, c); m.MY = w3(s.MY, r, c); m.MX = w3(s.MX, r, c);
    }
    return m;
  }
  // p1: car planes of every owned word, arrival masks for the colour pass, RNG draws, leak count
  FA_HD void p1(int tid, int nt, u32 *leak_plane = nullptr) {
    unsigned int nleak = 0;
    for (int i = tid; i < TR * TW; i += nt) {
      int tr = i / TW, tc = i - tr * TW;
      int r = tr + 1, c = tc + 1;
      int y = y0 + tr, w = w0 + tc;
      if (y >= g.gh || w >= g.pitch) { s.AY[i] = 0; s.AX[i] = 0; continue; }
      size_t go = (size_t)y * g.pitch + w;
      {  // fast path: no car in this word, above, below, or in the two adjacent cells -> nothing moves here
        int ci = idx(r, c);
        u32 cars = s.DYN[ci] | s.DXN[ci] | s.DYN[ci - SW] | s.DXN[ci - SW] | s.DYN[ci + SW] | s.DXN[ci + SW] |
                   ((s.DYN[ci - 1] | s.DXN[ci - 1]) >> 31) | ((s.DYN[ci + 1] | s.DXN[ci + 1]) & 1u);
        if (cars == 0u) {
          U4 z = {0, 0, 0, 0};
          g.head_out[go] = z;
          s.AY[i] = 0; s.AX[i] = 0;
          if (leak_plane) leak_plane[go] = 0;
          continue;
        }
      }
      Edge e = edge_masks(y, w, g.gh, g.gw);
      bool diag = (s.DYN[idx(r, c)] & s.DXN[idx(r, c)]) != 0u;
      MPre p = m_pre(mrow_side(r - 1, c, diag), mrow_center(r, c), mrow_side(r + 1, c, diag), e);
      u32 CH = 0;
      u32 fk = p.fork & e.valid;
      if (fk) {
        if (g.chooser == CHOOSE_X) CH = fk;
        else if (g.chooser == CHOOSE_TAPE) CH = g.tape_bit ? fk : 0u;
        else if (g.chooser == CHOOSE_RANDOM) {
          u32 m = fk;
          while (m) {  // flowamok.fut:10-14 on the destination cell's own rng (steppers.fut:103,112)
            int b = ffs32(m);
            m &= m - 1;
            size_t ci = (size_t)y * ((size_t)g.pitch * 32) + (size_t)w * 32 + b;
            u64 st = g.rng[ci];
            if (draw_choice01(st, g.rng_inc)) CH |= 1u << b;
            g.rng[ci] = st;
          }
        }
      }
      MOut o = m_post(p, CH);
      U4 hd = {o.DYN, o.DYS, o.DXN, o.DXS};
      g.head_out[go] = hd;
      s.AY[i] = o.AY; s.AX[i] = o.AX;
      if (leak_plane) leak_plane[go] = o.leak;
      nleak += (unsigned int)popc32(o.leak);
    }
    if (nleak) FA_ATOMIC_ADD_U32(&s.leaks, nleak);
  }
  // p2: colour payload (steppers.fut:109): 4 cells (one 128-bit access) per step
  FA_HD void p2(int tid, int nt) {
    const int q_per_row = TW * 8;  // uint4 groups per tile row
    const size_t cpitch = (size_t)g.pitch * 32;
    for (int i = tid; i < TR * q_per_row; i += nt) {
      int tr = i / q_per_row, q = i - tr * q_per_row;
      int tc = q >> 3, nib = q & 7;
      int y = y0 + tr, w = w0 + tc;
      if (y >= g.gh || w >= g.pitch) continue;
      size_t base = (size_t)y * cpitch + (size_t)w * 32 + (size_t)nib * 4;
      U4 v = *reinterpret_cast<const U4 *>(g.color_in + base);
      int wi = tr * TW + tc;
      u32 ay = (s.AY[wi] >> (nib * 4)) & 15u, ax = (s.AX[wi] >> (nib * 4)) & 15u;
      if (ay | ax) {
        int si = idx(tr + 1, tc + 1);
        u32 uys = (s.UYS[si] >> (nib * 4)) & 15u, uxs = (s.UXS[si] >> (nib * 4)) & 15u;
        u32 *pv = &v.x;
        for (int k = 0; k < 4; k++) {
          u32 bit = 1u << k;
          if (ay & bit) pv[k] = g.color_in[(uys & bit) ? base + k + cpitch : base + k - cpitch];
          else if (ax & bit) pv[k] = g.color_in[(uxs & bit) ? base + k + 1 : base + k - 1];
        }
      }
      *reinterpret_cast<U4 *>(g.color_out + base) = v;
    }
  }
};

}  // namespace fa
