// flowamok_b200 -- streaming strip programs of the per-tick update (the bodies of k_u_local and k_move).
//
// One WARP walks a strip of the grid top to bottom: lane l holds plane word w0 - 1 + l of the current row, so lanes
// 1..SW own SW consecutive words (32*SW cells) and lanes 0 and SW+1 carry the neighbouring words read-only.  A row's
// x-neighbours are one warp shuffle away, its y-neighbours are the previous / next row, which the lane itself still
// holds in registers (sliding window): every plane word is loaded once per strip with fully coalesced 128-bit
// accesses (requested into L2 a few rows ahead), the control planes are never staged in shared memory and the strip
// loop has no block barrier.  The bit logic is the one of fa_bits.h.
//
//   UStrip : update_can_be_moved_to_from (steppers.fut:9-56), the first K Jacobi levels of the fixpoint (:147-164) in a
//            K + 2 stage row pipeline, publication of calc / my|mx / next_preference_flow_x, and the worklist entries +
//            frontier for k_u_iter.
//   MTile  : move_cell + reset_cells (steppers.fut:58-145), per-cell RNG draws at forks, leak count (phase A, a warp
//            per strip) and the colour payload through the lazy double buffer (phase B, the whole CTA).
//
// Written once for device and host: tests/emul runs these very functions on the CPU with the lanes of a warp as
// coroutines (fa_simt.h) and compares every tick with the oracle.
#pragma once
#include "fa_simt.h"
#include "fa_tiles.h"

// tuning knobs (rows between the L2 prefetch and the register load of a row; loads in flight per thread in phase B)
#ifndef FA_U_AHEAD
#define FA_U_AHEAD 6
#endif
#ifndef FA_M_AHEAD
#define FA_M_AHEAD 6
#endif
#ifndef FA_B1_UN
#define FA_B1_UN 8
#endif
#ifndef FA_B2_UN
#define FA_B2_UN 8
#endif

namespace fa {

// per-lane column constants of a strip
struct StripCol {
  int w;            // plane word of this lane (may lie outside [0, pitch): then it reads as zeros)
  bool w_ok;        // 0 <= w < pitch
  bool owner;       // lane owns the word: it is inside the strip and w_ok
  u32 colFirst, colLast, colValid;   // Edge masks that depend on the column only
};
template <int SW>
FA_HD StripCol strip_col(const GridView &g, int sx, int lane) {
  StripCol c;
  c.w = sx * SW - 1 + lane;
  c.w_ok = c.w >= 0 && c.w < g.pitch;
  c.owner = lane >= 1 && lane <= SW && c.w_ok;
  Edge e = edge_masks(0, c.w, 1, g.gw);
  c.colFirst = e.colFirst; c.colLast = e.colLast; c.colValid = c.w_ok ? e.valid : 0u;
  return c;
}
FA_HD W3 simt_w3(u32 v) {
  W3 r;
  r.l = simt_up1(v); r.c = v; r.r = simt_dn1(v);
  return r;
}
FA_HD W3 w3_c(u32 v) {
  W3 r = {0u, v, 0u};
  return r;
}

// =============================================================================================
// U strip.  Levels (pure Jacobi, so every level is a function of the raw planes only and a strip can recompute its
// neighbours' rows exactly):
//   L0(y)   = terminals:              R = T, MY = T & GY, MX = T & GX                  needs raw rows y-1 .. y+1
//   Lk(y)   = relax(Lk-1(y-1..y+1))   k = 1..K; LK is published as calc / my|mx         needs raw rows y-k-1 .. y+k+1
//   LK+1(y) = relax(LK(y-1..y+1))     only compared with LK(y): a word with an unresolved waiting car goes on the
//           frontier iff LK+1 != LK; otherwise it is DORMANT -- consistent with all its neighbours' published values,
//           it can only change after one of them does, and that neighbour's visit will wake it (k_u_iter).
// A queue of c cars behind a free cell is resolved here for c <= K; only longer queues (and cars behind a full ring)
// reach k_u_iter.  The row pipeline has K + 2 stages: when raw row j arrives, the statics and L0 of row j-1, Lk of
// row j-1-k and the verdict on row j-2-K are computed.
// Cars with a diagonal heading (utils.fut:35 allows them) are never relaxed here: their word always starts on the
// frontier.  Row bands: the ghost rows own0-1 / own1 stay at L0 (the raw rows beyond are not held), which is
// published as this band's lower bound of the neighbour's row; the first / last owned row then always starts on the
// frontier.
// Frontier: the strip collects its entries in a private segment (qseg) and appends them to the device-wide list
// with ONE atomic at its end -- per-row atomics on one counter serialise in L2 and dominate the kernel.
// =============================================================================================
template <int SW, int K, bool BAND>
struct UStrip {
  struct Raw {
    U4 road, head;
    u32 pref;
  };
  struct Stat {
    u32 T, GY, GX, LN, LS, LW, LE, DG;
  };
  struct Env {
    const GridView &g;
    StripCol col;
    int lane;
    int ya, yb;     // owned rows of the strip
    int ylo, yhi;   // local rows that exist in the planes and lie inside the global grid
    UEntry *table;
    unsigned int *qseg;   // this strip's private frontier segment (SW * (yb - ya) slots)
  };
  FA_HD static void load_raw(const Env &E, int y, Raw &q) {
    const U4 z = {0, 0, 0, 0};
    q.road = z; q.head = z; q.pref = 0u;
    if (E.col.w_ok && y >= E.ylo && y < E.yhi) {
      const u32 o = (u32)y * (u32)E.g.pitch + (u32)E.col.w;
      q.road = E.g.road[o]; q.head = E.g.head_in[o]; q.pref = E.g.pref_in[o];
    }
  }
  FA_HD static void prefetch_raw(const Env &E, int y) {
    if (FA_U_AHEAD > 0 && E.col.w_ok && y >= E.ylo && y < E.yhi) {
      const u32 o = (u32)y * (u32)E.g.pitch + (u32)E.col.w;
      fa_prefetch_l2(&E.g.road[o]); fa_prefetch_l2(&E.g.head_in[o]);
      if ((E.lane & 7) == 1) fa_prefetch_l2(&E.g.pref_in[o]);   // 4-byte words: one prefetch per 32-byte sector
    }
  }
  // The statics are stored masked with the valid cells (T, GY, GX; links only exist on valid cells), so every level
  // stays inside them without further masking.
  FA_HD static UState relax(const Stat &s, const UState &a, const UState &b, const UState &c) {
    UStatic st;
    st.T = s.T; st.GYr = s.GY; st.GXr = s.GX; st.LN = s.LN; st.LS = s.LS; st.LW = s.LW; st.LE = s.LE;
    st.LNW = st.LNE = st.LSW = st.LSE = 0u;
    return u_relax_card(st, a.R, simt_w3(b.R), c.R, a.MY, c.MY, simt_w3(b.MX));
  }
  // `calculated` of the next level only: a cell's my / mx can only change together with its `calculated` (they are 0
  // until it is set and final afterwards), so this is all the dormancy test needs
  FA_HD static u32 relax_r(const Stat &s, u32 aR, u32 bR, u32 cR) {
    const W3 R = simt_w3(bR);
    return s.T | (s.LN & aR) | (s.LS & cR) | (s.LW & west(R)) | (s.LE & east(R));
  }

  // One pipeline step: raw row j has arrived.  rA / rB = raw rows j-1 / j; roadc2 / dyn2 = road presence and DYN of
  // row j-2.  S[i] = statics of row j-1-i (shifted here).  Lv[k] holds level k of three consecutive rows in a ring
  // whose slots rotate with the step's phase P = (step number) mod 3: N is written now (row j-1-k), M and O hold rows
  // j-2-k and j-3-k -- the callers unroll three phases, so the rings are pure register renaming.
  // Rows outside the grid hold zeros and produce values nobody reads (a car heading out of the grid is a terminal).
  template <int P>
  FA_HD static void step(const Env &E, int j, const Raw &rA, const Raw &rB, u32 &roadc2, u32 &dyn2, Stat (&S)[K + 2],
                         UState (&Lv)[K + 1][3], u32 &nq, u32 &prefD) {
    constexpr int N = (3 - P) % 3, M = (N + 1) % 3, O = (N + 2) % 3;
    const GridView &g = E.g;
    const u32 valid = E.col.colValid;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = K + 1; i >= 1; i--) S[i] = S[i - 1];
    // ---- statics and L0 of row j-1
    {
      const int yg = j - 1 + g.row0;
      const u32 roadc1 = rA.road.x | rA.road.z, roadc0 = rB.road.x | rB.road.z;
      W3 road_up = w3_c(roadc2), road_dn = w3_c(roadc0);
      if (simt_any((rA.head.x & rA.head.z) != 0u)) {   // a diagonal heading looks at the corner cells
        road_up = simt_w3(roadc2); road_dn = simt_w3(roadc0);
      }
      Edge e;
      e.rowFirst = (yg == 0) ? 0xFFFFFFFFu : 0u;
      e.rowLast = (yg == g.gh_global - 1) ? 0xFFFFFFFFu : 0u;
      e.colFirst = E.col.colFirst; e.colLast = E.col.colLast; e.valid = valid;
      const UStatic st = u_statics(rA.road.x, rA.road.y, rA.road.z, rA.road.w, rA.head.x, rA.head.y, rA.head.z, rA.head.w,
                                   rA.pref, road_up, simt_w3(roadc1), road_dn, dyn2, rB.head.x, simt_w3(rA.head.z), e);
      Stat &sN = S[0];
      sN.T = st.T & valid; sN.GY = st.GYr & valid; sN.GX = st.GXr & valid;
      sN.LN = st.LN; sN.LS = st.LS; sN.LW = st.LW; sN.LE = st.LE;
      sN.DG = st.LNW | st.LNE | st.LSW | st.LSE;
      UState &a = Lv[0][N];
      a.R = sN.T; a.MY = sN.T & sN.GY; a.MX = sN.T & sN.GX;
      roadc2 = roadc1; dyn2 = rA.head.x;
    }
    // ---- level k of row j-1-k; a band's ghost rows stay at L0 (the raw rows beyond are not held)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k <= K; k++) {
      const int y = j - 1 - k;
      const UState n = relax(S[k], Lv[k - 1][O], Lv[k - 1][M], Lv[k - 1][N]);
      UState &b = Lv[k][N];
      b = n;
      if (BAND) {
        const bool ghost = y < g.own0 || y >= g.own1;
        b.R = ghost ? Lv[k - 1][M].R : n.R; b.MY = ghost ? Lv[k - 1][M].MY : n.MY; b.MX = ghost ? Lv[k - 1][M].MX : n.MX;
      }
    }
    if (BAND) {   // ghost row: this band's lower bound of the neighbour's row
      const int y = j - 1 - K;
      if (((y == g.own0 - 1 && E.ya == g.own0) || (y == g.own1 && E.yb == g.own1)) && E.col.owner && y >= E.ylo && y < E.yhi) {
        const u32 o = (u32)y * (u32)g.pitch + (u32)E.col.w;
        g.mymx[o] = pack_mymx(Lv[K][N].MY, Lv[K][N].MX);
        g.calc[o] = Lv[K][N].R;
      }
    }
    // ---- publish row j-2-K and decide its worklist status
    {
      const int y = j - 2 - K;
      const Stat &sO = S[K + 1];
      const UState &pub = Lv[K][M];
      const u32 nR = relax_r(sO, Lv[K][O].R, pub.R, Lv[K][N].R);
      const u32 unresolved = ~pub.R & ~sO.T & valid;
      bool front = false;
      const u32 o = (u32)y * (u32)g.pitch + (u32)E.col.w;
      if (E.col.owner && y >= E.ya && y < E.yb) {
        g.mymx[o] = pack_mymx(pub.MY, pub.MX);
        g.calc[o] = pub.R;
        // next_preference_flow_x (steppers.fut:41-50) as far as the grants are known now; k_u_iter rewrites the words
        // it changes.  prefD = the row's pref, re-read one step ago (it left the registers K + 2 steps ago).
        g.pref_out[o] = u_pref_from_grants(sO.GY, sO.GX, prefD, pub.MY, pub.MX) & valid;
        u32 us = U_NONE;
        if (unresolved) {
          front = nR != pub.R || (sO.DG & unresolved) != 0u || (BAND && (y == g.own0 || y == g.own1 - 1));
          UEntry en;
          en.wi = o; en.T = sO.T | ~valid; en.GY = sO.GY; en.GX = sO.GX;   // cells past the grid count as terminals
          E.table[o] = en;
          us = front ? (o | U_QUEUED) : o;
        }
        g.ustate[o] = us;
      }
      const u32 fm = simt_ballot(front);
      if (front) E.qseg[nq + (u32)popc32(fm & ((1u << E.lane) - 1u))] = o;
      nq += (u32)popc32(fm);
      // the next step publishes row y + 1
      prefD = 0u;
      if (E.col.owner && y + 1 >= E.ya && y + 1 < E.yb) prefD = g.pref_in[o + (u32)g.pitch];
    }
  }

  // rows [ya, yb) of strip column sx.  qseg: private staging of SW * (yb - ya) slots; qlist / qn: the device-wide
  // frontier list and its length.
  FA_HD static void run(const GridView &g, int sx, int ya, int yb, int lane, UEntry *table, unsigned int *qseg,
                        unsigned int *qlist, unsigned int *qn) {
    const int ylo = g.row0 < 0 ? -g.row0 : 0;
    const int yhi = g.gh < g.gh_global - g.row0 ? g.gh : g.gh_global - g.row0;
    const Env E = {g, strip_col<SW>(g, sx, lane), lane, ya, yb, ylo, yhi, table, qseg};
    // Raw rows ya-K-2 .. yb+K+1 are read.  Slots rotate with period 3: at step j, R0/R1/R2 hold raw rows j-1, j, j+1.
    const int j0 = ya - K - 1, jend = yb + K + 1;   // the step at jend publishes row yb-1
    Raw R0, R1, R2;
    load_raw(E, j0 - 1, R0); load_raw(E, j0, R1); load_raw(E, j0 + 1, R2);
    u32 roadc2 = 0u, dyn2 = 0u, nq = 0u, prefD = 0u;
    Stat S[K + 2];
    UState Lv[K + 1][3];
    {
      const Stat zs = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      const UState z3 = {0u, 0u, 0u};
      for (int i = 0; i < K + 2; i++) S[i] = zs;
      for (int k = 0; k <= K; k++)
        for (int i = 0; i < 3; i++) Lv[k][i] = z3;
    }
    constexpr int AHEAD = FA_U_AHEAD;   // rows between the L2 prefetch and the register load of a row
    for (int y = j0 + 2; y < j0 + 2 + AHEAD; y++) prefetch_raw(E, y);
    // the last one or two steps of the last trip may run past jend: they publish nothing (the stores check y < yb)
    for (int j = j0; j <= jend; j += 3) {
      step<0>(E, j, R0, R1, roadc2, dyn2, S, Lv, nq, prefD);
      load_raw(E, j + 2, R0); prefetch_raw(E, j + 2 + AHEAD);
      step<1>(E, j + 1, R1, R2, roadc2, dyn2, S, Lv, nq, prefD);
      load_raw(E, j + 3, R1); prefetch_raw(E, j + 3 + AHEAD);
      step<2>(E, j + 2, R2, R0, roadc2, dyn2, S, Lv, nq, prefD);
      load_raw(E, j + 4, R2); prefetch_raw(E, j + 4 + AHEAD);
    }
    // ---- append the strip's frontier to the device-wide list: one atomic per strip
    if (nq) {   // warp-uniform
      u32 base = 0u;
      if (lane == 0) base = FA_ATOMIC_ADD_U32(qn, nq);
      base = simt_bcast(base, 0);
      simt_sync();   // the staging stores of all lanes are visible to all lanes
      for (u32 k = (u32)lane; k < nq; k += 32u) qlist[base + k] = qseg[k];
    }
  }
};

// =============================================================================================
// M tile: move_cell + reset_cells (steppers.fut:58-145).  A CTA of NWARP warps takes NWARP strips side by side.
//
//   phase A (one warp per strip, sliding window like UStrip): new car planes, per-cell RNG draws at forks, leak
//            count -- and, per plane word, a 16-byte colour descriptor {T, A, V, S} in shared
//            memory: T = which of the word's four 32-byte colour sectors phase B must copy, A = arrival cells,
//            V = arrival comes along y, S = the predecessor sits on the + side (row y+1 / column x+1).
//   phase B (all threads of the CTA, after a block barrier): the colour payload (steppers.fut:109).
//            B0 scans the descriptors and compacts the touched sectors and the arrivals into two shared-memory lists;
//            B1 copies the listed sectors color_in -> color_out, a lane pair per sector, eight independent 16-byte
//               loads in flight per thread; B2 gives every listed arrival the predecessor's colour.
//            Dense lists keep every load slot useful: the phase is bound by memory latency, and the bytes in flight
//            per SM (Little's law) are what sets its speed.  While one CTA waits in phase B, the others compute.
//
// LAZY double buffer.  Invariant at the start of a tick: color_out equals color_in everywhere except, possibly, in
// the sectors where a car arrived during the previous tick, when color_out was that tick's input (dirty_in, one byte
// per plane word).  So a sector of color_out is rewritten iff it received a car in this tick or the previous one; all
// other sectors -- the empty road and every queue of standing cars -- are already right and cost no traffic.  An
// arrival reads color_in, which nobody writes during the tick.
// =============================================================================================
template <int SW, int NWARP, int CAP = 2048>
struct MTile {
  static constexpr int TW = SW * NWARP;        // plane words per tile row
  static constexpr int ARR_CAP = CAP;          // arrival list capacity; a tile with more falls back to a direct pass
  struct Row {
    U4 road, head;
    u32 my, mx;
  };
  struct Aux {
    u32 dirty;
  };
  // shared memory of one tile of `rs` rows (all offsets in 32-bit words)
  struct Smem {
    u32 *base;
    int rs;
    FA_HD U4 *desc() const { return reinterpret_cast<U4 *>(base); }                       // [rs][TW]
    FA_HD unsigned short *sect() const { return reinterpret_cast<unsigned short *>(base + (size_t)rs * TW * 4); }   // [rs*TW*4]
    FA_HD u32 *arrl() const { return base + (size_t)rs * TW * 4 + (size_t)rs * TW * 2; }   // [ARR_CAP]
    FA_HD u32 *cnt() const { return arrl() + ARR_CAP; }                                   // nsect, narr
    static size_t bytes(int rs) { return ((size_t)rs * TW * 6 + ARR_CAP + 4) * 4; }
  };
  struct Env {
    const GridView &g;
    StripCol col;
    int sx, lane, wi;   // strip column, lane, warp index inside the CTA
    int ya;
    int ylo, yhi;       // local rows that exist in the planes and lie inside the global grid
    u32 *leak_plane;
    Smem sm;
  };
  FA_HD static void load_row(const Env &E, int y, Row &q) {
    const U4 z = {0, 0, 0, 0};
    q.road = z; q.head = z; q.my = 0u; q.mx = 0u;
    if (E.col.w_ok && y >= E.ylo && y < E.yhi) {
      const u32 o = (u32)y * (u32)E.g.pitch + (u32)E.col.w;
      q.road = E.g.road[o]; q.head = E.g.head_in[o];
      const u64 m = E.g.mymx[o];
      q.my = (u32)m; q.mx = (u32)(m >> 32);
    }
  }
  FA_HD static void load_aux(const Env &E, int y, Aux &a) {
    a.dirty = 0u;
    if (E.col.owner && y < E.g.own1) a.dirty = E.g.dirty_in[(u32)y * (u32)E.g.pitch + (u32)E.col.w];
  }
  FA_HD static void prefetch_row(const Env &E, int y) {
    if (FA_M_AHEAD > 0 && E.col.w_ok && y >= E.ylo && y < E.yhi) {
      const u32 o = (u32)y * (u32)E.g.pitch + (u32)E.col.w;
      fa_prefetch_l2(&E.g.road[o]); fa_prefetch_l2(&E.g.head_in[o]);
      if ((E.lane & 3) == 1) fa_prefetch_l2(&E.g.mymx[o]);
    }
  }
  FA_HD static MRow side_row(const Row &q, bool diag) {
    MRow m;
    m.UYN = w3_c(q.road.x); m.UYS = w3_c(q.road.y); m.UXN = w3_c(q.road.z); m.UXS = w3_c(q.road.w);
    m.DYN = w3_c(q.head.x); m.DYS = w3_c(q.head.y); m.DXN = w3_c(q.head.z); m.DXS = w3_c(q.head.w);
    m.MY = w3_c(q.my); m.MX = w3_c(q.mx);
    if (diag) {   // warp-uniform: a diagonal heading looks at the corner cells
      m.UYN = simt_w3(q.road.x); m.UXN = simt_w3(q.road.z); m.MY = simt_w3(q.my); m.MX = simt_w3(q.mx);
    }
    return m;
  }

  // one owned row y: up / c / dn = rows y-1, y, y+1; ax = the row's dirty colour sectors; live = false: a padding step past
  // the strip's last row (warp-uniform), which must not store anything
  FA_HD static void step(const Env &E, int y, bool live, const Row &up, const Row &c, const Row &dn, const Aux &ax, unsigned int &nleak) {
    const GridView &g = E.g;
    const StripCol &col = E.col;
    const bool diag = simt_any((c.head.x & c.head.z) != 0u);
    MRow mc;
    mc.UYN = simt_w3(c.road.x); mc.UYS = w3_c(c.road.y); mc.UXN = simt_w3(c.road.z); mc.UXS = simt_w3(c.road.w);
    mc.DYN = simt_w3(c.head.x); mc.DYS = simt_w3(c.head.y); mc.DXN = simt_w3(c.head.z); mc.DXS = simt_w3(c.head.w);
    mc.MY = simt_w3(c.my); mc.MX = simt_w3(c.mx);
    const MRow mu = side_row(up, diag), md = side_row(dn, diag);
    Edge e;
    const int yg = y + g.row0;
    e.rowFirst = (yg == 0) ? 0xFFFFFFFFu : 0u;
    e.rowLast = (yg == g.gh_global - 1) ? 0xFFFFFFFFu : 0u;
    e.colFirst = col.colFirst; e.colLast = col.colLast; e.valid = col.colValid;   // y is an owned row: inside the grid
    const MPre p = m_pre(mu, mc, md, e);
    U4 d = {0u, 0u, 0u, 0u};
    if (col.owner && live) {   // lanes 0 and SW+1 (and lanes past the planes) only feed their neighbours
      const u32 o = (u32)y * (u32)g.pitch + (u32)col.w;
      // RNG draws at forks (flowamok.fut:10-14) on the destination cell's own state (steppers.fut:103,112)
      u32 CH = 0u;
      const u32 fk = p.fork & e.valid;
      if (fk) {
        if (g.chooser == CHOOSE_X) CH = fk;
        else if (g.chooser == CHOOSE_TAPE) CH = g.tape_bit ? fk : 0u;
        else if (g.chooser == CHOOSE_RANDOM) {
          u32 m = fk;
          while (m) {
            const int b = ffs32(m);
            m &= m - 1;
            const size_t ci = (size_t)o * 32 + b;
            u64 st = g.rng[ci];
            if (draw_choice01(st, g.rng_inc)) CH |= 1u << b;
            g.rng[ci] = st;
          }
        }
      }
      const MOut out = m_post(p, CH);
      const U4 hd = {out.DYN, out.DYS, out.DXN, out.DXS};
      g.head_out[o] = hd;
      if (E.leak_plane) {   // flowamok_step_leaks: the leak mask, and (second plane) the cells that drew from their rng in this tick
        E.leak_plane[o] = out.leak;
        E.leak_plane[o + (size_t)g.gh * g.pitch] = g.chooser == CHOOSE_RANDOM ? out.fork : 0u;
      }
      nleak += (unsigned int)popc32(out.leak);
      d.y = out.AY | out.AX;
      u32 arrS = 0u;   // the word's 32-byte colour sectors that receive a car (bit k = cells 8k .. 8k+7)
      for (int k = 0; k < 4; k++) arrS |= ((d.y >> (8 * k)) & 0xFFu) ? (1u << k) : 0u;
      g.dirty_out[o] = (unsigned char)arrS;
      d.x = arrS | ax.dirty;   // sectors phase B1 copies: color_out is stale where last tick's cars arrived
      d.z = out.AY;
      d.w = (out.AY & p.UYS) | (out.AX & p.UXS);
    }
    if (E.lane >= 1 && E.lane <= SW && live) E.sm.desc()[(size_t)(y - E.ya) * TW + E.wi * SW + (E.lane - 1)] = d;
  }

  // ---- phase A: rows [ya, yb) of strip column sx, by one warp; returns this lane's number of leaked cars.
  // has_strip = false (warp-uniform): the CTA's tile is wider than the grid here; the warp only clears its descriptors.
  FA_HD static unsigned int phaseA(const GridView &g, bool has_strip, int sx, int wi, int ya, int yb, int lane, u32 *leak_plane, Smem sm) {
    if (!has_strip) {
      const U4 z = {0u, 0u, 0u, 0u};
      if (lane >= 1 && lane <= SW)
        for (int r = 0; r < yb - ya; r++) sm.desc()[(size_t)r * TW + wi * SW + (lane - 1)] = z;
      return 0u;
    }
    const int ylo = g.row0 < 0 ? -g.row0 : 0;
    const int yhi = g.gh < g.gh_global - g.row0 ? g.gh : g.gh_global - g.row0;
    const Env E = {g, strip_col<SW>(g, sx, lane), sx, lane, wi, ya, ylo, yhi, leak_plane, sm};
    unsigned int nleak = 0;
    // Slots rotate with period 4: at the start of step y, R0/R1/R2 hold rows y-1, y, y+1 and R3 is loaded with row y+2
    Row R0, R1, R2, R3;
    Aux X0, X1;
    load_row(E, ya - 1, R0); load_row(E, ya, R1); load_row(E, ya + 1, R2);
    load_aux(E, ya, X0);
    constexpr int AHEAD = FA_M_AHEAD;   // rows between the L2 prefetch and the register load of a row
    for (int y = ya + 2; y < ya + 2 + AHEAD && y <= yb; y++) prefetch_row(E, y);
    // the last trip may run up to three rows past yb: those steps compute but store nothing (live = false)
    for (int y = ya; y < yb; y += 4) {
      load_row(E, y + 2, R3); load_aux(E, y + 1, X1); if (y + 2 + AHEAD <= yb) prefetch_row(E, y + 2 + AHEAD);
      step(E, y, true, R0, R1, R2, X0, nleak);
      load_row(E, y + 3, R0); load_aux(E, y + 2, X0); if (y + 3 + AHEAD <= yb) prefetch_row(E, y + 3 + AHEAD);
      step(E, y + 1, y + 1 < yb, R1, R2, R3, X1, nleak);
      load_row(E, y + 4, R1); load_aux(E, y + 3, X1); if (y + 4 + AHEAD <= yb) prefetch_row(E, y + 4 + AHEAD);
      step(E, y + 2, y + 2 < yb, R2, R3, R0, X0, nleak);
      load_row(E, y + 5, R2); load_aux(E, y + 4, X0); if (y + 5 + AHEAD <= yb) prefetch_row(E, y + 5 + AHEAD);
      step(E, y + 3, y + 3 < yb, R3, R0, R1, X1, nleak);
    }
    return nleak;
  }

  // ---- phase B.  tsx = first strip column of the tile, rows = yb - ya; tid in [0, nt).  Block barriers separate
  // clear | A | B0 | B1 | B2 (the caller places them).
  FA_HD static void phaseB_clear(Smem sm, int tid) {
    if (tid < 2) sm.cnt()[tid] = 0u;
  }
  // arrival list entry: row (10 bits) | cell inside the tile row (17 bits) | vertical (1) | plus side (1)
  FA_HD static u32 arr_entry(int row, int wd, int b, const U4 &d) {
    const u32 bit = 1u << b;
    return ((u32)row << 19) | ((u32)(wd * 32 + b) << 2) | ((d.z & bit) ? 2u : 0u) | ((d.w & bit) ? 1u : 0u);
  }
  FA_HD static void patch_addr(const GridView &g, int tsx, int ya, u32 en, size_t &dst, size_t &src) {
    const size_t cpitch = (size_t)g.pitch * 32;
    dst = ((size_t)(ya + (int)(en >> 19)) * g.pitch + (size_t)tsx * SW) * 32 + ((en >> 2) & 0x1FFFFu);
    if (en & 2u) src = (en & 1u) ? dst + cpitch : dst - cpitch;
    else src = (en & 1u) ? dst + 1 : dst - 1;
  }
  FA_HD static void patch(const GridView &g, int tsx, int ya, u32 en) {
    size_t dst, src;
    patch_addr(g, tsx, ya, en, dst, src);
    g.color_out[dst] = g.color_in[src];
  }
  FA_HD static void phaseB0(Smem sm, int rows, int tid, int nt) {
    const U4 *desc = sm.desc();
    unsigned short *sect = sm.sect();
    u32 *arrl = sm.arrl();
    for (int idx = tid; idx < rows * TW; idx += nt) {
      const U4 d = desc[idx];
      if (d.x == 0u) continue;
      const int row = idx / TW, wd = idx - row * TW;
      const u32 smk = d.x;
      u32 at = FA_ATOMIC_ADD_U32(&sm.cnt()[0], (u32)popc32(smk));
      for (int k = 0; k < 4; k++)
        if (smk & (1u << k)) sect[at++] = (unsigned short)(row * (TW * 4) + wd * 4 + k);
      if (d.y) {
        u32 aa = FA_ATOMIC_ADD_U32(&sm.cnt()[1], (u32)popc32(d.y));
        u32 m = d.y;
        while (m) {
          const int b = ffs32(m);
          m &= m - 1;
          if (aa < (u32)ARR_CAP) arrl[aa] = arr_entry(row, wd, b, d);
          aa++;
        }
      }
    }
  }
  // copy the listed sectors: thread pair (2p, 2p+1) moves the two 16-byte halves of sector entries p, p + nt/2, ...
  FA_HD static void phaseB1(const GridView &g, Smem sm, int tsx, int ya, int tid, int nt) {
    const unsigned short *sect = sm.sect();
    const int ns = (int)sm.cnt()[0], half = tid & 1, np = nt >> 1;
    constexpr int UN = FA_B1_UN;
    for (int e0 = tid >> 1; e0 < ns; e0 += np * UN) {
      U4 v[UN];
      size_t off[UN];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < UN; k++) {
        const int e = e0 + k * np;
        if (e < ns) {
          const int id = sect[e], row = id / (TW * 4), sc = id - row * (TW * 4);
          off[k] = ((size_t)(ya + row) * g.pitch + (size_t)tsx * SW) * 32 + (size_t)sc * 8 + half * 4;
          v[k] = *reinterpret_cast<const U4 *>(g.color_in + off[k]);
        }
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int k = 0; k < UN; k++)
        if (e0 + k * np < ns) *reinterpret_cast<U4 *>(g.color_out + off[k]) = v[k];
    }
  }
  // every arrival cell takes the predecessor's colour (steppers.fut:109)
  FA_HD static void phaseB2(const GridView &g, Smem sm, int tsx, int ya, int rows, int tid, int nt) {
    const int na = (int)sm.cnt()[1];
    if (na <= ARR_CAP) {
      const u32 *arrl = sm.arrl();
      constexpr int UN = FA_B2_UN;   // independent gathers in flight per thread
      for (int a0 = tid; a0 < na; a0 += nt * UN) {
        u32 v[UN];
        size_t dst[UN];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < UN; k++)
          if (a0 + k * nt < na) {
            size_t src;
            patch_addr(g, tsx, ya, arrl[a0 + k * nt], dst[k], src);
            v[k] = g.color_in[src];
          }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < UN; k++)
          if (a0 + k * nt < na) g.color_out[dst[k]] = v[k];
      }
    } else {   // more arrivals than the list holds (a jam dissolving everywhere at once): straight from the descriptors
      const U4 *desc = sm.desc();
      for (int idx = tid; idx < rows * TW; idx += nt) {
        const U4 d = desc[idx];
        const int row = idx / TW, wd = idx - row * TW;
        u32 m = d.y;
        while (m) {
          const int b = ffs32(m);
          m &= m - 1;
          patch(g, tsx, ya, arr_entry(row, wd, b, d));
        }
      }
    }
  }
};

// =============================================================================================
// Speculative move + fix-up.  k_move is a pure gather of radius one (steppers.fut:58-125), and the only part of its input
// that the global fixpoint still changes after k_u_local is my|mx -- of a few thousand words.  So the fixpoint keeps its
// grants in a delta plane (GridView::mymx_B), k_move runs on the plane k_u_local published BESIDE k_u_iter, and afterwards
// fix_word recomputes every word in the 3 x 3 neighbourhood of a word with a delta from the final my|mx (plane | delta).
// What makes this exact:
//   * a cell's (calc, my, mx) is written once, so the final planes only ADD bits: every arrival the
//     speculative pass saw stays what it was, new ones appear;
//   * the new heading plane and the dirty-sector byte are simply rewritten; a new arrival's colour is patched in (its sector of
//     color_out equals color_in unless it was copied already: the lazy buffer's invariant);
//   * the per-cell RNG must draw once per fork arrival: arrivals the speculative pass saw are recognised by recomputing
//     the word from the plane without the delta, and their choice is read back from the heading it wrote; only NEW fork
//     arrivals draw;
//   * leaks are counted as the difference between the two recomputations.
// One thread per word; claim[] (tagged with the tick) makes sure a word is fixed once even if several changed words surround it.
// =============================================================================================
struct FixNbh {
  MRow up, c, dn;
};
// rows y-1 .. y+1, words w-1 .. w+1: statics and headings once, my|mx without (nL) and with (nF) the fixpoint's delta
FA_HD void fix_load(const GridView &g, int y, int w, int ylo, int yhi, FixNbh &nL, FixNbh &nF) {
  MRow *rowsL[3] = {&nL.up, &nL.c, &nL.dn}, *rowsF[3] = {&nF.up, &nF.c, &nF.dn};
  for (int r = 0; r < 3; r++) {
    u32 v[12][3];
    for (int k = 0; k < 3; k++) {
      const int yy = y + r - 1, ww = w + k - 1;
      U4 rd = {0u, 0u, 0u, 0u}, hd = {0u, 0u, 0u, 0u};
      u64 m = 0ull, d = 0ull;
      if (yy >= ylo && yy < yhi && ww >= 0 && ww < g.pitch) {
        const size_t o = (size_t)yy * g.pitch + ww;
        rd = g.road[o]; hd = g.head_in[o]; m = g.mymx[o]; d = g.mymx_B[o];
      }
      v[0][k] = rd.x; v[1][k] = rd.y; v[2][k] = rd.z; v[3][k] = rd.w;
      v[4][k] = hd.x; v[5][k] = hd.y; v[6][k] = hd.z; v[7][k] = hd.w;
      v[8][k] = (u32)m; v[9][k] = (u32)(m >> 32);
      v[10][k] = (u32)(m | d); v[11][k] = (u32)((m | d) >> 32);
    }
    for (int pass = 0; pass < 2; pass++) {
      MRow &q = pass ? *rowsF[r] : *rowsL[r];
      W3 *f[10] = {&q.UYN, &q.UYS, &q.UXN, &q.UXS, &q.DYN, &q.DYS, &q.DXN, &q.DXS, &q.MY, &q.MX};
      for (int i = 0; i < 10; i++) {
        const int j = (pass && i >= 8) ? i + 2 : i;
        f[i]->l = v[j][0]; f[i]->c = v[j][1]; f[i]->r = v[j][2];
      }
    }
  }
}
// returns the number of additional leaks
FA_HD unsigned int fix_word(const GridView &g, int y, int w) {
  const int ylo = g.row0 < 0 ? -g.row0 : 0;
  const int yhi = g.gh < g.gh_global - g.row0 ? g.gh : g.gh_global - g.row0;
  const size_t o = (size_t)y * g.pitch + w;
  const Edge e = edge_of(g, y, w);
  FixNbh nL, nF;
  fix_load(g, y, w, ylo, yhi, nL, nF);
  const MPre pL = m_pre(nL.up, nL.c, nL.dn, e), pF = m_pre(nF.up, nF.c, nF.dn, e);
  const u32 fkL = pL.fork & e.valid, fkF = pF.fork & e.valid;
  // choices: the speculative pass's for the forks it saw (a fork arrival's new heading has DXN set iff the choice was x)
  u32 CH = g.head_out[o].z & fkL;
  u32 m = fkF & ~fkL;
  if (m) {
    if (g.chooser == CHOOSE_X) CH |= m;
    else if (g.chooser == CHOOSE_TAPE) CH |= g.tape_bit ? m : 0u;
    else if (g.chooser == CHOOSE_RANDOM)
      while (m) {
        const int b = ffs32(m);
        m &= m - 1;
        const size_t ci = o * 32 + b;
        u64 st = g.rng[ci];
        if (draw_choice01(st, g.rng_inc)) CH |= 1u << b;
        g.rng[ci] = st;
      }
  }
  const MOut out = m_post(pF, CH);
  const U4 hd = {out.DYN, out.DYS, out.DXN, out.DXS};
  g.head_out[o] = hd;
  const u32 arrF = out.AY | out.AX, arrL = (pL.AY | pL.AX) & e.valid;
  u32 arrS = 0u;
  for (int k = 0; k < 4; k++) arrS |= ((arrF >> (8 * k)) & 0xFFu) ? (1u << k) : 0u;
  g.dirty_out[o] = (unsigned char)arrS;
  const size_t cpitch = (size_t)g.pitch * 32;
  u32 na = arrF & ~arrL;
  while (na) {   // a new arrival takes its predecessor's colour (steppers.fut:109)
    const int b = ffs32(na);
    na &= na - 1;
    const u32 bit = 1u << b;
    const size_t dst = o * 32 + b;
    size_t src;
    if (out.AY & bit) src = (pF.UYS & bit) ? dst + cpitch : dst - cpitch;
    else src = (pF.UXS & bit) ? dst + 1 : dst - 1;
    g.color_out[dst] = g.color_in[src];
  }
  return (unsigned int)(popc32(out.leak) - popc32(pL.leak & e.valid));
}
// item t of the fix-up: neighbour t % 9 of listed word t / 9; claim: one u32 per plane word, tag = tick number + 1
FA_HD unsigned int fix_item(const GridView &g, u32 *claim, u32 tag, unsigned int t) {
  const u32 wi = g.chg_list[t / 9u];
  const int k = (int)(t % 9u);
  const int y = (int)(wi / (u32)g.pitch) + k / 3 - 1, w = (int)(wi % (u32)g.pitch) + k % 3 - 1;
  if (y < g.own0 || y >= g.own1 || w < 0 || w >= g.pitch) return 0u;   // only owned words are moved
  const size_t o = (size_t)y * g.pitch + w;
#if defined(__CUDA_ARCH__)
  if (atomicExch(&claim[o], tag) == tag) return 0u;
#else
  if (claim[o] == tag) return 0u;
  claim[o] = tag;
#endif
  return fix_word(g, y, w);
}

}  // namespace fa
