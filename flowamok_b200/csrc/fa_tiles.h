// flowamok_b200 -- tile programs of the per-tick update, written once for device and host.
//
// Each program is a sequence of phases; inside a phase every "thread" tid in [0, nt) works on a
// strided set of words and phases are separated by a block barrier.  The CUDA kernels
// (fa_kernels.cu) call phase(threadIdx.x, blockDim.x) with __syncthreads() between phases; the CPU
// emulation used by tests/emul runs the same phases with a loop over tid.  So the bit logic, the
// indexing and the worklist protocol are all unit-tested
// on the CPU against the oracle before they ever reach a GPU.
//
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
  int gh, gw;        // rows held in this device's planes (a band plus its halo rows) / cells per row
  int pitch;         // words per plane row (multiple of 4 -> 16-byte aligned rows)
  int wvalid;        // ceil(gw / 32): words that hold cells
  int gh_global;     // rows of the whole grid (== gh on a single GPU)
  int row0;          // global row of local row 0 (0 on a single GPU; -halo for the first band)
  int own0, own1;    // local rows [own0, own1) are owned (stepped and written) by this device
  const U4 *road;    // {UYN, UYS, UXN, UXS} per word                       [gh][pitch]
  const U4 *head_in; // {DYN, DYS, DXN, DXS} per word, state before the tick [gh][pitch]
  U4 *head_out;      // state after the tick
  const u32 *pref_in;
  u32 *pref_out;
  u32 *calc;               // per-tick plane `calculated` (steppers.fut:140-145 resets it: here it is scratch)
  u64 *mymx;               // per-tick planes direction.y (low word) / direction.x (high word), one 64-bit access
  u32 *ustate;             // per word: U_NONE, or its worklist entry index (| U_QUEUED while on the next list)
  const u32 *color_in;     // [gh][pitch*32] argb payload
  u32 *color_out;
  u64 *rng;                // [gh][pitch*32] pcg32 state of each cell (inc is shared, utils.fut:15)
  u64 rng_inc;
  int chooser;             // CHOOSE_*
  u32 tape_bit;            // CHOOSE_TAPE: this tick's choice
};

#if defined(__CUDA_ARCH__)
#define FA_ATOMIC_ADD_U32(p, v) atomicAdd((p), (v))
#define FA_ATOMIC_ADD_U64(p, v) atomicAdd((p), (v))
#define FA_ATOMIC_OR_U32(p, v) atomicOr((p), (v))
#define FA_ATOMIC_AND_U32(p, v) atomicAnd((p), (v))
#define FA_ATOMIC_CAS_U32(p, c, v) atomicCAS((p), (c), (v))
#define FA_ATOMIC_OR_U64(p, v) atomicOr((unsigned long long *)(p), (unsigned long long)(v))
#else
template <class T> inline T fa_host_add(T *p, T v) { T o = *p; *p = o + v; return o; }
#define FA_ATOMIC_ADD_U32(p, v) fa_host_add<unsigned int>((p), (v))
#define FA_ATOMIC_ADD_U64(p, v) fa_host_add<unsigned long long>((p), (v))
template <class T> inline T fa_host_or(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T fa_host_and(T *p, T v) { T o = *p; *p = o & v; return o; }
#define FA_ATOMIC_OR_U32(p, v) fa_host_or<unsigned int>((p), (v))
#define FA_ATOMIC_AND_U32(p, v) fa_host_and<unsigned int>((p), (v))
inline unsigned int fa_host_cas(unsigned int *p, unsigned int c, unsigned int v) { unsigned int o = *p; if (o == c) *p = v; return o; }
#define FA_ATOMIC_CAS_U32(p, c, v) fa_host_cas((p), (c), (v))
#define FA_ATOMIC_OR_U64(p, v) fa_host_or<unsigned long long>((unsigned long long *)(p), (unsigned long long)(v))
#endif

// local row y exists in the planes AND lies inside the global grid
FA_HD bool row_in(const GridView &g, int y) { return y >= 0 && y < g.gh && y + g.row0 >= 0 && y + g.row0 < g.gh_global; }
FA_HD Edge edge_of(const GridView &g, int y, int w) { return edge_masks(y + g.row0, w, g.gh_global, g.gw); }

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
// U phase: least fixpoint of update_can_be_moved_to_from (steppers.fut:9-56, :147-164)
//
//   UInitTile   : every word once (k_u_init): statics, initial state (terminals calculated), a few
//                 relaxation sweeps inside the block in shared memory, and a worklist entry for each
//                 word that still holds an unresolved waiting car ("link" cell: occupied, successor in
//                 bounds and on a road).
//   u_iter_entry: one monotone relaxation of a worklist entry against the CURRENT global planes
//                 (k_u_iter, frontier discipline).
//
// Only link cells can change after initialisation, so the iterations touch nothing else.  Confluence
//  makes any relaxation order reach the reference's Jacobi result.  A word's state is
// published my/mx first, calc last, and read calc first: whoever sees calc = 1 sees the final my/mx.
// =============================================================================================
FA_HD void fa_fence_device() {
#if defined(__CUDA_ARCH__)
  __threadfence();
#endif
}
FA_HD u64 pack_mymx(u32 my, u32 mx) { return (u64)my | ((u64)mx << 32); }

constexpr u32 U_NONE = 0xFFFFFFFFu;    // word has no unresolved waiting car
constexpr u32 U_QUEUED = 0x80000000u;  // word is already on the next iteration's list

struct alignas(16) UEntry {  // 16 bytes: one 128-bit access
  u32 wi;        // plane word index
  u32 T, GY, GX; // statics of the word (UStatic.T, GYr, GXr)
};

// k_u_init block program: one block owns TR x TW words (one per thread) and sees one halo word all round.
//   p0: ONE round of global loads: every thread keeps its own word's raw planes in registers; the
//       first N_RING threads additionally fetch one word of the two rings around the tile.  The
//       road-presence / DYN / DXN planes of the region plus one more ring go to shared memory.
//   pA: statics + initial state (terminals calculated) of every region word, in shared memory
//   sweeps (<= NS): relax the owned words against the shared-memory neighbourhood; halo words keep
//       their initial state, so everything stays a valid lower bound of the fixpoint
//   pC: publish the owned words; a word that still holds an uncalculated link cell gets a worklist
//       entry.  It starts on the frontier only if it touches the halo or the block did not converge;
//       otherwise it is dormant: consistent with all its neighbours, it can only change after one of
//       them does, and that neighbour will wake it.
template <int TR_, int TW_>
struct UInitTile {
  static constexpr int TR = TR_, TW = TW_, SR = TR + 2, SW = TW + 2, NW = SR * SW, NT = TR * TW, NS = 8;
  static constexpr int ER = TR + 4, EW = TW + 4, NE = ER * EW;  // region + one more ring: raw planes the statics look at
  static constexpr int N_HALO = 2 * SW + 2 * TR;                 // ring 1 (region border): statics needed
  static constexpr int N_OUTER = 2 * EW + 2 * SR;                // ring 2 (extended border): raw planes only
  static constexpr int N_RING = N_HALO + N_OUTER;
  struct Smem {
    u32 T[NW], GY[NW], GX[NW], R[NW], MY[NW], MX[NW];
    u32 ROAD[NE], DYN[NE], DXN[NE];
  };
  struct Raw {  // raw planes of one word, in registers
    bool in;
    int y, w;
    U4 road, head;
    u32 pref;
  };
  struct Own {  // registers of the thread that owns an interior word
    bool in;
    int y, w, si;
    UStatic st;
    u32 valid;
    UState cur;
  };
  const GridView &g;
  Smem &s;
  int y0, w0;
  FA_HD UInitTile(const GridView &g_, Smem &s_, int ty, int tx) : g(g_), s(s_), y0(g_.own0 + ty * TR), w0(tx * TW) {}
  FA_HD static int idx(int r, int c) { return r * SW + c; }
  FA_HD static int eidx(int r, int c) { return r * EW + c; }   // region (r, c) sits at extended (r + 1, c + 1)
  FA_HD W3 w3(const u32 *p, int r, int c) const {
    W3 v = {p[idx(r, c - 1)], p[idx(r, c)], p[idx(r, c + 1)]};
    return v;
  }
  FA_HD W3 e3(const u32 *p, int er, int ec) const {
    W3 v = {p[eidx(er, ec - 1)], p[eidx(er, ec)], p[eidx(er, ec + 1)]};
    return v;
  }
  // extended coordinates of ring word k: first ring 1 (region border), then ring 2 (extended border)
  FA_HD static void ring_pos(int k, int &er, int &ec) {
    if (k < N_HALO) {
      int r, c;
      if (k < SW) { r = 0; c = k; }
      else if (k < 2 * SW) { r = SR - 1; c = k - SW; }
      else if (k < 2 * SW + TR) { r = k - 2 * SW + 1; c = 0; }
      else { r = k - 2 * SW - TR + 1; c = SW - 1; }
      er = r + 1; ec = c + 1;
    } else {
      k -= N_HALO;
      if (k < EW) { er = 0; ec = k; }
      else if (k < 2 * EW) { er = ER - 1; ec = k - EW; }
      else if (k < 2 * EW + SR) { er = k - 2 * EW + 1; ec = 0; }
      else { er = k - 2 * EW - SR + 1; ec = EW - 1; }
    }
  }
  FA_HD Raw load_raw(int er, int ec, bool want_pref) const {
    Raw q;
    q.y = y0 - 2 + er; q.w = w0 - 2 + ec;
    q.in = row_in(g, q.y) && q.w >= 0 && q.w < g.pitch;
    U4 z = {0, 0, 0, 0};
    q.road = z; q.head = z; q.pref = 0;
    if (q.in) {
      size_t o = (size_t)q.y * g.pitch + q.w;
      q.road = g.road[o]; q.head = g.head_in[o];
      if (want_pref) q.pref = g.pref_in[o];
    }
    return q;
  }
  FA_HD void stage(int er, int ec, const Raw &q) {
    int i = eidx(er, ec);
    s.ROAD[i] = q.road.x | q.road.z; s.DYN[i] = q.head.x; s.DXN[i] = q.head.z;
  }
  // p0: owner thread t loads its interior word; threads t < N_RING also load ring word t
  FA_HD void p0(int t, Raw &own, Raw &ring) {
    int tr = t / TW, tc = t - tr * TW;
    own = load_raw(tr + 2, tc + 2, true);
    if (t < N_RING) {
      int er, ec;
      ring_pos(t, er, ec);
      ring = load_raw(er, ec, t < N_HALO);
      stage(er, ec, ring);
    }
    if (N_RING > NT)  // tiny test tiles only: more ring words than threads
      for (int k = t + NT; k < N_RING; k += NT) {
        int er, ec;
        ring_pos(k, er, ec);
        stage(er, ec, load_raw(er, ec, false));
      }
    stage(tr + 2, tc + 2, own);
  }
  // statics + initial state of region word (r, c) from its raw planes and the staged neighbourhood
  FA_HD UStatic pA_word(int r, int c, const Raw &q, Edge &ed, Heading &h) {
    int i = idx(r, c), er = r + 1, ec = c + 1;
    UStatic st = {};
    u32 R = 0, MY = 0, MX = 0;
    ed = edge_of(g, q.y, q.w);
    h = heading(q.head.x, q.head.y, q.head.z, q.head.w);
    if (q.in) {
      st = u_statics(q.road.x, q.road.y, q.road.z, q.road.w, q.head.x, q.head.y, q.head.z, q.head.w, q.pref,
                     e3(s.ROAD, er - 1, ec), e3(s.ROAD, er, ec), e3(s.ROAD, er + 1, ec), s.DYN[eidx(er - 1, ec)],
                     s.DYN[eidx(er + 1, ec)], e3(s.DXN, er, ec), ed);
      R = st.T & ed.valid; MY = R & st.GYr; MX = R & st.GXr;
    }
    s.T[i] = st.T; s.GY[i] = st.GYr; s.GX[i] = st.GXr;
    s.R[i] = R; s.MY[i] = MY; s.MX[i] = MX;
    return st;
  }
  FA_HD void pA(int t, const Raw &own, const Raw &ring, Own &o) {
    int tr = t / TW, tc = t - tr * TW;
    o.si = idx(tr + 1, tc + 1);
    o.y = own.y; o.w = own.w; o.in = own.in && own.y >= g.own0 && own.y < g.own1;   // rows past the band are halo-like
    Edge ed; Heading h;
    o.st = pA_word(tr + 1, tc + 1, own, ed, h);
    o.valid = ed.valid;
    u32 link = ~o.st.T;
    o.st.LN = link & h.cN; o.st.LS = link & h.cS; o.st.LW = link & h.cW; o.st.LE = link & h.cE;
    o.st.LNW = link & h.dNW; o.st.LNE = link & h.dNE; o.st.LSW = link & h.dSW; o.st.LSE = link & h.dSE;
    o.cur.R = s.R[o.si]; o.cur.MY = s.MY[o.si]; o.cur.MX = s.MX[o.si];
    if (t < N_HALO) {
      int er, ec;
      ring_pos(t, er, ec);
      Edge e2; Heading h2;
      (void)pA_word(er - 1, ec - 1, ring, e2, h2);
    }
    if (N_HALO > NT)  // tiny test tiles only
      for (int k = t + NT; k < N_HALO; k += NT) {
        int er, ec;
        ring_pos(k, er, ec);
        Edge e2; Heading h2;
        (void)pA_word(er - 1, ec - 1, load_raw(er, ec, true), e2, h2);
      }
  }
  // one sweep: compute from shared memory (no writes) ...
  FA_HD bool sweep_compute(Own &o, UState &n) const {
    n = o.cur;
    // nothing to do unless an owned link cell is still uncalculated: a calculated cell is final (commit is atomic
    // with respect to the sweep barrier, so R = 1 always comes with its final MY/MX)
    if (!o.in || (~o.cur.R & ~o.st.T & o.valid) == 0u) return false;
    int r = o.si / SW, c = o.si - r * SW;
    if ((o.st.LNW | o.st.LNE | o.st.LSW | o.st.LSE) == 0u)
      n = u_relax_card(o.st, s.R[idx(r - 1, c)], w3(s.R, r, c), s.R[idx(r + 1, c)], s.MY[idx(r - 1, c)], s.MY[idx(r + 1, c)],
                       w3(s.MX, r, c));
    else
      n = u_relax(o.st, w3(s.R, r - 1, c), w3(s.R, r, c), w3(s.R, r + 1, c), w3(s.MY, r - 1, c), w3(s.MY, r + 1, c),
                  w3(s.MX, r, c), w3(s.MX, r - 1, c), w3(s.MX, r + 1, c));
    n.R &= o.valid; n.MY &= o.valid; n.MX &= o.valid;
    return n.R != o.cur.R || n.MY != o.cur.MY || n.MX != o.cur.MX;
  }
  // ... then commit
  FA_HD void sweep_commit(Own &o, const UState &n) {
    if (!o.in) return;
    o.cur = n;
    s.R[o.si] = n.R; s.MY[o.si] = n.MY; s.MX[o.si] = n.MX;
  }
  // Row bands: the rows just outside the owned range (own0-1 and own1) belong to the neighbouring device; the
  // fixpoint needs a lower bound of their state, which is the initial state this block computed anyway.
  FA_HD void ghost_word(int r, int c) const {
    int y = y0 - 1 + r, w = w0 - 1 + c;
    if ((y != g.own0 - 1 && y != g.own1) || !row_in(g, y) || w < 0 || w >= g.pitch) return;
    int i = idx(r, c);
    size_t off = (size_t)y * g.pitch + w;
    g.mymx[off] = pack_mymx(s.MY[i], s.MX[i]);
    g.calc[off] = s.R[i];
  }
  FA_HD void ghost_publish(int t) const {
    int tr = t / TW, tc = t - tr * TW;
    ghost_word(tr + 1, tc + 1);
    for (int k = t; k < N_HALO; k += NT) {
      int er, ec;
      ring_pos(k, er, ec);
      ghost_word(er - 1, ec - 1);
    }
  }
  // publish; returns 0 = no entry, 1 = dormant entry, 2 = entry that starts on the frontier
  FA_HD int pC(int t, const Own &o, bool converged, UEntry &e) const {
    if (!o.in) return 0;
    size_t off = (size_t)o.y * g.pitch + o.w;
    g.mymx[off] = pack_mymx(o.cur.MY, o.cur.MX);
    g.calc[off] = o.cur.R;
    e.wi = (u32)off; e.T = o.st.T; e.GY = o.st.GYr; e.GX = o.st.GXr;
    if ((~o.cur.R & ~o.st.T & o.valid) == 0u) return 0;
    int tr = t / TW, tc = t - tr * TW;
    bool border = tr == 0 || tr == TR - 1 || tc == 0 || tc == TW - 1;
    // a band's first/last owned row looks at a ghost row that the neighbour refines after this kernel
    if (g.gh != g.gh_global && (o.y == g.own0 || o.y == g.own1 - 1)) border = true;
    return (border || !converged) ? 2 : 1;
  }
};

struct UPlanes {  // calc / my / mx of one word
  u32 R, MY, MX;
};
FA_HD size_t u_off(const GridView &g, int y, int w, bool &ok) {
  ok = row_in(g, y) && w >= 0 && w < g.pitch;
  return ok ? (size_t)y * g.pitch + w : 0;
}
FA_HD u32 u_ld_calc(const GridView &g, size_t o, bool ok) {
  if (!ok) return 0u;
#if defined(__CUDA_ARCH__)
  return *(volatile const u32 *)&g.calc[o];
#else
  return g.calc[o];
#endif
}
FA_HD void u_ld_mymx(const GridView &g, size_t o, bool ok, UPlanes &p) {
  u64 m = 0;
  if (ok) {
#if defined(__CUDA_ARCH__)
    m = *(volatile const u64 *)&g.mymx[o];
#else
    m = g.mymx[o];
#endif
  }
  p.MY = (u32)m; p.MX = (u32)(m >> 32);
}

// Put word (y, w) on the next list unless it has nothing unresolved or is queued already.
// Returns its entry index, or U_NONE.
FA_HD u32 u_try_queue(const GridView &g, int y, int w) {
  if (y < g.own0 || y >= g.own1 || w < 0 || w >= g.pitch) return U_NONE;   // only owned words have entries
  size_t o = (size_t)y * g.pitch + w;
#if defined(__CUDA_ARCH__)
  u32 sv = *(volatile const u32 *)&g.ustate[o];
#else
  u32 sv = g.ustate[o];
#endif
  if (sv == U_NONE || (sv & U_QUEUED)) return U_NONE;
  u32 old = FA_ATOMIC_OR_U32(&g.ustate[o], U_QUEUED);
  if (old == U_NONE || (old & U_QUEUED)) return U_NONE;
  return old;
}

// One relaxation of worklist entry `e` against the current planes (frontier discipline):
//   clear QUEUED, re-read the neighbourhood, OR the new facts in (my/mx before calc);
//   if anything changed, wake the eight neighbouring words (their waiting cars may point here)
//   and this word itself while it still holds an uncalculated link cell;
//   a word nobody changes is left dormant -- cars queued behind a full ring cost one visit per tick.
// push[] receives up to 9 entry indices for the next list; returns their number.
FA_HD int u_iter_entry(const GridView &g, const UEntry &e, u32 idx, u32 *push) {
  int y = (int)(e.wi / (u32)g.pitch), w = (int)(e.wi - (u32)y * (u32)g.pitch);
  FA_ATOMIC_AND_U32(&g.ustate[e.wi], ~U_QUEUED);
  fa_fence_device();
  U4 hd = g.head_in[e.wi];
  Heading h = heading(hd.x, hd.y, hd.z, hd.w);
  UStatic st;
  u32 link = ~e.T;
  st.T = e.T; st.GYr = e.GY; st.GXr = e.GX;
  st.LN = link & h.cN; st.LS = link & h.cS; st.LW = link & h.cW; st.LE = link & h.cE;
  st.LNW = link & h.dNW; st.LNE = link & h.dNE; st.LSW = link & h.dSW; st.LSE = link & h.dSE;
  bool diag = (st.LNW | st.LNE | st.LSW | st.LSE) != 0u;
  // which of the 9 words are needed; all `calculated` words first, ONE fence, then their my/mx
  bool need[9];
  need[4] = true;
  need[3] = (st.LW & 1u) || diag; need[5] = (st.LE >> 31) || diag;
  need[1] = (st.LN | st.LNW | st.LNE) != 0u; need[7] = (st.LS | st.LSW | st.LSE) != 0u;
  need[0] = need[2] = need[6] = need[8] = diag;
  UPlanes nb[9];
  size_t off[9];
  bool ok[9];
  for (int k = 0; k < 9; k++) {
    nb[k].R = nb[k].MY = nb[k].MX = 0u;
    ok[k] = false; off[k] = 0;
    if (need[k]) { off[k] = u_off(g, y + k / 3 - 1, w + k % 3 - 1, ok[k]); nb[k].R = u_ld_calc(g, off[k], ok[k]); }
  }
  fa_fence_device();
  for (int k = 0; k < 9; k++)
    if (need[k]) u_ld_mymx(g, off[k], ok[k], nb[k]);
  const UPlanes &ul = nb[0], &up = nb[1], &ur = nb[2], &l = nb[3], &c = nb[4], &r = nb[5], &dl = nb[6], &dn = nb[7], &dr = nb[8];
  W3 R_up = {ul.R, up.R, ur.R}, R_c = {l.R, c.R, r.R}, R_dn = {dl.R, dn.R, dr.R};
  W3 MY_up = {ul.MY, up.MY, ur.MY}, MY_dn = {dl.MY, dn.MY, dr.MY};
  W3 MX_c = {l.MX, c.MX, r.MX}, MX_up = {ul.MX, up.MX, ur.MX}, MX_dn = {dl.MX, dn.MX, dr.MX};
  const u32 valid = edge_of(g, y, w).valid;
  UState n = u_relax(st, R_up, R_c, R_dn, MY_up, MY_dn, MX_c, MX_up, MX_dn);
  n.R = (n.R & valid) | c.R; n.MY = (n.MY & valid) | c.MY; n.MX = (n.MX & valid) | c.MX;   // monotone
  bool changed = n.R != c.R || n.MY != c.MY || n.MX != c.MX;
  bool unready = (~n.R & ~e.T) != 0u;
  int np = 0;
  if (changed) {
    FA_ATOMIC_OR_U64(&g.mymx[e.wi], pack_mymx(n.MY, n.MX));
    fa_fence_device();
    FA_ATOMIC_OR_U32(&g.calc[e.wi], n.R);
    fa_fence_device();
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        if (dy == 0 && dx == 0 && !unready) continue;
        u32 q = u_try_queue(g, y + dy, w + dx);
        if (q != U_NONE) push[np++] = q & ~U_QUEUED;
      }
  }
  // resolved for good: retire the word -- with a CAS, because a neighbour may be queueing it right now (then it
  // stays queued, gets one idle visit next iteration and is retired there)
  if (!unready) FA_ATOMIC_CAS_U32(&g.ustate[e.wi], idx, U_NONE);
  return np;
}

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
  FA_HD MTile(const GridView &g_, Smem &s_, int tile_y, int tile_w) : g(g_), s(s_), y0(g_.own0 + tile_y * TR), w0(tile_w * TW) {}
  FA_HD static int idx(int r, int c) { return r * SW + c; }

  FA_HD void p0(int tid, int nt) {
    if (tid == 0) s.leaks = 0;
    for (int i = tid; i < NW; i += nt) {
      int r = i / SW, c = i - r * SW;
      int y = y0 - 1 + r, w = w0 - 1 + c;
      U4 rd = {0, 0, 0, 0}, hd = {0, 0, 0, 0};
      u32 my = 0, mx = 0;
      if (row_in(g, y) && w >= 0 && w < g.pitch) {
        size_t o = (size_t)y * g.pitch + w;
        rd = g.road[o]; hd = g.head_in[o];
        u64 m = g.mymx[o];
        my = (u32)m; mx = (u32)(m >> 32);
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
      if (y >= g.own1 || w >= g.pitch) { s.AY[i] = 0; s.AX[i] = 0; continue; }
      size_t go = (size_t)y * g.pitch + w;
      {  // next_preference_flow_x (steppers.fut:41-50): decided by the fixpoint's own grants only; cells a ring
         // resolution marked (steppers.fut:189-224) have calc = 0 and keep their preference
        int ci = idx(r, c);
        u32 cal = g.calc[go];
        UState f = {cal, s.MY[ci] & cal, s.MX[ci] & cal};
        UStatic dummy = {};
        Edge ev = edge_of(g, y, w);
        g.pref_out[go] = u_pref_next(s.UYN[ci] | s.UXN[ci], g.pref_in[go], f, dummy) & ev.valid;
      }
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
      Edge e = edge_of(g, y, w);
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
      if (y >= g.own1 || w >= g.pitch) continue;
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
