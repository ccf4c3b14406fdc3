// flowamok_b200 -- device-resident grid view and the worklist protocol of the U fixpoint, written once for device
// and host.
//
// The streaming strip programs (fa_stream.h) and the frontier visit below are plain functions of a GridView; the CUDA
// kernels (fa_kernels.cuh) call them per thread, the CPU emulation used by tests/emul runs the same functions with
// loops (and coroutines for the warp programs) in place of the hardware threads.  So the bit logic, the indexing and
// the worklist protocol are all unit-tested on the CPU against the oracle before they ever reach a GPU.
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
  int gh_global;     // rows of the whole grid (a band is recognised by its halo rows, never by gh != gh_global:
                     // a band's rows plus its halo rows can add up to exactly the grid's height)
  int row0;          // global row of local row 0 (0 on a single GPU; -halo for the first band)
  int own0, own1;    // local rows [own0, own1) are owned (stepped and written) by this device
  const U4 *road;    // {UYN, UYS, UXN, UXS} per word                       [gh][pitch]
  const U4 *head_in; // {DYN, DYS, DXN, DXS} per word, state before the tick [gh][pitch]
  U4 *head_out;      // state after the tick
  const u32 *pref_in;
  u32 *pref_out;
  u32 *calc;               // per-tick plane `calculated` (steppers.fut:140-145 resets it: here it is scratch)
  u64 *mymx;               // per-tick planes direction.y (low word) / direction.x (high word), one 64-bit access
  // Speculative move (fa_stream.h, fix_word): with mymx_B set, the global fixpoint leaves `mymx` -- as k_u_local, ring
  // resolution and a band's first boundary swap wrote it -- alone and ORs its grants into the delta plane mymx_B (zero
  // between ticks); k_move gathers from `mymx` BESIDE the fixpoint, and every word whose delta becomes non-zero is listed once
  // in chg_list for the fix-up.  The final planes are mymx | mymx_B.  All three null: the kernels run one after the other.
  u64 *mymx_B;
  u32 *chg_list;
  unsigned int *chg_n;
  u32 *ustate;             // per word: U_NONE, or its worklist entry index (| U_QUEUED while on the next list)
  // colour payload [gh][pitch*32], LAZY double buffer (fa_stream.h, MTile): color_out already equals color_in
  // everywhere except in the 32-byte sectors that received a car during the previous tick (dirty_in: one byte per plane
  // word, bit k = sector k); k_move records this tick's in dirty_out
  const u32 *color_in;
  u32 *color_out;
  const unsigned char *dirty_in;
  unsigned char *dirty_out;
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
//   UStrip (fa_stream.h, k_u_local): every word once: statics, terminals, K Jacobi levels, and a worklist entry for
//                 each word that still holds an unresolved waiting car ("link" cell: occupied, successor in
//                 bounds and on a road).
//   u_iter_entry: one monotone relaxation of a worklist entry against the CURRENT global planes
//                 (k_u_iter, frontier discipline).
//
// Only link cells can change after initialisation, so the iterations touch nothing else.  Confluence
//  makes any relaxation order reach the reference's Jacobi result.  A word's state is
// published my/mx first, calc last, and read calc first: whoever sees calc = 1 sees the final my/mx.
// =============================================================================================
// CTA = true: every thread that touches the fixpoint planes right now belongs to this CTA (the single-CTA tail of
// k_u_iter), so a block-scope fence is enough -- it costs a fraction of the device-scope one (no L1 invalidation)
template <bool CTA>
FA_HD void fa_fence() {
#if defined(__CUDA_ARCH__)
  if (CTA) __threadfence_block();
  else __threadfence();
#endif
}
FA_HD void fa_fence_device() { fa_fence<false>(); }
FA_HD u64 pack_mymx(u32 my, u32 mx) { return (u64)my | ((u64)mx << 32); }
// The fixpoint (a visit, or a band's merge of the neighbour's row) grants `add` in word o whose my|mx currently reads `cur`:
// into the plane itself, or -- speculative move -- into the delta plane; the first delta of a word lists it for the fix-up.
FA_HD void u_grant(const GridView &g, size_t o, u64 cur, u64 add) {
  if (!g.mymx_B) { FA_ATOMIC_OR_U64(&g.mymx[o], add); return; }
  const u64 d = add & ~cur;
  if (d == 0ull) return;
  const u64 old = FA_ATOMIC_OR_U64(&g.mymx_B[o], d);
  if (old == 0ull) g.chg_list[FA_ATOMIC_ADD_U32(g.chg_n, 1u)] = (u32)o;
}

constexpr u32 U_NONE = 0xFFFFFFFFu;    // word has no unresolved waiting car
constexpr u32 U_QUEUED = 0x80000000u;  // word is already on the next iteration's list

struct alignas(16) UEntry {  // 16 bytes: one 128-bit access
  u32 wi;        // plane word index
  u32 T, GY, GX; // statics of the word (UStatic.T, GYr, GXr)
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
    if (g.mymx_B) m |= *(volatile const u64 *)&g.mymx_B[o];
#else
    m = g.mymx[o];
    if (g.mymx_B) m |= g.mymx_B[o];
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

// Row bands: the state of ghost word wg (the neighbour band's boundary row; side 0: the row above the band, side 1:
// the row below) changed at the cells `chg`.  Does owned word wg + dx of the adjacent owned row hold a waiting car that
// reads one of them?  Only cars heading towards the ghost row do: straight ones read the same bit of the same word,
// diagonal ones the bit beside it (possibly in the next word).  Everything else in the row is unaffected by the
// neighbour band -- waking it anyway would cost every tick one more convergence round of the band protocol.
FA_HD bool band_word_reads_ghost(const GridView &g, const UEntry *table, int side, int own_row, int wg, int dx, u32 chg) {
  const int w = wg + dx;
  if (w < 0 || w >= g.pitch) return false;
  const size_t o = (size_t)own_row * g.pitch + w;
#if defined(__CUDA_ARCH__)
  const u32 sv = *(volatile const u32 *)&g.ustate[o];
#else
  const u32 sv = g.ustate[o];
#endif
  if (sv == U_NONE) return false;   // no unresolved waiting car in the word
  const U4 hd = g.head_in[o];
  const Heading h = heading(hd.x, hd.y, hd.z, hd.w);
  const u32 link = ~table[o].T;
  const u32 straight = link & (side == 0 ? h.cN : h.cS);
  const u32 dl = link & (side == 0 ? h.dNW : h.dSW);   // reads the ghost cell at x - 1
  const u32 dr = link & (side == 0 ? h.dNE : h.dSE);   // reads the ghost cell at x + 1
  if (dx == 0) return ((straight & chg) | (dl & (chg << 1)) | (dr & (chg >> 1))) != 0u;
  if (dx == 1) return (dl & 1u & (chg >> 31)) != 0u;    // my first cell looks at the ghost word's last cell
  return ((dr >> 31) & chg & 1u) != 0u;                 // my last cell looks at the ghost word's first cell
}

// One relaxation of worklist entry `e` against the current planes (frontier discipline):
//   clear QUEUED, re-read the neighbourhood, OR the new facts in (my/mx before calc);
//   if anything changed, wake the neighbouring words whose waiting cars may point at a changed cell (the words above
//   and below; the words to the left / right and the corner words only if the changed cell is the word's first /
//   last one) and this word itself while it still holds an uncalculated link cell;
//   a word nobody changes is left dormant -- cars queued behind a full ring cost one visit per tick.
// The memory operations are grouped so that independent round trips to L2 overlap: a visit is a chain of about eight
// of them, and the fixpoint's tail is bound by exactly that latency.
// push[] receives up to 9 entry indices for the next list; returns their number.
template <bool CTA = false>
FA_HD int u_iter_entry(const GridView &g, const UEntry &e, u32 idx, u32 *push) {
  const int y = (int)(e.wi / (u32)g.pitch), w = (int)(e.wi - (u32)y * (u32)g.pitch);
  const U4 hd = g.head_in[e.wi];   // static during the fixpoint: no ordering needed
  const u32 pref = g.pref_in[e.wi];
  FA_ATOMIC_AND_U32(&g.ustate[e.wi], ~U_QUEUED);
  fa_fence<CTA>();
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
  fa_fence<CTA>();
  for (int k = 0; k < 9; k++)
    if (need[k]) u_ld_mymx(g, off[k], ok[k], nb[k]);
  const UPlanes &ul = nb[0], &up = nb[1], &ur = nb[2], &l = nb[3], &c = nb[4], &r = nb[5], &dl = nb[6], &dn = nb[7], &dr = nb[8];
  W3 R_up = {ul.R, up.R, ur.R}, R_c = {l.R, c.R, r.R}, R_dn = {dl.R, dn.R, dr.R};
  W3 MY_up = {ul.MY, up.MY, ur.MY}, MY_dn = {dl.MY, dn.MY, dr.MY};
  W3 MX_c = {l.MX, c.MX, r.MX}, MX_up = {ul.MX, up.MX, ur.MX}, MX_dn = {dl.MX, dn.MX, dr.MX};
  const u32 valid = edge_of(g, y, w).valid;
  UState n = u_relax(st, R_up, R_c, R_dn, MY_up, MY_dn, MX_c, MX_up, MX_dn);
  n.R = (n.R & valid) | c.R; n.MY = (n.MY & valid) | c.MY; n.MX = (n.MX & valid) | c.MX;   // monotone
  // cars queued along x inside this word wait on cells of the word itself: relax again on the word's own new values
  // (a few ALU operations) instead of one more grid-wide iteration per car
  for (int again = 0; again < 4 && (n.R != R_c.c || n.MX != MX_c.c); again++) {
    R_c.c = n.R; MX_c.c = n.MX;
    const UState m = u_relax(st, R_up, R_c, R_dn, MY_up, MY_dn, MX_c, MX_up, MX_dn);
    n.R |= m.R & valid; n.MY |= m.MY & valid; n.MX |= m.MX & valid;
  }
  const u32 chg = (n.R ^ c.R) | (n.MY ^ c.MY) | (n.MX ^ c.MX);
  bool unready = (~n.R & ~e.T) != 0u;
  int np = 0;
  if (chg) {
    // next_preference_flow_x follows the grants (k_u_local wrote it for the grants it knew); one visitor per word
    // (grants of the FIXPOINT only, i.e. of calculated cells: ring resolution may already have put its grants into my|mx,
    // and it never touches the preference, steppers.fut:205-211)
    g.pref_out[e.wi] = u_pref_from_grants(e.GY, e.GX, pref, n.MY & n.R, n.MX & n.R) & valid;
    u_grant(g, e.wi, pack_mymx(c.MY, c.MX), pack_mymx(n.MY, n.MX));
    fa_fence<CTA>();
    FA_ATOMIC_OR_U32(&g.calc[e.wi], n.R);
    fa_fence<CTA>();
    // whom to wake: k = 3 * (dy + 1) + (dx + 1)
    const bool first = (chg & 1u) != 0u, last = (chg >> 31) != 0u;
    bool cand[9] = {first, true, last, first, unready, last, first, true, last};
    u32 sv[9];
    for (int k = 0; k < 9; k++) {
      const int yy = y + k / 3 - 1, ww = w + k % 3 - 1;
      cand[k] = cand[k] && yy >= g.own0 && yy < g.own1 && ww >= 0 && ww < g.pitch;   // only owned words have entries
      sv[k] = U_NONE;
      if (cand[k]) {
#if defined(__CUDA_ARCH__)
        sv[k] = *(volatile const u32 *)&g.ustate[(size_t)yy * g.pitch + ww];
#else
        sv[k] = g.ustate[(size_t)yy * g.pitch + ww];
#endif
      }
    }
    for (int k = 0; k < 9; k++) {   // claim: whoever flips QUEUED from 0 to 1 puts the word on the list
      cand[k] = cand[k] && sv[k] != U_NONE && !(sv[k] & U_QUEUED);
      if (cand[k]) sv[k] = FA_ATOMIC_OR_U32(&g.ustate[(size_t)(y + k / 3 - 1) * g.pitch + (w + k % 3 - 1)], U_QUEUED);
    }
    for (int k = 0; k < 9; k++)
      if (cand[k] && sv[k] != U_NONE && !(sv[k] & U_QUEUED)) push[np++] = sv[k];
  }
  // resolved for good: retire the word -- with a CAS, because a neighbour may be queueing it right now (then it
  // stays queued, gets one idle visit next iteration and is retired there)
  if (!unready) FA_ATOMIC_CAS_U32(&g.ustate[e.wi], idx, U_NONE);
  return np;
}

}  // namespace fa
