// CPU emulation of the CUDA kernels' programs (TEST INFRASTRUCTURE).
//
// Compiles flowamok_b200/csrc/fa_stream.h and fa_tiles.h for the host and runs the very same functions the
// kernels run: the warp strip programs of k_u_local / k_move with the 32 lanes of a warp as coroutines
// (simt_host.h), the frontier visit of k_u_iter with a loop over the list, and the same row-band protocol
// as the flowamok_band_* calls of fa_api.cu.  tests/test_emul.py and tests/test_bands_cpu.py compare it
// with the oracle.  It is never part of the product library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../flowamok_b200/csrc/fa_stream.h"
#include "simt_host.h"

using namespace fa;

namespace {

struct Rings {
  std::vector<int64_t> off, y, x;   // y is GLOBAL
  std::vector<int8_t> cy, cx;
  std::vector<uint8_t> grant;
};

struct HostGrid {
  int gh, gw, pitch, wvalid, gh_global, row0, own0, own1, halo;
  int variant;
  std::vector<U4> road, head[2];
  std::vector<u32> pref[2], calc, color[2], ustate;
  std::vector<unsigned char> dirty[2];
  std::vector<u64> mymx, rng;
  u64 rng_inc = 0;
  int cur = 0;
  // fixpoint state of the current tick
  std::vector<UEntry> table;
  std::vector<u32> frontier;
  unsigned iters = 0, relaxed = 0;
  Rings rings;
  bool has_rings = false;
  std::vector<u32> rbits;   // verdicts of this band on the rings that straddle its cuts
  // speculative move (fa_stream.h fix_word): the fixpoint's delta plane, the words that have a delta, fix-up claims
  std::vector<u64> mymx_B;
  std::vector<u32> chg_list, claim;
  unsigned int chg_n = 0;
  u32 tick = 0;
  int64_t leaks_spec = 0;
};

HostGrid *make(int gh_global, int gw, int row_begin, int row_end, int halo, int variant) {
  HostGrid *h = new HostGrid();
  h->gh_global = gh_global; h->gw = gw; h->halo = halo; h->variant = variant;
  h->row0 = row_begin - halo; h->own0 = halo; h->own1 = halo + (row_end - row_begin);
  h->gh = h->own1 + halo;
  h->wvalid = (gw + 31) / 32;
  h->pitch = ((h->wvalid + 3) / 4) * 4;
  size_t nw = (size_t)h->gh * h->pitch, nc = nw * 32;
  h->road.assign(nw, U4{0, 0, 0, 0});
  for (int k = 0; k < 2; k++) { h->head[k].assign(nw, U4{0, 0, 0, 0}); h->pref[k].assign(nw, 0); h->color[k].assign(nc, 0); h->dirty[k].assign(nw, 0); }
  h->calc.assign(nw, 0); h->mymx.assign(nw, 0); h->ustate.assign(nw, U_NONE);
  h->mymx_B.assign(nw, 0); h->chg_list.assign(nw, 0); h->claim.assign(nw, 0);
  h->rng.assign(nc, 0);
  return h;
}

// spec: 0 = what runs before the speculative move reads the plane (or a tick without one), 2 = what runs after: grants go
// into the delta plane, changed words are listed
GridView view(HostGrid *h, int chooser, u32 tape_bit, int spec = 0) {
  GridView g;
  g.mymx_B = spec == 2 ? h->mymx_B.data() : nullptr;
  g.chg_list = spec == 2 ? h->chg_list.data() : nullptr;
  g.chg_n = spec == 2 ? &h->chg_n : nullptr;
  g.gh = h->gh; g.gw = h->gw; g.pitch = h->pitch; g.wvalid = h->wvalid;
  g.gh_global = h->gh_global; g.row0 = h->row0; g.own0 = h->own0; g.own1 = h->own1;
  g.road = h->road.data();
  g.head_in = h->head[h->cur].data(); g.head_out = h->head[h->cur ^ 1].data();
  g.pref_in = h->pref[h->cur].data(); g.pref_out = h->pref[h->cur ^ 1].data();
  g.calc = h->calc.data(); g.mymx = h->mymx.data(); g.ustate = h->ustate.data();
  g.color_in = h->color[h->cur].data(); g.color_out = h->color[h->cur ^ 1].data();
  g.dirty_in = h->dirty[h->cur].data(); g.dirty_out = h->dirty[h->cur ^ 1].data();
  g.rng = h->rng.data(); g.rng_inc = h->rng_inc;
  g.chooser = chooser; g.tape_bit = tape_bit;
  return g;
}

// k_u_local: one warp (32 coroutines) per strip
template <int SW, int K>
struct UArgs { HostGrid *h; const GridView *g; int sx, ya, yb; unsigned int qn; std::vector<u32> *ql, *qs; };
template <int SW, int K>
void u_lane(void *p, int lane) {
  UArgs<SW, K> *a = (UArgs<SW, K> *)p;
  if (a->h->halo != 0) UStrip<SW, K, true>::run(*a->g, a->sx, a->ya, a->yb, lane, a->h->table.data(), a->qs->data(), a->ql->data(), &a->qn);
  else UStrip<SW, K, false>::run(*a->g, a->sx, a->ya, a->yb, lane, a->h->table.data(), a->qs->data(), a->ql->data(), &a->qn);
}
template <int SW, int K>
void u_init(HostGrid *h, const GridView &g, int rs) {
  const size_t nw = (size_t)h->gh * h->pitch;
  h->table.resize(nw);
  h->frontier.clear();
  std::vector<u32> ql(nw), qs((size_t)SW * rs);
  const int nsx = (h->pitch + SW - 1) / SW, nsy = (h->own1 - h->own0 + rs - 1) / rs;
  for (int sy = 0; sy < nsy; sy++)
    for (int sx = 0; sx < nsx; sx++) {
      UArgs<SW, K> a = {h, &g, sx, h->own0 + sy * rs, std::min(h->own1, h->own0 + (sy + 1) * rs), 0u, &ql, &qs};
      simt_host::run_warp(u_lane<SW, K>, &a);
      for (unsigned k = 0; k < a.qn; k++) h->frontier.push_back(ql[k]);
    }
}

// k_u_iter: frontier iterations until nothing is queued; `order` varies the visiting order
void u_iter(HostGrid *h, const GridView &g, int order) {
  std::vector<u32> next;
  while (!h->frontier.empty()) {
    h->iters++;
    next.clear();
    size_t n = h->frontier.size();
    for (size_t k = 0; k < n; k++) {
      size_t kk = order == 2 ? n - 1 - k : (order == 1 && n % 7919u != 0 ? (k * 7919u) % n : k);
      u32 push[9];
      int np = u_iter_entry(g, h->table[h->frontier[kk]], h->frontier[kk], push);
      h->relaxed++;
      for (int j = 0; j < np; j++) next.push_back(push[j]);
    }
    h->frontier.swap(next);
  }
}

// steppers.fut:189-224, sparse form.  The ring list is GLOBAL; a grid judges and marks the cells of its own
// rows.  kind: 0 = judge every ring; rings wholly inside the owned rows are marked at once, rings that straddle a band
// cut only get their verdict bit cleared on an objection (h->rbits, bit k = ring k, 1 = no objection);
// 1 = mark the straddling rings whose bit is set in `pass`.
void do_rings(HostGrid *h, const GridView &g, int kind = 0, const u32 *pass = nullptr) {
  if (!h->has_rings) return;
  const Rings &R = h->rings;
  const size_t n = R.off.size() - 1;
  const int64_t lo = h->row0 + h->own0, hi = h->row0 + h->own1;
  if (h->rbits.size() != (n + 31) / 32) h->rbits.assign((n + 31) / 32, 0xFFFFFFFFu);
  for (size_t k = 0; k < n; k++) {
    bool in = false, out = false, ok = true;
    for (int64_t t = R.off[k]; t < R.off[k + 1]; t++) {
      if (R.y[t] < lo || R.y[t] >= hi) { out = true; continue; }
      in = true;
      size_t o = (size_t)(R.y[t] - h->row0) * h->pitch + (R.x[t] >> 5);
      u32 bit = 1u << (R.x[t] & 31);
      U4 hd = g.head_in[o];
      int dy = (hd.x & bit) ? ((hd.y & bit) ? -1 : 1) : 0, dx = (hd.z & bit) ? ((hd.w & bit) ? -1 : 1) : 0;
      if ((g.calc[o] & bit) || dy != R.cy[t] || dx != R.cx[t]) ok = false;
    }
    if (!in) continue;
    bool mark;
    if (kind == 0) {
      if (out && !ok) h->rbits[k >> 5] &= ~(1u << (k & 31));
      mark = ok && !out;
    } else {
      mark = out && ((pass[k >> 5] >> (k & 31)) & 1u);
    }
    if (!mark) continue;
    for (int64_t t = R.off[k]; t < R.off[k + 1]; t++) {
      if (R.y[t] < lo || R.y[t] >= hi) continue;
      size_t o = (size_t)(R.y[t] - h->row0) * h->pitch + (R.x[t] >> 5);
      u32 bit = 1u << (R.x[t] & 31);
      const u64 add = R.grant[t] == 1 ? (u64)bit : (u64)bit << 32;
      u_grant(g, o, g.mymx[o] | (g.mymx_B ? g.mymx_B[o] : 0ull), add);   // plane, or delta plane + fix-up list
    }
  }
}

// fix-up of the speculative move: every item of the list, as k_move_fix does
int64_t do_fix(HostGrid *h, const GridView &g) {
  int64_t leaks = 0;
  for (unsigned int t = 0; t < h->chg_n * 9u; t++) leaks += fix_item(g, h->claim.data(), h->tick + 1, t);
  for (unsigned int k = 0; k < h->chg_n; k++) h->mymx_B[h->chg_list[k]] = 0ull;   // k_u_iter's prologue of the next tick
  h->chg_n = 0;
  return leaks;
}

template <int SW, int NW, int CAP>
struct MArgs { const GridView *g; bool has; int sx, wi, ya, yb; typename MTile<SW, NW, CAP>::Smem sm; unsigned int leaks[32]; };
template <int SW, int NW, int CAP>
void m_lane(void *p, int lane) {
  MArgs<SW, NW, CAP> *a = (MArgs<SW, NW, CAP> *)p;
  a->leaks[lane] = MTile<SW, NW, CAP>::phaseA(*a->g, a->has, a->sx, a->wi, a->ya, a->yb, lane, nullptr, a->sm);
}
// k_move: one "CTA" per tile: phase A warp by warp (coroutines), phase B with a loop over the thread ids.
// flip = false: the speculative pass (the buffers are flipped after the fix-up)
template <int SW, int NW, int CAP>
int64_t do_move(HostGrid *h, const GridView &g, int rs, int nt, bool flip) {
  typedef MTile<SW, NW, CAP> MTL;
  int64_t leaks = 0;
  const int nsx = (h->pitch + SW - 1) / SW, ntx = (nsx + NW - 1) / NW, nty = (h->own1 - h->own0 + rs - 1) / rs;
  std::vector<u32> smem(MTL::Smem::bytes(rs) / 4 + 4);
  u32 *base = (u32 *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
  for (int ty = 0; ty < nty; ty++)
    for (int tx = 0; tx < ntx; tx++) {
      const typename MTL::Smem sm = {base, rs};
      const int ya = h->own0 + ty * rs, yb = std::min(h->own1, ya + rs);
      std::fill(smem.begin(), smem.end(), 0xDEADBEEFu);   // shared memory starts undefined
      for (int tid = 0; tid < nt; tid++) MTL::phaseB_clear(sm, tid);
      for (int wi = 0; wi < NW; wi++) {
        MArgs<SW, NW, CAP> a = {&g, tx * NW + wi < nsx, tx * NW + wi, wi, ya, yb, sm, {}};
        simt_host::run_warp(m_lane<SW, NW, CAP>, &a);
        for (int l = 0; l < 32; l++) leaks += a.leaks[l];
      }
      for (int tid = 0; tid < nt; tid++) MTL::phaseB0(sm, yb - ya, tid, nt);
      for (int tid = 0; tid < nt; tid++) MTL::phaseB1(g, sm, tx * NW, ya, tid, nt);
      for (int tid = 0; tid < nt; tid++) MTL::phaseB2(g, sm, tx * NW, ya, yb - ya, tid, nt);
    }
  if (flip) h->cur ^= 1;
  return leaks;
}

// variant: 0 = production strip shape, 1 = narrow strips / short rows + reversed frontier order,
//          2 = odd strip shape + scrambled order
void v_u_init(HostGrid *h, const GridView &g) {
  switch (h->variant) {
  case 1: u_init<1, 1>(h, g, 2); break;
  case 2: u_init<3, 2>(h, g, 5); break;
  default: u_init<30, 3>(h, g, 32); break;
  }
}
int v_order(const HostGrid *h) { return h->variant == 1 ? 2 : h->variant == 2 ? 1 : 0; }
int64_t v_move(HostGrid *h, const GridView &g, bool flip) {
  switch (h->variant) {
  case 1: return do_move<2, 3, 8>(h, g, 3, 6, flip);      // tiny arrival list: the overflow path runs
  case 2: return do_move<5, 2, 64>(h, g, 4, 64, flip);
  default: return do_move<30, 4, 2048>(h, g, 16, 128, flip);
  }
}
// the speculative pass: k_move gathering from the plane as it is now (the fixpoint's later grants go into the delta plane)
int64_t v_move_spec(HostGrid *h, int chooser, u32 tape_bit) {
  GridView gs = view(h, chooser, tape_bit, 0);
  return v_move(h, gs, false);
}

}  // namespace

extern "C" {

void *emul_new(int gh, int gw) { return make(gh, gw, 0, gh, 0, 0); }
void *emul_band_new(int gh_global, int gw, int row_begin, int row_end, int variant) {
  return make(gh_global, gw, row_begin, row_end, 2, variant);
}
void emul_free(void *p) { delete (HostGrid *)p; }

// host arrays are WHOLE-grid [gh_global][gw]; the rows this grid holds (band + halo) are taken
void emul_upload(void *p, const int8_t *uy, const int8_t *ux, const int8_t *dy, const int8_t *dx,
                 const uint32_t *color, const uint8_t *pref, const uint64_t *rng_state, uint64_t rng_inc) {
  HostGrid *h = (HostGrid *)p;
  size_t cp = (size_t)h->pitch * 32;
  for (auto &v : h->road) v = U4{0, 0, 0, 0};
  for (auto &v : h->head[h->cur]) v = U4{0, 0, 0, 0};
  for (auto &v : h->pref[h->cur]) v = 0;
  for (int yl = 0; yl < h->gh; yl++) {
    int y = yl + h->row0;
    if (y < 0 || y >= h->gh_global) continue;
    for (int x = 0; x < h->gw; x++) {
      size_t i = (size_t)y * h->gw + x, o = (size_t)yl * h->pitch + (x >> 5);
      u32 bit = 1u << (x & 31);
      if (uy[i]) { h->road[o].x |= bit; if (uy[i] < 0) h->road[o].y |= bit; }
      if (ux[i]) { h->road[o].z |= bit; if (ux[i] < 0) h->road[o].w |= bit; }
      U4 &hd = h->head[h->cur][o];
      if (dy[i]) { hd.x |= bit; if (dy[i] < 0) hd.y |= bit; }
      if (dx[i]) { hd.z |= bit; if (dx[i] < 0) hd.w |= bit; }
      if (pref[i]) h->pref[h->cur][o] |= bit;
      h->color[h->cur][(size_t)yl * cp + x] = color[i];
      h->rng[(size_t)yl * cp + x] = rng_state[i];
    }
  }
  h->rng_inc = rng_inc;
  h->color[h->cur ^ 1] = h->color[h->cur];   // lazy double buffer: both identical, nothing dirty
  for (int k = 0; k < 2; k++) std::fill(h->dirty[k].begin(), h->dirty[k].end(), 0);
}

// writes the OWNED rows into whole-grid arrays
void emul_download(void *p, int8_t *dy, int8_t *dx, uint32_t *color, uint8_t *pref, uint64_t *rng_state) {
  HostGrid *h = (HostGrid *)p;
  size_t cp = (size_t)h->pitch * 32;
  for (int yl = h->own0; yl < h->own1; yl++) {
    int y = yl + h->row0;
    for (int x = 0; x < h->gw; x++) {
      size_t i = (size_t)y * h->gw + x, o = (size_t)yl * h->pitch + (x >> 5);
      u32 bit = 1u << (x & 31);
      const U4 &hd = h->head[h->cur][o];
      dy[i] = (hd.x & bit) ? ((hd.y & bit) ? -1 : 1) : 0;
      dx[i] = (hd.z & bit) ? ((hd.w & bit) ? -1 : 1) : 0;
      pref[i] = (h->pref[h->cur][o] & bit) ? 1 : 0;
      color[i] = h->color[h->cur][(size_t)yl * cp + x];
      rng_state[i] = h->rng[(size_t)yl * cp + x];
    }
  }
}

void emul_set_rings(void *p, int64_t n_rings, const int64_t *off, const int64_t *ry, const int64_t *rx, const int8_t *cy,
                    const int8_t *cx, const uint8_t *grant) {
  HostGrid *h = (HostGrid *)p;
  h->has_rings = n_rings >= 0 && off;
  if (!h->has_rings) return;
  Rings &R = h->rings;
  int64_t tot = off[n_rings];
  R.off.assign(off, off + n_rings + 1);
  R.y.assign(ry, ry + tot); R.x.assign(rx, rx + tot);
  R.cy.assign(cy, cy + tot); R.cx.assign(cx, cx + tot); R.grant.assign(grant, grant + tot);
}

// one whole tick of a single (non-band) grid
int64_t emul_step(void *p, int variant, int nthreads, int chooser, uint32_t tape_bit, int64_t n_rings,
                  const int64_t *off, const int64_t *ry, const int64_t *rx, const int8_t *cy,
                  const int8_t *cx, const uint8_t *grant, unsigned *stats) {
  HostGrid *h = (HostGrid *)p;
  h->variant = variant;
  emul_set_rings(p, n_rings, off, ry, rx, cy, cx, grant);
  h->iters = h->relaxed = 0;
  h->chg_n = 0;
  int64_t leaks;
  if (variant == 2 && n_rings < 0) {   // the kernels one after the other (what flowamok_step_profile / _step_leaks run)
    GridView g = view(h, chooser, tape_bit, 0);
    v_u_init(h, g);
    u_iter(h, g, v_order(h));
    leaks = v_move(h, g, true);
  } else {   // the production order: snapshot writers, speculative move, fixpoint with listing, fix-up
    GridView g1 = view(h, chooser, tape_bit, 0), g2 = view(h, chooser, tape_bit, 2);
    v_u_init(h, g1);
    do_rings(h, g1);                               // ring resolution does not depend on the fixpoint: before the snapshot is read
    leaks = v_move_spec(h, chooser, tape_bit);     // on the GPU: beside k_u_iter
    u_iter(h, g2, v_order(h));
    leaks += do_fix(h, g2);
    h->cur ^= 1;
  }
  h->tick++;
  (void)nthreads;
  if (stats) { stats[0] += h->iters; stats[1] += h->relaxed; }
  return leaks;
}

// ---- row-band protocol: same messages and phases as flowamok_band_* (fa_api.cu) ----
void emul_band_msg_bytes(void *p, int64_t *state_bytes, int64_t *u_bytes) {
  HostGrid *h = (HostGrid *)p;
  *state_bytes = (int64_t)h->pitch * (2 * 16 + 2 * 4 + 128);
  *u_bytes = (int64_t)h->pitch * (4 + 8);
}
void emul_band_pack_state(void *p, int side, void *buf) {
  HostGrid *h = (HostGrid *)p;
  const size_t P = (size_t)h->pitch;
  const int r2 = side == 0 ? h->own0 : h->own1 - 2, rc = side == 0 ? h->own0 : h->own1 - 1;
  char *b = (char *)buf;
  memcpy(b, h->head[h->cur].data() + (size_t)r2 * P, 2 * P * 16);
  memcpy(b + 2 * P * 16, h->pref[h->cur].data() + (size_t)r2 * P, 2 * P * 4);
  memcpy(b + 2 * P * 20, h->color[h->cur].data() + (size_t)rc * P * 32, P * 128);
}
void emul_band_unpack_state(void *p, int side, const void *buf) {
  HostGrid *h = (HostGrid *)p;
  const size_t P = (size_t)h->pitch;
  const int r2 = side == 0 ? h->own0 - 2 : h->own1, rc = side == 0 ? h->own0 - 1 : h->own1;
  const char *b = (const char *)buf;
  memcpy(h->head[h->cur].data() + (size_t)r2 * P, b, 2 * P * 16);
  memcpy(h->pref[h->cur].data() + (size_t)r2 * P, b + 2 * P * 16, 2 * P * 4);
  memcpy(h->color[h->cur].data() + (size_t)rc * P * 32, b + 2 * P * 20, P * 128);
}
// k_u_local, and at once the speculative k_move on the snapshot it leaves: everything the band protocol does afterwards
// (merges of the neighbours' rows, frontier iterations, ring resolution) lists the words it changes, and emul_band_move
// is the fix-up.  Stricter than the GPU path, which reads the snapshot after the first boundary swap.
void emul_band_u_init(void *p, int chooser, uint32_t tape_bit) {
  HostGrid *h = (HostGrid *)p;
  GridView g = view(h, chooser, tape_bit, 0);
  v_u_init(h, g);
  h->leaks_spec = v_move_spec(h, chooser, tape_bit);
}
void emul_band_pack_u(void *p, int side, void *buf) {
  HostGrid *h = (HostGrid *)p;
  const size_t P = (size_t)h->pitch;
  const int r = side == 0 ? h->own0 : h->own1 - 1;
  char *b = (char *)buf;
  memcpy(b, h->calc.data() + (size_t)r * P, P * 4);
  u64 *dm = (u64 *)(b + P * 4);
  for (size_t w = 0; w < P; w++) dm[w] = h->mymx[(size_t)r * P + w] | h->mymx_B[(size_t)r * P + w];
}
int64_t emul_band_merge_u(void *p, int side, const void *buf, int wake) {  // k_band_merge / band_exchange_u
  HostGrid *h = (HostGrid *)p;
  GridView g = view(h, 0, 0, 2);
  const size_t P = (size_t)h->pitch;
  const int ghost = side == 0 ? h->own0 - 1 : h->own1, own = side == 0 ? h->own0 : h->own1 - 1;
  const u32 *rc = (const u32 *)buf;
  const u64 *rm = (const u64 *)((const char *)buf + P * 4);
  for (int w = 0; w < h->pitch; w++) {
    size_t o = (size_t)ghost * P + w;
    const u64 om = h->mymx[o] | h->mymx_B[o];
    u32 nc = h->calc[o] | rc[w];
    u64 nm = om | rm[w];
    if (nc == h->calc[o] && nm == om) continue;
    const u32 chg = (h->calc[o] ^ nc) | (u32)(om ^ nm) | (u32)((om ^ nm) >> 32);
    u_grant(g, o, om, nm);
    h->calc[o] = nc;
    if (!wake) continue;
    for (int dx = -1; dx <= 1; dx++) {
      if (!band_word_reads_ghost(g, h->table.data(), side, own, w, dx, chg)) continue;
      u32 q = u_try_queue(g, own, w + dx);
      if (q != U_NONE) h->frontier.push_back(q & ~U_QUEUED);
    }
  }
  return (int64_t)h->frontier.size();
}
// debugging aid: calc / my / mx of the cell (global y, x) as this band sees it (owned, ghost or halo row)
int emul_debug_cell(void *p, int y, int x) {
  HostGrid *h = (HostGrid *)p;
  const int yl = y - h->row0;
  if (yl < 0 || yl >= h->gh) return -1;
  const size_t o = (size_t)yl * h->pitch + (x >> 5);
  const u32 bit = 1u << (x & 31);
  return ((h->calc[o] & bit) ? 1 : 0) | (((u32)(h->mymx[o] | h->mymx_B[o]) & bit) ? 2 : 0) | (((u32)((h->mymx[o] | h->mymx_B[o]) >> 32) & bit) ? 4 : 0) |
         ((h->pref[h->cur][o] & bit) ? 8 : 0) | ((h->ustate[o] != U_NONE) ? 16 : 0) |
         ((h->head[h->cur][o].x & bit) ? 32 : 0) | ((h->head[h->cur][o].z & bit) ? 64 : 0) | ((h->road[o].x & bit) ? 128 : 0) |
         ((h->road[o].z & bit) ? 256 : 0);
}
void emul_band_u_resume(void *p) {
  HostGrid *h = (HostGrid *)p;
  GridView g = view(h, 0, 0, 2);
  u_iter(h, g, v_order(h));
}
void emul_band_rings(void *p) {
  HostGrid *h = (HostGrid *)p;
  GridView g = view(h, 0, 0, 2);
  do_rings(h, g);
}
int64_t emul_band_ring_words(void *p) {
  HostGrid *h = (HostGrid *)p;
  return h->has_rings ? (int64_t)((h->rings.off.size() - 1 + 31) / 32) : 0;
}
void emul_band_ring_verdicts(void *p, uint32_t *out) {   // copy and reset
  HostGrid *h = (HostGrid *)p;
  for (size_t i = 0; i < h->rbits.size(); i++) { out[i] = h->rbits[i]; h->rbits[i] = 0xFFFFFFFFu; }
}
void emul_band_rings_apply(void *p, const uint32_t *pass) {
  HostGrid *h = (HostGrid *)p;
  GridView g = view(h, 0, 0, 2);
  do_rings(h, g, 1, pass);
}
int64_t emul_band_move(void *p, int chooser, uint32_t tape_bit) {
  HostGrid *h = (HostGrid *)p;
  GridView g = view(h, chooser, tape_bit, 2);
  const int64_t leaks = h->leaks_spec + do_fix(h, g);
  h->cur ^= 1;
  h->tick++;
  return leaks;
}
}
