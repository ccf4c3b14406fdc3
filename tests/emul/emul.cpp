// CPU emulation of the CUDA tile programs (TEST INFRASTRUCTURE).
//
// Compiles flowamok_b200/csrc/fa_tiles.h for the host and runs the very same phase functions the
// kernels run, with a loop over "thread ids" in place of a thread block and the same multi-pass
// pending/changed protocol as the host driver in fa_api.cu.  tests/test_emul.py compares it with
// the oracle.  It is never part of the product library.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../flowamok_b200/csrc/fa_tiles.h"

using namespace fa;

namespace {

struct HostGrid {
  int gh, gw, pitch, wvalid;
  std::vector<U4> road, head[2];
  std::vector<u32> pref[2], calc, color[2];
  std::vector<u64> mymx;
  std::vector<u32> ustate;
  std::vector<u64> rng;
  u64 rng_inc;
  int cur = 0;
};

HostGrid *make(int gh, int gw) {
  HostGrid *h = new HostGrid();
  h->gh = gh; h->gw = gw;
  h->wvalid = (gw + 31) / 32;
  h->pitch = ((h->wvalid + 3) / 4) * 4;
  if (h->pitch == 0) h->pitch = 4;
  size_t nw = (size_t)gh * h->pitch, nc = nw * 32;
  h->road.assign(nw, U4{0, 0, 0, 0});
  for (int k = 0; k < 2; k++) { h->head[k].assign(nw, U4{0, 0, 0, 0}); h->pref[k].assign(nw, 0); h->color[k].assign(nc, 0); }
  h->calc.assign(nw, 0); h->mymx.assign(nw, 0); h->ustate.assign(nw, 0);
  h->rng.assign(nc, 0);
  h->rng_inc = 0;
  return h;
}

GridView view(HostGrid *h, int chooser, u32 tape_bit) {
  GridView g;
  g.gh = h->gh; g.gw = h->gw; g.pitch = h->pitch; g.wvalid = h->wvalid;
  g.road = h->road.data();
  g.head_in = h->head[h->cur].data(); g.head_out = h->head[h->cur ^ 1].data();
  g.pref_in = h->pref[h->cur].data(); g.pref_out = h->pref[h->cur ^ 1].data();
  g.calc = h->calc.data(); g.mymx = h->mymx.data(); g.ustate = h->ustate.data();
  g.color_in = h->color[h->cur].data(); g.color_out = h->color[h->cur ^ 1].data();
  g.rng = h->rng.data(); g.rng_inc = h->rng_inc;
  g.chooser = chooser; g.tape_bit = tape_bit;
  return g;
}

struct Rings {
  std::vector<int64_t> off;
  std::vector<int64_t> y, x;
  std::vector<int8_t> cy, cx;
  std::vector<uint8_t> grant;
};

// Frontier discipline of k_u_init / k_u_iter: iteration 0 visits every entry, later iterations visit
// what the previous one queued; stop on an empty list.  `order` varies the processing order inside
// an iteration (the result must not depend on it).
template <class MT, class UIT>
int64_t tick(HostGrid *h, int chooser, u32 tape_bit, const Rings *rings, int nt, unsigned *stats, int order) {
  GridView g = view(h, chooser, tape_bit);
  std::vector<UEntry> table;
  std::vector<u32> cur, next;
  {  // k_u_init: one "block" per UIT tile, NT threads emulated in lock step
    typename UIT::Smem *sm = new typename UIT::Smem();
    std::vector<typename UIT::Own> own(UIT::NT);
    std::vector<typename UIT::Raw> rawo(UIT::NT), rawr(UIT::NT);
    std::vector<UState> nn(UIT::NT);
    int nty = (h->gh + UIT::TR - 1) / UIT::TR, ntx = (h->pitch + UIT::TW - 1) / UIT::TW;
    for (int ty = 0; ty < nty; ty++)
      for (int tx = 0; tx < ntx; tx++) {
        UIT tile(g, *sm, ty, tx);
        for (int t = 0; t < UIT::NT; t++) tile.p0(t, rawo[t], rawr[t]);
        for (int t = 0; t < UIT::NT; t++) tile.pA(t, rawo[t], rawr[t], own[t]);
        bool converged = false;
        for (int k = 0; k < UIT::NS; k++) {
          bool any = false;
          for (int t = 0; t < UIT::NT; t++) any |= tile.sweep_compute(own[t], nn[t]);
          if (!any) { converged = true; break; }
          for (int t = 0; t < UIT::NT; t++) tile.sweep_commit(own[t], nn[t]);
        }
        for (int t = 0; t < UIT::NT; t++) {
          UEntry e;
          int kind = tile.pC(t, own[t], converged, e);
          if (kind >= 1) {
            u32 pos = (u32)table.size();
            table.push_back(e);
            g.ustate[e.wi] = kind == 2 ? (pos | U_QUEUED) : pos;
            if (kind == 2) cur.push_back(pos);
          } else if (own[t].in) g.ustate[(size_t)own[t].y * h->pitch + own[t].w] = U_NONE;
        }
      }
    delete sm;
  }
  unsigned iters = 0, relaxed = 0;
  while (!cur.empty()) {
    iters++;
    next.clear();
    size_t n = cur.size();
    for (size_t k = 0; k < n; k++) {
      size_t kk = order == 2 ? n - 1 - k : (order == 1 ? (k * 7919u) % n : k);
      if (order == 1 && n % 7919u == 0) kk = k;
      u32 push[9];
      int np = u_iter_entry(g, table[cur[kk]], cur[kk], push);
      relaxed++;
      for (int j = 0; j < np; j++) next.push_back(push[j]);
    }
    cur.swap(next);
  }
  (void)nt;
  if (rings) {  // steppers.fut:189-224, sparse form
    size_t n = rings->off.size() - 1;
    std::vector<char> pass(n);
    for (size_t k = 0; k < n; k++) {
      bool ok = true;
      for (int64_t t = rings->off[k]; t < rings->off[k + 1] && ok; t++) {
        size_t o = (size_t)rings->y[t] * h->pitch + (rings->x[t] >> 5);
        u32 bit = 1u << (rings->x[t] & 31);
        U4 hd = g.head_in[o];
        int dy = (hd.x & bit) ? ((hd.y & bit) ? -1 : 1) : 0, dx = (hd.z & bit) ? ((hd.w & bit) ? -1 : 1) : 0;
        if ((g.calc[o] & bit) || dy != rings->cy[t] || dx != rings->cx[t]) ok = false;
      }
      pass[k] = ok;
    }
    for (size_t k = 0; k < n; k++)
      if (pass[k])
        for (int64_t t = rings->off[k]; t < rings->off[k + 1]; t++) {
          size_t o = (size_t)rings->y[t] * h->pitch + (rings->x[t] >> 5);
          u32 bit = 1u << (rings->x[t] & 31);
          if (rings->grant[t] == 1) g.mymx[o] |= (u64)bit; else g.mymx[o] |= (u64)bit << 32;
        }
  }
  int64_t leaks = 0;
  int mty = (h->gh + MT::TR - 1) / MT::TR, mtx = (h->pitch + MT::TW - 1) / MT::TW;
  typename MT::Smem *sm = new typename MT::Smem();
  for (int ty = 0; ty < mty; ty++)
    for (int tx = 0; tx < mtx; tx++) {
      MT t(g, *sm, ty, tx);
      for (int tid = 0; tid < nt; tid++) t.p0(tid, nt);
      for (int tid = 0; tid < nt; tid++) t.p1(tid, nt);
      for (int tid = 0; tid < nt; tid++) t.p2(tid, nt);
      leaks += sm->leaks;
    }
  delete sm;
  h->cur ^= 1;
  if (stats) { stats[0] += iters; stats[1] += relaxed; }
  return leaks;
}

}  // namespace

extern "C" {

void *emul_new(int gh, int gw) { return make(gh, gw); }
void emul_free(void *p) { delete (HostGrid *)p; }

void emul_upload(void *p, const int8_t *uy, const int8_t *ux, const int8_t *dy, const int8_t *dx,
                 const uint32_t *color, const uint8_t *pref, const uint64_t *rng_state, uint64_t rng_inc) {
  HostGrid *h = (HostGrid *)p;
  size_t cp = (size_t)h->pitch * 32;
  for (auto &v : h->road) v = U4{0, 0, 0, 0};
  for (auto &v : h->head[h->cur]) v = U4{0, 0, 0, 0};
  for (auto &v : h->pref[h->cur]) v = 0;
  for (int y = 0; y < h->gh; y++)
    for (int x = 0; x < h->gw; x++) {
      size_t i = (size_t)y * h->gw + x, o = (size_t)y * h->pitch + (x >> 5);
      u32 bit = 1u << (x & 31);
      if (uy[i]) { h->road[o].x |= bit; if (uy[i] < 0) h->road[o].y |= bit; }
      if (ux[i]) { h->road[o].z |= bit; if (ux[i] < 0) h->road[o].w |= bit; }
      U4 &hd = h->head[h->cur][o];
      if (dy[i]) { hd.x |= bit; if (dy[i] < 0) hd.y |= bit; }
      if (dx[i]) { hd.z |= bit; if (dx[i] < 0) hd.w |= bit; }
      if (pref[i]) h->pref[h->cur][o] |= bit;
      h->color[h->cur][(size_t)y * cp + x] = color[i];
      h->rng[(size_t)y * cp + x] = rng_state[i];
    }
  h->rng_inc = rng_inc;
}

void emul_download(void *p, int8_t *dy, int8_t *dx, uint32_t *color, uint8_t *pref, uint64_t *rng_state) {
  HostGrid *h = (HostGrid *)p;
  size_t cp = (size_t)h->pitch * 32;
  for (int y = 0; y < h->gh; y++)
    for (int x = 0; x < h->gw; x++) {
      size_t i = (size_t)y * h->gw + x, o = (size_t)y * h->pitch + (x >> 5);
      u32 bit = 1u << (x & 31);
      const U4 &hd = h->head[h->cur][o];
      dy[i] = (hd.x & bit) ? ((hd.y & bit) ? -1 : 1) : 0;
      dx[i] = (hd.z & bit) ? ((hd.w & bit) ? -1 : 1) : 0;
      pref[i] = (h->pref[h->cur][o] & bit) ? 1 : 0;
      color[i] = h->color[h->cur][(size_t)y * cp + x];
      rng_state[i] = h->rng[(size_t)y * cp + x];
    }
}

// variant: 0 = production tile shapes, 1 = tiny tiles (stress halo / pending protocol), 2 = medium
int64_t emul_step(void *p, int variant, int nthreads, int chooser, uint32_t tape_bit, int64_t n_rings,
                  const int64_t *off, const int64_t *ry, const int64_t *rx, const int8_t *cy,
                  const int8_t *cx, const uint8_t *grant, unsigned *stats) {
  HostGrid *h = (HostGrid *)p;
  Rings R, *rp = nullptr;
  if (n_rings >= 0 && off) {
    R.off.assign(off, off + n_rings + 1);
    int64_t tot = off[n_rings];
    R.y.assign(ry, ry + tot); R.x.assign(rx, rx + tot);
    R.cy.assign(cy, cy + tot); R.cx.assign(cx, cx + tot); R.grant.assign(grant, grant + tot);
    rp = &R;
  }
  switch (variant) {  // variant: tile shapes / processing order of the frontier
  case 1: return tick<MTile<3, 1>, UInitTile<2, 1>>(h, chooser, tape_bit, rp, nthreads, stats, 2);
  case 2: return tick<MTile<4, 2>, UInitTile<5, 3>>(h, chooser, tape_bit, rp, nthreads, stats, 1);
  default: return tick<MTile<8, 32>, UInitTile<16, 32>>(h, chooser, tape_bit, rp, nthreads, stats, 0);
  }
}
}
