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
  std::vector<u32> pref[2], calc, my, mx, color[2];
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
  h->calc.assign(nw, 0); h->my.assign(nw, 0); h->mx.assign(nw, 0);
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
  g.calc = h->calc.data(); g.my = h->my.data(); g.mx = h->mx.data();
  g.color_in = h->color[h->cur].data(); g.color_out = h->color[h->cur ^ 1].data();
  g.rng = h->rng.data(); g.rng_inc = h->rng_inc;
  g.chooser = chooser; g.tape_bit = tape_bit;
  return g;
}

template <class UT>
bool run_u_tile(const GridView &g, int ty, int tx, bool first, int nt, int *res_out, unsigned *iters) {
  typename UT::Smem *sm = new typename UT::Smem();
  UT t(g, *sm, ty, tx, first);
  for (int tid = 0; tid < nt; tid++) t.p0(tid, nt);
  for (int tid = 0; tid < nt; tid++) t.p1(tid, nt);
  for (int tid = 0; tid < nt; tid++) t.p2(tid, nt);
  std::vector<typename UT::Pend> pd(nt);
  int cur = 0;
  for (;;) {
    int n = (int)sm->wn[cur];
    if (n == 0) break;
    bool ch = false;
    for (int base = 0; base < n; base += nt) {
      for (int tid = 0; tid < nt; tid++) t.p3a(tid, cur, base, pd[tid]);   // all reads ...
      for (int tid = 0; tid < nt; tid++) ch |= t.p3b(cur, pd[tid]);        // ... then all writes
    }
    (*iters)++;
    sm->wn[cur] = 0;
    cur ^= 1;
    if (!ch) break;
  }
  int res = 0;
  for (int tid = 0; tid < nt; tid++) res |= t.p4(tid, nt);
  bool pending = false;
  if ((res & 1) && sm->wn[cur] != 0) {
    for (int tid = 0; tid < nt; tid++) t.p5a(tid, nt);
    int n = (int)sm->wn[cur];
    std::vector<int> iw(nt);
    std::vector<u32> xv(nt);
    for (;;) {
      bool ch = false;
      for (int base = 0; base < n; base += nt) {
        for (int tid = 0; tid < nt; tid++) xv[tid] = t.p5b_compute(tid, cur, base, iw[tid]);
        for (int tid = 0; tid < nt; tid++) ch |= t.p5b_commit(iw[tid], xv[tid]);
      }
      if (!ch) break;
    }
    for (int tid = 0; tid < nt; tid++) pending |= t.p5c(tid, nt, cur);
  }
  *res_out = res;
  delete sm;
  return pending;
}

struct Rings {
  std::vector<int64_t> off;
  std::vector<int64_t> y, x;
  std::vector<int8_t> cy, cx;
  std::vector<uint8_t> grant;
};

template <class UT, class MT>
int64_t tick(HostGrid *h, int chooser, u32 tape_bit, const Rings *rings, int nt, unsigned *stats) {
  GridView g = view(h, chooser, tape_bit);
  int nty = (h->gh + UT::TR - 1) / UT::TR, ntx = (h->pitch + UT::TW - 1) / UT::TW;
  std::vector<int> pend, next;
  unsigned iters = 0, passes = 1;
  for (int ty = 0; ty < nty; ty++)
    for (int tx = 0; tx < ntx; tx++) {
      int res;
      if (run_u_tile<UT>(g, ty, tx, true, nt, &res, &iters)) pend.push_back(ty * ntx + tx);
    }
  while (!pend.empty()) {
    bool changed = false;
    next.clear();
    passes++;
    for (int id : pend) {
      int res;
      bool p = run_u_tile<UT>(g, id / ntx, id % ntx, false, nt, &res, &iters);
      if (res & 2) changed = true;
      if (p) next.push_back(id);
    }
    if (!changed) break;
    pend.swap(next);
  }
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
          if (rings->grant[t] == 1) g.my[o] |= bit; else g.mx[o] |= bit;
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
  if (stats) { stats[0] += passes; stats[1] += iters; }
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
  switch (variant) {
  case 1: return tick<UTile<2, 1, 1>, MTile<3, 1>>(h, chooser, tape_bit, rp, nthreads, stats);
  case 2: return tick<UTile<8, 2, 2>, MTile<4, 2>>(h, chooser, tape_bit, rp, nthreads, stats);
  default: return tick<UTile<48, 32, 4>, MTile<16, 32>>(h, chooser, tape_bit, rp, nthreads, stats);
  }
}
}
