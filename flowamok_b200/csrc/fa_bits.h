// flowamok_b200 -- bit-plane cell logic shared by every kernel.
//
// The grid is stored as bit planes: bit i of 32-bit word w of row y is cell (y, 32*w + i).
// i8 fields of the reference cell record (src/model.fut:5-24) become two planes each:
//   *N = "non-zero", *S = "negative" (only meaningful where *N is set; kept 0 otherwise).
//     road  u.y -> UYN,UYS   u.x -> UXN,UXS      (underlying.direction, static within a tick)
//     car   d.y -> DYN,DYS   d.x -> DXN,DXS      (direction)
//     PREF  = can_be_moved_to_from.next_preference_flow_x (persistent)
//     CALC, MY, MX = can_be_moved_to_from.{calculated,direction.y,direction.x} (per tick)
// One 32-bit logic op therefore advances 32 cells; x-neighbours are funnel shifts, y-neighbours
// are the same word of the row above/below.
//
// All functions are pure and __host__ __device__ so the exact code the kernels run is also unit
// tested on the CPU (tests/emul).  Citations: /root/reference/src/steppers.fut unless noted.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD inline
#endif

namespace fa {

typedef uint32_t u32;
typedef uint64_t u64;

// A word with its west (l) and east (r) neighbour words.
struct W3 {
  u32 l, c, r;
};
// value of the EAST neighbour cell (x+1) aligned to this cell's bit
FA_HD u32 east(const W3 &v) { return (v.c >> 1) | (v.r << 31); }
// value of the WEST neighbour cell (x-1) aligned to this cell's bit
FA_HD u32 west(const W3 &v) { return (v.c << 1) | (v.l >> 31); }

// Heading decomposition of a word of cars.
struct Heading {
  u32 cN, cS, cW, cE;      // cardinal: exactly one axis set
  u32 dNW, dNE, dSW, dSE;  // diagonal: utils.fut:35 allows both axes non-zero
  u32 hN, hS, hW, hE;      // per-axis components (cardinal or diagonal)
  u32 occ;                 // helpers.fut:21 is_occupied
};
FA_HD Heading heading(u32 DYN, u32 DYS, u32 DXN, u32 DXS) {
  Heading h;
  h.hN = DYN & DYS; h.hS = DYN & ~DYS; h.hW = DXN & DXS; h.hE = DXN & ~DXS;
  h.cN = h.hN & ~DXN; h.cS = h.hS & ~DXN; h.cW = h.hW & ~DYN; h.cE = h.hE & ~DYN;
  h.dNW = h.hN & h.hW; h.dNE = h.hN & h.hE; h.dSW = h.hS & h.hW; h.dSE = h.hS & h.hE;
  h.occ = DYN | DXN;
  return h;
}

// Position masks of one word: rowFirst/rowLast are all-ones on the first/last grid row,
// colFirst/colLast have the single bit of x==0 / x==gw-1 if it lies in this word,
// valid has the bits with x < gw.
struct Edge {
  u32 rowFirst, rowLast, colFirst, colLast, valid;
};
FA_HD Edge edge_masks(int y, int w, int gh, int gw) {
  Edge e;
  e.rowFirst = (y == 0) ? 0xFFFFFFFFu : 0u;
  e.rowLast = (y == gh - 1) ? 0xFFFFFFFFu : 0u;
  e.colFirst = (w == 0) ? 1u : 0u;
  int lastw = (gw - 1) >> 5;
  e.colLast = (w == lastw) ? (1u << ((gw - 1) & 31)) : 0u;
  int x0 = w * 32;
  if (y < 0 || y >= gh || x0 >= gw || w < 0) e.valid = 0u;
  else if (x0 + 32 <= gw) e.valid = 0xFFFFFFFFu;
  else e.valid = (1u << (gw - x0)) - 1u;
  return e;
}

// successor cell (y+d.y, x+d.x) out of bounds (helpers.fut:27; steppers.fut:20, :68, :117)
FA_HD u32 succ_oob(const Heading &h, const Edge &e) {
  return (h.hN & e.rowFirst) | (h.hS & e.rowLast) | (h.hW & e.colFirst) | (h.hE & e.colLast);
}

// Value of field F at each cell's successor.  up/cur/dn are the rows y-1, y, y+1.
FA_HD u32 succ_card(const Heading &h, u32 up_c, const W3 &cur, u32 dn_c) {
  return (h.cN & up_c) | (h.cS & dn_c) | (h.cW & west(cur)) | (h.cE & east(cur));
}
FA_HD u32 succ_diag(const Heading &h, const W3 &up, const W3 &dn) {
  return (h.dNW & west(up)) | (h.dNE & east(up)) | (h.dSW & west(dn)) | (h.dSE & east(dn));
}
FA_HD u32 any_diag(const Heading &h) { return h.dNW | h.dNE | h.dSW | h.dSE; }

// ---------------------------------------------------------------------------------------------
// U phase statics (steppers.fut:17-50).  Everything here is constant during the fixpoint.
// ---------------------------------------------------------------------------------------------
struct UStatic {
  u32 LN, LS, LW, LE;          // non-terminal car whose successor is N/S/W/E (cardinal headings)
  u32 LNW, LNE, LSW, LSE;      // same for diagonal headings
  u32 GYr, GXr;                // grant to the y / x predecessor if the cell itself is free (:35-53)
  u32 T;                       // terminal: ready and base are both true without looking anywhere
};

// road_* : UYN|UXN (helpers.fut:18) of the neighbourhood rows; dyn_up/dyn_dn: DYN of rows y-1/y+1;
// dxn: DXN of this row with neighbours.
FA_HD UStatic u_statics(u32 UYN, u32 UYS, u32 UXN, u32 UXS, u32 DYN, u32 DYS, u32 DXN_c, u32 DXS,
                        u32 PREF, const W3 &road_up, const W3 &road, const W3 &road_dn, u32 dyn_up,
                        u32 dyn_dn, const W3 &dxn, const Edge &e) {
  UStatic s;
  Heading h = heading(DYN, DYS, DXN_c, DXS);
  u32 oob = succ_oob(h, e);
  u32 sroad = succ_card(h, road_up.c, road, road_dn.c) | succ_diag(h, road_up, road_dn);
  // :18-28: empty, successor out of bounds, or successor without road -> (true, true)
  s.T = ~h.occ | oob | ~sroad;
  u32 link = ~s.T;
  s.LN = link & h.cN; s.LS = link & h.cS; s.LW = link & h.cW; s.LE = link & h.cE;
  s.LNW = link & h.dNW; s.LNE = link & h.dNE; s.LSW = link & h.dSW; s.LSE = link & h.dSE;
  // :33-40: is the upstream cell along each road axis occupied by a car moving on that axis?
  // (y - u.y): u.y = -1 -> row y+1.  Out-of-bounds rows/words are zero-filled by the loader.
  u32 py = UYN & ((UYS & dyn_dn) | (~UYS & dyn_up));
  u32 px = UXN & ((UXS & east(dxn)) | (~UXS & west(dxn)));
  // :41-47 as boolean algebra
  u32 gy = py & (~PREF | ~px);
  u32 gx = px & (PREF | ~py);
  u32 roadc = UYN | UXN;
  s.GYr = gy | ~roadc;  // :51-53: a cell without road accepts from both axes
  s.GXr = gx | ~roadc;
  return s;
}

// One monotone relaxation step of a word (steppers.fut:15-55 applied to 32 cells).
// R = calculated, MY/MX = can_be_moved_to_from.direction.{y,x}; *_up/*_dn/W3 are the
// neighbours' CURRENT values (any mix of old and new is fine: the system is confluent).
struct UState {
  u32 R, MY, MX;
};
FA_HD UState u_relax(const UStatic &s, const W3 &R_up, const W3 &R, const W3 &R_dn, const W3 &MY_up,
                     const W3 &MY_dn, const W3 &MX, const W3 &MX_up, const W3 &MX_dn) {
  UState o;
  // :23 ready = successor calculated
  u32 r = s.T | (s.LN & R_up.c) | (s.LS & R_dn.c) | (s.LW & west(R)) | (s.LE & east(R));
  // :24-25 base = (d.y != 0 && next.my) || (d.x != 0 && next.mx)
  u32 b = s.T | (s.LN & MY_up.c) | (s.LS & MY_dn.c) | (s.LW & west(MX)) | (s.LE & east(MX));
  if (s.LNW | s.LNE | s.LSW | s.LSE) {
    r |= (s.LNW & west(R_up)) | (s.LNE & east(R_up)) | (s.LSW & west(R_dn)) | (s.LSE & east(R_dn));
    b |= (s.LNW & (west(MY_up) | west(MX_up))) | (s.LNE & (east(MY_up) | east(MX_up))) |
         (s.LSW & (west(MY_dn) | west(MX_dn))) | (s.LSE & (east(MY_dn) | east(MX_dn)));
  }
  u32 free_ = r & b;  // base only counts once the successor is calculated
  o.R = r;
  o.MY = free_ & s.GYr;
  o.MX = free_ & s.GXr;
  return o;
}

// The same step for a word without diagonal headings: only 10 neighbour words are needed.
FA_HD UState u_relax_card(const UStatic &s, u32 R_up, const W3 &R, u32 R_dn, u32 MY_up, u32 MY_dn, const W3 &MX) {
  UState o;
  u32 r = s.T | (s.LN & R_up) | (s.LS & R_dn) | (s.LW & west(R)) | (s.LE & east(R));
  u32 b = s.T | (s.LN & MY_up) | (s.LS & MY_dn) | (s.LW & west(MX)) | (s.LE & east(MX));
  u32 free_ = r & b;
  o.R = r;
  o.MY = free_ & s.GYr;
  o.MX = free_ & s.GXr;
  return o;
}

// next_preference_flow_x after the fixpoint (:41-50): only road cells that were calculated AND free re-arbitrate;
// ring resolution (:189-224) never touches it.  On a road cell my and mx exclude each other and MY | MX can both be 0
// while free (nobody upstream), in which case pref stays (:47); MY=1 -> 1 (:44,:45), MX=1 -> 0 (:43,:46).  A cell has
// a road iff its static grants GYr and GXr are not both set (u_statics: a cell without road accepts from both axes).
FA_HD u32 u_pref_from_grants(u32 GYr, u32 GXr, u32 PREF, u32 MY, u32 MX) {
  u32 upd = ~(GYr & GXr) & (MY | MX);
  return (~upd & PREF) | (upd & MY);
}

// ---------------------------------------------------------------------------------------------
// M phase (steppers.fut:58-125) + reset (:140-145): pure gather, expressed per word.
// ---------------------------------------------------------------------------------------------
struct MRow {  // everything M needs from one neighbourhood row, centre word plus x-neighbours
  W3 UYN, UYS, UXN, UXS, DYN, DYS, DXN, DXS, MY, MX;
};
struct MOut {
  u32 DYN, DYS, DXN, DXS;  // new car planes
  u32 AY, AX;              // arrival from the y / x predecessor (colour source)
  u32 fork;                // arrivals that must draw a direction (:102-103)
  u32 leak;                // source cells whose car leaked (:116-124)
};

// First half: everything that does not depend on the random choices.
struct MPre {
  u32 AY, AX, fork, selYnf, selXnf, keep, PDYN, PDYS, PDXN, PDXS, V, leak, UYS, UXS, valid;
  u32 DYN, DYS, DXN, DXS;
};
FA_HD MPre m_pre(const MRow &up, const MRow &c, const MRow &dn, const Edge &e) {
  MPre p;
  const u32 UYN = c.UYN.c, UYS = c.UYS.c, UXN = c.UXN.c, UXS = c.UXS.c;
  const u32 MY = c.MY.c, MX = c.MX.c;
  Heading h = heading(c.DYN.c, c.DYS.c, c.DXN.c, c.DXS.c);
  // :76-83  which predecessor, and does its road continue into this cell
  u32 pym = (UYS & dn.UYN.c & dn.UYS.c) | (~UYS & up.UYN.c & ~up.UYS.c);
  u32 pxm = (UXS & east(c.UXN) & east(c.UXS)) | (~UXS & west(c.UXN) & ~west(c.UXS));
  p.AY = MY & UYN & pym;
  p.AX = ~MY & MX & UXN & pxm;
  u32 ARR = p.AY | p.AX;
  // :85-100  do the two onward roads continue
  u32 inbDY = (UYS & ~e.rowFirst) | (~UYS & ~e.rowLast);
  u32 nUYN = (UYS & up.UYN.c) | (~UYS & dn.UYN.c);
  u32 nUYS = (UYS & up.UYS.c) | (~UYS & dn.UYS.c);
  u32 nRdY = (UYS & (up.UYN.c | up.UXN.c)) | (~UYS & (dn.UYN.c | dn.UXN.c));
  u32 okY = nUYN & ~(nUYS ^ UYS);
  u32 fy = UYN & inbDY & okY, fyi = UYN & inbDY & (okY | ~nRdY);
  u32 inbDX = (UXS & ~e.colFirst) | (~UXS & ~e.colLast);
  W3 roadc = {c.UYN.l | c.UXN.l, c.UYN.c | c.UXN.c, c.UYN.r | c.UXN.r};
  u32 nUXN = (UXS & west(c.UXN)) | (~UXS & east(c.UXN));
  u32 nUXS = (UXS & west(c.UXS)) | (~UXS & east(c.UXS));
  u32 nRdX = (UXS & west(roadc)) | (~UXS & east(roadc));
  u32 okX = nUXN & ~(nUXS ^ UXS);
  u32 fx = UXN & inbDX & okX, fxi = UXN & inbDX & (okX | ~nRdX);
  // :101-108
  u32 c1 = fy & fx;
  u32 c2 = ~c1 & ~fx & fyi;
  u32 c3 = ~c1 & ~c2 & ~fy & fxi;
  u32 c4 = ~c1 & ~c2 & ~c3;
  p.fork = ARR & c1;
  p.selYnf = ARR & c2;
  p.selXnf = ARR & c3;
  p.keep = ARR & c4;
  // :108 the arriving car's own heading
  p.PDYN = (p.AY & ((UYS & dn.DYN.c) | (~UYS & up.DYN.c))) | (p.AX & ((UXS & east(c.DYN)) | (~UXS & west(c.DYN))));
  p.PDYS = (p.AY & ((UYS & dn.DYS.c) | (~UYS & up.DYS.c))) | (p.AX & ((UXS & east(c.DYS)) | (~UXS & west(c.DYS))));
  p.PDXN = (p.AY & ((UYS & dn.DXN.c) | (~UYS & up.DXN.c))) | (p.AX & ((UXS & east(c.DXN)) | (~UXS & west(c.DXN))));
  p.PDXS = (p.AY & ((UYS & dn.DXS.c) | (~UYS & up.DXS.c))) | (p.AX & ((UXS & east(c.DXS)) | (~UXS & west(c.DXS))));
  // :67-74 handle_cell_moved_to_cell_next, and :116-124 the leak test
  u32 oob = succ_oob(h, e);
  u32 sMY = succ_card(h, up.MY.c, c.MY, dn.MY.c);
  u32 sMX = succ_card(h, up.MX.c, c.MX, dn.MX.c);
  W3 road_up = {up.UYN.l | up.UXN.l, up.UYN.c | up.UXN.c, up.UYN.r | up.UXN.r};
  W3 road_dn = {dn.UYN.l | dn.UXN.l, dn.UYN.c | dn.UXN.c, dn.UYN.r | dn.UXN.r};
  u32 sRoad = succ_card(h, road_up.c, roadc, road_dn.c);
  if (any_diag(h)) {
    sMY |= succ_diag(h, up.MY, dn.MY);
    sMX |= succ_diag(h, up.MX, dn.MX);
    sRoad |= succ_diag(h, road_up, road_dn);
  }
  p.V = oob | (sMY & UYN) | (sMX & UXN);
  p.leak = h.occ & ~oob & ~sRoad & ((sMY & c.DYN.c) | (sMX & c.DXN.c));
  p.UYS = UYS; p.UXS = UXS; p.valid = e.valid;
  p.DYN = c.DYN.c; p.DYS = c.DYS.c; p.DXN = c.DXN.c; p.DXS = c.DXS.c;
  return p;
}
// Second half: CH has a 1 where the drawn choice is "x" (flowamok.fut:12-14), only read under fork.
FA_HD MOut m_post(const MPre &p, u32 CH) {
  MOut o;
  u32 ARR = p.AY | p.AX;
  u32 selY = p.selYnf | (p.fork & ~CH);
  u32 selX = p.selXnf | (p.fork & CH);
  u32 stay = ~ARR & ~p.V;
  o.DYN = (selY | (p.keep & p.PDYN) | (stay & p.DYN)) & p.valid;
  o.DYS = ((selY & p.UYS) | (p.keep & p.PDYS) | (stay & p.DYS)) & p.valid;
  o.DXN = (selX | (p.keep & p.PDXN) | (stay & p.DXN)) & p.valid;
  o.DXS = ((selX & p.UXS) | (p.keep & p.PDXS) | (stay & p.DXS)) & p.valid;
  o.AY = p.AY & p.valid; o.AX = p.AX & p.valid; o.fork = p.fork & p.valid; o.leak = p.leak & p.valid;
  return o;
}

// ---------------------------------------------------------------------------------------------
// per-cell RNG: diku-dk/cpprandom pcg32 + uniform_int_distribution i8 (0,1), as drawn by
// flowamok.fut:10-14.  Not in the reference tree; restated from the published algorithm
// ("parity unpinned").  Returns the choice (0 -> y, 1 -> x) and advances state.
// ---------------------------------------------------------------------------------------------
FA_HD u32 pcg32_next(u64 &state, u64 inc) {
  u64 old = state;
  state = old * 6364136223846793005ULL + (inc | 1ULL);
  u32 xs = (u32)(((old >> 18) ^ old) >> 27);
  u32 rot = (u32)(old >> 59);
  return (xs >> rot) | (xs << ((0u - rot) & 31u));
}
FA_HD u32 draw_choice01(u64 &state, u64 inc) {
  // range = 2, secure_max = 0xFFFFFFFF - 0xFFFFFFFF % 2 = 0xFFFFFFFE, result = x / (secure_max / 2)
  u32 x = pcg32_next(state, inc);
  while (x >= 0xFFFFFFFEu) x = pcg32_next(state, inc);
  return x / 0x7FFFFFFFu;
}
FA_HD u32 hash32(u32 x) {
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  x = ((x >> 16) ^ x) * 0x45d9f3bu;
  return (x >> 16) ^ x;
}
// pcg32.rng_from_seed [seed] and element i of split_rng (utils.fut:15)
FA_HD void pcg32_from_seed(int32_t seed, u64 &state, u64 &inc) {
  state = 0; inc = (0xda3e39cb94b95bdbULL << 1) | 1ULL;
  (void)pcg32_next(state, inc);
  state += (u64)(int64_t)seed;
  (void)pcg32_next(state, inc);
}
FA_HD u64 pcg32_split_state(u64 state, u64 inc, int64_t i) {
  u64 s = state ^ (u64)(int64_t)(int32_t)hash32((u32)(int32_t)i);
  (void)pcg32_next(s, inc);
  return s;
}

// splitmix64, used only by the synthetic road-network generator
FA_HD u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

}  // namespace fa
