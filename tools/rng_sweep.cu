// tools/rng_sweep.cu -- offline research tool (not product, not test): brute-force search for the derivation of the
// per-cell pcg32 state that reproduces the README-GIF fork stream (VERDICT r1 "missing" item 1).
//
// diku-dk/cpprandom 1.3.1 is not vendored in the reference (futhark.pkg:3), so `rng_from_seed` and `split_rng` are
// only known by recollection.  Hypothesis space swept here (everything that keeps the engine PCG-XSH-RR 64/32 and the
// per-cell state a function of ONE 32-bit index hash h and a seed-derived master state):
//
//     state_434 = advance^d( master  op  ext(h) )        h in [0, 2^32)   (any 32-bit hash of any index)
//       master : ~100 seed derivations (initseq shifted or not, seed added / xored / hashed, 0..2 advances around it)
//       op     : xor, mul, add, sub, h alone
//       ext    : zero- or sign-extension of h to 64 bits
//       d      : -2..3 (tape starts before / after the first output; also covers "output from the new state")
//       draw   : top-bit rule (x / 0x7FFFFFFF) or low bit (x % 2), plain or complemented
//
// A hit must reproduce >= 60 consecutive tape bits (false-positive rate ~2^-60 per candidate); hits are printed with
// their parameters and re-checked against all 398 bits on the host.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/rng_sweep tools/rng_sweep.cu && tools/rng_sweep
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;

static const char *TAPE_GIF =
    "0111111001011111111101100111110101100100011010011001100101000011101010011110001011000000001100011100"
    "1000101011010101101101110111010000101101101001001000010011010001100101110011100110110110110100000101"
    "0110101100110001000010001100011011010011100000111011100011110001001011011000110010000111101010010011"
    "00111100111110010100111010010101111101001010111000001101101001111110111101011100100000111011000001";
static char TAPE[512];
#define MULT 6364136223846793005ULL
#define NSHIFT 6   // d = -2 .. 3
#define NSTEPS 66

struct Master { u64 state, inc; int id; };
struct Hit { u64 state0, inc; u32 h; int master, op, ext, flag, pad; };

__constant__ u32 c_expect[NSTEPS];   // bit k of c_expect[t] = tape[t - (k - 2)] for shift index k; bit 8+k = index in range
__constant__ Master c_master[256];

__host__ __device__ inline u32 pcg_out(u64 old) {
  u32 xs = (u32)(((old >> 18) ^ old) >> 27);
  u32 rot = (u32)(old >> 59);
  return (xs >> rot) | (xs << ((0u - rot) & 31u));
}

// returns the surviving (shift, draw rule) flags of the stream that starts at state s
__device__ __forceinline__ u32 match_stream(u64 s, u64 inc) {
  u32 alive = 0xFFFFFFu;
  for (int t = 0; t < NSTEPS && alive; t++) {
    const u32 x = pcg_out(s);
    s = s * MULT + inc;
    const u32 e = c_expect[t], in = (e >> 8) & 0x3Fu, ex = e & 0x3Fu;
    const u32 top = (x >= 0x7FFFFFFFu) ? 0x3Fu : 0u, low = (x & 1u) ? 0x3Fu : 0u;
    const u32 mt = ~(top ^ ex) | ~in, ml = ~(low ^ ex) | ~in;
    const u32 mtc = (top ^ ex) | ~in, mlc = (low ^ ex) | ~in;
    alive &= (mt & 0x3Fu) | ((mtc & 0x3Fu) << 6) | ((ml & 0x3Fu) << 12) | ((mlc & 0x3Fu) << 18);
  }
  return alive;
}

// Second hypothesis space: the index hash is one of a few known candidates (c_hx), the SEED-derived master state is
// a function of a free 32-bit value x (any seed, any 32-bit seed hash):
//     master(x) = post( mix( pre(st0), x ) ),  state_434 = advance^d( master(x) op hx )
__constant__ u64 c_hx[64];
__constant__ u64 c_inc[4];
__global__ void k_sweep2(int nhx, int ninc, u64 x_begin, u64 x_count, Hit *hits, unsigned int *nhits, int cap) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= x_count) return;
  const u32 x32 = (u32)(x_begin + i);
  for (int ii = 0; ii < ninc; ii++) {
    const u64 inc = c_inc[ii] | 1ULL;
    for (int v = 0; v < 48; v++) {   // st0 (2) x pre (2) x mix (3) x post (2) x ext (2)
      const int st0 = v & 1, pre = (v >> 1) & 1, post = (v >> 2) & 1, ext = (v >> 3) & 1, mix = v >> 4;
      if ((ext && (int32_t)x32 >= 0) || st0) continue;   // st0: pcg's default initstate -- dropped to bound the run time
      const u64 x = ext ? (u64)(int64_t)(int32_t)x32 : (u64)x32;
      u64 m = st0 ? 0x853c49e6748fea9bULL : 0ULL;
      if (pre) m = m * MULT + inc;
      m = mix == 0 ? m + x : mix == 1 ? m ^ x : x;
      if (mix == 2 && (st0 || pre)) continue;
      if (post) m = m * MULT + inc;
      for (int k = 0; k < nhx; k++) {
        const u64 hx = c_hx[k];
        for (int op = 0; op < 4; op++) {
          const u64 s0 = op == 0 ? m ^ hx : op == 1 ? m * hx : op == 2 ? m + hx : m - hx;
          const u32 alive = match_stream(s0, inc);
          if (alive) {
            unsigned int q = atomicAdd(nhits, 1u);
            if (q < (unsigned)cap) { Hit h = {s0, inc, x32, 1000 + v, op, k, (int)alive, ii}; hits[q] = h; }
          }
        }
      }
    }
  }
}

__global__ void k_sweep(int nmaster, u64 h_begin, u64 h_count, Hit *hits, unsigned int *nhits, int cap) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h_count) return;
  const u32 h = (u32)(h_begin + i);
  for (int m = 0; m < nmaster; m++) {
    const u64 ms = c_master[m].state, inc = c_master[m].inc | 1ULL;
    for (int ext = 0; ext < 2; ext++) {
      const u64 hx = ext ? (u64)(int64_t)(int32_t)h : (u64)h;
      if (ext && (int32_t)h >= 0) continue;   // same as zero extension
      for (int op = 0; op < 5; op++) {
        if (op == 4 && m != 0) continue;      // h alone does not depend on the master state (inc: master 0's)
        u64 s = op == 0 ? ms ^ hx : op == 1 ? ms * hx : op == 2 ? ms + hx : op == 3 ? ms - hx : hx;
        const u64 s0 = s;
        // alive flags: [0..5] top-bit rule, [6..11] top complemented, [12..17] low bit, [18..23] low complemented
        const u32 alive = match_stream(s, inc);
        if (alive) {
          unsigned int k = atomicAdd(nhits, 1u);
          if (k < (unsigned)cap) { Hit q = {s0, inc, h, c_master[m].id, op, ext, (int)alive, 0}; hits[k] = q; }
        }
      }
    }
  }
}

static u64 adv(u64 s, u64 inc) { return s * MULT + (inc | 1ULL); }
static u32 hash32(u32 x) { x = ((x >> 16) ^ x) * 0x45d9f3bu; x = ((x >> 16) ^ x) * 0x45d9f3bu; return (x >> 16) ^ x; }
static int32_t hash32s(int32_t x) {   // the same on i32 with arithmetic shifts (Futhark >> on signed types)
  x = (int32_t)((u32)((x >> 16) ^ x) * 0x45d9f3bu);
  x = (int32_t)((u32)((x >> 16) ^ x) * 0x45d9f3bu);
  return (x >> 16) ^ x;
}

int main(int argc, char **argv) {
  const int seed = argc > 1 ? atoi(argv[1]) : 123;
  const bool selftest = argc > 2 && !strcmp(argv[2], "selftest");
  strcpy(TAPE, TAPE_GIF);
  if (selftest) {   // plant a known answer: the recollected derivation (xor with the sign-extended hash of 434, one advance)
    const u64 inc = (0xda3e39cb94b95bdbULL << 1) | 1ULL;
    u64 s = adv(adv(0, inc) + (u64)(int64_t)seed, inc);
    s = adv(s ^ (u64)(int64_t)(int32_t)hash32(434u), inc);
    for (int t = 0; t < 398; t++) { TAPE[t] = pcg_out(s) >= 0x7FFFFFFFu ? '1' : '0'; s = adv(s, inc); }
    printf("selftest: expecting a hit with h = 0x%08x, shift 1\n", hash32(434u));
  }
  const int n = (int)strlen(TAPE);
  // expectation table
  u32 expect[NSTEPS];
  for (int t = 0; t < NSTEPS; t++) {
    u32 e = 0;
    for (int k = 0; k < NSHIFT; k++) {
      const int d = k - 2, i = t - d;   // generator output t is tape draw t - d
      if (i >= 0 && i < n) e |= (1u << (8 + k)) | ((u32)(TAPE[i] - '0') << k);
    }
    expect[t] = e;
  }
  cudaMemcpyToSymbol(c_expect, expect, sizeof expect);
  // master states
  std::vector<Master> masters;
  const u64 initseq = 0xda3e39cb94b95bdbULL, initstate = 0x853c49e6748fea9bULL;
  const u64 incs[3] = {(initseq << 1) | 1ULL, initseq | 1ULL, (1442695040888963407ULL)};
  const u64 seeds[6] = {(u64)(int64_t)seed, (u64)hash32((u32)seed), (u64)(int64_t)hash32s(seed), 0, 1, (u64)(u32)seed << 32};
  int id = 0;
  for (int ii = 0; ii < 3; ii++)
    for (int si = 0; si < 6; si++)
      for (int mix = 0; mix < 3; mix++)        // add, xor, replace
        for (int pre = 0; pre < 2; pre++)
          for (int post = 0; post < 2; post++)   // further advances are covered by the shift d
            for (int st0 = 0; st0 < 2; st0++) {
              u64 s = st0 ? initstate : 0ULL;
              if (pre) s = adv(s, incs[ii]);
              s = mix == 0 ? s + seeds[si] : mix == 1 ? s ^ seeds[si] : seeds[si];
              if (post) s = adv(s, incs[ii]);
              bool dup = false;
              for (auto &q : masters) dup |= q.state == s && q.inc == incs[ii];
              if (!dup && masters.size() < 256) masters.push_back({s, incs[ii], id});
              id++;
            }
  printf("seed %d: %zu distinct master states\n", seed, masters.size());
  cudaMemcpyToSymbol(c_master, masters.data(), masters.size() * sizeof(Master));
  Hit *d_hits; unsigned int *d_n;
  const int cap = 4096;
  cudaMalloc(&d_hits, cap * sizeof(Hit)); cudaMalloc(&d_n, 4); cudaMemset(d_n, 0, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const u64 total = 1ULL << 32, chunk = 1ULL << 26;
  for (u64 b = 0; b < total && !(argc > 2 && !strcmp(argv[2], "seedspace")); b += chunk) {
    if (selftest && !(hash32(434u) >= b && hash32(434u) < b + chunk)) continue;   // only the chunk that holds the planted answer
    k_sweep<<<(unsigned)(chunk / 256), 256>>>((int)masters.size(), b, chunk, d_hits, d_n, cap);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  }
  const bool mode2 = argc > 2 && !strcmp(argv[2], "seedspace");
  if (mode2) {
    cudaMemset(d_n, 0, 4);
    std::vector<u64> hx;
    const int64_t idxs[] = {434, 433, 435};
    for (int64_t ix : idxs) {
      const u32 hu = hash32((u32)ix);
      const int32_t hs = hash32s((int32_t)ix);
      const u64 c[] = {(u64)ix, (u64)ix + 1, (u64)hu, (u64)(int64_t)(int32_t)hu, (u64)(u32)hs, (u64)(int64_t)hs};
      for (u64 q : c) { bool dup = false; for (u64 o : hx) dup |= o == q; if (!dup && hx.size() < 64) hx.push_back(q); }
    }
    cudaMemcpyToSymbol(c_hx, hx.data(), hx.size() * 8);
    cudaMemcpyToSymbol(c_inc, incs, sizeof incs);
    cudaEventRecord(e0);
    for (u64 b = 0; b < total; b += chunk) {
      k_sweep2<<<(unsigned)(chunk / 256), 256>>>((int)hx.size(), 3, b, chunk, d_hits, d_n, cap);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms2 = 0; cudaEventElapsedTime(&ms2, e0, e1);
    unsigned int nh2 = 0; cudaMemcpy(&nh2, d_n, 4, cudaMemcpyDeviceToHost);
    printf("seedspace: swept %.3g (x, master variant, index hash, op) combinations x 24 variants in %.1f s: %u hits\n",
           4294967296.0 * 3 * 15 * hx.size() * 4, ms2 * 1e-3, nh2);
    std::vector<Hit> h2(nh2 < (unsigned)cap ? nh2 : cap);
    cudaMemcpy(h2.data(), d_hits, h2.size() * sizeof(Hit), cudaMemcpyDeviceToHost);
    for (auto &q : h2)
      printf("HIT seedspace x 0x%08x variant %d op %d hx# %d state0 0x%016llx inc 0x%016llx flags 0x%x\n", q.h, q.master - 1000, q.op, q.ext,
             (unsigned long long)q.state0, (unsigned long long)q.inc, q.flag);
    return 0;
  }
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  unsigned int nh = 0; cudaMemcpy(&nh, d_n, 4, cudaMemcpyDeviceToHost);
  std::vector<Hit> hits(nh < (unsigned)cap ? nh : cap);
  cudaMemcpy(hits.data(), d_hits, hits.size() * sizeof(Hit), cudaMemcpyDeviceToHost);
  const double cands = (double)masters.size() * 2 * 4 * 4294967296.0;
  printf("swept %.3g (master, op, ext, h) combinations x 24 (shift, draw rule) variants in %.1f s: %u hits\n", cands, ms * 1e-3, nh);
  for (auto &q : hits) {
    // replay all 398 bits for every flag that survived
    for (int f = 0; f < 24; f++) {
      if (!(q.flag & (1 << f))) continue;
      const int k = f % 6, d = k - 2, rule = f / 6;
      u64 s = q.state0;
      int agree = 0, tot = 0;
      for (int t = 0; t < n + 4; t++) {
        u32 x = pcg_out(s); s = s * MULT + q.inc;
        int i = t - d;
        if (i < 0 || i >= n) continue;
        int bit = (rule < 2) ? (x >= 0x7FFFFFFFu) : (int)(x & 1u);
        if (rule & 1) bit ^= 1;
        agree += bit == TAPE[i] - '0'; tot++;
      }
      printf("HIT master %d op %d ext %d h 0x%08x state0 0x%016llx inc 0x%016llx shift %d rule %d: %d/%d tape bits\n", q.master, q.op,
             q.ext, q.h, (unsigned long long)q.state0, (unsigned long long)q.inc, d, rule, agree, tot);
    }
  }
  return 0;
}
