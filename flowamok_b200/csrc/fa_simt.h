// flowamok_b200 -- the handful of warp primitives the streaming strip programs (fa_stream.h) use.
//
// Device: the sm_100a intrinsics.  Host: the same names are provided by the fibre-based warp emulator of the test
// suite (tests/emul/simt_host.h), which runs the 32 lanes of a warp as coroutines in lock step -- so the strip programs
// are executed on the CPU exactly as written and compared with the oracle before they reach a GPU.  The product
// library never defines the host versions (and never calls a strip program on the host).
#pragma once
#include "fa_bits.h"

namespace fa {

#if defined(__CUDA_ARCH__)
FA_HD int simt_lane() { return (int)(threadIdx.x & 31u); }
// value held by lane-1 / lane+1 (lane 0 / lane 31 get their own value back)
FA_HD u32 simt_up1(u32 v) { return __shfl_up_sync(0xFFFFFFFFu, v, 1); }
FA_HD u32 simt_dn1(u32 v) { return __shfl_down_sync(0xFFFFFFFFu, v, 1); }
// value held by lane `src` (src may differ per lane)
FA_HD u32 simt_bcast(u32 v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
FA_HD u32 simt_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
FA_HD bool simt_any(bool p) { return __any_sync(0xFFFFFFFFu, p) != 0; }
// orders the memory accesses of the warp's lanes (a store before it is visible to every lane after it)
FA_HD void simt_sync() { __syncwarp(); }
// hint: bring the line at p into L2 (no register, no stall); the streaming programs run it a few rows ahead
FA_HD void fa_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
int simt_lane();
u32 simt_up1(u32 v);
u32 simt_dn1(u32 v);
u32 simt_bcast(u32 v, int src);
u32 simt_ballot(bool p);
bool simt_any(bool p);
void simt_sync();
inline void fa_prefetch_l2(const void *) {}
#endif

}  // namespace fa
