// Fibre-based warp emulator (TEST INFRASTRUCTURE): runs the 32 lanes of a warp as coroutines in lock step, so code
// written against flowamok_b200/csrc/fa_simt.h (shuffles, ballots) executes on the CPU as written.
//
// Every warp primitive is built on one exchange: a lane deposits its value, yields to the scheduler, and is resumed
// after all other lanes have deposited theirs.  The scheduler resumes the lanes round-robin, so as long as all lanes
// execute the same sequence of primitives (the CUDA requirement for full-mask *_sync intrinsics) two alternating
// slots are enough.  Context switches: a six-register stack switch on x86-64, ucontext elsewhere.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../flowamok_b200/csrc/fa_simt.h"

#if defined(__x86_64__)
extern "C" void fa_simt_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.hidden fa_simt_switch
.globl fa_simt_switch
.type fa_simt_switch,@function
fa_simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size fa_simt_switch,.-fa_simt_switch
)");
#else
#include <ucontext.h>
#endif

namespace simt_host {

struct Warp {
  static constexpr size_t STACK = 256 * 1024;
  char *stacks;   // 32 lane stacks, allocated once per host thread
#if defined(__x86_64__)
  void *sched_sp, *lane_sp[32];
#else
  ucontext_t sched, ctx[32];
#endif
  uint32_t slot[2][32];
  unsigned phase[32];
  bool done[32];
  int cur;
  void (*body)(void *, int);
  void *arg;
  void to_lane(int l) {
    cur = l;
#if defined(__x86_64__)
    fa_simt_switch(&sched_sp, lane_sp[l]);
#else
    swapcontext(&sched, &ctx[l]);
#endif
  }
  void to_sched(int l) {
#if defined(__x86_64__)
    fa_simt_switch(&lane_sp[l], sched_sp);
#else
    swapcontext(&ctx[l], &sched);
#endif
  }
};
static thread_local Warp *tl_warp = nullptr;

static void trampoline() {
  Warp *w = tl_warp;
  const int l = w->cur;
  w->body(w->arg, l);
  w->done[l] = true;
  w->to_sched(l);   // never resumed
  abort();
}

// run body(arg, lane) for lanes 0..31 as one warp
static void run_warp(void (*body)(void *, int), void *arg) {
  static thread_local char *stack_pool = nullptr;
  if (!stack_pool) stack_pool = (char *)malloc(Warp::STACK * 32 + 64);
  Warp *w = new Warp();
  w->stacks = stack_pool;
  w->body = body; w->arg = arg;
  Warp *outer = tl_warp;
  tl_warp = w;
  for (int l = 0; l < 32; l++) {
    w->phase[l] = 0; w->done[l] = false;
    char *lo = w->stacks + Warp::STACK * l;
#if defined(__x86_64__)
    uintptr_t top = ((uintptr_t)(lo + Warp::STACK)) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;                    // fake return address of the trampoline: keeps the ABI stack alignment
    *--sp = (void *)&trampoline;        // `ret` of the first switch jumps here
    for (int k = 0; k < 6; k++) *--sp = nullptr;   // rbp, rbx, r12..r15
    w->lane_sp[l] = sp;
#else
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = lo;
    w->ctx[l].uc_stack.ss_size = Warp::STACK;
    w->ctx[l].uc_link = &w->sched;
    makecontext(&w->ctx[l], trampoline, 0);
#endif
  }
  for (bool any = true; any;) {
    any = false;
    for (int l = 0; l < 32; l++) {
      if (w->done[l]) continue;
      w->to_lane(l);
      any = any || !w->done[l];
    }
  }
  tl_warp = outer;
  delete w;
}

// deposit v, wait for the other lanes, return the 32 deposited values
static const uint32_t *exchange(uint32_t v) {
  Warp *w = tl_warp;
  const int l = w->cur;
  const unsigned ph = w->phase[l]++ & 1u;
  w->slot[ph][l] = v;
  w->to_sched(l);
  return w->slot[ph];
}

}  // namespace simt_host

namespace fa {
int simt_lane() { return simt_host::tl_warp->cur; }
u32 simt_up1(u32 v) { int l = simt_lane(); const uint32_t *a = simt_host::exchange(v); return l > 0 ? a[l - 1] : v; }
u32 simt_dn1(u32 v) { int l = simt_lane(); const uint32_t *a = simt_host::exchange(v); return l < 31 ? a[l + 1] : v; }
u32 simt_bcast(u32 v, int src) { const uint32_t *a = simt_host::exchange(v); return a[src & 31]; }
u32 simt_ballot(bool p) {
  const uint32_t *a = simt_host::exchange(p ? 1u : 0u);
  u32 m = 0;
  for (int i = 0; i < 32; i++) m |= (a[i] & 1u) << i;
  return m;
}
bool simt_any(bool p) { return simt_ballot(p) != 0u; }
void simt_sync() { (void)simt_host::exchange(0u); }
}  // namespace fa
