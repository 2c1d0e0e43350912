// warp_emu.cpp — see warp_emu.h.  TEST-ONLY.
#include "warp_emu.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../mcp_context_forge_b200/csrc/warp_prims.h"

#if !defined(__x86_64__)
#error "the warp emulator's context switch is written for x86-64"
#endif

extern "C" void wemu_swap(void** save_sp, void* load_sp);
asm(R"(
.text
.globl wemu_swap
.type wemu_swap,@function
wemu_swap:
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
.size wemu_swap,.-wemu_swap
)");

namespace wemu {
namespace {
enum Op : int { OP_NONE = 0, OP_BALLOT, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SYNC };
const size_t STACK = 512 * 1024;

struct Warp {
  void* sched_sp = nullptr;
  void* lane_sp[32];
  bool done[32];
  int op[32];
  uint32_t val[32], arg[32], res[32];
  const std::function<void(uint32_t)>* body = nullptr;
  uint64_t ncoll = 0;
  std::vector<char> stacks;
};
thread_local Warp* W = nullptr;
thread_local uint32_t LANE = 0;
thread_local uint64_t LAST_COLL = 0;

void tramp() {
  Warp* w = W;
  const uint32_t l = LANE;
  (*w->body)(l);
  w->done[l] = true;
  w->op[l] = OP_NONE;
  wemu_swap(&w->lane_sp[l], w->sched_sp);
  abort();   // a finished lane is never resumed
}

uint32_t collective(int op, uint32_t v, uint32_t a) {
  Warp* w = W;
  const uint32_t l = LANE;
  w->op[l] = op; w->val[l] = v; w->arg[l] = a;
  wemu_swap(&w->lane_sp[l], w->sched_sp);
  LANE = l;
  return w->res[l];
}
}  // namespace

void run_warp(const std::function<void(uint32_t)>& body, int order) {
  Warp w;
  w.body = &body;
  w.stacks.resize(32 * STACK + 64);
  char* base = (char*)(((uintptr_t)w.stacks.data() + 63) & ~(uintptr_t)63);
  for (uint32_t l = 0; l < 32; ++l) {
    w.done[l] = false; w.op[l] = OP_NONE;
    uintptr_t top = ((uintptr_t)(base + (l + 1) * STACK)) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of tramp's caller
    *--sp = (void*)&tramp;           // `ret` of the first switch lands here
    for (int i = 0; i < 6; ++i) *--sp = nullptr;
    w.lane_sp[l] = sp;
  }
  Warp* prev = W;
  W = &w;
  while (true) {
    uint32_t live = 0;
    for (uint32_t i = 0; i < 32; ++i) {
      const uint32_t l = order ? 31 - i : i;
      if (w.done[l]) continue;
      LANE = l;
      wemu_swap(&w.sched_sp, w.lane_sp[l]);
      if (!w.done[l]) ++live;
    }
    if (!live) break;
    // every live lane sits at a collective; they must agree on which one
    int op = OP_NONE;
    for (uint32_t l = 0; l < 32; ++l) {
      if (w.done[l]) continue;
      if (op == OP_NONE) op = w.op[l];
      else if (op != w.op[l]) { fprintf(stderr, "warp_emu: divergent collectives (%d vs %d at lane %u)\n", op, w.op[l], l); abort(); }
    }
    ++w.ncoll;
    switch (op) {
      case OP_BALLOT: {
        uint32_t m = 0;
        for (uint32_t l = 0; l < 32; ++l) if (!w.done[l] && w.val[l]) m |= 1u << l;
        for (uint32_t l = 0; l < 32; ++l) w.res[l] = m;
        break;
      }
      case OP_SHFL:
        for (uint32_t l = 0; l < 32; ++l) w.res[l] = w.val[w.arg[l] & 31u];
        break;
      case OP_SHFL_UP:
        for (uint32_t l = 0; l < 32; ++l) w.res[l] = l >= w.arg[l] ? w.val[l - w.arg[l]] : w.val[l];
        break;
      case OP_SHFL_DOWN:
        for (uint32_t l = 0; l < 32; ++l) w.res[l] = l + w.arg[l] < 32 ? w.val[l + w.arg[l]] : w.val[l];
        break;
      default: break;
    }
  }
  LAST_COLL = w.ncoll;
  W = prev;
}

uint64_t collectives() { return LAST_COLL; }
}  // namespace wemu

namespace tpw {
uint32_t lane() { return wemu::LANE; }
uint32_t ballot(bool p) { return wemu::collective(wemu::OP_BALLOT, p ? 1u : 0u, 0); }
bool any(bool p) { return ballot(p) != 0; }
uint32_t shfl(uint32_t v, uint32_t src) { return wemu::collective(wemu::OP_SHFL, v, src); }
uint32_t shfl_up(uint32_t v, uint32_t d) { return wemu::collective(wemu::OP_SHFL_UP, v, d); }
uint32_t shfl_down(uint32_t v, uint32_t d) { return wemu::collective(wemu::OP_SHFL_DOWN, v, d); }
void sync() { wemu::collective(wemu::OP_SYNC, 0, 0); }
}  // namespace tpw
