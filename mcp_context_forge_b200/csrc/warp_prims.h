// warp_prims.h — the handful of warp-level primitives the token-parallel JSON kernels (json_tp.h) are written
// against.  Under nvcc they are the CUDA intrinsics; in the TEST-ONLY host build (tests/hostsim, -DCF_WARP_EMU)
// they are implemented by a 32-fibre warp emulator (tests/hostsim/warp_emu.cpp) so that the very same kernel
// source runs — lane for lane, collective for collective — on a box without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TP_FN __device__ __forceinline__
#define TP_SLOW __device__ __noinline__
namespace tpw {
static const uint32_t FULL = 0xFFFFFFFFu;
TP_FN uint32_t lane() { return threadIdx.x & 31u; }
TP_FN uint32_t ballot(bool p) { return __ballot_sync(FULL, p); }
TP_FN bool any(bool p) { return __any_sync(FULL, p) != 0; }
TP_FN uint32_t shfl(uint32_t v, uint32_t src) { return __shfl_sync(FULL, v, (int)src); }
TP_FN uint32_t shfl_up(uint32_t v, uint32_t d) { return __shfl_up_sync(FULL, v, d); }
TP_FN uint32_t shfl_down(uint32_t v, uint32_t d) { return __shfl_down_sync(FULL, v, d); }
TP_FN void sync() { __syncwarp(); }
TP_FN uint32_t popc(uint32_t v) { return (uint32_t)__popc(v); }
TP_FN uint32_t clz(uint32_t v) { return (uint32_t)__clz((int)v); }
TP_FN uint32_t ffs(uint32_t v) { return (uint32_t)__ffs((int)v); }   // 1-based, 0 for v == 0
}  // namespace tpw
#else
#define TP_FN inline
#define TP_SLOW inline
struct uint4 { uint32_t x, y, z, w; };
namespace tpw {
static const uint32_t FULL = 0xFFFFFFFFu;
uint32_t lane();
uint32_t ballot(bool p);
bool any(bool p);
uint32_t shfl(uint32_t v, uint32_t src);
uint32_t shfl_up(uint32_t v, uint32_t d);
uint32_t shfl_down(uint32_t v, uint32_t d);
void sync();
inline uint32_t popc(uint32_t v) { return (uint32_t)__builtin_popcount(v); }
inline uint32_t clz(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }
inline uint32_t ffs(uint32_t v) { return (uint32_t)__builtin_ffs((int)v); }
}  // namespace tpw
#endif

namespace tpw {
// inclusive prefix sum over the warp
TP_FN uint32_t scan_incl(uint32_t v) {
  const uint32_t l = lane();
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1) {
    const uint32_t t = shfl_up(v, d);
    if (l >= d) v += t;
  }
  return v;
}
TP_FN uint32_t lt_mask() { return (1u << lane()) - 1u; }
}  // namespace tpw
