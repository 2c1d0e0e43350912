// cfgpu.cu — device half of libcfgpu.so (include/cfgpu.h): contexts, table upload, batches and
// the sm_100a kernels of the plugin hook-chain hot path.
//
// Kernel inventory
//   prep_kernel      per scan: bitmap initialisation (always-match bits), tile -> first-unit index
//   scan_kernel      the fused multi-pattern scan (harmful + deny + regex_filter dirty detection):
//                    persistent CTAs, TMA (cp.async.bulk) tile ring in shared memory, per byte one
//                    LDS + 4 integer ops shift-AND prefilter, rare candidates verified by an
//                    anchored class DFA (scan_core.h) cooperatively within the warp.
// This is HBM-bound byte work: no tensor cores by design.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "cf_host.h"
#include "scan_core.h"

// ------------------------------------------------------------------------------------------------
// host-side objects
// ------------------------------------------------------------------------------------------------
struct cf_ctx {
  int device = 0;
  int sm_count = 0;
  std::string err;
  uint64_t launches = 0;
  uint64_t* d_counters = nullptr;   // [2] candidates, verify steps
  uint32_t* d_tile_unit = nullptr;  // per-tile first unit index (grown on demand)
  uint64_t tile_unit_cap = 0;
};

struct DevDfa {
  cf::DfaTables t;
  std::vector<void*> allocs;
};

struct cf_prog {
  cf_ctx* ctx = nullptr;
  uint32_t npat = 0, W = 1;
  DevDfa search;
  uint32_t* d_E = nullptr;
  uint64_t* d_always = nullptr;
  bool any_always = false;
  bool search_empty = false;       // every pattern is "always" -> no automaton work at all
  std::vector<DevDfa> ordered;
  std::vector<uint32_t*> d_ordered_E;
  std::vector<int> ordered_pat;    // pattern index of each ordered rule
  std::vector<uint8_t*> d_repl;
  std::vector<uint32_t> repl_len;
};

struct cf_batch {
  cf_ctx* ctx = nullptr;
  uint8_t* d_buf = nullptr;        // FRONT_PAD + stream + tail pad
  uint64_t* d_offsets = nullptr;
  uint64_t cap_bytes = 0;
  uint32_t cap_units = 0;
  uint64_t nbytes = 0;
  uint32_t n = 0;
};

static const uint32_t SCAN_WARPS = 8;
static const uint32_t LANE_BYTES = 64;
static const uint32_t TILE = SCAN_WARPS * 32 * LANE_BYTES;   // 16 KiB
static const uint32_t HALO = 16;
static const uint32_t STAGES = 3;

#define CF_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return CF_E_CUDA;                                                                     \
    }                                                                                       \
  } while (0)

// ------------------------------------------------------------------------------------------------
// device helpers: mbarrier + 1-D TMA bulk copy
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// prep kernel: bitmaps := always-bits; tile_unit[i] := index of the unit containing byte i*TILE
// ------------------------------------------------------------------------------------------------
__global__ void prep_kernel(uint64_t* __restrict__ bitmaps, const uint64_t* __restrict__ always, uint32_t W,
                            uint32_t n_units, const uint64_t* __restrict__ offsets,
                            uint32_t* __restrict__ tile_unit, uint64_t ntiles, uint64_t* __restrict__ counters) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { counters[0] = 0; counters[1] = 0; }
  uint64_t nb = (uint64_t)n_units * W;
  if (i < nb) bitmaps[i] = always ? always[i % W] : 0ull;
  if (i <= ntiles) {
    uint64_t pos = i * TILE;
    uint32_t lo = 0, hi = n_units;   // largest u with offsets[u] <= pos (offsets[0] == 0)
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (offsets[mid] <= pos) lo = mid; else hi = mid;
    }
    tile_unit[i] = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// scan kernel
// ------------------------------------------------------------------------------------------------
struct ScanParams {
  const uint8_t* stream;     // device pointer to stream byte 0 (FRONT_PAD bytes of 0xFF precede it)
  uint64_t nbytes;           // stream length including terminators
  uint64_t ntiles;
  const uint64_t* offsets;
  const uint32_t* tile_unit;
  uint32_t n_units;
  const uint32_t* E;
  cf::DfaTables dfa;
  unsigned long long* bitmaps;
  unsigned long long* counters;
};

struct ScanSmem {
  alignas(128) uint8_t tile[STAGES][TILE + HALO];
  uint32_t E[256];
  alignas(8) uint64_t full[STAGES];
  alignas(8) uint64_t empty[STAGES];
};

__device__ __forceinline__ void verify_candidate(const ScanParams& P, uint64_t start, uint32_t& steps) {
  if (start >= P.nbytes) return;                       // pad region
  const uint8_t* s = P.stream;
  if ((s[start] & 0xC0) == 0x80) return;               // not a character boundary
  uint64_t tix = start / TILE;
  uint32_t lo = P.tile_unit[tix], hi = P.tile_unit[tix + 1] + 1;   // unit in [lo, hi)
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (P.offsets[mid] <= start) lo = mid; else hi = mid;
  }
  uint64_t ustart = P.offsets[lo], uend = P.offsets[lo + 1] - 1;
  const cf::DfaTables& t = P.dfa;
  uint32_t ctx = (start == ustart) ? (uint32_t)cf::P_START : cf::prev_context(t, s, ustart, start);
  uint32_t S = t.start_state[ctx];
  uint64_t q = start;
  while (S != cf::DEAD) {
    uint32_t col, len = 0;
    if (q >= uend) col = t.ncols - 1;
    else col = cf::classify(t, cf::utf8_decode(s, q, uend, &len));
    uint32_t e = t.trans[(uint64_t)S * t.ncols + col];
    uint32_t a = e >> cf::ACC_SHIFT;
    if (a)
      for (uint32_t w = 0; w < t.W; ++w) {
        unsigned long long v = t.accsets[(uint64_t)a * t.W + w];
        if (v) atomicOr(&P.bitmaps[(uint64_t)lo * t.W + w], v);
      }
    S = e & 0xFFFFu;
    ++steps;
    if (q >= uend) break;
    q += len;
  }
}

#define FEED(word, k)                                                        \
  {                                                                          \
    uint32_t b_ = __byte_perm((word), 0, 0x4440 + (k));                      \
    acc = ((acc >> 8) | 0xFF000000u) & sE[b_];                               \
    hit |= acc;                                                              \
  }
#define FEED4(word) FEED(word, 0) FEED(word, 1) FEED(word, 2) FEED(word, 3)
#define FEED16(v) FEED4((v).x) FEED4((v).y) FEED4((v).z) FEED4((v).w)

__global__ void __launch_bounds__(SCAN_WARPS * 32, 3) scan_kernel(const ScanParams P) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ScanSmem& sm = *reinterpret_cast<ScanSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* __restrict__ sE = sm.E;

  sm.E[tid] = P.E[tid];   // blockDim == 256
  if (tid == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], SCAN_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t first = blockIdx.x, stride = gridDim.x;
  // prologue: fill STAGES-1 slots
  if (tid == 0) {
    for (uint32_t j = 0; j < STAGES - 1; ++j) {
      uint64_t tj = first + (uint64_t)j * stride;
      if (tj < P.ntiles) {
        mbar_expect_tx(&sm.full[j], TILE + HALO);
        tma_load_1d(sm.tile[j], P.stream + tj * TILE - HALO, TILE + HALO, &sm.full[j]);
      }
    }
  }

  uint32_t cand_total = 0, step_total = 0;
  uint32_t it = 0;
  for (uint64_t t = first; t < P.ntiles; t += stride, ++it) {
    const uint32_t slot = it % STAGES, phase = (it / STAGES) & 1;
    if (tid == 0) {   // keep STAGES-1 tiles in flight
      uint32_t j = it + STAGES - 1;
      uint64_t tj = first + (uint64_t)j * stride;
      if (tj < P.ntiles) {
        uint32_t sj = j % STAGES;
        if (j >= STAGES) mbar_wait(&sm.empty[sj], ((j / STAGES) - 1) & 1);
        mbar_expect_tx(&sm.full[sj], TILE + HALO);
        tma_load_1d(sm.tile[sj], P.stream + tj * TILE - HALO, TILE + HALO, &sm.full[sj]);
      }
    }
    mbar_wait(&sm.full[slot], phase);

    // this lane's 64 bytes (+ the word holding its 3 look-back bytes)
    const uint32_t chunk = (warp * 32 + lane) * LANE_BYTES;          // tile-relative
    const uint8_t* base = sm.tile[slot] + HALO + chunk;
    const uint32_t back = *reinterpret_cast<const uint32_t*>(base - 4);
    const uint4 v0 = *reinterpret_cast<const uint4*>(base);
    const uint4 v1 = *reinterpret_cast<const uint4*>(base + 16);
    const uint4 v2 = *reinterpret_cast<const uint4*>(base + 32);
    const uint4 v3 = *reinterpret_cast<const uint4*>(base + 48);
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[slot]);   // slot may be refilled: data is in registers

    uint32_t acc = 0, hit = 0;
    FEED(back, 1) FEED(back, 2) FEED(back, 3)
    hit = 0;   // look-back windows belong to the previous lane
    FEED16(v0) FEED16(v1) FEED16(v2) FEED16(v3)

    // rare path: some lane saw an admissible 4-byte window
    uint32_t any = __ballot_sync(0xFFFFFFFFu, (hit & 0xFFu) != 0);
    while (any) {
      const uint32_t src = __ffs(any) - 1;
      any &= any - 1;
      // the whole warp re-filters the 64 bytes of lane `src`, two feed positions per lane
      const uint64_t cbase = t * TILE + (uint64_t)(warp * 32 + src) * LANE_BYTES;
      const uint8_t* g = P.stream + cbase + 2 * lane;
      uint32_t a2 = 0;
      a2 = cf::filter_step(a2, sE[g[-3]]);
      a2 = cf::filter_step(a2, sE[g[-2]]);
      a2 = cf::filter_step(a2, sE[g[-1]]);
      a2 = cf::filter_step(a2, sE[g[0]]);
      const bool c0 = (a2 & 0xFF) != 0;               // start = cbase + 2*lane - 2
      a2 = cf::filter_step(a2, sE[g[1]]);
      const bool c1 = (a2 & 0xFF) != 0;               // start = cbase + 2*lane - 1
      uint32_t steps = 0;
      if (c0) { verify_candidate(P, cbase + 2 * lane - 2, steps); ++cand_total; }
      if (c1) { verify_candidate(P, cbase + 2 * lane - 1, steps); ++cand_total; }
      step_total += steps;
      __syncwarp();
    }
  }
  // statistics (one atomic per warp)
  for (int o = 16; o; o >>= 1) {
    cand_total += __shfl_xor_sync(0xFFFFFFFFu, cand_total, o);
    step_total += __shfl_xor_sync(0xFFFFFFFFu, step_total, o);
  }
  if (lane == 0 && (cand_total | step_total)) {
    atomicAdd(&P.counters[0], (unsigned long long)cand_total);
    atomicAdd(&P.counters[1], (unsigned long long)step_total);
  }
}

// ------------------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------------------
template <typename T>
static int upload(cf_ctx* ctx, const std::vector<T>& v, const T** out, std::vector<void*>& allocs) {
  void* d = nullptr;
  size_t bytes = (v.size() ? v.size() : 1) * sizeof(T);
  CF_CUDA(ctx, cudaMalloc(&d, bytes));
  allocs.push_back(d);
  if (!v.empty()) CF_CUDA(ctx, cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)d;
  return CF_OK;
}

static int upload_dfa(cf_ctx* ctx, const cfre::DfaOut& d, DevDfa& o) {
  int rc;
  if ((rc = upload(ctx, d.ascii_cls, &o.t.ascii_cls, o.allocs))) return rc;
  if ((rc = upload(ctx, d.range_start, &o.t.range_start, o.allocs))) return rc;
  if ((rc = upload(ctx, d.range_cls, &o.t.range_cls, o.allocs))) return rc;
  if ((rc = upload(ctx, d.cls_ctx, &o.t.cls_ctx, o.allocs))) return rc;
  if ((rc = upload(ctx, d.trans, &o.t.trans, o.allocs))) return rc;
  if ((rc = upload(ctx, d.accsets, &o.t.accsets, o.allocs))) return rc;
  o.t.nranges = (uint32_t)d.range_start.size();
  o.t.ncols = d.ncols;
  o.t.W = d.W;
  for (int i = 0; i < 4; ++i) o.t.start_state[i] = d.start_state[i];
  return CF_OK;
}

extern "C" {

int cf_init(int device_ordinal, cf_ctx** out) {
  if (!out) return CF_E_BADARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_ordinal >= n) return CF_E_NOGPU;
  cf_ctx* ctx = new (std::nothrow) cf_ctx();
  if (!ctx) return CF_E_NOMEM;
  ctx->device = device_ordinal;
  *out = ctx;
  CF_CUDA(ctx, cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  CF_CUDA(ctx, cudaGetDeviceProperties(&prop, device_ordinal));
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) { ctx->err = "libcfgpu.so is built for sm_100a (Blackwell) only"; return CF_E_NOGPU; }
  CF_CUDA(ctx, cudaMalloc(&ctx->d_counters, 2 * sizeof(uint64_t)));
  CF_CUDA(ctx, cudaMemset(ctx->d_counters, 0, 2 * sizeof(uint64_t)));
  CF_CUDA(ctx, cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScanSmem)));
  return CF_OK;
}

void cf_shutdown(cf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_counters);
  cudaFree(ctx->d_tile_unit);
  delete ctx;
}

const char* cf_last_error(cf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

uint64_t cf_kernel_launches(const cf_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cf_compile(cf_ctx* ctx, cf_builder* b, cf_prog** out) {
  if (!ctx || !b || !out) return CF_E_BADARG;
  int rc = cf_builder_compile_host(b, nullptr);
  if (rc) { ctx->err = b->err; return rc; }
  for (size_t i = 0; i < b->pats.size(); ++i)
    if (b->ordered[i] && !b->has_repl[i]) { ctx->err = "ordered pattern without replacement"; return CF_E_BADARG; }
  CF_CUDA(ctx, cudaSetDevice(ctx->device));
  cf_prog* p = new (std::nothrow) cf_prog();
  if (!p) return CF_E_NOMEM;
  p->ctx = ctx;
  p->npat = (uint32_t)b->pats.size();
  p->W = b->out.search.W;
  *out = p;
  if ((rc = upload_dfa(ctx, b->out.search, p->search))) return rc;
  CF_CUDA(ctx, cudaMalloc(&p->d_E, 256 * 4));
  CF_CUDA(ctx, cudaMemcpy(p->d_E, b->out.filter.E, 256 * 4, cudaMemcpyHostToDevice));
  CF_CUDA(ctx, cudaMalloc(&p->d_always, p->W * 8));
  CF_CUDA(ctx, cudaMemcpy(p->d_always, b->out.always_bits.data(), p->W * 8, cudaMemcpyHostToDevice));
  for (uint64_t v : b->out.always_bits) if (v) p->any_always = true;
  p->search_empty = true;
  for (int i = 0; i < 256; ++i) if (b->out.filter.E[i]) p->search_empty = false;
  size_t oi = 0;
  for (size_t i = 0; i < b->pats.size(); ++i) {
    if (!b->ordered[i]) continue;
    p->ordered.emplace_back();
    if ((rc = upload_dfa(ctx, b->out.ordered[oi], p->ordered.back()))) return rc;
    uint32_t* dE = nullptr;
    CF_CUDA(ctx, cudaMalloc(&dE, 256 * 4));
    CF_CUDA(ctx, cudaMemcpy(dE, b->out.ordered_filter[oi].E, 256 * 4, cudaMemcpyHostToDevice));
    p->d_ordered_E.push_back(dE);
    p->ordered_pat.push_back((int)i);
    uint8_t* dr = nullptr;
    size_t rl = b->repl[i].size();
    CF_CUDA(ctx, cudaMalloc(&dr, rl ? rl : 1));
    if (rl) CF_CUDA(ctx, cudaMemcpy(dr, b->repl[i].data(), rl, cudaMemcpyHostToDevice));
    p->d_repl.push_back(dr);
    p->repl_len.push_back((uint32_t)rl);
    ++oi;
  }
  return CF_OK;
}

void cf_free_prog(cf_prog* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  for (void* a : p->search.allocs) cudaFree(a);
  for (auto& d : p->ordered) for (void* a : d.allocs) cudaFree(a);
  for (auto* e : p->d_ordered_E) cudaFree(e);
  for (auto* r : p->d_repl) cudaFree(r);
  cudaFree(p->d_E);
  cudaFree(p->d_always);
  delete p;
}

uint32_t cf_prog_words(const cf_prog* p) { return p ? p->W : 0; }
uint32_t cf_prog_patterns(const cf_prog* p) { return p ? p->npat : 0; }

static uint64_t ntiles_for(uint64_t nbytes) { return (nbytes + 2 + TILE - 1) / TILE; }

int cf_batch_create(cf_ctx* ctx, uint64_t max_stream_bytes, uint32_t max_units, cf_batch** out) {
  if (!ctx || !out) return CF_E_BADARG;
  CF_CUDA(ctx, cudaSetDevice(ctx->device));
  cf_batch* b = new (std::nothrow) cf_batch();
  if (!b) return CF_E_NOMEM;
  b->ctx = ctx;
  *out = b;
  b->cap_bytes = max_stream_bytes;
  b->cap_units = max_units;
  uint64_t total = cf::FRONT_PAD + (ntiles_for(max_stream_bytes) + 1) * TILE;
  CF_CUDA(ctx, cudaMalloc(&b->d_buf, total));
  CF_CUDA(ctx, cudaMemset(b->d_buf, 0xFF, total));
  CF_CUDA(ctx, cudaMalloc(&b->d_offsets, ((uint64_t)max_units + 1) * 8));
  return CF_OK;
}

void cf_batch_free(cf_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaFree(b->d_buf);
  cudaFree(b->d_offsets);
  delete b;
}

uint32_t cf_batch_units(const cf_batch* b) { return b ? b->n : 0; }
uint64_t cf_batch_bytes(const cf_batch* b) { return b ? b->nbytes : 0; }

int cf_batch_upload(cf_ctx* ctx, cf_batch* b, const uint8_t* stream, uint64_t stream_bytes,
                    const uint64_t* offsets, uint32_t n_units, void* cuda_stream) {
  if (!ctx || !b || !stream || !offsets || !n_units) return CF_E_BADARG;
  if (stream_bytes > b->cap_bytes || n_units > b->cap_units) { ctx->err = "batch capacity exceeded"; return CF_E_CAPACITY; }
  if (offsets[0] != 0 || offsets[n_units] != stream_bytes) { ctx->err = "offsets[0] must be 0 and offsets[n] the stream length"; return CF_E_BADARG; }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  uint8_t* d_stream = b->d_buf + cf::FRONT_PAD;
  CF_CUDA(ctx, cudaMemcpyAsync(d_stream, stream, stream_bytes, cudaMemcpyHostToDevice, st));
  // re-arm the tail padding that a previous, longer upload may have overwritten
  uint64_t end = (ntiles_for(b->nbytes > stream_bytes ? b->nbytes : stream_bytes) + 1) * TILE;
  CF_CUDA(ctx, cudaMemsetAsync(d_stream + stream_bytes, 0xFF, end - stream_bytes, st));
  CF_CUDA(ctx, cudaMemcpyAsync(b->d_offsets, offsets, ((uint64_t)n_units + 1) * 8, cudaMemcpyHostToDevice, st));
  b->nbytes = stream_bytes;
  b->n = n_units;
  return CF_OK;
}

int cf_scan(cf_ctx* ctx, cf_prog* p, cf_batch* b, uint64_t* d_bitmaps, void* cuda_stream) {
  if (!ctx || !p || !b || !d_bitmaps || !b->n) return CF_E_BADARG;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  uint64_t ntiles = ntiles_for(b->nbytes);
  if (ntiles + 2 > ctx->tile_unit_cap) {
    CF_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(ctx->d_tile_unit);
    ctx->d_tile_unit = nullptr;
    ctx->tile_unit_cap = (ntiles + 2) * 2;
    CF_CUDA(ctx, cudaMalloc(&ctx->d_tile_unit, ctx->tile_unit_cap * 4));
  }
  uint64_t work = (uint64_t)b->n * p->W;
  if (ntiles + 1 > work) work = ntiles + 1;
  uint32_t pb = 256;
  prep_kernel<<<(unsigned)((work + pb - 1) / pb), pb, 0, st>>>(d_bitmaps, p->any_always ? p->d_always : nullptr, p->W,
                                                               b->n, b->d_offsets, ctx->d_tile_unit, ntiles,
                                                               ctx->d_counters);
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  if (p->search_empty) return CF_OK;
  ScanParams P;
  P.stream = b->d_buf + cf::FRONT_PAD;
  P.nbytes = b->nbytes;
  P.ntiles = ntiles;
  P.offsets = b->d_offsets;
  P.tile_unit = ctx->d_tile_unit;
  P.n_units = b->n;
  P.E = p->d_E;
  P.dfa = p->search.t;
  P.bitmaps = (unsigned long long*)d_bitmaps;
  P.counters = (unsigned long long*)ctx->d_counters;
  uint64_t grid = (uint64_t)ctx->sm_count * 3;
  if (grid > ntiles) grid = ntiles;
  scan_kernel<<<(unsigned)grid, SCAN_WARPS * 32, sizeof(ScanSmem), st>>>(P);
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  return CF_OK;
}

int cf_scan_host(cf_ctx* ctx, cf_prog* p, cf_batch* b, const uint8_t* stream, uint64_t stream_bytes,
                 const uint64_t* offsets, uint32_t n_units, uint64_t* h_bitmaps) {
  if (!h_bitmaps) return CF_E_BADARG;
  int rc = cf_batch_upload(ctx, b, stream, stream_bytes, offsets, n_units, nullptr);
  if (rc) return rc;
  uint64_t* d_bm = nullptr;
  size_t bytes = (size_t)n_units * p->W * 8;
  CF_CUDA(ctx, cudaMallocAsync((void**)&d_bm, bytes, 0));
  rc = cf_scan(ctx, p, b, d_bm, nullptr);
  if (rc == CF_OK) {
    cudaError_t e = cudaMemcpyAsync(h_bitmaps, d_bm, bytes, cudaMemcpyDeviceToHost, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = CF_E_CUDA; }
  }
  cudaFreeAsync(d_bm, 0);
  return rc;
}

int cf_scan_counters(cf_ctx* ctx, uint64_t out[2]) {
  if (!ctx || !out) return CF_E_BADARG;
  CF_CUDA(ctx, cudaMemcpy(out, ctx->d_counters, 16, cudaMemcpyDeviceToHost));
  return CF_OK;
}

}  // extern "C"
