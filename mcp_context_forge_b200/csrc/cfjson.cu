// cfjson.cu — the JSON kernels of libcfgpu.so and their C-ABI entry points (include/cfgpu.h): structural index
// (json_index.h), toon_encoder (token-parallel json_tp.h + sequential json_toon.h for the units it hands over),
// request_logging_masking (json_mask.h).  Split from cfgpu.cu so that either half rebuilds on its own.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "cf_internal.h"
#include "json_index.h"
#include "json_mask.h"
#include "json_toon.h"
#include "json_tp.h"

// ------------------------------------------------------------------------------------------------
// JSON structural index (csrc/json_index.h), one WARP per unit:
//   stage 1   one ballot per byte class and 32 bytes (bytes prefetched eight chunks ahead): escape parity,
//             in-string mask by prefix XOR, structural characters, scalar starts -> token positions
//   stage 1b  (CF_INDEX_CLASSIFY) lane-parallel over the tokens: strings validated + their predicates/hash,
//             scalars validated — the same functions the sequential parser uses
// Measured as a front end of the TOON / masking kernels (index kernel + token-driven DOM build per lane)
// it LOSES to the sequential per-lane parser on B200 at large batches (15.8 vs 10.1 ms for 32 768 x 16 KiB:
// stage 1b is issue-bound at ~12 warp-instructions per byte and the tokens triple the memory traffic), so
// those kernels keep json_parse; the index stands alone as a reusable op (string extraction, length
// guards) and as the first stage of the token-parallel design the next round needs.
// ------------------------------------------------------------------------------------------------
static const uint32_t IDX_AHEAD = 8;     // chunks of 32 bytes in flight per warp
__device__ __forceinline__ uint32_t warp_index(const uint8_t* __restrict__ s, uint32_t n, cfx::Tok* __restrict__ tk, bool* unterminated,
                                               uint32_t lane, const uint8_t* __restrict__ cls) {
  cfx::IndexCarry cy;
  cy.init();
  uint32_t nt = 0;
  const uint32_t below = (1u << lane) - 1u;
  for (uint32_t base0 = 0; base0 < n; base0 += 32 * IDX_AHEAD) {
    uint32_t cs[IDX_AHEAD];
#pragma unroll
    for (uint32_t k = 0; k < IDX_AHEAD; ++k) {
      const uint32_t p = base0 + 32 * k + lane;
      cs[k] = p < n ? (uint32_t)s[p] : (uint32_t)' ';
    }
#pragma unroll
    for (uint32_t k = 0; k < IDX_AHEAD; ++k) {
      const uint32_t base = base0 + 32 * k;
      if (base >= n) break;
      const uint32_t kc = cls[cs[k]];            // byte class from a 256-byte shared table: 1 LDS instead of ~14 compares
      const uint32_t bs = __ballot_sync(0xFFFFFFFFu, (kc & 1u) != 0);
      const uint32_t qm = __ballot_sync(0xFFFFFFFFu, (kc & 2u) != 0);
      const uint32_t st = __ballot_sync(0xFFFFFFFFu, (kc & 4u) != 0);
      const uint32_t ws = __ballot_sync(0xFFFFFFFFu, (kc & 8u) != 0);
      uint32_t esc = 0;
      if (bs | cy.bs_parity) {
        esc = __ballot_sync(0xFFFFFFFFu, cfx::escaped_bit(bs, lane, cy.bs_parity) != 0);
        cy.bs_parity = cfx::next_bs_parity(bs, cy.bs_parity);
      }
      uint32_t close;
      const uint32_t tm = cfx::index_chunk(qm & ~esc, st, ws, cy, &close);
      if ((tm >> lane) & 1u) {
        cfx::Tok t;
        t.pos = (base + lane) | (((close >> lane) & 1u) ? cfx::T_CLOSE : 0u);
        t.aux = 0;
        tk[nt + __popc(tm & below)] = t;
      }
      nt += __popc(tm);
    }
  }
  *unterminated = cy.in_string != 0;
  return nt;
}

static const uint32_t NTOK_UNTERMINATED = 0x80000000u;
// tokens of unit u at toks + offsets[u] (capacity len + 1: offsets count one terminator per unit); ntok[u] = count | flag
__global__ void __launch_bounds__(128) json_index_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ offsets, uint32_t n_units,
                                                          cfx::Tok* __restrict__ toks, uint32_t* __restrict__ ntok, uint64_t max_len,
                                                          uint32_t flags) {
  __shared__ uint8_t cls[256];       // bit 0 backslash, 1 quote, 2 structural, 3 JSON whitespace
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x)
    cls[i] = (uint8_t)((i == '\\' ? 1u : 0u) | (i == '"' ? 2u : 0u) | (cfx::is_structural(i) ? 4u : 0u) | (cfj::j_ws(i) ? 8u : 0u));
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_units) return;
  const uint64_t b = offsets[u];
  const uint64_t len64 = offsets[u + 1] - b - 1;
  if (len64 > max_len) { if (lane == 0) ntok[u] = 0; return; }
  const uint32_t len = (uint32_t)len64;
  cfx::Tok* tk = toks + b;
  bool unt;
  const uint32_t nt = warp_index(stream + b, len, tk, &unt, lane, cls);
  __syncwarp();
  if (flags & CF_INDEX_CLASSIFY)
    for (uint32_t t = lane; t < nt; t += 32) tk[t].aux = cfx::classify_token(stream + b, len, tk[t].pos);
  if (lane == 0) ntok[u] = nt | (unt ? NTOK_UNTERMINATED : 0u);
}

// ------------------------------------------------------------------------------------------------
// toon_encoder, token-parallel (csrc/json_tp.h): one WARP per unit, two passes over the unit's bytes (the second
// one comes from L2), 8-byte tokens through HBM scratch, no DOM.  Units the fast path does not cover get status
// TS_FALLBACK (reason in out_len) and are re-done by the sequential toon_kernel below (TOON_ONLY_FALLBACK).
// ------------------------------------------------------------------------------------------------
static const uint32_t TP_WARPS = 8;
static const uint32_t TP_TOK_SLACK = 64;               // token capacity of unit u: len/2 + TP_TOK_SLACK
static const uint32_t TP_WARP_SMEM = (uint32_t)sizeof(cftp::Shared) + cftp::STAGE;   // container stack + token ring | staging buffer
static const uint32_t TP_SMEM = TP_WARPS * TP_WARP_SMEM;                              // 110 592 B: two CTAs per SM
__global__ void __launch_bounds__(TP_WARPS * 32, 2) toon_tp_kernel(const uint8_t* __restrict__ stream, const uint64_t* __restrict__ offsets, uint32_t n_units,
                                                                    cftp::GTok* __restrict__ toks, uint8_t* __restrict__ out, uint32_t* __restrict__ out_len,
                                                                    int32_t* __restrict__ status, uint32_t flags, const uint8_t* __restrict__ unit_stages) {
  extern __shared__ __align__(16) uint8_t tp_smem[];
  const uint32_t lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const uint32_t u = blockIdx.x * TP_WARPS + wic;
  if (u >= n_units) return;
  if (unit_stages && !(unit_stages[u] & CF_STAGE_TOON)) { if (lane == 0) { status[u] = CF_TOON_SKIPPED; out_len[u] = 0; } return; }
  cftp::Shared& sh = *reinterpret_cast<cftp::Shared*>(tp_smem + (size_t)wic * TP_WARP_SMEM);
  uint8_t* stage = tp_smem + (size_t)wic * TP_WARP_SMEM + sizeof(cftp::Shared);
  if (lane == 0) {                                   // the warp's mbarrier for its bulk TMA loads into the staging buffer
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sh.sbar_bar)));
    sh.sbar_phase = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t b = offsets[u];
  const uint64_t len64 = offsets[u + 1] - b - 1;
  if (len64 > 0x7FFFFFFFull) { if (lane == 0) { status[u] = cfj::TS_UNSUPPORTED; out_len[u] = 0; } return; }
  const uint32_t len = (uint32_t)len64;
  cftp::GTok* my = toks + (b >> 1) + (uint64_t)TP_TOK_SLACK * u;
  uint32_t ol = 0;
  const int st = cftp::toon_unit(stream + b, len, my, len / 2 + TP_TOK_SLACK, out + b, len ? len - 1 : 0, &ol, sh, stage, (flags & 1u) != 0);
  if (lane == 0) {
    status[u] = st & 0xFF;
    out_len[u] = (st & 0xFF) == cfj::TS_CONVERTED ? ol : (uint32_t)st >> 8;
  }
}

// gather per-unit results (unit u at src + mul*offsets[u] + add*u, out_len[u] bytes) into one contiguous buffer
__global__ void compact_kernel(const uint8_t* __restrict__ src, uint32_t mul, uint32_t add, const uint64_t* __restrict__ offsets,
                               const uint32_t* __restrict__ out_len, const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out,
                               uint32_t n_units) {
  const uint32_t u = blockIdx.x;
  if (u >= n_units) return;
  const uint8_t* s = src + (uint64_t)mul * offsets[u] + (uint64_t)add * u;
  uint8_t* dst = out + out_off[u];
  for (uint32_t i = threadIdx.x; i < out_len[u]; i += blockDim.x) dst[i] = s[i];
}



int cf_dev_reserve(cf_ctx* ctx, cf_ctx::DevBuf& b, size_t need) {
  if (need <= b.cap) return CF_OK;
  cudaFree(b.p);
  b.p = nullptr; b.cap = 0;
  const size_t c = need + need / 4 + 256;
  CF_CUDA(ctx, cudaMalloc(&b.p, c));
  b.cap = c;
  return CF_OK;
}
int cf_stage_reserve(cf_ctx* ctx, size_t need) {
  if (need <= ctx->h_stage_bytes) return CF_OK;
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  ctx->h_stage = nullptr; ctx->h_stage_bytes = 0;
  const size_t c = need + need / 4 + 4096;
  CF_CUDA(ctx, cudaHostAlloc(&ctx->h_stage, c, cudaHostAllocDefault));
  ctx->h_stage_bytes = c;
  return CF_OK;
}
// units per warp for the thread-per-unit JSON kernels: fill the GPU with warps first (about 12 resident
// warps per SM at their register footprint), only then put several units into one warp
static uint32_t units_per_warp(const cf_ctx* ctx, uint32_t n) {
  const uint32_t warps = (uint32_t)ctx->sm_count * 12u;
  uint32_t u = 1;
  while (u < 32 && (n + u - 1) / u > warps) u <<= 1;
  return u;
}
static uint32_t json_blocks(uint32_t n, uint32_t upw) { return ((n + upw - 1) / upw + 1) / 2; }   // two warps per block
extern "C" {
int cf_json_index(cf_ctx* ctx, cf_batch* b, uint32_t flags, cf_json_token* d_tokens, uint32_t* d_counts, void* cuda_stream) {
  if (!ctx || !b || !d_tokens || !d_counts) return CF_E_BADARG;
  if (b->n == 0) return CF_OK;
  static_assert(sizeof(cf_json_token) == sizeof(cfx::Tok), "token layout");
  json_index_kernel<<<(b->n + 3) / 4, 128, 0, (cudaStream_t)cuda_stream>>>(b->d_buf + cf::FRONT_PAD, b->d_offsets, b->n, (cfx::Tok*)d_tokens, d_counts,
                                                                            0x7FFFFFFFull, flags);
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  return CF_OK;
}

int cf_json_index_host(cf_ctx* ctx, cf_batch* b, uint32_t flags, const uint8_t* stream, uint64_t stream_bytes, const uint64_t* offsets,
                       uint32_t n_units, cf_json_token* tokens, uint32_t* counts) {
  if (!ctx || !b || !tokens || !counts) return CF_E_BADARG;
  int rc = cf_batch_upload(ctx, b, stream, stream_bytes, offsets, n_units, nullptr);
  if (rc) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->d_tok, (size_t)(stream_bytes + 64) * sizeof(cfx::Tok)))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->d_ntok, (size_t)(n_units + 1) * 4))) return rc;
  if ((rc = cf_json_index(ctx, b, flags, (cf_json_token*)ctx->d_tok.p, (uint32_t*)ctx->d_ntok.p, nullptr))) return rc;
  CF_CUDA(ctx, cudaMemcpy(counts, ctx->d_ntok.p, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
  CF_CUDA(ctx, cudaMemcpy(tokens, ctx->d_tok.p, (size_t)stream_bytes * sizeof(cfx::Tok), cudaMemcpyDeviceToHost));
  return CF_OK;
}

static int toon_launch(cf_ctx* ctx, cf_batch* b, uint32_t flags, uint8_t* d_out, uint32_t* d_out_len, int32_t* d_status, const uint8_t* d_unit_stages,
                       cudaStream_t st) {
  uint64_t need = (b->nbytes / 2 + 4ull * b->n + 8) * sizeof(cfj::JNode);
  const uint64_t need_tp = (b->nbytes / 2 + (uint64_t)TP_TOK_SLACK * b->n + 8) * sizeof(cftp::GTok);
  if (need_tp > need) need = need_tp;
  if (need > ctx->toon_scratch_bytes) {
    CF_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(ctx->d_toon_scratch);
    ctx->d_toon_scratch = nullptr;
    ctx->toon_scratch_bytes = 0;
    CF_CUDA(ctx, cudaMalloc(&ctx->d_toon_scratch, need + need / 4));
    ctx->toon_scratch_bytes = need + need / 4;
  }
  const bool prof = ctx->prof_on && (size_t)ctx->prof_used + 2 <= ctx->prof_ev.size();
  if (prof) cudaEventRecord(ctx->prof_ev[ctx->prof_used], st);
  if (!(flags & (CF_TOON_PARSE_ONLY | CF_TOON_SEQUENTIAL))) {
    static bool smem_set = false;
    if (!smem_set) { CF_CUDA(ctx, cudaFuncSetAttribute(toon_tp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM)); smem_set = true; }
    toon_tp_kernel<<<(b->n + TP_WARPS - 1) / TP_WARPS, TP_WARPS * 32, TP_SMEM, st>>>(b->d_buf + cf::FRONT_PAD, b->d_offsets, b->n, (cftp::GTok*)ctx->d_toon_scratch,
                                                                            d_out, d_out_len, d_status, flags, d_unit_stages);
    ctx->launches++;
    CF_CUDA(ctx, cudaGetLastError());
    // the units the fast path handed over: sequential encoder, one unit per warp (they are few)
    if (!(flags & CF_TOON_NO_HANDOVER)) cf_launch_toon_seq(json_blocks(b->n, 1), st, b->d_buf + cf::FRONT_PAD, b->d_offsets, b->n, (cfj::JNode*)ctx->d_toon_scratch, d_out, d_out_len, d_status,
                                                     (flags & 1u) | TOON_ONLY_FALLBACK, 1);
  } else {
    if (d_unit_stages) { ctx->err = "per-unit stage masks need the token-parallel encoder"; return CF_E_BADARG; }
    const uint32_t upw = units_per_warp(ctx, b->n);
    cf_launch_toon_seq(json_blocks(b->n, upw), st, b->d_buf + cf::FRONT_PAD, b->d_offsets, b->n, (cfj::JNode*)ctx->d_toon_scratch, d_out, d_out_len,
                                                       d_status, flags & ~TOON_ONLY_FALLBACK, upw);
  }
  if (prof) { cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], st); ctx->prof_used += 2; }
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  return CF_OK;
}

int cf_toon(cf_ctx* ctx, cf_batch* b, uint32_t flags, uint8_t* d_out, uint32_t* d_out_len, int32_t* d_status, void* cuda_stream) {
  if (!ctx || !b || !b->n || !d_out || !d_out_len || !d_status) return CF_E_BADARG;
  return toon_launch(ctx, b, flags, d_out, d_out_len, d_status, nullptr, (cudaStream_t)cuda_stream);
}

int cf_chain(cf_ctx* ctx, cf_prog* prog, cf_batch* b, uint32_t stage_mask, uint32_t toon_flags, uint64_t* d_bitmaps, const uint8_t* d_unit_stages,
             uint8_t* d_out, uint32_t* d_out_len, int32_t* d_status, void* cuda_stream) {
  if (!ctx || !b || !b->n) return CF_E_BADARG;
  if (stage_mask & ~(CF_STAGE_SCAN | CF_STAGE_TOON)) { ctx->err = "cf_chain runs CF_STAGE_SCAN / CF_STAGE_TOON; the other stages need cf_run_batch"; return CF_E_BADARG; }
  int rc;
  if (stage_mask & CF_STAGE_SCAN) {
    if (!prog || !d_bitmaps) return CF_E_BADARG;
    if ((rc = cf_scan(ctx, prog, b, d_bitmaps, cuda_stream))) return rc;
  }
  if (stage_mask & CF_STAGE_TOON) {
    if (!d_out || !d_out_len || !d_status) return CF_E_BADARG;
    if ((rc = toon_launch(ctx, b, toon_flags, d_out, d_out_len, d_status, d_unit_stages, (cudaStream_t)cuda_stream))) return rc;
  }
  return CF_OK;
}

int cf_toon_host(cf_ctx* ctx, cf_batch* b, uint32_t flags, const uint8_t* stream, uint64_t stream_bytes, const uint64_t* offsets,
                 uint32_t n_units, uint8_t* out_stream, uint32_t* out_len, int32_t* status) {
  if (!out_stream || !out_len || !status) return CF_E_BADARG;
  int rc = cf_batch_upload(ctx, b, stream, stream_bytes, offsets, n_units, nullptr);
  if (rc) return rc;
  // device: encode in the input's layout, then gather the converted texts so that only they cross PCIe
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[0], stream_bytes + 16))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[1], (size_t)n_units * 4))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[2], (size_t)n_units * 4))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[3], ((size_t)n_units + 1) * 8))) return rc;
  uint8_t* d_out = (uint8_t*)ctx->tmp[0].p;
  uint32_t* d_len = (uint32_t*)ctx->tmp[1].p;
  int32_t* d_st = (int32_t*)ctx->tmp[2].p;
  uint64_t* d_ooff = (uint64_t*)ctx->tmp[3].p;
  rc = cf_toon(ctx, b, flags, d_out, d_len, d_st, nullptr);
  if (rc) return rc;
  CF_CUDA(ctx, cudaMemcpy(out_len, d_len, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
  CF_CUDA(ctx, cudaMemcpy(status, d_st, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
  if (flags & CF_TOON_PARSE_ONLY) return CF_OK;
  std::vector<uint64_t> ooff((size_t)n_units + 1);
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_units; ++i) { ooff[i] = total; total += out_len[i]; }
  ooff[n_units] = total;
  if (!total) return CF_OK;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[4], total))) return rc;
  if ((rc = cf_stage_reserve(ctx, total))) return rc;
  CF_CUDA(ctx, cudaMemcpy(d_ooff, ooff.data(), ((size_t)n_units + 1) * 8, cudaMemcpyHostToDevice));
  compact_kernel<<<n_units, 128>>>(d_out, 1, 0, b->d_offsets, d_len, d_ooff, (uint8_t*)ctx->tmp[4].p, n_units);
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  CF_CUDA(ctx, cudaMemcpy(ctx->h_stage, ctx->tmp[4].p, total, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n_units; ++i)
    if (out_len[i]) memcpy(out_stream + offsets[i], (const uint8_t*)ctx->h_stage + ooff[i], out_len[i]);
  return CF_OK;
}

static int cf_mask_resident(cf_ctx* ctx, cf_batch* b, int max_depth, uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets, int32_t* status,
                            uint64_t* out_needed);
int cf_mask_host(cf_ctx* ctx, cf_batch* b, const uint8_t* stream, uint64_t stream_bytes, const uint64_t* offsets, uint32_t n_units,
                 int max_depth, uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets, int32_t* status, uint64_t* out_needed) {
  if (!ctx || !b || !out_offsets || !status) return CF_E_BADARG;
  int rc = cf_batch_upload(ctx, b, stream, stream_bytes, offsets, n_units, nullptr);
  if (rc) return rc;
  return cf_mask_resident(ctx, b, max_depth, out_bytes, out_cap, out_offsets, status, out_needed);
}
// masking of the batch already uploaded
static int cf_mask_resident(cf_ctx* ctx, cf_batch* b, int max_depth, uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets, int32_t* status,
                            uint64_t* out_needed) {
  int rc;
  const uint64_t stream_bytes = b->nbytes;
  const uint32_t n_units = b->n;
  const uint64_t nnodes = stream_bytes / 2 + 4ull * n_units + 8;
  const uint64_t need = nnodes * sizeof(cfj::JNode);
  if (need > ctx->toon_scratch_bytes) {
    cudaFree(ctx->d_toon_scratch);
    ctx->d_toon_scratch = nullptr;
    ctx->toon_scratch_bytes = 0;
    CF_CUDA(ctx, cudaMalloc(&ctx->d_toon_scratch, need + need / 4));
    ctx->toon_scratch_bytes = need + need / 4;
  }
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[0], 5 * stream_bytes + 32ull * n_units + 64))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[1], (size_t)n_units * 4))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[2], (size_t)n_units * 4))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[3], ((size_t)n_units + 1) * 8))) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[5], nnodes * 4))) return rc;
  uint8_t* d_arena = (uint8_t*)ctx->tmp[0].p;
  uint32_t* d_len = (uint32_t*)ctx->tmp[1].p;
  int32_t* d_st = (int32_t*)ctx->tmp[2].p;
  uint64_t* d_ooff = (uint64_t*)ctx->tmp[3].p;
  uint32_t* d_idx = (uint32_t*)ctx->tmp[5].p;
  std::vector<uint32_t> lens(n_units);
  const uint32_t upw = units_per_warp(ctx, n_units);
  cf_launch_mask_seq(json_blocks(n_units, upw), b->d_buf + cf::FRONT_PAD, b->d_offsets, n_units, (cfj::JNode*)ctx->d_toon_scratch, d_idx, d_arena, d_len,
                                                 d_st, max_depth, upw);
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  CF_CUDA(ctx, cudaMemcpy(lens.data(), d_len, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
  CF_CUDA(ctx, cudaMemcpy(status, d_st, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_units; ++i) { out_offsets[i] = total; total += lens[i]; }
  out_offsets[n_units] = total;
  if (out_needed) *out_needed = total;
  if (total > out_cap || (!out_bytes && total)) { ctx->err = "output buffer too small"; return CF_E_CAPACITY; }
  if (total) {
    if ((rc = cf_dev_reserve(ctx, ctx->tmp[4], total))) return rc;
    if ((rc = cf_stage_reserve(ctx, total))) return rc;
    CF_CUDA(ctx, cudaMemcpy(d_ooff, out_offsets, ((size_t)n_units + 1) * 8, cudaMemcpyHostToDevice));
    compact_kernel<<<n_units, 128>>>(d_arena, 5, 32, b->d_offsets, d_len, d_ooff, (uint8_t*)ctx->tmp[4].p, n_units);
    ctx->launches++;
    CF_CUDA(ctx, cudaGetLastError());
    CF_CUDA(ctx, cudaMemcpy(ctx->h_stage, ctx->tmp[4].p, total, cudaMemcpyDeviceToHost));
    memcpy(out_bytes, ctx->h_stage, total);
  }
  return CF_OK;
}

int cf_classify_keys_host(cf_ctx* ctx, cf_batch* b, const uint8_t* stream, uint64_t stream_bytes, const uint64_t* offsets, uint32_t n_units,
                          uint8_t* sensitive) {
  if (!ctx || !b || !sensitive) return CF_E_BADARG;
  int rc = cf_batch_upload(ctx, b, stream, stream_bytes, offsets, n_units, nullptr);
  if (rc) return rc;
  if ((rc = cf_dev_reserve(ctx, ctx->tmp[6], n_units))) return rc;
  uint8_t* d = (uint8_t*)ctx->tmp[6].p;
  cf_launch_classify_keys((n_units + 127) / 128, b->d_buf + cf::FRONT_PAD, b->d_offsets, n_units, d);
  ctx->launches++;
  CF_CUDA(ctx, cudaGetLastError());
  CF_CUDA(ctx, cudaMemcpy(sensitive, d, n_units, cudaMemcpyDeviceToHost));
  return CF_OK;
}


// ---- the fused chain with host buffers (include/cfgpu.h): one H2D of the stream, every stage on the resident batch, then
// verdicts + only the produced texts cross PCIe back
int cf_run_batch(cf_ctx* ctx, cf_prog* prog, cf_batch* b, const uint8_t* stream, uint64_t stream_bytes, const uint64_t* offsets, uint32_t n_units,
                 uint32_t stage_mask, const uint8_t* unit_stages, uint32_t toon_flags, int mask_max_depth, cf_verdict* verdicts, uint64_t* bitmaps_full,
                 uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets, uint64_t* out_needed) {
  if (!ctx || !b || !offsets || !n_units || !verdicts || !out_offsets) return CF_E_BADARG;
  if ((stage_mask & (CF_STAGE_SCAN | CF_STAGE_SUB)) && !prog) return CF_E_BADARG;
  if ((stage_mask & CF_STAGE_TOON) && (stage_mask & CF_STAGE_MASK)) { ctx->err = "CF_STAGE_TOON and CF_STAGE_MASK both produce the unit's output: two calls"; return CF_E_BADARG; }
  if (stage_mask & CF_STAGE_SUB) stage_mask |= CF_STAGE_SCAN;
  struct Nvtx { Nvtx(const char* n) { nvtxRangePushA(n); } ~Nvtx() { nvtxRangePop(); } } nvtx_call("cf_run_batch");   // ranges: assemble (caller) | h2d | kernels | d2h
  int rc = CF_OK;
  if (stream) {
    nvtxRangePushA("cf_run_batch:h2d");
    rc = cf_batch_upload(ctx, b, stream, stream_bytes, offsets, n_units, nullptr);
    nvtxRangePop();
  } else if (b->n != n_units || b->nbytes != stream_bytes) { ctx->err = "resident run: the batch on the device is a different one"; return CF_E_BADARG; }
  if (rc) return rc;
  const uint32_t W = prog ? prog->W : 1;
  std::vector<uint64_t> bm;
  std::vector<uint32_t> tlen;
  std::vector<int32_t> tst;
  uint8_t* d_us = nullptr;
  if (unit_stages) {
    if ((rc = cf_dev_reserve(ctx, ctx->tmp[7], n_units))) return rc;
    d_us = (uint8_t*)ctx->tmp[7].p;
    CF_CUDA(ctx, cudaMemcpyAsync(d_us, unit_stages, n_units, cudaMemcpyHostToDevice, 0));
  }
  // ---- launches, back to back
  nvtxRangePushA("cf_run_batch:kernels");
  if (stage_mask & CF_STAGE_SCAN) {
    if ((rc = cf_dev_reserve(ctx, ctx->tmp[6], (size_t)n_units * W * 8))) return rc;
    if ((rc = cf_scan(ctx, prog, b, (uint64_t*)ctx->tmp[6].p, nullptr))) return rc;
  }
  if (stage_mask & CF_STAGE_TOON) {
    if ((rc = cf_dev_reserve(ctx, ctx->tmp[0], stream_bytes + 16))) return rc;
    if ((rc = cf_dev_reserve(ctx, ctx->tmp[1], (size_t)n_units * 4))) return rc;
    if ((rc = cf_dev_reserve(ctx, ctx->tmp[2], (size_t)n_units * 4))) return rc;
    if ((rc = toon_launch(ctx, b, toon_flags & ~(CF_TOON_PARSE_ONLY | CF_TOON_SEQUENTIAL | CF_RUN_OUTPUTS_RESIDENT), (uint8_t*)ctx->tmp[0].p, (uint32_t*)ctx->tmp[1].p, (int32_t*)ctx->tmp[2].p, d_us, 0))) return rc;
  }
  nvtxRangePop();
  // ---- results of the launches
  Nvtx nvtx_d2h("cf_run_batch:d2h+verdicts");
  for (uint32_t i = 0; i < n_units; ++i) { verdicts[i].match_bitmap = 0; verdicts[i].flags = 0; verdicts[i].out_len = 0; verdicts[i].aux = 0; verdicts[i].reserved = 0; }
  std::vector<uint32_t> dirty;
  if (stage_mask & CF_STAGE_SCAN) {
    bm.resize((size_t)n_units * W);
    CF_CUDA(ctx, cudaMemcpy(bm.data(), ctx->tmp[6].p, bm.size() * 8, cudaMemcpyDeviceToHost));
    if (bitmaps_full) memcpy(bitmaps_full, bm.data(), bm.size() * 8);
    std::vector<uint64_t> rule_mask(W, 0);
    for (int pi : prog->ordered_pat) rule_mask[(size_t)pi / 64] |= 1ull << (pi % 64);
    for (uint32_t i = 0; i < n_units; ++i) {
      verdicts[i].match_bitmap = bm[(size_t)i * W];
      if ((stage_mask & CF_STAGE_SUB) && (!unit_stages || (unit_stages[i] & CF_STAGE_SUB))) {
        bool d = false;
        for (uint32_t w = 0; w < W; ++w) if (bm[(size_t)i * W + w] & rule_mask[w]) { d = true; break; }
        if (d) dirty.push_back(i);
      }
    }
  }
  if (stage_mask & CF_STAGE_TOON) {
    tlen.resize(n_units); tst.resize(n_units);
    CF_CUDA(ctx, cudaMemcpy(tlen.data(), ctx->tmp[1].p, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
    CF_CUDA(ctx, cudaMemcpy(tst.data(), ctx->tmp[2].p, (size_t)n_units * 4, cudaMemcpyDeviceToHost));
  }
  // ---- regex_filter rewriting of the (few) units a rule matched
  std::vector<uint8_t> sub_bytes;
  std::vector<uint64_t> sub_off;
  if (!dirty.empty()) {
    sub_off.assign(dirty.size() + 1, 0);
    uint64_t need = 0;
    sub_bytes.resize(1 << 16);
    rc = cf_sub_host(ctx, prog, b, dirty.data(), (uint32_t)dirty.size(), sub_bytes.data(), sub_bytes.size(), sub_off.data(), &need);
    if (rc == CF_E_CAPACITY && need > sub_bytes.size()) {
      sub_bytes.resize(need);
      rc = cf_sub_host(ctx, prog, b, dirty.data(), (uint32_t)dirty.size(), sub_bytes.data(), sub_bytes.size(), sub_off.data(), &need);
    }
    if (rc) return rc;
    for (size_t k = 0; k < dirty.size(); ++k) {
      const uint32_t i = dirty[k];
      verdicts[i].flags |= CF_V_REWRITTEN;
      verdicts[i].out_len = (uint32_t)(sub_off[k + 1] - sub_off[k]);
      if ((stage_mask & CF_STAGE_TOON) && (!unit_stages || (unit_stages[i] & CF_STAGE_TOON))) { verdicts[i].flags |= CF_V_RESUBMIT; tst[i] = CF_TOON_SKIPPED; tlen[i] = 0; }
    }
  }
  if (stage_mask & CF_STAGE_TOON)
    for (uint32_t i = 0; i < n_units; ++i) {
      verdicts[i].aux = tst[i];
      if (tst[i] == CF_TOON_CONVERTED && !(verdicts[i].flags & CF_V_REWRITTEN)) { verdicts[i].flags |= CF_V_TOON; verdicts[i].out_len = tlen[i]; }
    }
  // ---- masking on the same upload (sequential kernel; its own gather)
  if (stage_mask & CF_STAGE_MASK) {
    std::vector<int32_t> mst(n_units);
    std::vector<uint64_t> moff((size_t)n_units + 1);
    uint64_t need = 0;
    rc = cf_mask_resident(ctx, b, mask_max_depth, out_bytes, out_cap, moff.data(), mst.data(), &need);
    if (out_needed) *out_needed = need;
    if (rc) return rc;
    for (uint32_t i = 0; i < n_units; ++i) {
      out_offsets[i] = moff[i];
      verdicts[i].aux = mst[i];
      if (mst[i] == CF_MASK_OK) { verdicts[i].flags |= CF_V_MASKED; verdicts[i].out_len = (uint32_t)(moff[i + 1] - moff[i]); }
    }
    out_offsets[n_units] = moff[n_units];
    return CF_OK;
  }
  // ---- pack the outputs: TOON texts gathered on the device (one D2H), rewritten texts from the substitution call
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_units; ++i) { out_offsets[i] = total; total += verdicts[i].out_len; }
  out_offsets[n_units] = total;
  if (out_needed) *out_needed = total;
  const bool keep = (toon_flags & CF_RUN_OUTPUTS_RESIDENT) != 0;
  ctx->run_out = nullptr; ctx->run_out_bytes = 0;
  if (!keep && (total > out_cap || (!out_bytes && total))) { ctx->err = "output buffer too small"; return CF_E_CAPACITY; }
  if (keep && total && (rc = cf_dev_reserve(ctx, ctx->tmp[4], total))) return rc;
  if ((stage_mask & CF_STAGE_TOON) && total) {
    // gather the TOON texts on the device AT THEIR FINAL OFFSETS (rewritten units leave holes), then ONE D2H straight into out_bytes
    std::vector<uint32_t> glen(n_units);
    bool any = false;
    for (uint32_t i = 0; i < n_units; ++i) { glen[i] = (verdicts[i].flags & CF_V_TOON) ? verdicts[i].out_len : 0u; any = any || glen[i]; }
    if (any) {
      if ((rc = cf_dev_reserve(ctx, ctx->tmp[3], ((size_t)n_units + 1) * 8))) return rc;
      if ((rc = cf_dev_reserve(ctx, ctx->tmp[4], total))) return rc;
      CF_CUDA(ctx, cudaMemcpyAsync(ctx->tmp[3].p, out_offsets, ((size_t)n_units + 1) * 8, cudaMemcpyHostToDevice, 0));
      CF_CUDA(ctx, cudaMemcpyAsync(ctx->tmp[1].p, glen.data(), (size_t)n_units * 4, cudaMemcpyHostToDevice, 0));
      compact_kernel<<<n_units, 128>>>((const uint8_t*)ctx->tmp[0].p, 1, 0, b->d_offsets, (const uint32_t*)ctx->tmp[1].p, (const uint64_t*)ctx->tmp[3].p, (uint8_t*)ctx->tmp[4].p, n_units);
      ctx->launches++;
      CF_CUDA(ctx, cudaGetLastError());
      if (!keep) CF_CUDA(ctx, cudaMemcpy(out_bytes, ctx->tmp[4].p, total, cudaMemcpyDeviceToHost));
    }
  }
  for (size_t k = 0; k < dirty.size(); ++k) {
    const uint32_t i = dirty[k];
    if (!verdicts[i].out_len) continue;
    // resident outputs: the rewritten text is still in the substitution call's device buffer (ctx->tmp[14], packed at sub_off)
    if (keep) CF_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ctx->tmp[4].p + out_offsets[i], (const uint8_t*)ctx->tmp[14].p + sub_off[k], verdicts[i].out_len, cudaMemcpyDeviceToDevice, 0));
    else memcpy(out_bytes + out_offsets[i], sub_bytes.data() + sub_off[k], verdicts[i].out_len);
  }
  if (keep) {
    CF_CUDA(ctx, cudaStreamSynchronize(0));
    ctx->run_out = total ? (const uint8_t*)ctx->tmp[4].p : nullptr;
    ctx->run_out_bytes = total;
  }
  return CF_OK;
}

int cf_run_batch_device_output(cf_ctx* ctx, const uint8_t** d_out, uint64_t* bytes) {
  if (!ctx || !d_out || !bytes) return CF_E_BADARG;
  *d_out = ctx->run_out;
  *bytes = ctx->run_out_bytes;
  return CF_OK;
}

int cf_copy_to_host(cf_ctx* ctx, void* host_dst, const void* device_src, uint64_t bytes) {
  if (!ctx || (bytes && (!host_dst || !device_src))) return CF_E_BADARG;
  if (bytes) CF_CUDA(ctx, cudaMemcpy(host_dst, device_src, bytes, cudaMemcpyDeviceToHost));
  return CF_OK;
}

int cf_profile_collect_each(cf_ctx* ctx, double* ms, uint32_t cap, uint32_t* n_launches) {
  if (!ctx || !ms || !n_launches) return CF_E_BADARG;
  uint32_t n = 0;
  for (uint32_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    CF_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[i + 1]));
    float t = 0;
    CF_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_ev[i], ctx->prof_ev[i + 1]));
    if (n < cap) ms[n] = t;
    ++n;
  }
  *n_launches = n;
  ctx->prof_used = 0;
  return CF_OK;
}

}  // extern "C"
