// cf_internal.h — objects shared by the translation units of libcfgpu.so (cfgpu.cu: contexts, batches, scan / substitution
// kernels; cfjson.cu: the JSON kernels — structural index, TOON, masking).  Not part of the ABI (include/cfgpu.h is).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "cf_host.h"
#include "scan_core.h"

// ------------------------------------------------------------------------------------------------
// host-side objects
// ------------------------------------------------------------------------------------------------
static const uint32_t LANE_BYTES = 64;
static const uint32_t WARP_BYTES = 32 * LANE_BYTES;          // 2 KiB per warp per tile
static const uint32_t MAX_WARPS = 32;   // padding granularity (>= every variant's tile)
static const uint32_t MAX_TILE = MAX_WARPS * WARP_BYTES;     // buffers are padded for the largest tile

struct cf_ctx {
  int device = 0;
  int sm_count = 0;
  std::string err;
  uint64_t launches = 0;
  uint64_t* d_qstate = nullptr;     // two {candidates appended, verify steps} pairs, used alternately
  uint32_t qphase = 0;
  uint64_t* d_queue = nullptr;      // candidate start positions
  uint32_t qcap = 1u << 20;
  void* d_toon_scratch = nullptr;   // DOM node arrays for toon_kernel (grown on demand)
  uint64_t toon_scratch_bytes = 0;
  struct DevBuf { void* p = nullptr; size_t cap = 0; };
  DevBuf tmp[16];                   // grow-only device scratch of the *_host entry points (no cudaMalloc per call)
  DevBuf d_tok, d_ntok;             // structural index of the current batch (json_index_kernel)
  void* h_stage = nullptr;          // pinned host staging for gathered results
  size_t h_stage_bytes = 0;
  const uint8_t* run_out = nullptr;   // device buffer of the last CF_RUN_OUTPUTS_RESIDENT call
  uint64_t run_out_bytes = 0;
  // optional per-launch timing of the dominant kernel (bench.py roofline): event pairs
  std::vector<cudaEvent_t> prof_ev;
  uint32_t prof_used = 0;
  bool prof_on = false;
  // scan kernel configuration (CF_SCAN_WARPS / CF_SCAN_ACC override the defaults; experiments)
  uint32_t scan_warps = 16;        // best of the measured variants (profiles/README.md)
  uint32_t scan_lane_bytes = 64;
  uint32_t scan_acc = 1;
  uint32_t scan_stages = 3;
  uint32_t scan_reserve_sms = 0;   // CF_SCAN_RESERVE_SMS: SMs the persistent scan grid leaves free
  uint32_t tile() const { return scan_warps * 32 * scan_lane_bytes; }
  uint32_t box_rows() const { uint32_t rows = tile() / 128, nbox = (rows + 255) / 256; return rows / nbox; }
};

struct DevDfa {
  cf::DfaTables t;
  std::vector<void*> allocs;
  uint64_t trans_bytes = 0, acc_bytes = 0, stage_bytes = 0;   // sizes for staging in shared memory
};

struct cf_prog {
  cf_ctx* ctx = nullptr;
  uint32_t npat = 0, W = 1;
  DevDfa search;
  uint32_t* d_E = nullptr;         // byte prefilter E[256], or the pair prefilter's T[PF_SLOTS] when use_pairs
  bool use_pairs = false;
  uint64_t* d_always = nullptr;
  bool any_always = false;
  bool search_empty = false;       // every pattern is "always" -> no automaton work at all
  std::vector<DevDfa> ordered;
  std::vector<uint32_t*> d_ordered_E;
  std::vector<int> ordered_pat;    // pattern index of each ordered rule
  std::vector<uint8_t*> d_repl;
  std::vector<uint32_t> repl_len;
  std::vector<uint32_t> ordered_minlen;   // minimum match length (code points) of each ordered rule
  struct RuleTmpl {                       // replacement template with group references (n_parts == 0: literal replacement)
    uint32_t *d_code = nullptr, *d_sets = nullptr, *d_parts = nullptr;
    uint32_t ninst = 0, wpc = 1, nslots = 2, n_parts = 0, nrefs = 0, lit_len = 0;
  };
  std::vector<RuleTmpl> tmpl;             // per ordered rule
  std::vector<uint64_t> h_offsets;        // host copy of the last batch's offsets (cf_sub_host sizing)
  const void* h_offsets_owner = nullptr;
  uint64_t h_offsets_gen = 0;
};

struct cf_batch {
  cf_ctx* ctx = nullptr;
  uint8_t* d_buf = nullptr;        // FRONT_PAD + stream + tail pad
  uint64_t* d_offsets = nullptr;
  uint32_t* d_coarse = nullptr;    // unit index at every 4 KiB of stream (built on upload)
  std::vector<uint32_t> h_coarse;
  uint64_t cap_bytes = 0;
  uint32_t cap_units = 0;
  uint64_t nbytes = 0;
  uint32_t n = 0;
  uint64_t generation = 0;         // bumped by every upload
  CUtensorMap tmap;                // 2-D view of d_buf: rows of 128 B, box = one scan tile, SWIZZLE_128B
};


#define CF_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
      return CF_E_CUDA;                                                                     \
    }                                                                                       \
  } while (0)


// grow-only device / pinned-host scratch of the *_host entry points (defined in cfjson.cu)
int cf_dev_reserve(cf_ctx* ctx, cf_ctx::DevBuf& b, size_t need);
int cf_stage_reserve(cf_ctx* ctx, size_t need);

// sequential JSON kernels (cfjson_seq.cu)
namespace cfj { struct JNode; }
static const int CF_TS_FALLBACK = 7;                   // == cftp::TS_FALLBACK (json_tp.h): unit handed to the sequential encoder
static const uint32_t TOON_ONLY_FALLBACK = 0x100u;     // internal flag of toon_kernel: only units with status TS_FALLBACK
void cf_launch_toon_seq(uint32_t blocks, cudaStream_t st, const uint8_t* stream, const uint64_t* offsets, uint32_t n_units, cfj::JNode* nodes, uint8_t* out,
                        uint32_t* out_len, int32_t* status, uint32_t flags, uint32_t upw);
void cf_launch_mask_seq(uint32_t blocks, const uint8_t* stream, const uint64_t* offsets, uint32_t n_units, cfj::JNode* nodes, uint32_t* idx, uint8_t* out,
                        uint32_t* out_len, int32_t* status, int max_depth, uint32_t upw);
void cf_launch_classify_keys(uint32_t blocks, const uint8_t* stream, const uint64_t* offsets, uint32_t n_units, uint8_t* sensitive);
