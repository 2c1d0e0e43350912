// host_sim.cpp — CPU simulation of the scan / sub kernels' table-driven algorithm.
// TEST-ONLY: built into tests/hostsim/libcfhostsim.so and loaded only by tests/ (-m "not gpu")
// to validate the compiled tables (prefilter + DFAs, shared scan_core.h routines) against
// CPython `re` on a box with no GPU.  The product library (libcfgpu.so) does not contain this.
#include <string.h>

#include <vector>

#include "../../mcp_context_forge_b200/csrc/cf_host.h"
#include "../../mcp_context_forge_b200/csrc/scan_core.h"

static cf::DfaTables tables_of(const cfre::DfaOut& d) {
  cf::DfaTables t;
  t.ascii_cls = d.ascii_cls.data();
  t.range_start = d.range_start.data();
  t.range_cls = d.range_cls.data();
  t.cls_ctx = d.cls_ctx.data();
  t.trans = d.trans.data();
  t.accsets = d.accsets.data();
  t.nranges = (uint32_t)d.range_start.size();
  t.ncols = d.ncols;
  t.W = d.W;
  for (int i = 0; i < 4; ++i) t.start_state[i] = d.start_state[i];
  return t;
}

extern "C" {

int cfh_scan(cf_builder* b, const uint8_t* stream, uint64_t nbytes, const uint64_t* offsets,
             uint32_t n, uint64_t* bitmaps, uint64_t* stats) {
  int rc = cf_builder_compile_host(b, nullptr);
  if (rc) return rc;
  const cfre::CompileOut& co = b->out;
  cf::DfaTables t = tables_of(co.search);
  uint32_t W = t.W;
  for (uint32_t u = 0; u < n; ++u)
    for (uint32_t w = 0; w < W; ++w) bitmaps[(uint64_t)u * W + w] = co.always_bits[w];
  std::vector<uint8_t> pad(cf::FRONT_PAD + nbytes + 64, cf::TERM);
  memcpy(pad.data() + cf::FRONT_PAD, stream, nbytes);
  const uint8_t* s = pad.data() + cf::FRONT_PAD;   // s[-FRONT_PAD..] valid
  uint32_t acc = 0;
  uint64_t ncand = 0, nsteps = 0;
  // feed from LOOKBACK bytes before the stream to START_OFF bytes after (the last start is nbytes-1)
  for (int64_t p = -(int64_t)cf::F_LOOKBACK; p < (int64_t)nbytes + (int64_t)cf::F_START_OFF; ++p) {
    acc = cf::filter_step(acc, co.filter.E[s[p]]);
    if (!(acc & cf::F_HIT)) continue;
    int64_t start = p - (int64_t)cf::F_START_OFF;
    if (start < 0 || start >= (int64_t)nbytes) continue;
    if ((s[start] & 0xC0) == 0x80) continue;   // not a character boundary
    ++ncand;
    // unit lookup: largest u with offsets[u] <= start
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (offsets[mid] <= (uint64_t)start) lo = mid; else hi = mid; }
    uint64_t ustart = offsets[lo], uend = offsets[lo + 1] - 1;
    nsteps += cf::verify_search(t, s, ustart, uend, (uint64_t)start, bitmaps + (uint64_t)lo * W);
  }
  if (stats) { stats[0] = ncand; stats[1] = nsteps; }
  return CF_OK;
}

// Apply ONE ordered rule (index into the ordered list, in add order) to one unit, Python
// `pattern.sub(repl, unit)` semantics for patterns that cannot match the empty string.
int cfh_sub(cf_builder* b, uint32_t ordered_index, const uint8_t* unit, uint64_t len,
            uint8_t* out, uint64_t cap, uint64_t* out_len, uint64_t* n_matches) {
  int rc = cf_builder_compile_host(b, nullptr);
  if (rc) return rc;
  if (ordered_index >= b->out.ordered.size()) return CF_E_BADARG;
  // find the pattern index of this ordered rule
  uint32_t pat = 0, seen = 0;
  for (; pat < b->ordered.size(); ++pat) if (b->ordered[pat] && seen++ == ordered_index) break;
  const std::vector<uint8_t>& repl = b->repl[pat];
  cf::DfaTables t = tables_of(b->out.ordered[ordered_index]);
  const uint32_t* E = b->out.ordered_filter[ordered_index].E;
  std::vector<uint8_t> pad(cf::FRONT_PAD + len + 64, cf::TERM);
  memcpy(pad.data() + cf::FRONT_PAD, unit, len);
  const uint8_t* s = pad.data() + cf::FRONT_PAD;
  uint64_t o = 0, cur = 0, nm = 0;
  auto put = [&](const uint8_t* p, uint64_t k) {
    if (o + k <= cap) memcpy(out + o, p, k);
    o += k;
  };
  uint32_t acc = 0;
  for (int64_t p = -(int64_t)cf::F_LOOKBACK; p < (int64_t)len + (int64_t)cf::F_START_OFF; ++p) {
    acc = cf::filter_step(acc, E[s[p]]);
    if (!(acc & cf::F_HIT)) continue;
    int64_t start = p - (int64_t)cf::F_START_OFF;
    if (start < 0 || start >= (int64_t)len) continue;
    if ((uint64_t)start < cur) continue;            // inside the previous match
    if ((s[start] & 0xC0) == 0x80) continue;
    uint64_t e = cf::match_first(t, s, 0, len, (uint64_t)start);
    if (e == ~0ull) continue;
    put(s + cur, (uint64_t)start - cur);
    put(repl.data(), repl.size());
    cur = e;
    ++nm;
  }
  put(s + cur, len - cur);
  *out_len = o;
  if (n_matches) *n_matches = nm;
  return o <= cap ? CF_OK : CF_E_CAPACITY;
}

}  // extern "C"

// ---- TOON: run the shared json_toon.h pipeline on the CPU (same code the CUDA kernel calls) ----
#include "../../mcp_context_forge_b200/csrc/json_toon.h"
extern "C" int cfh_toon(const uint8_t* text, uint32_t n, uint8_t* out, uint32_t out_cap, uint32_t* out_len) {
  std::vector<cfj::JNode> nodes(n / 2 + 4);
  cfj::Big big;
  std::vector<uint8_t> digits(1240);
  *out_len = 0;
  return cfj::toon_process(text, n, nodes.data(), (uint32_t)nodes.size(), out, out_cap, out_len, &big, digits.data(), (uint32_t)digits.size(), false);
}

// ---- masking: shared json_mask.h pipeline on the CPU ----
#include "../../mcp_context_forge_b200/csrc/json_mask.h"
extern "C" int cfh_mask(const uint8_t* text, uint32_t n, int max_depth, uint8_t* out, uint32_t out_cap, uint32_t* out_len) {
  std::vector<cfj::JNode> nodes(n / 2 + 4);
  std::vector<uint32_t> idx(n / 2 + 4);
  cfj::Big big;
  std::vector<uint8_t> digits(1240);
  cfm::NumWork w{&big, &big, digits.data(), (uint32_t)digits.size()};
  *out_len = 0;
  return cfm::mask_process(text, n, nodes.data(), (uint32_t)nodes.size(), idx.data(), (uint32_t)idx.size(), out, out_cap, out_len, max_depth, w);
}
extern "C" int cfh_key_sensitive(const uint8_t* key, uint32_t n) {
  cfj::JNode k{cfj::J_KEY, 0, n, 0};   // raw (unescaped) key bytes
  return cfm::key_sensitive(key, k) ? 1 : 0;
}
