// re_backend.h — host-side regex back-end: serialized AST -> Thompson NFA -> code-point-class
// DFA (+ \b / ^ contexts) -> prefilter table.  The front-end (Python, regex_frontend.py) parses
// patterns with CPython's own sre parser and resolves IGNORECASE / \w / \d / \s to explicit
// code-point interval sets, so this back-end never needs a Unicode database: Python `re`
// semantics (the reference's matcher: plugins/regex_filter/search_replace.py:71,
// plugins/harmful_content_detector/harmful_content_detector.py:76,85) are exact by construction.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace cfre {

// ---- serialized AST (uint32 words, prefix order) ----
//   A_EMPTY
//   A_SET    nranges (lo hi)*          one code point in the (sorted, disjoint, inclusive) ranges
//   A_CAT    n child*                  concatenation
//   A_ALT    n child*                  ordered alternation (leftmost-first)
//   A_REPEAT min max greedy child      max = 0xFFFFFFFF means unbounded
//   A_ASSERT kind
//   A_GROUP  index child               capturing group (emitted for substitution rules only; index >= 1)
enum : uint32_t { A_EMPTY = 0, A_SET = 1, A_CAT = 2, A_ALT = 3, A_REPEAT = 4, A_ASSERT = 5, A_GROUP = 6 };
enum : uint32_t {
  AS_WORD_B = 1, AS_NOT_WORD_B = 2, AS_BEGIN_STRING = 3, AS_BEGIN_LINE = 4,
  AS_END_STRING = 5, AS_END_LINE = 6, AS_END_DOLLAR = 7
};
static const uint32_t REPEAT_INF = 0xFFFFFFFFu;

struct Interval { uint32_t lo, hi; };
typedef std::vector<Interval> CharSet;

struct PatternIn {
  std::vector<uint32_t> ast;
};

// One compiled anchored DFA, host copy (flat arrays mirror cf::DfaTables).
struct DfaOut {
  std::vector<uint16_t> ascii_cls;    // 128
  std::vector<uint32_t> range_start;  // non-ASCII ranges
  std::vector<uint16_t> range_cls;
  std::vector<uint8_t>  cls_ctx;      // per class
  std::vector<uint32_t> trans;        // nstates * ncols
  std::vector<uint64_t> accsets;      // naccs * W
  uint32_t ncols = 0, nstates = 0, W = 1;
  uint32_t nl_cls = 0, nlf_cls = 0xFFFFFFFFu;   // see cf::DfaTables
  uint32_t start_state[4] = {0, 0, 0, 0};
  uint32_t start_adv[4] = {0, 0, 0, 0};   // ordered DFAs: start states that do not accept a zero-length match (re.sub's must_advance)
};

struct FilterOut {
  uint32_t E[256];           // P3<<20 | P2<<15 | P1<<10 | P0<<5 | N (five buckets per field; scan_core.h)
  std::vector<int> bucket_of_pattern;
  // pair prefilter (scan_core.h), built when the byte filter's expected candidate rate is too high
  bool use_pairs = false;
  std::vector<uint32_t> pairT;          // [cf::PF_SLOTS] when use_pairs
  double byte_cost = 0;                 // expected candidates per byte under the byte-frequency prior
};

// The Thompson NFA of one substitution rule, for the capture pass (cf::pike_captures in scan_core.h): three words per
// instruction {op | arg << 8, x, y}: N_CHAR (arg = set, x = next), N_SPLIT (x preferred, y), N_ASSERT (arg = kind, x),
// N_SAVE (arg = slot, x), N_MATCH.  setbits[set * wpc + (cls >> 5)] bit (cls & 31): the set contains code-point class cls.
struct NfaOut {
  std::vector<uint32_t> code;
  std::vector<uint32_t> setbits;
  uint32_t ninst = 0, start = 0, wpc = 1, ngroups = 0;
};

struct PatternInfo {
  bool nullable_always = false;   // matches the empty string with no assertion (or: substitution rule that can match "") -> bit set for every unit
  uint32_t min_len_chars = 0;     // minimum match length in code points
};

struct CompileOut {
  DfaOut search;                        // unordered union DFA over all patterns (existence bitmaps)
  FilterOut filter;                     // prefilter for `search`
  std::vector<DfaOut> ordered;          // one leftmost-first DFA per pattern flagged `want_ordered`
  std::vector<FilterOut> ordered_filter;// per-ordered-pattern prefilter (own bucket 0)
  std::vector<NfaOut> ordered_nfa;      // per-ordered-pattern NFA (group captures for replacement templates)
  std::vector<PatternInfo> info;
  std::vector<uint64_t> always_bits;    // W words: patterns that match every unit
};

// word_set: code points that are \w for the running interpreter (str patterns, UNICODE).
// want_ordered[i] != 0 -> also build a leftmost-first DFA for pattern i (regex_filter rules).
// Returns 0 or a negative error; `err` receives a message.  Error codes mirror include/cfgpu.h.
int compile(const std::vector<PatternIn>& pats, const std::vector<uint8_t>& want_ordered,
            const CharSet& word_set, CompileOut* out, std::string* err);

}  // namespace cfre
