// cf_host.cpp — host-only half of the C ABI: the program builder (include/cfgpu.h,
// "program construction").  No CUDA here, so it also links into the CPU-test helper library.
#include "cf_host.h"

#include <string.h>

extern "C" {

int cf_builder_new(cf_builder** out) {
  if (!out) return CF_E_BADARG;
  *out = new (std::nothrow) cf_builder();
  return *out ? CF_OK : CF_E_NOMEM;
}

void cf_builder_free(cf_builder* b) { delete b; }

const char* cf_builder_last_error(cf_builder* b) { return b ? b->err.c_str() : "null builder"; }

int cf_builder_set_word_set(cf_builder* b, const uint32_t* ranges, uint32_t nranges) {
  if (!b || (!ranges && nranges)) return CF_E_BADARG;
  b->word.clear();
  for (uint32_t i = 0; i < nranges; ++i) {
    cfre::Interval iv{ranges[2 * i], ranges[2 * i + 1]};
    if (iv.lo > iv.hi || iv.hi > 0x10FFFF || (!b->word.empty() && iv.lo <= b->word.back().hi)) {
      b->err = "word set must be sorted disjoint inclusive ranges";
      return CF_E_BADARG;
    }
    b->word.push_back(iv);
  }
  b->compiled = false;
  return CF_OK;
}

int cf_builder_add_pattern(cf_builder* b, const uint32_t* ast, uint32_t nwords, uint32_t flags,
                           uint32_t* out_index) {
  if (!b || !ast || !nwords) return CF_E_BADARG;
  cfre::PatternIn p;
  p.ast.assign(ast, ast + nwords);
  b->pats.push_back(std::move(p));
  b->ordered.push_back((flags & CF_PAT_ORDERED) ? 1 : 0);
  b->repl.emplace_back();
  b->has_repl.push_back(0);
  b->tmpl.emplace_back();
  if (out_index) *out_index = (uint32_t)b->pats.size() - 1;
  b->compiled = false;
  return CF_OK;
}

int cf_builder_set_replacement(cf_builder* b, uint32_t idx, const uint8_t* repl, uint32_t len) {
  if (!b || idx >= b->pats.size() || (!repl && len)) return CF_E_BADARG;
  if (!b->ordered[idx]) { b->err = "replacement on a non-ordered pattern"; return CF_E_BADARG; }
  b->repl[idx].assign(repl, repl + len);
  b->has_repl[idx] = 1;
  b->tmpl[idx].clear();
  return CF_OK;
}

int cf_builder_set_template(cf_builder* b, uint32_t idx, const uint8_t* literals, uint32_t literals_len, const uint32_t* parts, uint32_t n_parts) {
  if (!b || idx >= b->pats.size() || (!literals && literals_len) || (!parts && n_parts)) return CF_E_BADARG;
  if (!b->ordered[idx]) { b->err = "replacement on a non-ordered pattern"; return CF_E_BADARG; }
  bool refs = false;
  for (uint32_t k = 0; k < n_parts; ++k) {
    const uint32_t kind = parts[3 * k], a = parts[3 * k + 1], n = parts[3 * k + 2];
    if (kind == 0) { if ((uint64_t)a + n > literals_len) { b->err = "template literal out of range"; return CF_E_BADARG; } }
    else if (kind == 1) { if (a > 31) { b->err = "template references group " + std::to_string(a) + " (at most 31 groups)"; return CF_E_UNSUPPORTED; } refs = true; }
    else { b->err = "template part kind"; return CF_E_BADARG; }
  }
  b->repl[idx].assign(literals, literals + literals_len);
  b->has_repl[idx] = 1;
  if (refs) b->tmpl[idx].assign(parts, parts + 3ull * n_parts);
  else {                                   // no group reference: a literal replacement after all
    std::vector<uint8_t> flat;
    for (uint32_t k = 0; k < n_parts; ++k) flat.insert(flat.end(), literals + parts[3 * k + 1], literals + parts[3 * k + 1] + parts[3 * k + 2]);
    b->repl[idx] = flat;
    b->tmpl[idx].clear();
  }
  b->compiled = false;
  return CF_OK;
}

int cf_builder_compile_host(cf_builder* b, cf_compile_stats* out) {
  if (!b) return CF_E_BADARG;
  if (!b->compiled) {
    b->out = cfre::CompileOut();
    int rc = cfre::compile(b->pats, b->ordered, b->word, &b->out, &b->err);
    if (rc) return rc;
    // a template may only name groups the pattern has
    for (size_t i = 0, r = 0; i < b->pats.size(); ++i) {
      if (!b->ordered[i]) continue;
      const uint32_t ng = b->out.ordered_nfa[r++].ngroups;
      for (size_t k = 0; k + 2 < b->tmpl[i].size(); k += 3)
        if (b->tmpl[i][k] == 1 && b->tmpl[i][k + 1] > ng) { b->err = "pattern " + std::to_string(i) + ": template references group " + std::to_string(b->tmpl[i][k + 1]) + " but the pattern has " + std::to_string(ng); return CF_E_BADARG; }
      if (!b->tmpl[i].empty() && ng > 31) { b->err = "pattern " + std::to_string(i) + ": more than 31 capturing groups"; return CF_E_UNSUPPORTED; }
    }
    b->compiled = true;
  }
  if (out) {
    memset(out, 0, sizeof(*out));
    out->n_patterns = (uint32_t)b->pats.size();
    out->words_per_bitmap = b->out.search.W;
    out->n_classes = b->out.search.ncols - 1;
    out->n_states = b->out.search.nstates;
    out->n_accsets = (uint32_t)(b->out.search.accsets.size() / b->out.search.W);
    out->n_ordered = (uint32_t)b->out.ordered.size();
    out->trans_bytes = (uint32_t)(b->out.search.trans.size() * 4);
    out->prefilter = b->out.filter.use_pairs ? 1u : 0u;
  }
  return CF_OK;
}

}  // extern "C"
