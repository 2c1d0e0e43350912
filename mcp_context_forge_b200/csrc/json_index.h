// json_index.h — structural index of a JSON text (simdjson-style stage 1) and the token-driven DOM
// builder that consumes it.  SURVEY.md §8(f)-2: one reusable op feeding TOON, masking and string extraction.
//
// The index is computed 32 bytes at a time from per-chunk bit masks; on the GPU a warp produces the
// masks with one ballot each (lane i <-> byte i), on the host (tests/hostsim) a loop does.  Everything
// after the masks is the same code on both sides.
//
//   stage 1   index_chunk     quotes (escape parity), in-string mask (prefix XOR), structural characters,
//                             starts of scalars  ->  token positions, in text order
//   stage 1b  classify_token  per token, independently (lane-parallel on the GPU): strings are validated and
//                             their emit-time predicates + hash computed by cfj::parse_string — the SAME
//                             function the sequential parser uses; scalars by cfj::scan_number
//   stage 2   json_build      grammar + DOM over the tokens only (no byte scanning): same node layout,
//                             duplicate-key rule and limits as cfj::json_parse
//
// json_build(index(text)) == json_parse(text) node for node; tests/test_json_index_cpu.py holds that on the
// golden corpus and fuzz.
#pragma once
#include <stdint.h>

#include "json_toon.h"

namespace cfx {

struct Tok { uint32_t pos; uint32_t aux; };
static const uint32_t T_CLOSE = 0x80000000u;     // pos flag: this quote closes a string
static const uint32_t T_POS = 0x7FFFFFFFu;

// aux of an opening quote: string flags (ESC, Q, CTRLERR, KEYOK -> 4 bits) + 28 bits of the FNV-1a hash
CF_HD uint32_t pack_str(uint32_t fl, uint32_t h) {
  uint32_t f = ((fl & cfj::JF_ESC) ? 1u : 0u) | ((fl & cfj::JF_Q) ? 2u : 0u) | ((fl & cfj::JF_CTRLERR) ? 4u : 0u) | ((fl & cfj::JF_KEYOK) ? 8u : 0u);
  return (h << 4) | f;
}
CF_HD uint32_t str_flags(uint32_t aux) {
  return ((aux & 1u) ? cfj::JF_ESC : 0u) | ((aux & 2u) ? cfj::JF_Q : 0u) | ((aux & 4u) ? cfj::JF_CTRLERR : 0u) | ((aux & 8u) ? cfj::JF_KEYOK : 0u);
}
CF_HD uint32_t str_hash(uint32_t aux) { return aux >> 4; }
// aux of a scalar: kind in the low 4 bits (J_NULL/J_FALSE/J_TRUE/J_NUM), number flags in bits 4-6, length above
static const uint32_t SC_MAXLEN = 1u << 25;
CF_HD uint32_t pack_scalar(uint32_t kind, uint32_t fl, uint32_t len) {
  return kind | ((fl & cfj::JF_NEG) ? 16u : 0u) | ((fl & cfj::JF_FRAC) ? 32u : 0u) | ((fl & cfj::JF_EXP) ? 64u : 0u) | (len << 7);
}
CF_HD uint32_t scalar_type(uint32_t aux) {
  return (aux & 15u) | ((aux & 16u) ? cfj::JF_NEG : 0u) | ((aux & 32u) ? cfj::JF_FRAC : 0u) | ((aux & 64u) ? cfj::JF_EXP : 0u);
}
CF_HD uint32_t scalar_len(uint32_t aux) { return aux >> 7; }
static const uint32_t AUX_BAD = 0xFFFFFFFFu;    // classify_token: invalid string / scalar  (never a valid packing: kind 15)
static const uint32_t AUX_BIG = 0xFFFFFFFEu;    // scalar longer than SC_MAXLEN: beyond the device limits

CF_HD uint32_t clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__clz((int)v);
#else
  return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}

// carries between 32-byte chunks
struct IndexCarry {
  uint32_t in_string;    // ~0 when the chunk starts inside a string
  uint32_t bs_parity;    // parity of the backslash run that ends at the chunk boundary
  uint32_t prev_other;   // the last byte of the previous chunk was part of a scalar
  CF_HD void init() { in_string = 0; bs_parity = 0; prev_other = 0; }
};

// Is byte i of the chunk escaped, i.e. preceded by an odd-length run of backslashes?  `bs` = backslash mask.
CF_HD uint32_t escaped_bit(uint32_t bs, uint32_t i, uint32_t carry_parity) {
  uint32_t run = 0;
  if (i) run = clz32(~(bs << (32 - i)));       // consecutive backslashes immediately below bit i (<= i by construction)
  return (run + (run == i ? carry_parity : 0u)) & 1u;
}
CF_HD uint32_t next_bs_parity(uint32_t bs, uint32_t carry_parity) {
  const uint32_t run = clz32(~bs);             // backslashes ending the chunk
  return run == 32 ? carry_parity : (run & 1u);
}

// From the chunk's masks — real (unescaped) quotes, structural characters { } [ ] : , and JSON whitespace
// (bytes past the end of the text count as whitespace) — to its token mask.  *close = quotes that close a string.
CF_HD uint32_t index_chunk(uint32_t quotes, uint32_t structural, uint32_t ws, IndexCarry& cy, uint32_t* close) {
  uint32_t x = quotes;
  x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;       // prefix XOR: parity of quotes at or below each bit
  const uint32_t instr = x ^ cy.in_string;      // inside a string: opening quote included, closing quote excluded
  cy.in_string = (instr & 0x80000000u) ? ~0u : 0u;
  const uint32_t openq = quotes & instr, closeq = quotes & ~instr;
  const uint32_t outside = ~instr & ~quotes;
  const uint32_t other = outside & ~structural & ~ws;                      // bytes of scalars (or garbage; stage 1b decides)
  const uint32_t starts = other & ~((other << 1) | cy.prev_other);
  cy.prev_other = other >> 31;
  *close = closeq;
  return (structural & outside) | openq | closeq | starts;
}

CF_HD bool is_structural(uint32_t c) { return c == '{' || c == '}' || c == '[' || c == ']' || c == ':' || c == ','; }

// stage 1b: aux of one token (pos as stored by stage 1).  Structural characters and closing quotes carry nothing.
CF_HD uint32_t classify_token(const uint8_t* s, uint32_t n, uint32_t tpos) {
  if (tpos & T_CLOSE) return 0;
  const uint32_t c = s[tpos];
  if (is_structural(c)) return 0;
  if (c == '"') {
    uint32_t p = tpos, fl, h;
    if (!cfj::parse_string(s, n, &p, &fl, &h)) return AUX_BAD;
    return pack_str(fl, h);
  }
  uint32_t p = tpos, fl = 0, kind;
  if (c == '-' || (c >= '0' && c <= '9')) {
    if (!cfj::scan_number(s, n, &p, &fl)) return AUX_BAD;
    kind = cfj::J_NUM;
  } else if (c == 't' && tpos + 4 <= n && s[tpos + 1] == 'r' && s[tpos + 2] == 'u' && s[tpos + 3] == 'e') { kind = cfj::J_TRUE; p += 4; }
  else if (c == 'f' && tpos + 5 <= n && s[tpos + 1] == 'a' && s[tpos + 2] == 'l' && s[tpos + 3] == 's' && s[tpos + 4] == 'e') { kind = cfj::J_FALSE; p += 5; }
  else if (c == 'n' && tpos + 4 <= n && s[tpos + 1] == 'u' && s[tpos + 2] == 'l' && s[tpos + 3] == 'l') { kind = cfj::J_NULL; p += 4; }
  else return AUX_BAD;
  // the scalar must be the whole run of non-structural bytes ("123abc", "truex" are not JSON)
  if (p < n) { const uint32_t e = s[p]; if (!(cfj::j_ws(e) || is_structural(e) || e == '"')) return AUX_BAD; }
  if (p - tpos >= SC_MAXLEN) return AUX_BIG;
  return pack_scalar(kind, fl, p - tpos);
}

// stage 2: DOM from the classified tokens.  `unterminated` = stage 1 ended inside a string.
// Same node layout, limits and duplicate-key rule as cfj::json_parse.
CF_HD int json_build(const uint8_t* s, uint32_t n, const Tok* tok, uint32_t ntok, bool unterminated, cfj::JNode* nodes, uint32_t cap,
                     uint32_t* out_count) {
  using namespace cfj;
  (void)n;
  if (unterminated || ntok == 0) return PARSE_ERROR;
  uint32_t nn = 0, ti = 0;
  uint32_t st_node[MAXD], st_last[MAXD], st_len[MAXD], st_kh[MAXD];
  uint32_t kh[KH_CAP];
  uint32_t khn = 0;
  uint64_t khbad = 0, objbits = 0;
  int sp = 0;
  bool member = false;
  uint32_t member_hash = 0;
  enum { M_VALUE, M_KEY, M_AFTER } mode = M_VALUE;
  while (true) {
    if (mode == M_VALUE || mode == M_KEY) {
      if (ti >= ntok) return PARSE_ERROR;
      if (nn + 2 > cap) return PARSE_UNSUPPORTED;
      const Tok t = tok[ti];
      if (t.pos & T_CLOSE) return PARSE_ERROR;
      if (t.aux == AUX_BAD) return PARSE_ERROR;
      if (t.aux == AUX_BIG) return PARSE_UNSUPPORTED;
      const uint32_t c = s[t.pos];
      const uint32_t idx = nn;
      if (mode == M_KEY) {
        if (c != '"') return PARSE_ERROR;
        if (ti + 2 >= ntok) return PARSE_ERROR;                 // closing quote, ':' and a value must follow
        const uint32_t endq = tok[ti + 1].pos & T_POS;
        const uint32_t h = str_hash(t.aux);
        nodes[nn++] = JNode{J_KEY | str_flags(t.aux), t.pos + 1, endq - t.pos - 1, 0};
        if (st_last[sp - 1]) nodes[st_last[sp - 1]].next = idx; else nodes[st_node[sp - 1]].off = idx;
        st_last[sp - 1] = idx;
        st_len[sp - 1]++;
        if (khn < KH_CAP) kh[khn++] = h; else khbad |= 1ull << (sp - 1);
        const Tok colon = tok[ti + 2];
        if ((colon.pos & T_CLOSE) || s[colon.pos] != ':') return PARSE_ERROR;
        ti += 3;
        member = true;
        member_hash = h;
        mode = M_VALUE;
        continue;
      }
      const uint32_t keep_next = member ? member_hash : 0;
      if (!member && sp > 0) {
        if (st_last[sp - 1]) nodes[st_last[sp - 1]].next = idx; else nodes[st_node[sp - 1]].off = idx;
        st_last[sp - 1] = idx;
        st_len[sp - 1]++;
      }
      member = false;
      if (c == '{' || c == '[') {
        if (sp >= MAXD) return PARSE_UNSUPPORTED;
        nodes[nn++] = JNode{c == '{' ? (uint32_t)J_OBJ : (uint32_t)J_ARR, 0, 0, keep_next};
        st_node[sp] = idx; st_last[sp] = 0; st_len[sp] = 0; st_kh[sp] = khn;
        if (c == '{') objbits |= 1ull << sp; else objbits &= ~(1ull << sp);
        ++sp;
        ++ti;
        if (ti < ntok && !(tok[ti].pos & T_CLOSE) && s[tok[ti].pos] == (c == '{' ? '}' : ']')) { ++ti; --sp; mode = M_AFTER; }
        else mode = (c == '{') ? M_KEY : M_VALUE;
        continue;
      }
      if (c == '"') {
        if (ti + 1 >= ntok) return PARSE_ERROR;
        const uint32_t endq = tok[ti + 1].pos & T_POS;
        nodes[nn++] = JNode{J_STR | str_flags(t.aux), t.pos + 1, endq - t.pos - 1, keep_next};
        ti += 2;
      } else if (is_structural(c)) {
        return PARSE_ERROR;                                     // , : } ] where a value must start
      } else {
        nodes[nn++] = JNode{scalar_type(t.aux), t.pos, scalar_len(t.aux), keep_next};
        ++ti;
      }
      mode = M_AFTER;
      continue;
    }
    // M_AFTER
    if (sp == 0) {
      if (ti != ntok) return PARSE_ERROR;
      *out_count = nn;
      return PARSE_OK;
    }
    if (ti >= ntok) return PARSE_ERROR;
    const Tok t = tok[ti];
    if (t.pos & T_CLOSE) return PARSE_ERROR;
    const bool is_obj = (objbits >> (sp - 1)) & 1;
    const uint32_t c = s[t.pos];
    if (c == ',') { ++ti; mode = is_obj ? M_KEY : M_VALUE; continue; }
    if (c != (uint32_t)(is_obj ? '}' : ']')) return PARSE_ERROR;
    ++ti;
    --sp;
    const uint32_t cidx = st_node[sp];
    uint32_t clen = st_len[sp];
    if (is_obj) {
      const uint32_t kb = st_kh[sp];
      bool maybe_dup = (khbad >> sp) & 1;
      khbad &= ~(1ull << sp);
      if (!maybe_dup)
        for (uint32_t i = kb + 1; i < khn && !maybe_dup; ++i)
          for (uint32_t j = kb; j < i; ++j) if (kh[j] == kh[i]) { maybe_dup = true; break; }
      khn = kb;
      if (maybe_dup && clen > 1) {
        const uint32_t first = nodes[cidx].off;
        uint32_t prev = first;
        for (uint32_t k = nodes[prev].next; k;) {
          const uint32_t nxt = nodes[k].next;
          bool dup = false;
          for (uint32_t i = first; i != k; i = nodes[i].next)
            if (nodes[i + 1].next == nodes[k + 1].next && keys_equal(s, nodes[i], nodes[k])) {
              const uint32_t hsh = nodes[i + 1].next;
              nodes[i + 1] = nodes[k + 1];
              nodes[i + 1].next = hsh;
              dup = true;
              break;
            }
          if (dup) { nodes[prev].next = nxt; clen--; }
          else prev = k;
          k = nxt;
        }
      }
    }
    nodes[cidx].len = clen;
    mode = M_AFTER;
  }
}

// Host-side stage 1 (tests): the masks a warp gets from ballots, built with a loop.  Returns the token count;
// *unterminated as for json_build.
inline uint32_t index_host(const uint8_t* s, uint32_t n, Tok* tok, bool* unterminated) {
  IndexCarry cy;
  cy.init();
  uint32_t nt = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    uint32_t bs = 0, qm = 0, st = 0, ws = 0;
    for (uint32_t i = 0; i < 32; ++i) {
      const uint32_t c = base + i < n ? s[base + i] : (uint32_t)' ';
      if (c == '\\') bs |= 1u << i;
      if (c == '"') qm |= 1u << i;
      if (is_structural(c)) st |= 1u << i;
      if (cfj::j_ws(c)) ws |= 1u << i;
    }
    uint32_t esc = 0;
    for (uint32_t i = 0; i < 32; ++i) esc |= escaped_bit(bs, i, cy.bs_parity) << i;
    cy.bs_parity = next_bs_parity(bs, cy.bs_parity);
    uint32_t close = 0;
    uint32_t tm = index_chunk(qm & ~esc, st, ws, cy, &close);
    while (tm) {
      const uint32_t i = (uint32_t)__builtin_ctz(tm);
      tm &= tm - 1;
      tok[nt].pos = (base + i) | (((close >> i) & 1u) ? T_CLOSE : 0u);
      tok[nt].aux = 0;
      ++nt;
    }
  }
  *unterminated = cy.in_string != 0;
  return nt;
}

}  // namespace cfx
