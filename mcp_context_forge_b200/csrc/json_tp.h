// json_tp.h — token-parallel JSON -> TOON, one WARP per payload, no DOM.
//
// The reference work this replaces (paths relative to /root/reference):
//   orjson.loads + toon.encode + "keep only if strictly smaller"   plugins/toon_encoder/toon_encoder.py:277-303
//   encoder rules                                                    plugins/toon_encoder/toon.py:82-565
//
// Shape of the computation:
//   pass 1  front end   32 bytes per step, lane i <-> byte i: byte classes by ballot, escape parity, in-string mask by
//                       prefix XOR -> SIGNIFICANT tokens (brackets, strings at their closing quote, scalar starts); commas
//                       and colons are not tokens, each token carries how many of them precede it.  Tokens go through a
//                       ring in shared memory.
//           batch       32 tokens at a time, lane i <-> token i: scalars / strings validated and classified in parallel
//                       (orjson's accept set), then ONE warp-uniform walk over the batch's brackets maintains the container
//                       stack (grammar, child counts, duplicate-key screen, columnar-table detection against the first row)
//                       and patches every opener's token with {child count, layout mode} when its closer arrives.
//                       Tokens are stored to HBM scratch (8 bytes each).
//   pass 2  emit        32 tokens at a time: the same bracket walk now carries the TOON frames (layout mode, prefix width,
//                       indent level); every token's output piece length is known locally, a warp prefix sum gives its
//                       offset, pieces are written in parallel (long spans by the whole warp).
// Whatever the fast path does not cover (duplicate keys, escaped keys, permuted table rows, numbers that need the exact
// big-integer formatter, ...) is reported as TS_FALLBACK and re-done by the sequential per-thread encoder (json_toon.h) —
// still on the GPU, never on the CPU.
#pragma once
#include <stdint.h>

#include "json_index.h"
#include "json_toon.h"
#include "warp_prims.h"

namespace cftp {

using cfj::TS_ATTR_ERROR;
using cfj::TS_CONVERTED;
using cfj::TS_NOT_JSON;
using cfj::TS_NOT_SMALLER;
using cfj::TS_UNSUPPORTED;
using cfj::TS_VALUE_ERROR;
enum : int { TS_FALLBACK = 7 };      // internal: bits 8.. carry the reason (FB_*), stripped at the ABI
enum : uint32_t { FB_NUM_EXACT = 1, FB_KEY_ESCAPE = 2, FB_TOK_CAP = 3, FB_KH_CAP = 4, FB_DUP_HASH = 5, FB_ROW_ORDER = 6, FB_MIXED_ITEM = 7 };

// ---- tokens --------------------------------------------------------------------------------------------------------
enum : uint32_t { K_OPEN_OBJ = 0, K_OPEN_ARR = 1, K_CLOSE_OBJ = 2, K_CLOSE_ARR = 3, K_STR = 4, K_KEY = 5, K_NUM = 6, K_LIT = 7 };
struct GTok { uint32_t pos, w; };                 // w = kind | flags << 3 | len << 8
static const uint32_t GT_MAXLEN = (1u << 24) - 1;
// flags (5 bits)
enum : uint32_t { SF_Q = 1, SF_ESCX = 2, SF_CTRLERR = 4 };        // K_STR: needs quotes / needs transcoding / holds a char TOON cannot quote
enum : uint32_t { KF_KEYOK = 1 };                                 // K_KEY: valid unquoted key
enum : uint32_t { AM_EMPTY = 0, AM_COLUMNAR = 1, AM_INLINE = 2, AM_ITEMS = 3, AF_MIXED = 4 };   // K_OPEN_ARR (patched at its closer)
TP_FN uint32_t gt_kind(uint32_t w) { return w & 7u; }
TP_FN uint32_t gt_flags(uint32_t w) { return (w >> 3) & 31u; }
TP_FN uint32_t gt_len(uint32_t w) { return w >> 8; }
TP_FN uint32_t gt_make(uint32_t kind, uint32_t fl, uint32_t len) { return kind | (fl << 3) | (len << 8); }

// ring record meta (front end -> batch)
enum : uint32_t { RM_KIND = 7u, RM_NCOMMA_SH = 3, RM_NCOLON_SH = 5, RM_SPECIAL = 1u << 7, RM_NONKEY = 1u << 8, RM_HI = 1u << 9, RM_BS = 1u << 10,
                  RM_OPENEND = 1u << 11 /* scalar run reaches the end of its 32-byte chunk: length still unknown */ };

// ---- per-warp shared memory ----------------------------------------------------------------------------------------
static const uint32_t RING = 128, MAXD = 64, KH_CAP = 256;
static const uint32_t UNSET = 0xFFFFFFFFu;
struct Shared {
  uint32_t ring_pos[RING], ring_len[RING], ring_meta[RING];
  // pass 1: container stack            | pass 2: frame stack (same storage)
  uint32_t open_idx[MAXD];             // token index of the opener      | frame mode
  uint32_t cnt[MAXD];                  // children (values) so far       | children so far
  uint32_t cfl[MAXD];                  // C_* flags                      | prefix width
  uint32_t khbase[MAXD];               // base of the key hashes         | indent level
  uint32_t row0_idx[MAXD];             // arrays: 1 + token index of the first element when it is an object
  uint32_t row0_n[MAXD];               // arrays: member count of that first object once it closed
  uint32_t kh[KH_CAP];                 // key hashes of the open objects (stack)
};
enum : uint32_t { C_OBJ = 1, C_ALL_SIMPLE = 2, C_ALL_OBJ = 4, C_COL_OK = 8, C_ALIGNED = 16, C_VALS_SIMPLE = 32 };

// byte classes of the front end
enum : uint32_t { BC_BS = 1, BC_QUOTE = 2, BC_STRUCT = 4, BC_COMMA = 8, BC_COLON = 16, BC_WS = 32, BC_CTRL = 64, BC_SPECIAL = 128, BC_NONKEY = 256, BC_HI = 512 };
TP_FN uint32_t byte_class(uint32_t b) {
  uint32_t k = 0;
  if (b == '\\') k |= BC_BS;
  if (b == '"') k |= BC_QUOTE;
  if (b == '{' || b == '}' || b == '[' || b == ']' || b == ':' || b == ',') k |= BC_STRUCT;
  if (b == ',') k |= BC_COMMA;
  if (b == ':') k |= BC_COLON;
  if (b == ' ' || b == '\t' || b == '\n' || b == '\r') k |= BC_WS;
  if (b < 0x20) k |= BC_CTRL;
  if (b == ',' || b == ':' || b == '[' || b == ']' || b == '{' || b == '}' || b == '-') k |= BC_SPECIAL;
  if (!((b >= 'A' && b <= 'Z') || (b >= 'a' && b <= 'z') || (b >= '0' && b <= '9') || b == '_' || b == '.')) k |= BC_NONKEY;
  if (b >= 0x80) k |= BC_HI;
  return k;
}

TP_FN uint32_t bits_below(uint32_t i) { return i >= 32 ? 0xFFFFFFFFu : ((1u << i) - 1u); }   // bits 0..i-1
TP_FN uint32_t range_mask(uint32_t lo, uint32_t hi) { return bits_below(hi) & ~bits_below(lo); }   // bits lo..hi-1
TP_FN uint32_t sat3(uint32_t v) { return v > 3u ? 3u : v; }

// ---- scalars ---------------------------------------------------------------------------------------------------------
// Number literal t[0..len) (already grammar-checked, flags from cfj::scan_number): the text toon._encode_float / str(int)
// produce is, on the cheap path, a PREFIX of the literal (possibly without its '-').  Returns false when the exact
// big-integer formatter is needed (json_toon.h emit_number's exact path).
TP_FN bool num_canon(const uint8_t* t, uint32_t len, uint32_t fl, uint32_t* eoff, uint32_t* elen) {
  const bool neg = (fl & cfj::JF_NEG) != 0;
  const uint8_t* dg = t + (neg ? 1 : 0);
  const uint32_t dl = len - (neg ? 1 : 0);
  if (!(fl & (cfj::JF_FRAC | cfj::JF_EXP))) {
    if (dl >= 19) return false;                                  // may leave the i64/u64 range: exact path decides
    if (dl == 1 && dg[0] == '0') { *eoff = neg ? 1 : 0; *elen = 1; return true; }   // "-0" -> int 0
    *eoff = 0; *elen = len;
    return true;
  }
  if (fl & cfj::JF_EXP) return false;
  uint32_t dot = 0;
  while (dg[dot] != '.') ++dot;
  uint32_t fe = dl;
  while (fe > dot + 1 && dg[fe - 1] == '0') --fe;
  const uint32_t nfrac = fe - dot - 1;
  const bool int_zero = (dot == 1 && dg[0] == '0');
  uint32_t lead_fz = 0;
  if (int_zero) while (lead_fz < nfrac && dg[dot + 1 + lead_fz] == '0') ++lead_fz;
  const uint32_t sigd = int_zero ? nfrac - lead_fz : dot + nfrac;
  const bool tiny_long = int_zero && lead_fz >= 4 && nfrac > 15;
  if (sigd > 15 || tiny_long) return false;
  if (nfrac == 0) {
    if (int_zero) { *eoff = neg ? 1 : 0; *elen = 1; return true; }     // +-0.0 -> "0"
    *eoff = 0; *elen = (neg ? 1 : 0) + dot;
    return true;
  }
  *eoff = 0; *elen = (neg ? 1 : 0) + dot + 1 + nfrac;
  return true;
}

// ---- strings ---------------------------------------------------------------------------------------------------------
TP_FN bool is_reserved(const uint8_t* b, uint32_t len) {
  if (len == 4) return (b[0] == 'n' && b[1] == 'u' && b[2] == 'l' && b[3] == 'l') || (b[0] == 't' && b[1] == 'r' && b[2] == 'u' && b[3] == 'e');
  if (len == 5) return b[0] == 'f' && b[1] == 'a' && b[2] == 'l' && b[3] == 's' && b[4] == 'e';
  return false;
}
// escapes other than \" \\ \n \r \t need transcoding (the TOON text differs from the JSON text)
TP_FN bool has_complex_escape(const uint8_t* b, uint32_t len) {
  for (uint32_t i = 0; i + 1 < len; ++i)
    if (b[i] == '\\') {
      const uint32_t e = b[i + 1];
      if (!(e == '"' || e == '\\' || e == 'n' || e == 'r' || e == 't')) return true;
      ++i;
    }
  return false;
}
TP_FN uint32_t fnv1a(const uint8_t* b, uint32_t len) {
  uint32_t h = 2166136261u;
  for (uint32_t i = 0; i < len; ++i) h = (h ^ b[i]) * 16777619u;
  return h;
}

// decoded length / bytes of a string that needs transcoding (JSON escapes -> TOON text), quoted or not
TP_FN uint32_t escx_len(const uint8_t* b, uint32_t len, bool quoted) {
  cfj::StrIter it{b, b + len};
  uint32_t n = quoted ? 2u : 0u;
  while (!it.done()) {
    const uint32_t cp = it.next();
    if (quoted && (cp == '\\' || cp == '"' || cp == '\n' || cp == '\r' || cp == '\t')) n += 2;
    else n += cp < 0x80 ? 1u : cp < 0x800 ? 2u : cp < 0x10000 ? 3u : 4u;
  }
  return n;
}

// ---- output ----------------------------------------------------------------------------------------------------------
struct Emit {
  uint8_t* out;
  uint32_t cap, o;
  TP_FN void put(uint32_t c) { if (o < cap) out[o] = (uint8_t)c; ++o; }
  TP_FN void span(const uint8_t* b, uint32_t len) { for (uint32_t i = 0; i < len; ++i) put(b[i]); }
  TP_FN void spaces(uint32_t k) { for (uint32_t i = 0; i < k; ++i) put(' '); }
  TP_FN void uint_dec(uint32_t v) {
    uint8_t b[10]; int k = 0;
    do { b[k++] = (uint8_t)('0' + v % 10); v /= 10; } while (v);
    while (k) put(b[--k]);
  }
  TP_FN void cp(uint32_t c) {
    if (c < 0x80) put(c);
    else if (c < 0x800) { put(0xC0 | (c >> 6)); put(0x80 | (c & 63)); }
    else if (c < 0x10000) { put(0xE0 | (c >> 12)); put(0x80 | ((c >> 6) & 63)); put(0x80 | (c & 63)); }
    else { put(0xF0 | (c >> 18)); put(0x80 | ((c >> 12) & 63)); put(0x80 | ((c >> 6) & 63)); put(0x80 | (c & 63)); }
  }
  TP_FN void escx(const uint8_t* b, uint32_t len, bool quoted) {
    cfj::StrIter it{b, b + len};
    if (quoted) put('"');
    while (!it.done()) {
      const uint32_t c = it.next();
      if (quoted && c == '\\') { put('\\'); put('\\'); }
      else if (quoted && c == '"') { put('\\'); put('"'); }
      else if (quoted && c == '\n') { put('\\'); put('n'); }
      else if (quoted && c == '\r') { put('\\'); put('r'); }
      else if (quoted && c == '\t') { put('\\'); put('t'); }
      else cp(c);
    }
    if (quoted) put('"');
  }
};
TP_FN uint32_t dec_digits(uint32_t v) { uint32_t d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

// frame modes of pass 2
enum : uint32_t { M_ROOT = 0, M_OBJ = 1, M_LIST_ITEM = 2, M_ROW = 3, M_ARR_ITEMS = 4, M_ARR_INLINE = 5, M_ARR_COL = 6, M_DEAD = 7 /* empty container */ };

struct P1State {
  uint32_t sp, root_cnt, ntok, last_was_key;
  int status;          // 0 while everything is fine
};

// ----------------------------------------------------------------------------------------------------------------------
// pass 1, one batch: tokens ring[head .. head+m), la_ncolon = colons in front of the token that follows the batch
// ----------------------------------------------------------------------------------------------------------------------
TP_FN void p1_batch(const uint8_t* s, uint32_t n, GTok* toks, uint32_t tok_cap, Shared& sh, P1State& st, uint32_t head, uint32_t m, uint32_t la_ncolon) {
  const uint32_t l = tpw::lane();
  const uint32_t ltm = tpw::lt_mask();
  const bool act = l < m;
  uint32_t pos = 0, len = 0, meta = 0;
  if (act) { const uint32_t r = (head + l) & (RING - 1); pos = sh.ring_pos[r]; len = sh.ring_len[r]; meta = sh.ring_meta[r]; }
  uint32_t kind = meta & RM_KIND;
  const uint32_t ncomma = (meta >> RM_NCOMMA_SH) & 3u, ncolon = (meta >> RM_NCOLON_SH) & 3u;
  uint32_t nxt_ncolon = tpw::shfl_down(ncolon, 1);
  if (l + 1 >= m) nxt_ncolon = la_ncolon;
  bool bad = false, unsup = false;
  uint32_t fb = 0;                                   // fallback reason (FB_*), 0 = none
  uint32_t fl = 0, hash = 0;
  const bool isK = act && kind == K_STR && ncolon == 0 && nxt_ncolon >= 1;

  // ---- phase 0: every token on its own lane
  if (act && kind == K_NUM) {                       // scalar run starting at pos
    if (meta & RM_OPENEND) {                        // the run left its chunk: find its end
      uint32_t e = pos + len;
      while (e < n) { const uint32_t k = byte_class(s[e]); if (k & (BC_STRUCT | BC_WS | BC_QUOTE)) break; ++e; }
      len = e - pos;
    }
    const uint32_t c0 = s[pos];
    if (c0 == 't') { if (len == 4 && s[pos + 1] == 'r' && s[pos + 2] == 'u' && s[pos + 3] == 'e') kind = K_LIT; else bad = true; }
    else if (c0 == 'f') { if (len == 5 && s[pos + 1] == 'a' && s[pos + 2] == 'l' && s[pos + 3] == 's' && s[pos + 4] == 'e') kind = K_LIT; else bad = true; }
    else if (c0 == 'n') { if (len == 4 && s[pos + 1] == 'u' && s[pos + 2] == 'l' && s[pos + 3] == 'l') kind = K_LIT; else bad = true; }
    else if (c0 == '-' || (c0 >= '0' && c0 <= '9')) {
      uint32_t p = pos, nf = 0;
      if (!cfj::scan_number(s, pos + len, &p, &nf) || p != pos + len) bad = true;
      else {
        uint32_t eoff, elen;
        if (num_canon(s + pos, len, nf, &eoff, &elen)) { pos += eoff; len = elen; }
        else fb = FB_NUM_EXACT;                     // exact formatter: sequential encoder
      }
    } else bad = true;
  } else if (act && kind == K_STR) {
    const uint8_t* b = s + pos;
    if (len > GT_MAXLEN) unsup = true;
    else if ((meta & (RM_BS | RM_HI)) || (len && b[0] >= '0' && b[0] <= '9')) {
      // escapes, non-ASCII or number-like candidates: the sequential validator decides (same function as json_toon.h)
      uint32_t p = pos - 1, sf = 0, h = 0;
      if (!cfj::parse_string(s, n, &p, &sf, &h) || p != pos + len + 1) bad = true;
      else {
        hash = h;
        if (isK) { if (sf & cfj::JF_ESC) fb = FB_KEY_ESCAPE; if (sf & cfj::JF_KEYOK) fl |= KF_KEYOK; }
        else {
          if (sf & cfj::JF_Q) fl |= SF_Q;
          if (sf & cfj::JF_CTRLERR) fl |= SF_CTRLERR;
          if ((sf & cfj::JF_ESC) && has_complex_escape(b, len)) fl |= SF_ESCX;
        }
      }
    } else {
      const bool res = is_reserved(b, len);
      if (isK) {
        const uint32_t f0 = len ? b[0] : 0u;
        const bool al = (f0 >= 'A' && f0 <= 'Z') || (f0 >= 'a' && f0 <= 'z') || f0 == '_';
        if (len && al && !(meta & RM_NONKEY) && !res) fl |= KF_KEYOK;
        hash = fnv1a(b, len);
      } else if (len == 0 || res || (meta & RM_SPECIAL) || b[0] == ' ' || b[len - 1] == ' ') fl |= SF_Q;
    }
    if (isK) kind = K_KEY;
  }
  if (len > GT_MAXLEN) unsup = true;
  // store the tokens (openers are patched when their closer arrives)
  const uint32_t idx = st.ntok + l;
  if (st.ntok + m > tok_cap) fb = FB_TOK_CAP;
  else if (act) { GTok t; t.pos = pos; t.w = gt_make(kind, fl, kind <= K_CLOSE_ARR ? 0u : len); toks[idx] = t; }
  tpw::sync();

  // ---- phase 1: one walk over the brackets of the batch
  const bool isV = act && kind >= K_STR && kind != K_KEY;           // value that is not a container
  uint32_t evm = tpw::ballot(act && kind <= K_CLOSE_ARR);
  const uint32_t Km = tpw::ballot(act && kind == K_KEY), Vm = tpw::ballot(isV);
  uint32_t prevK = tpw::shfl_up(kind == K_KEY ? 1u : 0u, 1);
  if (l == 0) prevK = st.last_was_key;
  uint32_t sp = st.sp, root_cnt = st.root_cnt;
  bool ubad = false, uunsup = false;
  uint32_t ufb = 0;                  // warp-uniform verdicts of the walk
  uint32_t cur = 0;
  while (true) {
    const uint32_t e = evm ? tpw::ffs(evm) - 1 : m;
    const uint32_t run = range_mask(cur, e);
    if (run) {
      const bool inrun = (run >> l) & 1u;
      if (sp == 0) {                       // a scalar document: exactly one token, no separators
        if (inrun && (kind == K_KEY || ncomma || ncolon || root_cnt + tpw::popc(run & ltm) != 0)) bad = true;
        root_cnt += tpw::popc(run);
      } else {
        const uint32_t top = sp - 1, tfl = sh.cfl[top], c0 = sh.cnt[top];
        const bool isobj = (tfl & C_OBJ) != 0;
        const uint32_t Vr = Vm & run, Kr = Km & run;
        const uint32_t ord = c0 + tpw::popc(Vr & ltm);
        if (inrun) {
          if (isobj) {
            if (kind == K_KEY) { if (ncolon != 0 || ncomma != (ord > 0 ? 1u : 0u)) bad = true; }
            else if (ncolon != 1 || ncomma != 0 || !prevK) bad = true;
          } else if (kind == K_KEY || ncolon != 0 || ncomma != (ord > 0 ? 1u : 0u)) bad = true;
        }
        tpw::sync();
        if (l == 0) {
          sh.cnt[top] = c0 + tpw::popc(Vr);
          if (!isobj && Vr) sh.cfl[top] = tfl & ~C_ALL_OBJ;
        }
        if (isobj && Kr) {
          // duplicate-key screen on the hashes (a repeated hash, real duplicate or not, goes to the sequential encoder)
          const uint32_t kb = sh.khbase[top];
          const bool mine = inrun && kind == K_KEY;
          if (mine) { if (kb + ord >= KH_CAP) fb = FB_KH_CAP; else sh.kh[kb + ord] = hash; }
          tpw::sync();
          if (mine && kb + ord < KH_CAP) for (uint32_t j = kb; j < kb + ord; ++j) if (sh.kh[j] == hash) { fb = FB_DUP_HASH; break; }
          // table detection: the keys of every later row against the first row's, position by position
          if (sp >= 2) {
            const uint32_t par = sp - 2, pfl = sh.cfl[par];
            if (!(pfl & C_OBJ) && (pfl & C_COL_OK) && sh.row0_n[par] != UNSET && sh.open_idx[top] + 1 != sh.row0_idx[par]) {
              const uint32_t r0 = sh.row0_idx[par] - 1, rn = sh.row0_n[par];
              bool mism = false;
              if (mine) {
                if (ord >= rn) mism = true;
                else {
                  const GTok r = toks[r0 + 1 + 2 * ord];
                  if (gt_kind(r.w) != K_KEY || gt_len(r.w) != len) mism = true;
                  else { const uint8_t* a = s + r.pos; const uint8_t* b = s + pos; for (uint32_t i = 0; i < len; ++i) if (a[i] != b[i]) { mism = true; break; } }
                }
              }
              if (tpw::any(mism) && l == 0) sh.cfl[top] = sh.cfl[top] & ~C_ALIGNED;
            }
          }
        }
        tpw::sync();
      }
    }
    if (e >= m) break;
    // ---- the bracket at lane e
    const uint32_t ek = tpw::shfl(kind, e), ecomma = tpw::shfl(ncomma, e), ecolon = tpw::shfl(ncolon, e), eprevK = tpw::shfl(prevK, e);
    const uint32_t eidx = st.ntok + e;
    if (ek <= K_OPEN_ARR) {
      uint32_t newkb = 0;
      if (sp == 0) { if (ecomma || ecolon || root_cnt) ubad = true; ++root_cnt; }
      else {
        const uint32_t top = sp - 1;
        uint32_t tfl = sh.cfl[top];
        const uint32_t ord = sh.cnt[top];
        const bool isobj = (tfl & C_OBJ) != 0;
        if (isobj) { if (ecolon != 1 || ecomma != 0 || !eprevK) ubad = true; }
        else if (ecolon != 0 || ecomma != (ord > 0 ? 1u : 0u)) ubad = true;
        tfl &= ~C_ALL_SIMPLE;
        if (ek == K_OPEN_ARR) tfl &= ~C_ALL_OBJ;
        if (isobj) tfl &= ~C_VALS_SIMPLE;
        newkb = sh.khbase[top] + (isobj ? ord + 1 : 0u);
        tpw::sync();
        if (l == 0) {
          sh.cnt[top] = ord + 1;
          sh.cfl[top] = tfl;
          if (!isobj && ord == 0 && ek == K_OPEN_OBJ) sh.row0_idx[top] = eidx + 1;
        }
      }
      if (sp >= MAXD) { uunsup = true; break; }
      if (l == 0) {
        sh.open_idx[sp] = eidx; sh.cnt[sp] = 0; sh.khbase[sp] = newkb < KH_CAP ? newkb : KH_CAP;
        sh.cfl[sp] = ek == K_OPEN_OBJ ? (C_OBJ | C_ALIGNED | C_VALS_SIMPLE) : (C_ALL_SIMPLE | C_ALL_OBJ | C_COL_OK);
        sh.row0_idx[sp] = 0; sh.row0_n[sp] = UNSET;
      }
      ++sp;
      tpw::sync();
    } else {
      if (sp == 0) { ubad = true; break; }
      const uint32_t top = sp - 1, tfl = sh.cfl[top], nn = sh.cnt[top], oi = sh.open_idx[top];
      const bool isobj = (tfl & C_OBJ) != 0;
      if ((ek == K_CLOSE_OBJ) != isobj || ecomma || ecolon) ubad = true;
      if (nn > GT_MAXLEN) uunsup = true;
      uint32_t w;
      if (isobj) w = gt_make(K_OPEN_OBJ, 0, nn & GT_MAXLEN);
      else {
        const uint32_t mode = nn == 0 ? AM_EMPTY : ((tfl & C_ALL_OBJ) && (tfl & C_COL_OK)) ? AM_COLUMNAR : (tfl & C_ALL_SIMPLE) ? AM_INLINE : AM_ITEMS;
        const uint32_t mixed = (sh.row0_idx[top] && !(tfl & C_ALL_OBJ)) ? AF_MIXED : 0u;
        w = gt_make(K_OPEN_ARR, mode | mixed, nn & GT_MAXLEN);
      }
      if (sp >= 2 && isobj) {
        const uint32_t par = sp - 2;
        uint32_t pfl = sh.cfl[par];
        if (!(pfl & C_OBJ)) {
          if (oi + 1 == sh.row0_idx[par]) {
            if (nn == 0 || !(tfl & C_VALS_SIMPLE)) pfl &= ~C_COL_OK;
            tpw::sync();
            if (l == 0) { sh.row0_n[par] = nn; sh.cfl[par] = pfl; }
          } else if (pfl & C_COL_OK) {
            const uint32_t rn = sh.row0_n[par];
            if (rn == UNSET || nn != rn || !(tfl & C_VALS_SIMPLE)) pfl &= ~C_COL_OK;
            else if (!(tfl & C_ALIGNED)) ufb = FB_ROW_ORDER;      // same size, other key order or other keys: the sequential encoder sorts it out
            tpw::sync();
            if (l == 0) sh.cfl[par] = pfl;
          }
        }
      }
      if (l == 0 && oi < tok_cap) toks[oi].w = w;
      --sp;
      tpw::sync();
    }
    cur = e + 1;
    evm &= evm - 1;
  }
  st.sp = sp; st.root_cnt = root_cnt;
  st.last_was_key = tpw::shfl(kind == K_KEY ? 1u : 0u, m - 1);
  st.ntok += m;
  const bool any_bad = tpw::any(bad) || ubad, any_unsup = tpw::any(unsup) || uunsup;
  const uint32_t fbm = tpw::ballot(fb != 0);
  if (any_bad) st.status = TS_NOT_JSON;
  else if (any_unsup) st.status = TS_UNSUPPORTED;
  else if (fbm) st.status = TS_FALLBACK | (int)(tpw::shfl(fb, tpw::ffs(fbm) - 1) << 8);
  else if (ufb) st.status = TS_FALLBACK | (int)(ufb << 8);
}

// ----------------------------------------------------------------------------------------------------------------------
// pass 1: front end + batches.  Returns 0 or a TS_* status (warp-uniform); *ntok_out = tokens stored.
// ----------------------------------------------------------------------------------------------------------------------
TP_FN int pass1(const uint8_t* s, uint32_t n, GTok* toks, uint32_t tok_cap, Shared& sh, const uint16_t* ctab, uint32_t* ntok_out) {
  const uint32_t l = tpw::lane();
  const uint32_t ltm = tpw::lt_mask();
  P1State st;
  st.sp = 0; st.root_cnt = 0; st.ntok = 0; st.last_was_key = 0; st.status = 0;
  uint32_t in_string = 0, bs_parity = 0, prev_other = 0, c_ncomma = 0, c_ncolon = 0, open_pos = 0;
  uint32_t c_special = 0, c_nonkey = 0, c_hi = 0, c_bs = 0;     // classes seen so far in the string that is open across chunks
  uint32_t head = 0, rcount = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t p = base + l;
    const uint32_t c = p < n ? (uint32_t)s[p] : (uint32_t)' ';
    const uint32_t kc = ctab[c];                                // byte_class(c) from a 256-entry table (shared memory on the GPU)
    const uint32_t bs = tpw::ballot(kc & BC_BS), qm = tpw::ballot(kc & BC_QUOTE), stc = tpw::ballot(kc & BC_STRUCT);
    const uint32_t cm = tpw::ballot(kc & BC_COMMA), co = tpw::ballot(kc & BC_COLON), ws = tpw::ballot(kc & BC_WS), ctl = tpw::ballot(kc & BC_CTRL);
    uint32_t esc = 0;
    if (bs | bs_parity) {
      esc = tpw::ballot(cfx::escaped_bit(bs, l, bs_parity) != 0);
      bs_parity = cfx::next_bs_parity(bs, bs_parity);
    }
    const uint32_t quotes = qm & ~esc;
    uint32_t x = quotes;
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
    const uint32_t instr = x ^ in_string;                       // opening quote included, closing quote excluded
    in_string = (instr & 0x80000000u) ? ~0u : 0u;
    const uint32_t openq = quotes & instr, closeq = quotes & ~instr;
    const uint32_t content = instr & ~openq;
    if (ctl & content) { st.status = TS_NOT_JSON; break; }      // raw control character inside a string
    const uint32_t outside = ~instr & ~closeq;
    const uint32_t st_out = stc & outside, comma_out = cm & outside, colon_out = co & outside;
    const uint32_t other = outside & ~stc & ~ws;
    const uint32_t starts = other & ~((other << 1) | prev_other);
    prev_other = other >> 31;
    const uint32_t brackets = st_out & ~comma_out & ~colon_out;
    const uint32_t T = brackets | closeq | starts;
    // string classes (only bytes inside strings matter)
    uint32_t m_special = 0, m_nonkey = 0, m_hi = 0;
    if (content) {
      m_special = tpw::ballot(kc & BC_SPECIAL) & content;
      m_nonkey = tpw::ballot(kc & BC_NONKEY) & content;
      m_hi = tpw::ballot(kc & BC_HI) & content;
    }
    const uint32_t m_bs = bs & content;
    if ((T >> l) & 1u) {
      const uint32_t prevT = T & ltm;
      uint32_t between, nc, nk;
      if (prevT) { between = ltm & ~bits_below(32 - tpw::clz(prevT)); nc = 0; nk = 0; }
      else { between = ltm; nc = c_ncomma; nk = c_ncolon; }
      nc = sat3(nc + tpw::popc(comma_out & between));
      nk = sat3(nk + tpw::popc(colon_out & between));
      uint32_t meta = (nc << RM_NCOMMA_SH) | (nk << RM_NCOLON_SH), tpos = p, tlen = 0;
      if ((brackets >> l) & 1u) meta |= c == '{' ? K_OPEN_OBJ : c == '[' ? K_OPEN_ARR : c == '}' ? K_CLOSE_OBJ : K_CLOSE_ARR;
      else if ((closeq >> l) & 1u) {
        const uint32_t qb = quotes & ltm;
        uint32_t span, sp_ = 0, nk_ = 0, hi_ = 0, b_ = 0, op;
        if (qb) { const uint32_t ob = 31 - tpw::clz(qb); op = base + ob; span = ltm & ~bits_below(ob + 1); }
        else { op = open_pos; span = ltm; sp_ = c_special; nk_ = c_nonkey; hi_ = c_hi; b_ = c_bs; }
        tpos = op + 1; tlen = p - op - 1;
        meta |= K_STR;
        if (sp_ | (m_special & span)) meta |= RM_SPECIAL;
        if (nk_ | (m_nonkey & span)) meta |= RM_NONKEY;
        if (hi_ | (m_hi & span)) meta |= RM_HI;
        if (b_ | (m_bs & span)) meta |= RM_BS;
      } else {
        meta |= K_NUM;
        const uint32_t e = ~other & ~bits_below(l + 1);           // first byte after the run, within the chunk
        if (e) tlen = tpw::ffs(e) - 1 - l; else { tlen = 32 - l; meta |= RM_OPENEND; }
      }
      const uint32_t r = (head + rcount + tpw::popc(prevT)) & (RING - 1);
      sh.ring_pos[r] = tpos; sh.ring_len[r] = tlen; sh.ring_meta[r] = meta;
    }
    // carries
    if (T) {
      const uint32_t after = ~bits_below(32 - tpw::clz(T));
      c_ncomma = sat3(tpw::popc(comma_out & after)); c_ncolon = sat3(tpw::popc(colon_out & after));
    } else { c_ncomma = sat3(c_ncomma + tpw::popc(comma_out)); c_ncolon = sat3(c_ncolon + tpw::popc(colon_out)); }
    if (in_string) {
      if (openq) {                                               // the string still open was opened in this chunk
        const uint32_t ob = 31 - tpw::clz(openq), after = ~bits_below(ob + 1);
        open_pos = base + ob;
        c_special = m_special & after; c_nonkey = m_nonkey & after; c_hi = m_hi & after; c_bs = m_bs & after;
      } else { c_special |= m_special; c_nonkey |= m_nonkey; c_hi |= m_hi; c_bs |= m_bs; }
    }
    rcount += tpw::popc(T);
    tpw::sync();
    while (rcount >= 33 && !st.status) {
      const uint32_t la = (sh.ring_meta[(head + 32) & (RING - 1)] >> RM_NCOLON_SH) & 3u;
      p1_batch(s, n, toks, tok_cap, sh, st, head, 32, la);
      head += 32; rcount -= 32;
    }
    if (st.status) break;
  }
  if (!st.status) {
    if (in_string || c_ncomma || c_ncolon) st.status = TS_NOT_JSON;     // unterminated string / separators after the last token
    while (rcount && !st.status) {
      const uint32_t m = rcount > 32 ? 32u : rcount;
      const uint32_t la = rcount > 32 ? ((sh.ring_meta[(head + 32) & (RING - 1)] >> RM_NCOLON_SH) & 3u) : 0u;
      p1_batch(s, n, toks, tok_cap, sh, st, head, m, la);
      head += m; rcount -= m;
    }
    if (!st.status && (st.sp != 0 || st.root_cnt != 1)) st.status = TS_NOT_JSON;
  }
  *ntok_out = st.ntok;
  return st.status;
}

// ----------------------------------------------------------------------------------------------------------------------
// pass 2: tokens -> TOON text
// ----------------------------------------------------------------------------------------------------------------------
struct Piece {
  uint32_t l0;        // literal before the line break: 0 or ':'
  uint32_t nl;        // 1 = line break + `spaces` blanks
  uint32_t spaces;
  uint32_t l1, l1n;   // literal after the indentation (up to 2 chars, low byte first)
  uint32_t body;      // B_*
  uint32_t tail;      // array header tail: T_*
};
enum : uint32_t { B_NONE = 0, B_SPAN = 1, B_QSPAN = 2, B_ESCX = 3, B_QESCX = 4, B_ARR = 5 };
enum : uint32_t { T_COLON = 0, T_COLON_SP = 1, T_COLUMNAR = 2 };
TP_FN uint32_t lit2(char a, char b) { return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8); }

TP_FN int pass2(const uint8_t* s, const GTok* toks, uint32_t ntok, uint8_t* out, uint32_t out_cap, uint32_t* out_len, Shared& sh, bool report_errors) {
  const uint32_t l = tpw::lane();
  const uint32_t ltm = tpw::lt_mask();
  uint32_t* f_mode = sh.open_idx; uint32_t* f_cnt = sh.cnt; uint32_t* f_pre = sh.cfl; uint32_t* f_ind = sh.khbase;
  uint32_t sp = 0, root_cnt = 0, ocur = 0;
  int status = 0;
  bool over = false;
  for (uint32_t base = 0; base < ntok; base += 32) {
    const uint32_t m = ntok - base > 32 ? 32u : ntok - base;
    const bool act = l < m;
    GTok t; t.pos = 0; t.w = K_CLOSE_ARR;
    if (act) t = toks[base + l];
    const uint32_t kind = gt_kind(t.w), fl = gt_flags(t.w), len = gt_len(t.w);
    uint32_t nk = tpw::shfl_down(kind, 1);                         // kind of the following token
    if (l + 1 >= m) nk = base + m < ntok ? gt_kind(toks[base + m].w) : (uint32_t)K_CLOSE_ARR;
    Piece pc; pc.l0 = 0; pc.nl = 0; pc.spaces = 0; pc.l1 = 0; pc.l1n = 0; pc.body = B_NONE; pc.tail = T_COLON;
    uint32_t err = 0;                                             // per-lane TS_* error of this token
    const bool isV = act && kind >= K_STR && kind != K_KEY;
    uint32_t evm = tpw::ballot(act && kind <= K_CLOSE_ARR);
    const uint32_t Vm = tpw::ballot(isV);
    uint32_t cur = 0;
    while (true) {
      const uint32_t e = evm ? tpw::ffs(evm) - 1 : m;
      const uint32_t run = range_mask(cur, e);
      if (run) {
        const bool inrun = (run >> l) & 1u;
        if (sp == 0) { root_cnt += tpw::popc(run); if (inrun) pc.body = B_SPAN; }
        else {
          const uint32_t top = sp - 1, mode = f_mode[top], pre = f_pre[top], ind = f_ind[top], c0 = f_cnt[top];
          const uint32_t Vr = Vm & run;
          const uint32_t ord = c0 + tpw::popc(Vr & ltm);
          if (inrun) {
            if (kind == K_KEY) {
              if (mode == M_OBJ) { pc.nl = (pre == 0 && ord == 0 && sp == 1) ? 2u : 1u; pc.spaces = pre; pc.body = B_SPAN; }
              else if (mode == M_LIST_ITEM) {
                pc.nl = 1; pc.spaces = pre + (ord == 0 ? 2 * ind : 2 * (ind + 1));
                if (ord == 0) { pc.l1 = lit2('-', ' '); pc.l1n = 2; }
                pc.body = B_SPAN;
              }
            } else {
              pc.body = B_SPAN;
              if (mode == M_OBJ || mode == M_LIST_ITEM) { pc.l1 = lit2(':', ' '); pc.l1n = 2; }
              else if (mode == M_ROW || mode == M_ARR_INLINE) { if (ord > 0) { pc.l1 = ','; pc.l1n = 1; } }
              else if (mode == M_ARR_ITEMS) { pc.nl = 1; pc.spaces = pre + 2 * (ind + 1); pc.l1 = lit2('-', ' '); pc.l1n = 2; }
            }
          }
          tpw::sync();
          if (l == 0) f_cnt[top] = c0 + tpw::popc(Vr);
          tpw::sync();
        }
      }
      if (e >= m) break;
      const uint32_t ek = tpw::shfl(kind, e);
      if (ek <= K_OPEN_ARR) {
        const uint32_t en = tpw::shfl(len, e), efl = tpw::shfl(fl, e), enk = tpw::shfl(nk, e);
        uint32_t pmode = M_ROOT, pre = 0, ind = 0, ord = 0;
        if (sp == 0) ++root_cnt;
        else { const uint32_t top = sp - 1; pmode = f_mode[top]; pre = f_pre[top]; ind = f_ind[top]; ord = f_cnt[top]; tpw::sync(); if (l == 0) f_cnt[top] = ord + 1; }
        uint32_t nmode = M_DEAD, npre = 0, nind = 0;
        Piece q; q.l0 = 0; q.nl = 0; q.spaces = 0; q.l1 = 0; q.l1n = 0; q.body = B_NONE; q.tail = T_COLON;
        uint32_t eerr = 0;
        if (ek == K_OPEN_OBJ) {
          if (pmode == M_ROOT) { nmode = M_OBJ; npre = 0; nind = 0; }
          else if (pmode == M_OBJ) { q.l1 = ':'; q.l1n = 1; nmode = M_OBJ; npre = pre + 2; nind = ind + 1; }
          else if (pmode == M_ARR_ITEMS) {
            if (en == 0) { q.nl = 1; q.spaces = pre + 2 * (ind + 1); q.l1 = '-'; q.l1n = 1; }
            else { nmode = M_LIST_ITEM; npre = pre; nind = ind + 1; }
          } else if (pmode == M_LIST_ITEM) {
            if (en == 0) { q.l1 = lit2(':', ' '); q.l1n = 2; }
            else { q.l1 = ':'; q.l1n = 1; nmode = M_OBJ; npre = pre + 2 * (ind + 1) + 2; nind = ind + 2; }
          } else if (pmode == M_ARR_COL) { q.nl = 1; q.spaces = pre; nmode = M_ROW; }
          if (en == 0) nmode = M_DEAD;
        } else {
          const uint32_t amode = efl & 3u;
          uint32_t apre = pre, aind = ind;                           // arguments of begin_array
          bool col_on_hyphen = false;
          q.body = B_ARR;
          if (pmode == M_ARR_ITEMS) { const uint32_t ci = 2 * (ind + 1); q.nl = 1; q.spaces = pre + ci; q.l1 = lit2('-', ' '); q.l1n = 2; apre = pre + ci + 2; aind = ind + 2; }
          else if (pmode == M_LIST_ITEM) {
            const uint32_t fi = 2 * (ind + 1);
            if (en == 0) { q.l1 = lit2(':', ' '); q.l1n = 2; }
            else {
              if (ord == 0) {                                       // toon.py:400-404: columnar attempt without a type check
                if (enk != K_OPEN_OBJ) eerr = TS_ATTR_ERROR;
                else if (efl & AF_MIXED) eerr = TS_FALLBACK | (FB_MIXED_ITEM << 8);
                else if (amode == AM_COLUMNAR) col_on_hyphen = true;
              }
              apre = pre + fi + 2; aind = ind + 2;
              if (!col_on_hyphen) { q.l0 = ':'; q.nl = 1; q.spaces = apre; }
            }
          }
          if (amode == AM_EMPTY) q.tail = T_COLON;
          else if (amode == AM_COLUMNAR) { q.tail = T_COLUMNAR; nmode = M_ARR_COL; npre = apre + 2; nind = aind; }
          else if (amode == AM_INLINE) { q.tail = T_COLON_SP; nmode = M_ARR_INLINE; }
          else { q.tail = T_COLON; nmode = M_ARR_ITEMS; npre = apre; nind = aind; }
          if (col_on_hyphen) npre = apre;
        }
        if (l == e) { pc = q; err = eerr; }
        if (sp >= MAXD) { status = TS_UNSUPPORTED; break; }
        tpw::sync();
        if (l == 0) { f_mode[sp] = nmode; f_pre[sp] = npre; f_ind[sp] = nind; f_cnt[sp] = 0; }
        ++sp;
        tpw::sync();
      } else {
        --sp;
      }
      cur = e + 1;
      evm &= evm - 1;
    }
    if (status) break;

    // ---- piece lengths
    uint32_t blen = 0;
    const uint8_t* src = s + t.pos;
    if (act && pc.body == B_SPAN) {
      if (kind == K_KEY) { if (!(fl & KF_KEYOK)) pc.body = B_QSPAN; }
      else if (kind == K_STR) {
        const bool q = (fl & SF_Q) != 0;
        if (q && (fl & SF_CTRLERR)) err = TS_VALUE_ERROR;
        pc.body = (fl & SF_ESCX) ? (q ? B_QESCX : B_ESCX) : (q ? B_QSPAN : B_SPAN);
      }
    }
    uint32_t hk = 0;                                              // columnar header: number of keys
    if (act) {
      if (pc.body == B_SPAN) blen = len;
      else if (pc.body == B_QSPAN) blen = len + 2;
      else if (pc.body == B_ESCX || pc.body == B_QESCX) blen = escx_len(src, len, pc.body == B_QESCX);
      else if (pc.body == B_ARR) {
        blen = 2 + dec_digits(len) + (pc.tail == T_COLON_SP ? 2u : 1u);
        if (pc.tail == T_COLUMNAR) {
          hk = gt_len(toks[base + l + 1].w);                      // members of the first row
          blen += 2 + (hk - 1);
          for (uint32_t j = 0; j < hk; ++j) blen += gt_len(toks[base + l + 2 + 2 * j].w);
        }
      }
    }
    const uint32_t plen = act ? ((pc.l0 ? 1u : 0u) + (pc.nl == 1 ? 1u : 0u) + (pc.nl ? pc.spaces : 0u) + pc.l1n + blen) : 0u;
    const uint32_t incl = tpw::scan_incl(plen);
    const uint32_t off = ocur + incl - plen;
    const uint32_t total = tpw::shfl(incl, 31);
    // first error / first overflow in token order
    const uint32_t errm = tpw::ballot(err != 0), ovm = tpw::ballot(plen && off + plen > out_cap);
    if (errm) {
      const uint32_t fe = tpw::ffs(errm) - 1;
      const bool ov_first = ovm && (tpw::ffs(ovm) - 1) < fe;
      if (report_errors || !(over || ov_first)) { status = (int)tpw::shfl(err, fe); break; }
    }
    if (ovm) { over = true; if (!report_errors) { status = TS_NOT_SMALLER; break; } }
    // ---- write
    const bool longspan = act && (pc.body == B_SPAN || pc.body == B_QSPAN) && len >= 64;
    if (act && plen) {
      Emit em; em.out = out; em.cap = out_cap; em.o = off;
      if (pc.l0) em.put(pc.l0);
      if (pc.nl == 1) em.put('\n');
      if (pc.nl) em.spaces(pc.spaces);
      if (pc.l1n >= 1) em.put(pc.l1 & 0xFF);
      if (pc.l1n >= 2) em.put((pc.l1 >> 8) & 0xFF);
      if (pc.body == B_SPAN) { if (!longspan) em.span(src, len); }
      else if (pc.body == B_QSPAN) { em.put('"'); if (!longspan) em.span(src, len); else em.o += len; em.put('"'); }
      else if (pc.body == B_ESCX || pc.body == B_QESCX) em.escx(src, len, pc.body == B_QESCX);
      else if (pc.body == B_ARR) {
        em.put('['); em.uint_dec(len); em.put(']');
        if (pc.tail == T_COLUMNAR) {
          em.put('{');
          for (uint32_t j = 0; j < hk; ++j) { const GTok k = toks[base + l + 2 + 2 * j]; if (j) em.put(','); em.span(s + k.pos, gt_len(k.w)); }
          em.put('}'); em.put(':');
        } else { em.put(':'); if (pc.tail == T_COLON_SP) em.put(' '); }
      }
    }
    uint32_t lm = tpw::ballot(longspan);
    while (lm) {                                                  // long spans: the whole warp copies
      const uint32_t j = tpw::ffs(lm) - 1;
      lm &= lm - 1;
      const uint32_t jpos = tpw::shfl(t.pos, j), jlen = tpw::shfl(len, j);
      const uint32_t jdst = tpw::shfl(off + plen - blen + (pc.body == B_QSPAN ? 1u : 0u), j);
      for (uint32_t i = l; i < jlen; i += 32) if (jdst + i < out_cap) out[jdst + i] = s[jpos + i];
    }
    ocur += total;
  }
  if (status) return status;
  if (over || ocur > out_cap) return TS_NOT_SMALLER;
  *out_len = ocur;
  return TS_CONVERTED;
}

// Whole per-unit pipeline (all 32 lanes call it with the same arguments).  out_cap = n - 1 in the product (a
// conversion is only kept when strictly smaller).  Returns a TS_* status, TS_FALLBACK when the sequential encoder has
// to redo the unit.
TP_FN int toon_unit(const uint8_t* s, uint32_t n, GTok* toks, uint32_t tok_cap, uint8_t* out, uint32_t out_cap, uint32_t* out_len, Shared& sh,
                    const uint16_t* ctab, bool report_errors) {
  uint32_t ntok = 0;
  const int st = pass1(s, n, toks, tok_cap, sh, ctab, &ntok);
  if (st) return st;
  tpw::sync();
  return pass2(s, toks, ntok, out, out_cap, out_len, sh, report_errors);
}

}  // namespace cftp
