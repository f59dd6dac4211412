// lex_core.cuh -- the GENERIC lexer: FALexTools_t::Process_int (FALexTools_t.h:205-400) for any
// [wbd] grammar -- left/right contexts, tag-less actions, several functions per action, the
// `once` rule, nested calls down to max-depth -- over the flattened tables.
//
// The reference recurses; here the recursion is an explicit frame stack so the same code runs
// as one GPU thread per document (lex_kernel.cu) and, for the CPU-side checks, inside the test
// twin.  Semantics are restated one-to-one; every branch cites the reference line it mirrors.
#pragma once

#include <cstdint>

#include "wp_core.cuh"   // BF_HD, bf_ldg, TeTraits, kNone32

namespace bfb200 {

constexpr int kMaxLexDepth = 8;

template <typename TE>
struct LexGlobal {
  const TE* trans;               // [NS][NC+1], IW_ANY fallback folded in
  const int32_t* ow_of_state;    // [NS]
  const int32_t* act_begin;      // [num_acts+1]
  const int32_t* act_data;
  const uint32_t* fn_ini;        // [fn_count] global state or kNone32
  int fn_count;
  uint32_t NC1, first_final, cls_caret, cls_dollar, initial;
  int max_depth, max_token_length;
};

struct LexFrame {
  uint32_t ini;
  int base;          // index of the span start in cls[] (== the reference's Offset)
  int n;             // span length (InSize)
  int from;          // FromPos
  int once;          // fOnce
  int state;         // 0 scan, 1 iterate functions, 2 returned from a call
  int fn_idx, act_end, fn_once, fn_from, to2, resume;
  int out_at_call;   // OutSize when the child was entered
};

// Runs Process(pIn, InSize, pOut, MaxOutSize) (FALexTools_t.h:403-421) on the class sequence
// cls[0..n).  Writes (Tag, From, To) triples, returns the number of ints written.
template <typename TE>
BF_HD int lex_process(const LexGlobal<TE>& g, const uint16_t* cls, int n, int32_t* out, int max_out) {
  LexFrame st[kMaxLexDepth];
  int depth = 0, out_size = 0;
  if (g.max_depth < 1) return 0;                                  // :222 at RecDepth 1
  st[0].ini = g.initial; st[0].base = 0; st[0].n = n; st[0].from = -1; st[0].once = 0; st[0].state = 0;
  while (depth >= 0) {
    LexFrame& F = st[depth];
    if (F.state == 2) {                                           // a call returned (:371-381)
      if (out_size - F.out_at_call > 0) {
        F.fn_from = out[out_size - 1] + 1 - F.base;
        if (F.fn_from > F.to2) F.fn_idx = F.act_end;
      }
      F.state = 1;
    }
    if (F.state == 1) {
      if (F.fn_idx < F.act_end) {                                 // next function of the action (:350-365)
        const int fn = bf_ldg(g.act_data + F.fn_idx);
        ++F.fn_idx;
        F.out_at_call = out_size;
        F.state = 2;
        if (g.max_depth < depth + 2) continue;                    // the callee returns 0 at once (:222)
        if (depth + 1 >= kMaxLexDepth) continue;
        const uint32_t ini = (fn >= 0 && fn < g.fn_count) ? bf_ldg(g.fn_ini + fn) : kNone32;
        LexFrame& C = st[depth + 1];
        C.ini = ini; C.base = F.base + F.fn_from; C.n = F.to2 - F.fn_from + 1; C.from = -1;
        C.once = fn == 0 ? 0 : F.fn_once; C.state = 0;
        ++depth;
        continue;
      }
      if (F.once) { --depth; continue; }                          // :385-387
      if (F.resume > F.from) F.from = F.resume;                   // :390-393
      ++F.from;
      F.state = 0;
    }
    // ---- scan: one start position (:229-397) ----
    if (F.from >= F.n || F.ini == kNone32) { --depth; continue; }
    const int from = F.from;
    uint32_t q = F.ini;
    int j = from;
    int bound = from + g.max_token_length;
    if (F.n < bound) bound = F.n;
    if (j == -1) {                                                // left anchor (:244-252)
      const uint32_t d = bf_ldg(g.trans + (size_t)q * g.NC1 + g.cls_caret);
      if (d == TeTraits<TE>::none) { ++F.from; continue; }
      q = d; j = 0;
    }
    uint32_t fq = kNone32;
    int fpos = -1;
    for (; j < bound; ++j) {                                      // :255-277
      const uint32_t d = bf_ldg(g.trans + (size_t)q * g.NC1 + cls[F.base + j]);
      if (d == TeTraits<TE>::none) break;
      if (d >= g.first_final) { fq = d; fpos = j; }
      q = d;
    }
    if (j == F.n) {                                               // right anchor (:280-290)
      const uint32_t d = bf_ldg(g.trans + (size_t)q * g.NC1 + g.cls_dollar);
      if (d != TeTraits<TE>::none && d >= g.first_final) { fq = d; fpos = j; }
    }
    if (fpos == -1) { ++F.from; continue; }
    const int ow = bf_ldg(g.ow_of_state + fq);                    // :297-305
    const int a0 = bf_ldg(g.act_begin + ow), a1 = bf_ldg(g.act_begin + ow + 1);
    const int left = bf_ldg(g.act_data + a0), right = bf_ldg(g.act_data + a0 + 1), tag = bf_ldg(g.act_data + a0 + 2);
    int from2 = from + left;                                      // :308-316
    if (from2 < 0) from2 = 0; else if (F.n <= from2) from2 = F.n - 1;
    int to2 = fpos - right;                                       // :319-327
    if (to2 < 0) to2 = 0; else if (F.n <= to2) to2 = F.n - 1;
    int fn_idx = a0 + 3;
    if (tag != 0) {                                               // :332-342
      if (out_size + 3 <= max_out) {
        out[out_size++] = tag; out[out_size++] = from2 + F.base; out[out_size++] = to2 + F.base;
      } else {
        return out_size;   // buffer full: nothing can be emitted any more, the result is final
      }
      fn_idx = a0 + 4;
    }
    F.fn_once = 1 < (a1 - fn_idx);                                // :345
    F.fn_idx = fn_idx; F.act_end = a1; F.fn_from = from2; F.to2 = to2; F.resume = fpos - right;
    F.state = 1;
  }
  return out_size;
}

// TextToIdsWithOffsets_wp's post-pass over the triples (blingfiretokdll.cpp:1207-1313), for
// lexer models outside the FastPath shape.  Returns the number of ids written (<= max_ids).
BF_HD int wp_postpass(const int32_t* res, int rn, int32_t* ids, int max_ids, int unk_id) {
  int out = 0;
  for (int i = 0; i < rn; i += 3) {
    const int tag = res[i];
    if (tag == 4) continue;                                       // WBD_IGNORE_TAG
    if (tag == 1) {                                               // WBD_WORD_TAG
      const int tfrom = res[i + 1], tto = res[i + 2];
      int j = i + 3, nsub = 0;
      bool covered = false;
      if (j < rn) {
        int expect = tfrom, stag = res[j], sfrom = res[j + 1], sto = res[j + 2];
        while (j <= rn && stag > 4 && expect == sfrom) {
          expect = sto + 1; ++nsub; j += 3;
          if (j < rn) { stag = res[j]; sfrom = res[j + 1]; sto = res[j + 2]; }
        }
        if (nsub > 0 && expect - 1 == tto) {
          for (int k = 0; k < nsub && out < max_ids; ++k) ids[out++] = res[(k + 1) * 3 + i];
          covered = true;
        }
      }
      if (!covered && out < max_ids) ids[out++] = unk_id;
      i = j - 3;
    }
    if (out >= max_ids) break;
  }
  return out;
}

}  // namespace bfb200
