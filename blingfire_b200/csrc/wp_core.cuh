// wp_core.cuh -- the per-chunk lexer + WordPiece routine of the fused kernel.
//
// One lane owns one "chunk": a maximal stretch of top-level start positions between two
// sync points (positions no top-level match can cross).  Inside its chunk the
// lane runs the reference's loop verbatim in structure:
//   outer loop  = FALexTools_t::Process_int at depth 1 (FALexTools_t.h:229-397)
//   inner loop  = the same function at depth 2 over the word span (the `_call FnTokWord`
//                 sub-grammar, FALexTools_t.h:350-382)
//   tiling rule = TextToIdsWithOffsets_wp's post-pass (blingfiretokdll.cpp:1221-1306)
// specialised by the load-time FastPath conditions (lexer_tables.h): zero contexts, top-level
// tags <= 4, one function per action, function actions = plain ids > 4.
//
// The routine is __host__ __device__ so that tests/twin can run the identical code over the
// identical flattened tables on the CPU; the product only ever calls it from wp_kernel.cu.
#pragma once

#include <climits>
#include <cstdint>

#if defined(__CUDACC__)
#define BF_HD __host__ __device__ __forceinline__
#else
#define BF_HD inline
#endif

namespace bfb200 {

// Small, read-mostly part of the model.  On the device every pointer below addresses shared
// memory (the blob is staged once per CTA with a bulk async copy); the twin points them at
// host memory.
struct WpTop {
  const uint16_t* ascii_cls;      // [128] class of code points < 128 (charmap + clamp folded)
  const uint8_t* tc_of_class;     // [NC+1]
  const uint8_t* ttop;            // [K*NT]
  const unsigned long long* cross;// [NT]
  const uint8_t* top_final;       // [K]
  const int32_t* top_tag;         // [K]
  const uint32_t* top_fn_root;    // [K] global id or none
  const uint32_t* top_fn_caret;   // [K] global id or none
  const int8_t* top_row_root;     // [K] index of the staged copy of trans[top_fn_root], -1 if not staged
  const int8_t* top_row_caret;    // [K] same for top_fn_caret
  const void* staged_rows;        // [R][NC+1] copies of hot transition rows
  const uint8_t* sync_start;      // [1<<sync_shift][1<<sync_shift] (previous top class, top class) -> a chunk may start here:
                                  //           no walk crosses the pair, and some match starts with the class
  int K, NT;
  uint8_t tc_caret, tc_dollar, tc_none, sync_shift;
};

template <typename TE>
struct WpGlobal {
  const TE* trans;                // [NS][NC+1] dense transition table in HBM
  const int32_t* tag_of_state;    // [NS]
  const uint16_t* cls_of_cp;      // [0x110000]
  uint32_t NC1;                   // NC + 1
  uint32_t first_final;
  uint32_t cls_caret, cls_dollar;
  int max_token_length;
};

template <typename TE> struct TeTraits;
template <> struct TeTraits<uint16_t> { static constexpr uint32_t none = 0xFFFFu; };
template <> struct TeTraits<uint32_t> { static constexpr uint32_t none = 0xFFFFFFFFu; };

constexpr uint32_t kNone32 = 0xFFFFFFFFu;

#if defined(__CUDA_ARCH__)
template <typename T> __device__ __forceinline__ T bf_ldg(const T* p) { return __ldg(p); }
#else
template <typename T> inline T bf_ldg(const T* p) { return *p; }
#endif

// One step in the global table.  `q` must be a valid state.
template <typename TE>
BF_HD uint32_t wp_step(const WpGlobal<TE>& g, uint32_t q, uint32_t c) {
  // 16-bit tables have < 65 536 states and <= 65 537 columns: the flat index fits 32 bits
  const uint32_t v = sizeof(TE) == 2 ? bf_ldg(g.trans + (uint32_t)(q * g.NC1 + c)) : bf_ldg(g.trans + ((size_t)q * g.NC1 + c));
  return v == TeTraits<TE>::none ? kNone32 : v;
}
template <typename TE>
BF_HD uint32_t wp_row_step(const void* rows, int row, uint32_t nc1, uint32_t c) {
  const uint32_t v = reinterpret_cast<const TE*>(rows)[(size_t)row * nc1 + c];
  return v == TeTraits<TE>::none ? kNone32 : v;
}

// ids_at[p] of a position where no piece starts (a real id is a rule tag or the caller's UnkId)
constexpr int32_t kNoPiece = INT_MIN;
constexpr uint8_t kTopFinal = 0x80;      // ttop entries: bit 7 = the destination is final (K <= 127 states)

// The function sub-grammar over the word span cls[w0..w1] (inclusive), i.e. Process_int with
// Initial = FnIni at RecDepth 2.  Pieces are written position-indexed: ids_at[p] for a piece starting
// at p (kNoPiece elsewhere).  Returns true when the pieces tile the word exactly.
template <typename TE>
BF_HD bool wp_word(const WpTop& t, const WpGlobal<TE>& g, const uint16_t* cls, int w0, int w1,
                   uint32_t root, uint32_t caret, int row_root, int row_caret,
                   int32_t* ids_at) {
  const int L = w1 - w0 + 1;
  int expect = w0;          // ExpectedFrom of the post-pass
  bool contiguous = true;   // still inside the leading run of gap-free sub-tokens
  int nsub = 0;
  for (int from = -1; from < L; ++from) {
    uint32_t q;
    int j = from;
    int bound = from + g.max_token_length;
    if (bound > L) bound = L;
    bool first_from_staged;
    int row;
    if (j == -1) {                       // left anchor only at from == -1 (FALexTools_t.h:244-252)
      if (caret == kNone32) continue;
      q = caret; j = 0; row = row_caret;
    } else {
      q = root; row = row_root;
    }
    first_from_staged = row >= 0;
    uint32_t fq = kNone32;
    int fpos = -1;
    if (j < bound) {
      // the first hop comes from the staged copy of the row when there is one; then the table
      uint32_t d = first_from_staged ? wp_row_step<TE>(t.staged_rows, row, g.NC1, cls[w0 + j]) : wp_step(g, q, cls[w0 + j]);
      first_from_staged = false;
      while (d != kNone32) {
        if (d >= g.first_final) { fq = d; fpos = j; }
        q = d;
        if (++j >= bound) break;
        d = wp_step(g, q, cls[w0 + j]);
      }
    }
    if (j == L) {                        // right anchor only when the walk consumed the span (:280-290)
      uint32_t d;
      if (first_from_staged) d = wp_row_step<TE>(t.staged_rows, row, g.NC1, g.cls_dollar);
      else d = wp_step(g, q, g.cls_dollar);
      if (d != kNone32 && d >= g.first_final) { fq = d; fpos = j; }
    }
    if (fpos == -1) continue;
    const int f2 = from < 0 ? 0 : from;                 // clamp(From + 0, 0, L-1)
    const int t2 = fpos > L - 1 ? L - 1 : fpos;         // clamp(FinalPos - 0, 0, L-1)
    const int32_t tag = bf_ldg(g.tag_of_state + fq);
    ids_at[w0 + f2] = tag;
    if (contiguous && w0 + f2 == expect) { expect = w0 + t2 + 1; ++nsub; }   // tag > 4 by FastPath
    else contiguous = false;
    if (fpos > from) from = fpos;                       // resume after the token (:389-393)
  }
  return nsub > 0 && expect - 1 == w1;
}

// Runs the top-level loop for start positions from_begin <= From < from_end.  `m` is the number
// of classes available in `cls`; `at_doc_end` says whether position m is the end of the
// document (right anchor) or just the end of the current window.  Returns the first From not
// processed (>= from_end).
template <typename TE>
BF_HD int wp_chunk(const WpTop& t, const WpGlobal<TE>& g, const uint16_t* cls, int m, bool at_doc_end,
                   int from_begin, int from_end, int unk_id, int32_t* ids_at, const uint8_t* tcs) {
  int from = from_begin;
  for (; from < from_end; ++from) {
    uint32_t q = 0;                      // local id of the initial state
    int j = from;
    int bound = from + g.max_token_length;
    if (bound > m) bound = m;
    if (j == -1) {
      const uint8_t d = t.ttop[t.tc_caret];              // row 0
      if (d == 0xFF) continue;
      q = d & ~kTopFinal; j = 0;
    }
    int fq = -1, fpos = -1;
    for (; j < bound; ++j) {
      const uint8_t d = t.ttop[q * t.NT + tcs[j]];
      if (d == 0xFF) break;
      q = d & ~kTopFinal;
      if (d & kTopFinal) { fq = (int)q; fpos = j; }
    }
    if (j == m && at_doc_end) {
      const uint8_t d = t.ttop[q * t.NT + t.tc_dollar];
      if (d != 0xFF && (d & kTopFinal)) { fq = d & ~kTopFinal; fpos = j; }
    }
    if (fpos == -1) continue;
    const int f2 = from < 0 ? 0 : from;
    const int t2 = fpos > m - 1 ? m - 1 : fpos;
    if (t.top_tag[fq] == 1) {            // WBD_WORD_TAG (blingfiretokdll.cpp:38, :1221)
      bool tiled = false;
      const uint32_t root = t.top_fn_root[fq];
      if (root != kNone32)
        tiled = wp_word<TE>(t, g, cls, f2, t2, root, t.top_fn_caret[fq], t.top_row_root[fq], t.top_row_caret[fq], ids_at);
      if (!tiled) {                      // not covered without gaps -> one UnkId (:1282-1301)
        for (int p = f2 + 1; p <= t2; ++p) ids_at[p] = kNoPiece;
        ids_at[f2] = unk_id;
      }
    }
    if (fpos > from) from = fpos;
  }
  return from;
}

}  // namespace bfb200
