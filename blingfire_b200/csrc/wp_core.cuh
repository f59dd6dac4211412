// wp_core.cuh -- the per-chunk lexer + WordPiece routines of the fused kernel.
//
// A "chunk" is a maximal stretch of top-level start positions between two sync points (positions no
// top-level match can cross).  Two ways to serve one, both exact:
//
//   * the reference's loops, verbatim in structure (wp_chunk / wp_word):
//       outer loop  = FALexTools_t::Process_int at depth 1 (FALexTools_t.h:229-397)
//       inner loop  = the same function at depth 2 over the word span (the `_call FnTokWord`
//                     sub-grammar, FALexTools_t.h:350-382)
//       tiling rule = TextToIdsWithOffsets_wp's post-pass (blingfiretokdll.cpp:1221-1306)
//     specialised by the load-time FastPath conditions (lexer_tables.h): zero contexts, top-level
//     tags <= 4, one function per action, function actions = plain ids > 4;
//
//   * a MEMO of those loops for the common case: a chunk whose positions all have the same top-level
//     class is classified by a per-class table computed at load time by running the top-level loop on
//     such runs (kind_of_tc: "emits nothing", or "one WORD token spanning the run"), and a WORD run of
//     at most kMaxFastLen positions is looked up as a whole in a table that holds every class sequence
//     for which wp_word yields exactly ONE piece (the packed sequence itself is the key, so a hit is
//     exact, not probabilistic).  Everything the memo does not hold goes through the loops above.
//
// The routines are __host__ __device__ so that tests/twin can run the identical code over the
// identical flattened tables on the CPU, and so that the load-time builder (wp_model.cpp) fills the
// memo by calling the very routine it abbreviates; the product only ever tokenizes from wp_kernel.cu.
#pragma once

#include <climits>
#include <cstdint>

#if defined(__CUDACC__)
#define BF_HD __host__ __device__ __forceinline__
#else
#define BF_HD inline
#endif

namespace bfb200 {

// chunk kinds per top-level class (WpTop::kind_of_tc)
constexpr uint32_t kKindInert = 1u;      // a run of this class, of any length, with or without anchors, emits no id
constexpr uint32_t kKindCaretOk = 2u;    // the left anchor (run at the start of the document) does not change the outcome
constexpr uint32_t kKindDollarOk = 4u;   // the right anchor (run at the end of the document) does not change the outcome
constexpr uint32_t kKindDead = 8u;       // positions of this class match nothing and start nothing (white space in bert_*)
constexpr int kKindLenShift = 8;         // bit (kKindLenShift + L): a run of L positions is ONE WORD token over [0, L) calling
                                         // the class's function pair; L = 1..kMaxFastLen
constexpr int kMaxFastLen = 12;
// sync_start entries
constexpr uint8_t kSyncStart = 1;        // a chunk may start at the second position of the pair
constexpr uint8_t kSyncGroupChange = 2;  // the two positions belong to different groups of top-level classes (wp_model.cpp)

// One slot of the whole-word table: the packed class sequence (with its length) and the piece id.
struct alignas(16) WpWordSlot {
  uint32_t kw[4];
  int32_t id;
  uint32_t pad[3];
};
static_assert(sizeof(WpWordSlot) == 32, "one 32-byte sector per slot");

// Cuckoo table: slot h1(key) of the first half or slot h2(key) of the second half.
struct WpWords {
  const WpWordSlot* slots;   // [2 << log2_size]
  uint32_t log2_size;        // slots per half (power of two)
  uint32_t cb;               // bits per class in the key
  uint32_t cpw;              // classes per 32-bit key word: 3, 2 or 1
  uint32_t max_len;          // longest run the key holds: min(kMaxFastLen, 4 * cpw); 0 = no table
  uint32_t mul[8];           // odd multipliers of the two hash functions
};

// Small, read-mostly part of the model.  On the device every pointer below addresses shared
// memory (the blob is staged once per CTA with a bulk async copy); the twin points them at
// host memory.
struct WpTop {
  const uint32_t* ascii_clsx;     // [128] class | top-level class << 16 of code points < 128 (charmap + clamp folded)
  const uint8_t* ttop;            // [K*NT] local transition; bit 7 = the destination is final; 0xFF = none
  const int32_t* top_tag;         // [K]
  const uint32_t* top_fn_root;    // [K] global id or none
  const uint32_t* top_fn_caret;   // [K] global id or none
  const uint8_t* sync_start;      // [1<<sync_shift][1<<sync_shift] (previous top class, top class) -> kSyncStart: a chunk may start
                                  //           here (no walk crosses the pair, and some match starts with the class);
                                  //           kSyncGroupChange: the classes belong to different groups
  const uint32_t* kind_of_tc;     // [NT] kKind* bits
  int K, NT;
  uint8_t tc_caret, tc_dollar, tc_none, sync_shift;
};

template <typename TE>
struct WpGlobal {
  const TE* trans;                // [NS][NC+1] dense transition table in HBM
  const int32_t* tag_of_state;    // [NS]
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

// ids_at[p] of a position where no piece starts (a real id is a rule tag or the caller's UnkId)
constexpr int32_t kNoPiece = INT_MIN;
constexpr uint8_t kTopFinal = 0x80;      // ttop entries: bit 7 = the destination is final (K <= 127 states)

// The function sub-grammar over the word span cls[w0..w1] (inclusive), i.e. Process_int with
// Initial = FnIni at RecDepth 2.  Pieces are written position-indexed: ids_at[p] for a piece starting
// at p (kNoPiece elsewhere).  Returns the number of pieces when they tile the word exactly, else 0.
template <typename TE>
BF_HD int wp_word(const WpGlobal<TE>& g, const uint16_t* cls, int w0, int w1, uint32_t root, uint32_t caret, int32_t* ids_at) {
  const int L = w1 - w0 + 1;
  int expect = w0;          // ExpectedFrom of the post-pass
  bool contiguous = true;   // still inside the leading run of gap-free sub-tokens
  int nsub = 0;
  for (int from = -1; from < L; ++from) {
    uint32_t q;
    int j = from;
    int bound = from + g.max_token_length;
    if (bound > L) bound = L;
    if (j == -1) {                       // left anchor only at from == -1 (FALexTools_t.h:244-252)
      if (caret == kNone32) continue;
      q = caret; j = 0;
    } else {
      q = root;
    }
    uint32_t fq = kNone32;
    int fpos = -1;
    while (j < bound) {
      const uint32_t d = wp_step(g, q, cls[w0 + j]);
      if (d == kNone32) break;
      if (d >= g.first_final) { fq = d; fpos = j; }
      q = d;
      ++j;
    }
    if (j == L) {                        // right anchor only when the walk consumed the span (:280-290)
      const uint32_t d = wp_step(g, q, g.cls_dollar);
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
  return (nsub > 0 && expect - 1 == w1) ? nsub : 0;
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
      int tiled = 0;
      const uint32_t root = t.top_fn_root[fq];
      if (root != kNone32) tiled = wp_word<TE>(g, cls, f2, t2, root, t.top_fn_caret[fq], ids_at);
      if (!tiled) {                      // not covered without gaps -> one UnkId (:1282-1301)
        for (int p = f2 + 1; p <= t2; ++p) ids_at[p] = kNoPiece;
        ids_at[f2] = unk_id;
      }
    }
    if (fpos > from) from = fpos;
  }
  return from;
}

// ---- the whole-word memo ----

// Packs the class sequence cls[0..L) and its length into four 32-bit words, CPW classes of `cb` bits per
// word (CPW * cb <= 30; bits 30-31 of words 0 and 1 hold L).  `lcap` >= L bounds the unrolled loop (on
// the device: the warp-wide maximum, so the lanes stay together).  Injective for L <= 4 * CPW.
template <int CPW>
BF_HD void wp_pack_key(const uint16_t* cls, int L, int lcap, uint32_t cb, uint32_t kw[4]) {
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint32_t v = 0;
    if (w * CPW < lcap) {
#pragma unroll
      for (int j = 0; j < CPW; ++j) {
        const int i = w * CPW + j;
        if (i < L) v += (uint32_t)cls[i] << (cb * (uint32_t)j);
      }
    }
    kw[w] = v;
  }
  kw[0] |= ((uint32_t)L & 3u) << 30;
  kw[1] |= ((uint32_t)L >> 2) << 30;
}
BF_HD void wp_pack_key_any(uint32_t cpw, const uint16_t* cls, int L, int lcap, uint32_t cb, uint32_t kw[4]) {
  if (cpw == 3) wp_pack_key<3>(cls, L, lcap, cb, kw);
  else if (cpw == 2) wp_pack_key<2>(cls, L, lcap, cb, kw);
  else wp_pack_key<1>(cls, L, lcap, cb, kw);
}

BF_HD uint32_t wp_key_hash(const uint32_t kw[4], const uint32_t* mul, uint32_t log2_size) {
  uint32_t h = kw[0] * mul[0] + kw[1] * mul[1] + kw[2] * mul[2] + kw[3] * mul[3];
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  return h >> (32u - log2_size);
}

// The piece id of the word with this key, or kNoPiece when the table does not hold it.
BF_HD int32_t wp_words_find(const WpWords& W, const uint32_t kw[4]) {
  const uint32_t h1 = wp_key_hash(kw, W.mul, W.log2_size);
  const uint32_t h2 = wp_key_hash(kw, W.mul + 4, W.log2_size) + (1u << W.log2_size);
#if defined(__CUDA_ARCH__)
  const uint4* s1 = reinterpret_cast<const uint4*>(W.slots + h1);
  const uint4* s2 = reinterpret_cast<const uint4*>(W.slots + h2);
  const uint4 a = __ldg(s1), b = __ldg(s2);
  const uint4 ai = __ldg(s1 + 1), bi = __ldg(s2 + 1);
  const bool m1 = ((a.x ^ kw[0]) | (a.y ^ kw[1]) | (a.z ^ kw[2]) | (a.w ^ kw[3])) == 0;
  const bool m2 = ((b.x ^ kw[0]) | (b.y ^ kw[1]) | (b.z ^ kw[2]) | (b.w ^ kw[3])) == 0;
  return m1 ? (int32_t)ai.x : (m2 ? (int32_t)bi.x : kNoPiece);
#else
  const WpWordSlot& a = W.slots[h1];
  const WpWordSlot& b = W.slots[h2];
  if (a.kw[0] == kw[0] && a.kw[1] == kw[1] && a.kw[2] == kw[2] && a.kw[3] == kw[3]) return a.id;
  if (b.kw[0] == kw[0] && b.kw[1] == kw[1] && b.kw[2] == kw[2] && b.kw[3] == kw[3]) return b.id;
  return kNoPiece;
#endif
}

// May the run [s, e) -- every position in one group of top-level classes, kind `kind` -- be served by the memo?  Returns
// 0 = no (run the loops), 1 = it emits nothing, 2 = look the word up.
BF_HD int wp_classify_run(uint32_t kind, int len, uint32_t max_len, bool at_doc_start, bool at_doc_end) {
  if (kind & kKindInert) return 1;       // established over the closure with both anchors
  if (at_doc_start && !(kind & kKindCaretOk)) return 0;
  if (at_doc_end && !(kind & kKindDollarOk)) return 0;
  if (len <= (int)max_len && ((kind >> (kKindLenShift + len)) & 1u)) return 2;
  return 0;
}

}  // namespace bfb200
