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
constexpr uint32_t kKindWordRun = 16u;   // a run of any length is ONE WORD token over the run, calling the group's function pair
constexpr uint32_t kKindWordOne = 32u;   // ... a run of length 1 is
constexpr int kMaxFastLen = 24;          // longest word the table's key holds (two halves of 12 classes at 10 bits)
// sync_start entries
constexpr uint8_t kSyncStart = 1;        // a chunk may start at the second position of the pair
constexpr uint8_t kSyncGroupChange = 2;  // the two positions belong to different groups of top-level classes (wp_model.cpp)

// One slot of the word table: the packed class sequence of a word (with its length) and its pieces.  Two 32-byte
// sectors; a word of at most half the maximal length with at most 3 pieces is served from the first one alone.
//   meta == 0            empty
//   meta == kSlotBusy    being written (run-time insertion)
//   meta & kSlotValid    n = meta & 7 pieces (0: the word is one UnkId), ids in id[0..3) and id_hi[0..3); piece 1 starts
//                        (meta >> 4) & 31 positions into the word, piece 2 (meta >> 9) & 31, pieces 3..5 at the bytes of offs_hi
struct alignas(32) WpWordSlot {
  uint32_t kw[4];
  int32_t id[3];
  uint32_t meta;
  uint32_t kw_hi[4];
  int32_t id_hi[3];
  uint32_t offs_hi;
};
static_assert(sizeof(WpWordSlot) == 64, "two 32-byte sectors per slot");
constexpr uint32_t kSlotBusy = 1u, kSlotValid = 0x80000000u;
constexpr int kMaxLearnPieces = 6;

// Two-choice table: slot h1(key) of the first half or slot h2(key) of the second half.  The single-piece words of the
// vocabulary are placed at load time (cuckoo, wp_model.cpp); words with several pieces (and words that are one
// UnkId) are added at RUN time by the warp that has just sent one through the lexer loops -- into an EMPTY candidate
// slot only, never over an entry, so a slot is written once and a reader can trust what it matches (wp_words_insert).
struct WpWords {
  WpWordSlot* slots;         // [2 << log2_size]
  uint32_t log2_size;        // slots per half (power of two)
  uint32_t cb;               // bits per class in the key
  uint32_t cpw;              // classes per 32-bit key word: 3, 2 or 1
  uint32_t max_len;          // longest run the key holds: min(kMaxFastLen, 8 * cpw); 0 = no table
  uint32_t mul[9];           // odd multipliers of the hash
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
  const uint32_t* fn_root_of_tc;  // [NT] function entry states of the class's group (word runs), or none
  const uint32_t* fn_caret_of_tc; // [NT]
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

// Packs the class sequence cls[0..L) and its length into eight 32-bit words, CPW classes of `cb` bits per word
// (CPW * cb <= 30; bits 30-31 of words 0..2 hold L).  Words 4..7 are only needed (and only computed) for words longer
// than 4 * CPW classes.  `lcap` >= L bounds the unrolled loop (on the device: the warp-wide maximum, so the lanes stay
// together).  Injective for L <= 8 * CPW.
template <int CPW>
BF_HD void wp_pack_key(const uint16_t* cls, int L, int lcap, uint32_t cb, uint32_t kw[8]) {
#pragma unroll
  for (int w = 0; w < 8; ++w) {
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
  kw[1] |= (((uint32_t)L >> 2) & 3u) << 30;
  kw[2] |= ((uint32_t)L >> 4) << 30;
}
BF_HD void wp_pack_key_any(uint32_t cpw, const uint16_t* cls, int L, int lcap, uint32_t cb, uint32_t kw[8]) {
  if (cpw == 3) wp_pack_key<3>(cls, L, lcap, cb, kw);
  else if (cpw == 2) wp_pack_key<2>(cls, L, lcap, cb, kw);
  else wp_pack_key<1>(cls, L, lcap, cb, kw);
}

// 32-bit mix of the key; the two slot indices are different bit ranges of two products of it.  `wide`: some word of
// the round is longer than half the maximum (the upper key words of the others are zero).
BF_HD uint32_t wp_key_hash(const uint32_t kw[8], const uint32_t* mul, bool wide) {
  uint32_t h = kw[0] * mul[0] + kw[1] * mul[1] + kw[2] * mul[2] + kw[3] * mul[3];
  if (wide) h += kw[4] * mul[5] + kw[5] * mul[6] + kw[6] * mul[7] + kw[7] * mul[8];
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  return h;
}
BF_HD uint32_t wp_slot1(const WpWords& W, uint32_t h) { return h >> (32u - W.log2_size); }
BF_HD uint32_t wp_slot2(const WpWords& W, uint32_t h) { return ((h * W.mul[4]) >> (32u - W.log2_size)) + (1u << W.log2_size); }

struct WpWordHit {
  uint32_t meta;       // 0: not in the table
  uint32_t offs_hi;
  int32_t id[6];
};

BF_HD bool wp_slot_matches(const WpWordSlot& a, const uint32_t kw[8], bool wide) {
  bool m = a.kw[0] == kw[0] && a.kw[1] == kw[1] && a.kw[2] == kw[2] && a.kw[3] == kw[3];
  if (wide) m = m && a.kw_hi[0] == kw[4] && a.kw_hi[1] == kw[5] && a.kw_hi[2] == kw[6] && a.kw_hi[3] == kw[7];
  return m;
}

// Looks a word up.  Slots are written once (meta goes 0 -> busy -> valid, the key and the ids are in place before
// it turns valid), so a valid slot whose key matches is complete; a slot seen half-way is simply "not found".
BF_HD WpWordHit wp_words_find(const WpWords& W, const uint32_t kw[8], bool wide) {
  const uint32_t h = wp_key_hash(kw, W.mul, wide);
  const uint32_t h1 = wp_slot1(W, h), h2 = wp_slot2(W, h);
  WpWordHit r;
#if defined(__CUDA_ARCH__)
  const uint4* s1 = reinterpret_cast<const uint4*>(W.slots + h1);
  const uint4* s2 = reinterpret_cast<const uint4*>(W.slots + h2);
  const uint4 a = s1[0], b = s2[0];
  const uint4 ai = s1[1], bi = s2[1];
  bool m1 = (((a.x ^ kw[0]) | (a.y ^ kw[1]) | (a.z ^ kw[2]) | (a.w ^ kw[3])) == 0) && (ai.w & kSlotValid);
  bool m2 = (((b.x ^ kw[0]) | (b.y ^ kw[1]) | (b.z ^ kw[2]) | (b.w ^ kw[3])) == 0) && (bi.w & kSlotValid);
  const uint4* sm = m1 ? s1 : s2;
  if (wide) {                    // (warp-uniform) the upper halves of the keys
    const uint4 c = __ldcg(s1 + 2), d = __ldcg(s2 + 2);      // (second sector: from L2 -- see below)
    m1 = m1 && (((c.x ^ kw[4]) | (c.y ^ kw[5]) | (c.z ^ kw[6]) | (c.w ^ kw[7])) == 0);
    m2 = m2 && (((d.x ^ kw[4]) | (d.y ^ kw[5]) | (d.z ^ kw[6]) | (d.w ^ kw[7])) == 0);
    sm = m1 ? s1 : s2;
  }
  r.meta = m1 ? ai.w : (m2 ? bi.w : 0u);
  r.id[0] = (int32_t)(m1 ? ai.x : bi.x); r.id[1] = (int32_t)(m1 ? ai.y : bi.y); r.id[2] = (int32_t)(m1 ? ai.z : bi.z);
  r.offs_hi = 0; r.id[3] = r.id[4] = r.id[5] = 0;
  if ((r.meta & 7u) > 3u) {      // rare: more than three pieces
    // The first sector (key, ids, meta) is one snapshot: a valid meta in it means the writer's earlier stores are in it too.
    // The second sector is read on its own, so it must not come from an L1 copy older than that snapshot: read it from L2.
    const uint4 e = __ldcg(sm + 3);
    r.id[3] = (int32_t)e.x; r.id[4] = (int32_t)e.y; r.id[5] = (int32_t)e.z; r.offs_hi = e.w;
  }
#else
  r.meta = 0; r.offs_hi = 0;
  for (int i = 0; i < 6; ++i) r.id[i] = 0;
  const uint32_t hs[2] = {h1, h2};
  for (int k = 0; k < 2 && !r.meta; ++k) {
    const WpWordSlot& a = W.slots[hs[k]];
    const uint32_t meta = __atomic_load_n(&a.meta, __ATOMIC_ACQUIRE);
    if ((meta & kSlotValid) && wp_slot_matches(a, kw, wide)) {
      r.meta = meta; r.offs_hi = a.offs_hi;
      for (int i = 0; i < 3; ++i) { r.id[i] = a.id[i]; r.id[3 + i] = a.id_hi[i]; }
    }
  }
#endif
  return r;
}

// Adds a word the lexer loops have just resolved (n pieces, n <= kMaxLearnPieces; n = 0: one UnkId) to an empty
// candidate slot.  Another warp may be adding the same word: whoever claims the slot first writes it, the other
// one finds it busy and leaves.  Nothing is ever overwritten.
BF_HD void wp_words_insert(const WpWords& W, const uint32_t kw[8], bool wide, int n, const int32_t* ids, const int* offs) {
  const uint32_t h = wp_key_hash(kw, W.mul, wide);
  const uint32_t hs[2] = {wp_slot1(W, h), wp_slot2(W, h)};
  const uint32_t meta = kSlotValid | (uint32_t)n | ((uint32_t)(n > 1 ? offs[1] : 0) << 4) | ((uint32_t)(n > 2 ? offs[2] : 0) << 9);
  const uint32_t offs_hi = (uint32_t)(n > 3 ? offs[3] : 0) | ((uint32_t)(n > 4 ? offs[4] : 0) << 8) | ((uint32_t)(n > 5 ? offs[5] : 0) << 16);
  for (int k = 0; k < 2; ++k) {
    WpWordSlot* s = W.slots + hs[k];
#if defined(__CUDA_ARCH__)
    const uint32_t old = atomicCAS(&s->meta, 0u, kSlotBusy);
#else
    uint32_t old = 0;
    __atomic_compare_exchange_n(&s->meta, &old, kSlotBusy, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
#endif
    if (old == 0u) {
      for (int i = 0; i < 4; ++i) { s->kw[i] = kw[i]; s->kw_hi[i] = wide ? kw[4 + i] : 0u; }
      for (int i = 0; i < 3; ++i) { s->id[i] = n > i ? ids[i] : 0; s->id_hi[i] = n > 3 + i ? ids[3 + i] : 0; }
      s->offs_hi = offs_hi;
#if defined(__CUDA_ARCH__)
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(&s->meta) = meta;
#else
      __atomic_store_n(&s->meta, meta, __ATOMIC_RELEASE);
#endif
      return;
    }
    // taken: by this very word (then it is in the table) or by another one (try the other slot)
    if ((old & kSlotValid) && wp_slot_matches(*s, kw, wide)) return;
  }
}

// Writes a table hit for the word starting at window position s.
BF_HD void wp_apply_hit(const WpWordHit& h, int s, int unk_id, int32_t* ids_at) {
  const uint32_t n = h.meta & 7u;
  ids_at[s] = n == 0 ? unk_id : h.id[0];
  if (n > 1) ids_at[s + (int)((h.meta >> 4) & 31u)] = h.id[1];
  if (n > 2) ids_at[s + (int)((h.meta >> 9) & 31u)] = h.id[2];
  if (n > 3) {
    ids_at[s + (int)(h.offs_hi & 255u)] = h.id[3];
    if (n > 4) ids_at[s + (int)((h.offs_hi >> 8) & 255u)] = h.id[4];
    if (n > 5) ids_at[s + (int)((h.offs_hi >> 16) & 255u)] = h.id[5];
  }
}

// May the run [s, e) -- every position in one group of top-level classes, kind `kind` -- be served by the memo?  Returns
// 0 = no (run the loops), 1 = it emits nothing, 2 = look the word up.
BF_HD int wp_classify_run(uint32_t kind, int len, uint32_t max_len, bool at_doc_start, bool at_doc_end) {
  if (kind & kKindInert) return 1;       // established over the closure with both anchors
  if (at_doc_start && !(kind & kKindCaretOk)) return 0;
  if (at_doc_end && !(kind & kKindDollarOk)) return 0;
  if (len <= (int)max_len && ((kind & kKindWordRun) || (len == 1 && (kind & kKindWordOne)))) return 2;
  return 0;
}

}  // namespace bfb200
