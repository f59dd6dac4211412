// lexer_tables.h -- flattened, HBM-ready form of a [wbd] lexer model.
//
// Built once at LoadModel time from the packed image (ldb.h).  It replaces the
// per-transition packed-record decoding of FARSDfa_pack_triv::GetDest /
// FAIwMap_pack::GetNewIw / FAState2Ow_pack_triv::GetOw / FAMultiMap_pack::Get
//  with direct table lookups:
//
//   cls_of_cp[cp]            code point -> class, with the charmap (FANormalize, 1->1 maps)
//                            and the "symbol < 3 -> 3" clamp (FALexTools_t.h:259-261) folded in
//   trans[state][class]      dense state x class transition table, IW_ANY fallback
//                            (FALexTools_t.h:266-270) folded in; column NC = "unmapped"
//   states are renumbered so that final states are exactly ids >= first_final
//   ow / action tables       rule id -> {LeftCx, RightCx, Tag, fns...}
//
// plus, when the model has the flat two-level WordPiece shape (see FastPath), the tiny
// top-level automaton in its own numbering for shared-memory residency.
#pragma once

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "ldb.h"

namespace bfb200 {

// A vector whose resize() leaves new elements uninitialised: the dense transition table (9.3 GB for
// bert_multi_cased) is first touched by several threads at once -- one thread faulting the pages in
// takes most of a minute.
template <class T>
struct NoInitAllocator {
  using value_type = T;
  NoInitAllocator() = default;
  template <class U> NoInitAllocator(const NoInitAllocator<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(::operator new(n * sizeof(T))); }
  void deallocate(T* p, size_t) { ::operator delete(p); }
  template <class U> void construct(U*) {}
  template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
  template <class U> bool operator==(const NoInitAllocator<U>&) const { return true; }
  template <class U> bool operator!=(const NoInitAllocator<U>&) const { return false; }
};
template <class T> using DenseVec = std::vector<T, NoInitAllocator<T>>;


constexpr int kMaxCodePoint = 0x10FFFF;
constexpr uint32_t kNoState = 0xFFFFFFFFu;     // "no transition" in host-side tables

// The flat two-level WordPiece shape (bert_*): the top-level grammar only emits WORD/IGNORE
// style tags (<= 4) with zero contexts and at most one function call, and the called
// function's actions are plain [0,0,id] with id > 4.  Under these conditions
// TextToIdsWithOffsets_wp's post-pass (blingfiretokdll.cpp:1207-1313) reduces to
// "a word's pieces if they tile it exactly, else UnkId", and the top-level scan can be cut
// at positions no match can cross ("sync points").
struct FastPath {
  bool ok = false;
  std::string why_not;                 // first violated condition, for diagnostics
  int K = 0;                           // top-level states (local ids 0..K-1, 0 = initial)
  int NT = 0;                          // top-level class-equivalence classes (incl. "none")
  std::vector<uint8_t> tc_of_class;    // [NC+1] class -> top-level class
  std::vector<uint8_t> ttop;           // [K*NT] local transition, 0xFF = none
  std::vector<uint64_t> cross;         // [NT] bit t2 set: some top-level walk can consume t1 then t2
  std::vector<uint8_t> top_final;      // [K]
  std::vector<int32_t> top_tag;        // [K] action tag for finals
  std::vector<uint32_t> top_fn_root;   // [K] global state id of FnIni[fn], kNoState if no call
  std::vector<uint32_t> top_fn_caret;  // [K] delta'(FnIni[fn], class(^)) or kNoState
  uint8_t tc_caret = 0, tc_dollar = 0; // top-level classes of the anchors
  uint8_t tc_none = 0;                 // top-level class of "unmapped"
  int max_piece_walk = 0;              // longest path in the function sub-automata (diagnostic)
};

struct LexerTables {
  // configuration ([wbd] section, FAWbdConfKeeper.cpp:56-232)
  int max_depth = 2;
  int max_token_length = 300;
  bool has_charmap = false;
  bool charmap_one_to_one = true;      // every charmap row maps to exactly one code point

  // classes
  int NC = 0;                          // number of real classes; class NC = "unmapped"
  uint32_t cls_caret = 0, cls_dollar = 0;   // classes of IW_L_ANCHOR / IW_R_ANCHOR (NC if unmapped)
  std::vector<uint16_t> cls_of_cp;     // [0x110000] combined charmap+clamp+class (1->1 charmaps only)
  std::vector<uint32_t> clsx_of_cp;    // [0x110000] cls_of_cp | top-level class << 16 (FastPath models only): one gather
                                       //            per code point in the fused kernel
  std::vector<uint16_t> cls_of_iw;     // [max_iw+1] plain class map (general path)
  // TextToWords' view of a code point (blingfiretokdll.cpp:475-482): NO charmap, U+0000 -> U+0020,
  // then the lexer's clamp and the class map
  std::vector<uint16_t> cls_words_of_cp;   // [0x110000]

  // general charmap (1->N); only filled when !charmap_one_to_one
  std::vector<uint8_t> norm_count;     // [0x110000] 0..10, 0xFF = unmapped (keep code point)
  std::vector<uint32_t> norm_first;    // [0x110000] index into norm_values
  std::vector<int32_t> norm_values;

  // automaton, renumbered: non-final states first
  int NS = 0;                          // number of states incl. the explicit dead sink
  uint32_t first_final = 0;
  uint32_t initial = 0;
  uint32_t dead = 0;                   // explicit sink standing for DFA_DEAD_STATE
  bool wide_states = false;            // NS >= 65535 -> 32-bit table entries
  DenseVec<uint16_t> trans16;          // [NS*(NC+1)] when !wide_states, 0xFFFF = none
  DenseVec<uint32_t> trans32;          // [NS*(NC+1)] when wide_states, kNoState = none
  std::vector<int32_t> ow_of_state;    // [NS] rule id for finals, -1 otherwise
  std::vector<int32_t> orig_offset;    // [NS] the state's id in the packed image (byte offset), -1 for the sink
  // the stored arcs in the new numbering (CSR over states): the load-time enumeration of the vocabulary
  // (wp_model.cpp, whole-word table) walks these instead of scanning dense rows
  std::vector<int64_t> arc_begin;      // [NS+1]
  std::vector<uint32_t> arc_label;     // class
  std::vector<uint32_t> arc_dst;

  // actions: rule id -> ints (FAMultiMap_pack rows), validated like FALexTools_t::Validate
  std::vector<int32_t> act_begin;      // [num_acts+1]
  std::vector<int32_t> act_data;
  std::vector<int32_t> tag_of_state;   // [NS] act[2] of the state's rule (0 when not final)
  std::vector<uint32_t> fn_ini;        // [fn count] global state id or kNoState

  FastPath fast;

  std::vector<uint32_t> any_dst;       // [NS] destination of the state's IW_ANY arc (the fallback of every other class), or kNoState
  bool dense_on_host = true;           // trans16 / trans32 are filled (false: only the stored arcs; the device builds the table)

  // the transition, from the dense table when it is here, else from the stored arcs (binary search + IW_ANY fallback)
  uint32_t next(uint32_t s, uint32_t c) const {
    if (dense_on_host) {
      const size_t i = (size_t)s * (NC + 1) + c;
      if (wide_states) return trans32[i];
      const uint16_t v = trans16[i];
      return v == 0xFFFF ? kNoState : v;
    }
    int64_t lo = arc_begin[s], hi = arc_begin[(size_t)s + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (arc_label[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (lo < arc_begin[(size_t)s + 1] && arc_label[lo] == c) return arc_dst[lo];
    return any_dst[s];
  }
  bool is_final(uint32_t s) const { return s != kNoState && s >= first_final; }
};

// Returns false with *err set if the model cannot be served (malformed, or uses a feature
// the reference itself would assert on).
// dense_wide = false: a table with 32-bit entries (9.3 GB for bert_multi_cased) is not built on the host; the caller fills
// the device copy from the stored arcs (capi.cu) and every host-side question goes through next().  Tables with 16-bit
// entries (129 MB for bert_base_tok) are always built.
bool build_lexer_tables(const LdbImage& ldb, LexerTables* out, std::string* err, bool dense_wide = true);

}  // namespace bfb200
