// wp_twin.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A sequential host driver around the product's own table builders (ldb.cpp,
// lexer_tables.cpp, wp_model.cpp) and the product's own __host__ __device__ chunk routine
// (wp_core.cuh).  It mirrors, step for step, what one warp of wp_kernel.cu does for one
// document (decode -> classes -> sync points -> chunks -> windows with carry -> compaction),
// so the algorithmic part of the CUDA path can be checked against the oracle on the CPU box,
// where no GPU exists.  It is compiled into tests/twin/libwp_twin.so, is never linked into
// the product library, and the product API has no route to it.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../blingfire_b200/csrc/ldb.h"
#include "../../blingfire_b200/csrc/lexer_tables.h"
#include "../../blingfire_b200/csrc/lex_core.cuh"
#include "../../blingfire_b200/csrc/wp_core.cuh"
#include "../../blingfire_b200/csrc/wp_model.h"

using namespace bfb200;

struct Twin {
  LdbImage ldb;
  LexerTables T;
  WpBlob blob;
  std::string err;
};

extern "C" {

void* twin_load(const char* path) {
  Twin* t = new Twin();
  if (!t->ldb.load_file(path)) { t->err = t->ldb.error(); return t; }
  if (!t->ldb.conf().get(kFuncWbd)) { t->err = "no [wbd]"; return t; }
  if (!build_lexer_tables(t->ldb, &t->T, &t->err)) return t;
  if (t->T.fast.ok && t->T.charmap_one_to_one) build_wp_blob(t->T, &t->blob);
  return t;
}
// the same model with the dense table left out on the host when it is wide (what LoadModel does): only the load-time
// products can be asked for (twin_info, twin_blob_digest)
void* twin_load_sparse(const char* path) {
  Twin* t = new Twin();
  if (!t->ldb.load_file(path)) { t->err = t->ldb.error(); return t; }
  if (!build_lexer_tables(t->ldb, &t->T, &t->err, /*dense_wide=*/false)) return t;
  if (t->T.fast.ok && t->T.charmap_one_to_one) build_wp_blob(t->T, &t->blob);
  return t;
}
// FNV-1a-64 over the staged blob, the word table and its parameters
uint64_t twin_blob_digest(void* h) {
  Twin* t = (Twin*)h;
  uint64_t x = 0xcbf29ce484222325ull;
  auto eat = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; ++i) { x ^= b[i]; x *= 0x100000001b3ull; } };
  eat(t->blob.bytes.data(), t->blob.bytes.size());
  eat(t->blob.word_slots.data(), t->blob.word_slots.size() * sizeof(WpWordSlot));
  eat(&t->blob.words.log2_size, 4); eat(&t->blob.words.cb, 4); eat(&t->blob.words.cpw, 4); eat(&t->blob.words.max_len, 4); eat(t->blob.words.mul, 36);
  return x;
}
int twin_dense_on_host(void* h) { return ((Twin*)h)->T.dense_on_host ? 1 : 0; }
void twin_free(void* h) { delete (Twin*)h; }
const char* twin_error(void* h) { return ((Twin*)h)->err.c_str(); }
int twin_fast_ok(void* h) { Twin* t = (Twin*)h; return t->err.empty() && t->T.fast.ok && t->T.charmap_one_to_one; }
const char* twin_fast_why(void* h) { return ((Twin*)h)->T.fast.why_not.c_str(); }

// table introspection for the exhaustive cross-check against the oracle's packed readers
int twin_info(void* h, int what) {
  Twin* t = (Twin*)h;
  switch (what) {
    case 0: return t->T.NS;
    case 1: return t->T.NC;
    case 2: return (int)t->T.first_final;
    case 3: return (int)t->T.initial;
    case 4: return (int)t->T.dead;
    case 5: return t->T.wide_states ? 1 : 0;
    case 6: return (int)t->T.cls_of_iw.size();
    case 7: return (int)t->T.fn_ini.size();
    case 8: return t->T.fast.K;
    case 9: return t->T.fast.NT;
    case 10: return t->T.max_depth;
    case 11: return t->T.max_token_length;
    case 12: return (int)t->blob.layout.total_bytes;
    case 13: return (int)t->blob.word_count;
    case 14: return (int)t->blob.words.log2_size;
    case 15: return (int)t->blob.words.max_len;
    default: return -1;
  }
}
const void* twin_ptr(void* h, int what) {
  Twin* t = (Twin*)h;
  switch (what) {
    case 0: return t->T.wide_states ? (const void*)t->T.trans32.data() : (const void*)t->T.trans16.data();
    case 1: return t->T.orig_offset.data();
    case 2: return t->T.cls_of_iw.data();
    case 3: return t->T.cls_of_cp.data();
    case 4: return t->T.ow_of_state.data();
    case 5: return t->T.tag_of_state.data();
    case 6: return t->T.fn_ini.data();
    default: return nullptr;
  }
}

// The generic lexer (lex_core.cuh) over UTF-32 input, whole document as one span.
// mode 0: TextToIds view of the symbols (charmap + clamp + class); mode 1: TextToWords view.
int twin_lex_process(void* h, const int* cps, int n, int32_t* out, int max_out, int mode) {
  Twin* t = (Twin*)h;
  if (!t->err.empty()) return -2;
  const LexerTables& T = t->T;
  if (mode == 0 && !T.charmap_one_to_one) return -2;
  std::vector<uint16_t> cls((size_t)n + 1);
  for (int i = 0; i < n; ++i) {
    const int cp = cps[i];
    const bool in_range = cp >= 0 && cp <= kMaxCodePoint;
    cls[i] = !in_range ? (uint16_t)T.NC : (mode == 0 ? T.cls_of_cp[cp] : T.cls_words_of_cp[cp]);
  }
  auto run = [&](auto& g, const auto* trans) {
    g.trans = trans; g.ow_of_state = T.ow_of_state.data(); g.act_begin = T.act_begin.data(); g.act_data = T.act_data.data();
    g.fn_ini = T.fn_ini.data(); g.fn_count = (int)T.fn_ini.size(); g.NC1 = (uint32_t)T.NC + 1; g.first_final = T.first_final;
    g.cls_caret = T.cls_caret; g.cls_dollar = T.cls_dollar; g.initial = T.initial; g.max_depth = T.max_depth;
    g.max_token_length = T.max_token_length;
    return lex_process(g, cls.data(), n, out, max_out);
  };
  if (T.wide_states) { LexGlobal<uint32_t> g{}; return run(g, T.trans32.data()); }
  LexGlobal<uint16_t> g{};
  return run(g, T.trans16.data());
}

int twin_wp_postpass(const int32_t* res, int rn, int32_t* ids, int max_ids, int unk) { return wp_postpass(res, rn, ids, max_ids, unk); }

// One document through the mirrored warp algorithm.  `window` plays the role of the
// kernel's per-warp window capacity (in code points); small values stress the carry logic.
int twin_text_to_ids_ex(void* h, const char* s, int n, int32_t* ids, int max_ids, int unk, int window, int use_memo, int64_t* stats);
int twin_text_to_ids(void* h, const char* s, int n, int32_t* ids, int max_ids, int unk, int window) {
  return twin_text_to_ids_ex(h, s, n, ids, max_ids, unk, window, 1, nullptr);
}
// use_memo = 0: every chunk through the lexer loops (the memo must not change a single id);
// stats[0..3] += single-piece words found in the table, words the table did not hold, chunks that emit nothing,
// words found with another number of pieces (learned at run time)
int twin_text_to_ids_ex(void* h, const char* s, int n, int32_t* ids, int max_ids, int unk, int window, int use_memo, int64_t* stats) {
  Twin* t = (Twin*)h;
  if (!twin_fast_ok(h)) return -2;
  if (n <= 0 || n > 1000000000 || !s) return 0;
  const LexerTables& T = t->T;
  const WpTop top = make_wp_top(t->blob.bytes.data(), t->blob.layout);
  const int max_tok = T.max_token_length;
  if (window < max_tok + 40) window = max_tok + 40;

  const uint8_t* b = (const uint8_t*)s;
  int lo = 0, hi = n;
  if (n >= 3 && b[0] == 0xEF && b[1] == 0xBB && b[2] == 0xBF) lo = 3;   // FAUtf8Utils.cpp:247-252
  if (lo >= hi) return 0;

  // strict UTF-8 validation of the whole document first (FAUtf8Utils.cpp:121-196)
  {
    int p = lo;
    while (p < hi) {
      const int c = b[p];
      int len, cp;
      if (c < 0x80) { p += 1; continue; }
      if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
      else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
      else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
      else return 0;
      if (p + len > hi) return 0;
      for (int k = 1; k < len; ++k) { if ((b[p + k] & 0xC0) != 0x80) return 0; cp = (cp << 6) | (b[p + k] & 0x3F); }
      const int need = cp <= 0x7F ? 1 : cp <= 0x7FF ? 2 : cp <= 0xFFFF ? 3 : cp <= 0x10FFFF ? 4 : 0;
      if (need != len) return 0;
      if ((cp & 0xFFFFF800) == 0xD800) return 0;
      p += len;
    }
  }

  std::vector<uint16_t> cls((size_t)window + 32);
  std::vector<int32_t> ids_at((size_t)window + 8);
  std::vector<uint8_t> tcs((size_t)window + 8);
  std::vector<unsigned> events;
  WpGlobal<uint16_t> g16{};
  WpGlobal<uint32_t> g32{};
  auto fill_g = [&](auto& g, const auto* trans) {
    g.trans = trans; g.tag_of_state = T.tag_of_state.data();
    g.NC1 = (uint32_t)T.NC + 1; g.first_final = T.first_final; g.cls_caret = T.cls_caret; g.cls_dollar = T.cls_dollar;
    g.max_token_length = max_tok;
  };
  if (T.wide_states) fill_g(g32, T.trans32.data()); else fill_g(g16, T.trans16.data());
  WpWords words = t->blob.words;
  words.slots = t->blob.word_slots.data();
  const bool learn = use_memo == 1;      // use_memo 2: the table as built at load time, nothing added
  if (!use_memo) words.max_len = 0;

  int m = 0, bpos = lo, out = 0;
  bool first = true;
  for (;;) {
    // fill the window with classes of further code points
    while (m < window && bpos < hi) {
      const int c = b[bpos];
      int len = c < 0x80 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : 4;
      int cp = c < 0x80 ? c : len == 2 ? c & 0x1F : len == 3 ? c & 0x0F : c & 0x07;
      for (int k = 1; k < len; ++k) cp = (cp << 6) | (b[bpos + k] & 0x3F);
      const uint32_t x = cp < 128 ? top.ascii_clsx[cp] : T.clsx_of_cp[cp];
      cls[m] = (uint16_t)x; tcs[m] = (uint8_t)(x >> 16); ids_at[m] = kNoPiece;
      ++m;
      bpos += len;
    }
    const bool at_end = bpos >= hi;
    if (m == 0) break;
    // events: chunk starts (sync points whose class can start a match; position 0 always) and changes of group
    events.clear();
    for (int p = 0; p < m; ++p) {
      const unsigned tc = tcs[p], tp = p > 0 ? tcs[p - 1] : tc;
      const unsigned sv = p == 0 ? (unsigned)kSyncStart : (unsigned)top.sync_start[(tp << top.sync_shift) | tc];
      if (sv) events.push_back((unsigned)p | ((sv & kSyncStart) ? 0x8000u : 0u));
    }
    events.push_back((unsigned)m | 0x8000u);
    events.push_back((unsigned)m | 0x8000u);
    const int limit = at_end ? m : m - max_tok;
    int carry = at_end ? m : 0;
    auto loops = [&](int s, int e) -> int {
      int fe = e > limit ? limit : e;
      const int fb = (s == 0 && first) ? -1 : s;
      if (fb >= fe) return 0;
      int r;
      if (T.wide_states) r = wp_chunk<uint32_t>(top, g32, cls.data(), m, at_end, fb, fe, unk, ids_at.data(), tcs.data());
      else r = wp_chunk<uint16_t>(top, g16, cls.data(), m, at_end, fb, fe, unk, ids_at.data(), tcs.data());
      if (r > carry) carry = r;
      return 0;
    };
    for (size_t i = 0; i + 2 < events.size(); ++i) {
      if (!(events[i] & 0x8000u)) continue;
      const int s = (int)(events[i] & 0x7FFFu);
      int we = (int)(events[i + 1] & 0x7FFFu), e = we;
      bool shaped = true;
      if (!(events[i + 1] & 0x8000u)) {
        e = (int)(events[i + 2] & 0x7FFFu);
        shaped = (events[i + 2] & 0x8000u) && (top.kind_of_tc[tcs[we]] & kKindDead);
        if (!shaped) {
          int p = we + 1;
          while (p < m && !(top.sync_start[((unsigned)tcs[p - 1] << top.sync_shift) | (unsigned)tcs[p]] & kSyncStart)) ++p;
          e = p;
        }
      }
      if (!shaped || (!at_end && e > limit)) {
        if (stats) ++stats[!shaped ? 4 : 5];
        const int c0 = carry;
        if (loops(s, e)) return -3;
        // exactness guard of the sync-point argument: a chunk must end exactly on the next start
        if (e < m && e <= limit && carry > c0 && carry != e && carry > e) return -3;
        continue;
      }
      const int r = wp_classify_run(top.kind_of_tc[tcs[s]], we - s, words.max_len, first && s == 0, at_end && we == m);
      if (r == 0) { if (stats) ++stats[(we - s) > (int)words.max_len && (top.kind_of_tc[tcs[s]] & kKindWordRun) ? 6 : 7]; if (loops(s, e)) return -3; continue; }
      if (e > carry) carry = e;
      if (r == 2) {
        uint32_t kw[8];
        const bool wide = we - s > (int)(4 * words.cpw);
        wp_pack_key_any(words.cpw, cls.data() + s, we - s, we - s, words.cb, kw);
        const WpWordHit hit = wp_words_find(words, kw, wide);
        if (hit.meta) { wp_apply_hit(hit, s, unk, ids_at.data()); if (stats) ++stats[(hit.meta & 7u) == 1 ? 0 : 3]; }
        else {
          // the kernel's slow_round for a word run: the function sub-grammar alone, then into the table
          if (stats) ++stats[1];
          const unsigned tc = tcs[s];
          int tiled;
          if (T.wide_states) tiled = wp_word<uint32_t>(g32, cls.data(), s, we - 1, top.fn_root_of_tc[tc], top.fn_caret_of_tc[tc], ids_at.data());
          else tiled = wp_word<uint16_t>(g16, cls.data(), s, we - 1, top.fn_root_of_tc[tc], top.fn_caret_of_tc[tc], ids_at.data());
          int n = 0, offs[kMaxLearnPieces];
          int32_t pids[kMaxLearnPieces];
          if (!tiled) { for (int p = s + 1; p < we; ++p) ids_at[p] = kNoPiece; ids_at[s] = unk; }
          else for (int p = s; p < we; ++p) if (ids_at[p] != kNoPiece) { if (n < kMaxLearnPieces) { pids[n] = ids_at[p]; offs[n] = p - s; } ++n; }
          if (learn && n <= kMaxLearnPieces) wp_words_insert(words, kw, wide, n, pids, offs);
        }
      } else if (stats) ++stats[2];
    }
    for (int p = 0; p < carry && p < m; ++p)
      if (ids_at[p] != kNoPiece) { if (out < max_ids) ids[out] = ids_at[p]; ++out; }
    if (at_end) break;
    std::memmove(cls.data(), cls.data() + carry, sizeof(uint16_t) * (size_t)(m - carry));
    std::memmove(tcs.data(), tcs.data() + carry, (size_t)(m - carry));
    m -= carry;
    for (int p = 0; p < m; ++p) ids_at[p] = kNoPiece;
    first = false;
  }
  return out < max_ids ? out : max_ids;
}

}  // extern "C"
