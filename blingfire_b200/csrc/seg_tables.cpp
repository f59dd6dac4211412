// seg_tables.cpp -- see seg_tables.h.  Host-only, runs once per LoadModel.
#include "seg_tables.h"

#include <algorithm>
#include <climits>
#include <cstring>
#include <map>

namespace bfb200 {

void flatten_charmap(const FixedMap& charmap, std::vector<uint8_t>* count, std::vector<uint32_t>* first,
                     std::vector<int32_t>* values) {
  count->assign((size_t)0x110000, 0xFF);
  first->assign((size_t)0x110000, 0);
  values->clear();
  int tmp[16];
  for (int cp = charmap.min_key; cp <= charmap.max_key && cp <= 0x10FFFF; ++cp) {
    const int c = charmap.get(cp, tmp, 10);
    if (c == -1) continue;
    // FANormalize (FAUtils_cl.h:311-369): counts outside 1..10 fire no branch: the symbol vanishes
    const int eff = (c >= 1 && c <= 10) ? c : 0;
    (*count)[cp] = (uint8_t)eff;
    (*first)[cp] = (uint32_t)values->size();
    for (int i = 0; i < eff; ++i) values->push_back(tmp[i]);
  }
}

bool build_seg_tables(const LdbImage& ldb, SegTables* T, std::string* err) {
  const std::vector<int>* sec = ldb.conf().get(kFuncPosDict);
  if (!sec) { *err = "no [pos-dict] section"; return false; }
  // FADictConfKeeper::Init (FADictConfKeeper.cpp:57-228)
  int fsm_type = kTypeMealyDfa, map_mode = kModePackTriv, fsm_dump = -1, i2info_dump = -1, charmap_dump = -1;
  const std::vector<int>& v = *sec;
  for (size_t i = 0; i < v.size(); ++i) {
    auto arg = [&](int* dst) { if (i + 1 >= v.size()) return false; *dst = v[++i]; return true; };
    int tmp = 0;
    bool ok = true;
    switch (v[i]) {
      case kParamNoTr: break;
      case kParamIgnoreCase: *err = "[pos-dict] ignore-case is not served"; return false;
      case kParamUseByteEncoding: T->use_raw_bytes = true; break;
      case kParamNoDummyPrefix: T->no_dummy_prefix = true; break;
      case kParamDirection: ok = arg(&tmp); break;
      case kParamTokenizationType: ok = arg(&T->tok_algo); break;
      case kParamIdOffset: ok = arg(&T->id_offset); break;
      case kParamFsmType: ok = arg(&fsm_type); break;
      case kParamMapMode: ok = arg(&map_mode); break;
      case kParamFsm: ok = arg(&fsm_dump); break;
      case kParamArray: ok = arg(&tmp); break;            // K2I: identity, never read on this path
      case kParamCharmap: ok = arg(&charmap_dump); break;
      case kParamMultiMap: ok = arg(&i2info_dump); break;
      default: *err = "[pos-dict] unknown parameter"; return false;
    }
    if (!ok) { *err = "[pos-dict] truncated parameter"; return false; }
  }
  if (fsm_type != kTypeMealyDfa || fsm_dump < 0 || i2info_dump < 0) { *err = "[pos-dict] needs a Mealy fsm and a multi-map"; return false; }
  if (T->tok_algo < 0 || T->tok_algo > 5) { *err = "[pos-dict] tokenization type"; return false; }

  Automaton A;
  if (!LdbImage::parse_automaton(ldb.dump(fsm_dump), /*mealy=*/true, &A, err)) return false;

  // ---- I2Info ----
  const bool unigram = !(T->tok_algo == kTokenizeBpe || T->tok_algo == kTokenizeBpeOpt || T->tok_algo == kTokenizeBpeOptWithMerges);
  const int need = (unigram || T->tok_algo == kTokenizeBpeOptWithMerges) ? 2 : 1;
  if (map_mode == kModePackFixed) {
    FixedMap fm;
    if (!LdbImage::parse_fixedmap(ldb.dump(i2info_dump), &fm, err)) return false;
    if (fm.size_of_value != 4) { *err = "I2Info is not int-valued"; return false; }   // Get(key,&ptr) needs ints
    if (fm.max_key > (1 << 26)) { *err = "I2Info too large"; return false; }
    T->info.assign((size_t)fm.max_key + 1, SegInfo{INT_MIN, 0.0f});
    int row[16];
    for (int k = fm.min_key; k <= fm.max_key; ++k) {
      const int c = fm.get(k, row, 16);
      if (c < need || c > 16) continue;
      SegInfo si; si.id = row[0]; si.score = 0.0f;
      if (c >= 2) std::memcpy(&si.score, &row[1], 4);
      T->info[k] = si;
    }
  } else if (map_mode == kModePackTriv) {
    MultiMap mm;
    if (!LdbImage::parse_multimap(ldb.dump(i2info_dump), &mm, err)) return false;
    if (!mm.ptr_interface_ok) { *err = "I2Info is not int-valued"; return false; }
    T->info.assign(mm.rows.size(), SegInfo{INT_MIN, 0.0f});
    for (size_t k = 0; k < mm.rows.size(); ++k) {
      if (!mm.present[k] || (int)mm.rows[k].size() < need) continue;
      SegInfo si; si.id = mm.rows[k][0]; si.score = 0.0f;
      if (mm.rows[k].size() >= 2) std::memcpy(&si.score, &mm.rows[k][1], 4);
      T->info[k] = si;
    }
  } else { *err = "I2Info container mode is not served"; return false; }

  // ---- alphabet ----
  std::vector<int> labels;
  for (const Arc& a : A.arcs) labels.push_back(a.label);
  std::sort(labels.begin(), labels.end());
  labels.erase(std::unique(labels.begin(), labels.end()), labels.end());
  if (labels.empty() || labels.size() >= 0xFFFF) { *err = "unsupported alphabet size"; return false; }
  T->alphabet = (int)labels.size();
  T->sym_of_cp.assign((size_t)0x110000, kNoSym);
  std::map<int, int> sidx;
  for (size_t i = 0; i < labels.size(); ++i) {
    sidx[labels[i]] = (int)i;
    if (labels[i] >= 0 && labels[i] <= 0x10FFFF) T->sym_of_cp[labels[i]] = (uint16_t)i;
  }

  // ---- depth / delimiter analysis ----
  const int n = A.num_states();
  std::vector<int> depth((size_t)n, -1);
  {
    std::vector<int> q{0};
    depth[0] = 0;
    for (size_t h = 0; h < q.size(); ++h) {
      const int s = q[h];
      for (int64_t k = A.arc_begin[s]; k < A.arc_begin[s + 1]; ++k) {
        const Arc& a = A.arcs[k];
        if (a.label == kSpDelim && s != 0) T->delim_inside_tokens = true;
        if (a.dst >= 0 && depth[a.dst] < 0) { depth[a.dst] = depth[s] + 1; q.push_back(a.dst); }
      }
    }
    // longest path from the root (the automaton of a finite vocabulary is acyclic); an iterative
    // DFS with colours, "unbounded" if a cycle shows up
    std::vector<int> longest((size_t)n, 0);
    std::vector<uint8_t> colour((size_t)n, 0);
    std::vector<std::pair<int, int64_t>> stk;
    bool cyclic = false;
    stk.push_back({0, A.arc_begin[0]});
    colour[0] = 1;
    while (!stk.empty() && !cyclic) {
      auto& top = stk.back();
      const int s = top.first;
      if (top.second < A.arc_begin[s + 1]) {
        const int d = A.arcs[top.second++].dst;
        if (d < 0) continue;
        if (colour[d] == 1) { cyclic = true; break; }
        if (colour[d] == 0) { colour[d] = 1; stk.push_back({d, A.arc_begin[d]}); }
        else longest[s] = std::max(longest[s], longest[d] + 1);
      } else {
        colour[s] = 2;
        stk.pop_back();
        if (!stk.empty()) { const int p = stk.back().first; longest[p] = std::max(longest[p], longest[s] + 1); }
      }
    }
    T->max_arc_len = cyclic ? (1 << 30) : longest[0];
    // the BPE engine splits documents at U+2581; that needs U+2581 to be a token start of its own
    if (A.dest(0, kSpDelim) < 0) T->delim_inside_tokens = true;
  }

  // ---- double-array placement (first fit; every base is unique and >= alphabet) ----
  // First fit in increasing base order.  Only bases whose slot for the state's smallest symbol is free
  // can fit, so the candidates are walked along a free-slot list (next-pointer with path compression)
  // instead of one by one: the table ends up > 90 % full and the plain scan took over a minute for the
  // 250k-token xlm-r vocabulary.
  const uint32_t Aw = (uint32_t)T->alphabet;
  std::vector<uint32_t> arc_sym(A.arcs.size());
  for (size_t k = 0; k < A.arcs.size(); ++k) arc_sym[k] = (uint32_t)sidx[A.arcs[k].label];
  std::vector<uint32_t> base((size_t)n, 0);
  std::vector<uint8_t> used_slot, used_base;
  std::vector<uint32_t> next_free;                      // next_free[i] == i: slot i is unused; else look further right
  auto ensure = [&](size_t need_size) {
    if (used_slot.size() >= need_size) return;
    const size_t old = used_slot.size(), grown = need_size * 2;
    used_slot.resize(grown, 0); used_base.resize(grown, 0); next_free.resize(grown);
    for (size_t i = old; i < grown; ++i) next_free[i] = (uint32_t)i;
  };
  auto find_free = [&](uint32_t i) -> uint32_t {
    ensure((size_t)i + 2);
    uint32_t r = i;
    while (next_free[r] != r) { r = next_free[r]; ensure((size_t)r + 2); }
    while (next_free[i] != r && i != r) { const uint32_t nx = next_free[i]; next_free[i] = r; i = nx; }
    return r;
  };
  ensure(A.arcs.size() * 2 + (size_t)Aw * 4 + 1024);
  uint32_t scan = Aw;   // slots below the alphabet size stay empty: base 0 (leaf) misses there
  for (int s = 0; s < n; ++s) {
    const int64_t b0 = A.arc_begin[s], b1 = A.arc_begin[s + 1];
    if (b0 == b1) continue;
    uint32_t first_sym = arc_sym[(size_t)b0];
    for (int64_t k = b0 + 1; k < b1; ++k) first_sym = std::min(first_sym, arc_sym[(size_t)k]);
    uint32_t slot = find_free(std::max(scan, Aw + first_sym));
    uint32_t b;
    for (;;) {
      b = slot - first_sym;
      ensure((size_t)b + Aw + 2);
      if (!used_base[b]) {
        bool fits = true;
        for (int64_t k = b0; k < b1 && fits; ++k) fits = !used_slot[b + arc_sym[(size_t)k]];
        if (fits) break;
      }
      slot = find_free(slot + 1);
    }
    base[s] = b;
    used_base[b] = 1;
    for (int64_t k = b0; k < b1; ++k) { const uint32_t x = b + arc_sym[(size_t)k]; used_slot[x] = 1; next_free[x] = x + 1; }
    scan = find_free(scan);
  }
  uint32_t max_slot = Aw;
  for (int s = 0; s < n; ++s) if (base[s]) max_slot = std::max(max_slot, base[s] + Aw);
  T->da.assign((size_t)max_slot + 1, DaEntry{0xFFFFFFFFu, 0, 0, 0});
  for (int s = 0; s < n; ++s) {
    for (int64_t k = A.arc_begin[s]; k < A.arc_begin[s + 1]; ++k) {
      const Arc& a = A.arcs[k];
      DaEntry& e = T->da[(size_t)base[s] + arc_sym[(size_t)k]];
      e.check = base[s];
      e.dst = a.dst >= 0 ? (base[a.dst] | (A.is_final[a.dst] ? kDaFinalBit : 0u)) : 0u;   // DEAD: a non-final leaf
      e.ow = a.ow;
    }
  }
  T->root = base[0];
  if (T->root == 0) { *err = "initial state has no transitions"; return false; }
  {
    const uint16_t ds = T->sym_of_cp[kSpDelim];
    T->delim_is_token = false;
    if (ds != kNoSym) {
      const DaEntry& e = T->da[(size_t)T->root + ds];
      T->delim_is_token = e.check == T->root && (e.dst & kDaFinalBit) != 0;
    }
  }

  // ---- charmap ----
  T->has_charmap = charmap_dump >= 0;
  if (T->has_charmap) {
    FixedMap cm;
    if (!LdbImage::parse_fixedmap(ldb.dump(charmap_dump), &cm, err)) return false;
    flatten_charmap(cm, &T->norm_count, &T->norm_first, &T->norm_values);
  }

  // ---- BPE: the arc order as one integer ----
  // The BPE family sorts a segment's arcs by (rank descending -- with-merges only), id, start
  // (FATokenSegmentationTools_1best_bpe_t.h:238-255, ..._with_merges_t.h:242-262).  rank is a function
  // of the key, so (rank, id) collapses into a dense per-key ordinal and a sort key is one integer.
  T->bpe_ord.clear(); T->bpe_id_of_ord.clear(); T->bpe_ord_ok = false; T->bpe_singles_first = false;
  const bool bpe = T->tok_algo == kTokenizeBpe || T->tok_algo == kTokenizeBpeOpt || T->tok_algo == kTokenizeBpeOptWithMerges;
  if (bpe && !T->info.empty()) {
    const bool merges = T->tok_algo == kTokenizeBpeOptWithMerges;
    const size_t n = T->info.size();
    std::vector<int> order; order.reserve(n);
    bool ok = true;
    for (size_t k = 0; k < n; ++k) {
      if (T->info[k].id == INT32_MIN) continue;              // unusable row: never an arc of a valid model
      if (T->info[k].score != T->info[k].score) ok = false;  // a NaN rank has no place in an order
      order.push_back((int)k);
    }
    std::sort(order.begin(), order.end(), [&](int a, int b) {
      const SegInfo& x = T->info[a]; const SegInfo& y = T->info[b];
      if (merges) { if (x.score > y.score) return true; if (x.score < y.score) return false; }
      return x.id < y.id;
    });
    T->bpe_ord.assign(n, -1);
    int ord = -1;
    for (size_t i = 0; i < order.size(); ++i) {
      const SegInfo& x = T->info[order[i]];
      const bool same = i > 0 && T->info[order[i - 1]].id == x.id && (!merges || T->info[order[i - 1]].score == x.score);
      if (!same) { ++ord; T->bpe_id_of_ord.push_back(x.id); }
      T->bpe_ord[order[i]] = ord;
    }
    T->bpe_ord_ok = ok && ord < (1 << 20) - 1;               // 20 bits in a sort key, all-ones reserved
    // do all one-symbol tokens sort before every longer token?  (byte-level BPE: the 256 bytes come
    // first.)  Then the greedy claim takes every one-symbol arc before anything is marked
    // intermediate, and a kernel may start from that state instead of sorting those arcs.
    std::vector<uint8_t> single(n, 0);
    int max_single = -1, min_multi = INT_MAX;
    for (int s = 0; s < T->alphabet; ++s) {
      const DaEntry& e = T->da[(size_t)T->root + (size_t)s];
      if (e.check == T->root && (e.dst & kDaFinalBit) && e.ow >= 0 && (size_t)e.ow < n) single[(size_t)e.ow] = 1;
    }
    for (size_t k = 0; k < n; ++k) {
      if (T->bpe_ord[k] < 0) continue;
      if (single[k]) max_single = std::max(max_single, T->bpe_ord[k]);
      else min_multi = std::min(min_multi, T->bpe_ord[k]);
    }
    T->bpe_singles_first = T->bpe_ord_ok && max_single < min_multi;
  }
  return true;
}

}  // namespace bfb200
