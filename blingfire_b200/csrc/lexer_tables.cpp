// lexer_tables.cpp -- see lexer_tables.h.  Host-only, runs once per LoadModel.
#include "lexer_tables.h"

#include <algorithm>
#include <thread>
#include <map>
#include <set>

namespace bfb200 {

namespace {

constexpr int kMaxTag = 65535;   // FALimits.h:49

struct WbdConf {
  int fsm_dump = -1, acts_dump = -1, charmap_dump = -1;
  int max_depth = 2;             // FALexTools_t.h:108 (DefMaxDepth)
  int max_token_length = 300;    // FALimits.h:35
  bool ignore_case = false;
};

// FAWbdConfKeeper::Initialize (FAWbdConfKeeper.cpp:56-232)
bool parse_wbd_conf(const std::vector<int>& v, WbdConf* c, std::string* err) {
  for (size_t i = 0; i < v.size(); ++i) {
    auto arg = [&](int* dst) { if (i + 1 >= v.size()) return false; *dst = v[++i]; return true; };
    int tmp = 0;
    switch (v[i]) {
      case kParamMapMode: if (!arg(&tmp) || tmp != kModePackTriv) { *err = "[wbd] map mode"; return false; } break;
      case kParamDepth: if (!arg(&c->max_depth) || c->max_depth < 0) { *err = "[wbd] depth"; return false; } break;
      case kParamMaxLength: if (!arg(&c->max_token_length) || c->max_token_length < 0) { *err = "[wbd] max length"; return false; } break;
      case kParamIgnoreCase: c->ignore_case = true; break;
      case kParamFsmType: if (!arg(&tmp) || tmp != kTypeMooreDfa) { *err = "[wbd] fsm type (only Moore DFA is served)"; return false; } break;
      case kParamFsm: if (!arg(&c->fsm_dump)) { *err = "[wbd] fsm"; return false; } break;
      case kParamMultiMap: if (!arg(&c->acts_dump)) { *err = "[wbd] multi-map"; return false; } break;
      case kParamCharmap: if (!arg(&c->charmap_dump)) { *err = "[wbd] charmap"; return false; } break;
      case kParamActData: case kParamPunkt: case kParamEos: case kParamEop: case kParamWord:
      case kParamXWord: case kParamSeg: case kParamIgnore: case kParamMaxTag:
        if (!arg(&tmp)) { *err = "[wbd] parameter"; return false; } break;
      default: *err = "[wbd] unknown parameter"; return false;
    }
  }
  if (c->fsm_dump < 0 || c->acts_dump < 0) { *err = "[wbd] needs fsm and multi-map"; return false; }
  if (c->ignore_case) { *err = "[wbd] ignore-case models are not served (no shipped model uses it)"; return false; }
  return true;
}

// where the function ids start in an action row, per FALexTools_t::Validate (FALexTools_t.h:157-202)
// returns -1 if the row is malformed
int action_fn_start(const std::vector<int>& a) {
  const int n = (int)a.size();
  if (n < 3) return -1;
  if (a[0] < -kMaxTag || a[0] > kMaxTag || a[1] < -kMaxTag || a[1] > kMaxTag) return -1;
  if (n == 3 && a[2] != 0) return 3;          // just one tag
  if (n > 3 && a[2] == 0) return 3;           // delimiter + fns, no tag
  if (n > 4 && a[3] == 0) return 4;           // tag, delimiter, fns
  return -1;
}

}  // namespace

bool build_lexer_tables(const LdbImage& ldb, LexerTables* T, std::string* err, bool dense_wide) {
  const std::vector<int>* sec = ldb.conf().get(kFuncWbd);
  if (!sec) { *err = "no [wbd] section"; return false; }
  WbdConf conf;
  if (!parse_wbd_conf(*sec, &conf, err)) return false;
  T->max_depth = conf.max_depth;
  T->max_token_length = conf.max_token_length;

  Automaton A;
  if (!LdbImage::parse_automaton(ldb.dump(conf.fsm_dump), /*mealy=*/false, &A, err)) return false;
  MultiMap acts;
  if (!LdbImage::parse_multimap(ldb.dump(conf.acts_dump), &acts, err)) return false;
  if (!acts.ptr_interface_ok) { *err = "action map is not int-valued"; return false; }
  FixedMap charmap;
  T->has_charmap = conf.charmap_dump >= 0;
  if (T->has_charmap && !LdbImage::parse_fixedmap(ldb.dump(conf.charmap_dump), &charmap, err)) return false;

  // ---- classes ----
  const int NC = A.num_classes;
  if (NC <= 0 || NC >= 65535) { *err = "unsupported class count"; return false; }
  T->NC = NC;
  auto cls = [&](int iw) -> uint32_t {   // GetNewIw; NC stands for "unmapped" (-1)
    const int c = A.class_of(iw);
    return (c < 0 || c >= NC) ? (uint32_t)NC : (uint32_t)c;
  };
  T->cls_caret = cls(kIwLAnchor);
  T->cls_dollar = cls(kIwRAnchor);
  const uint32_t cls_any = cls(kIwAny);

  // ---- function initial states: FAWbdConfKeeper::CalcFnIniStates (FAWbdConfKeeper.cpp:246-314) ----
  const int a_initial = 0;   // BFS id of the initial state
  {
    int state_r = (T->cls_dollar < (uint32_t)NC) ? A.dest(a_initial, (int)T->cls_dollar) : -1;
    int max_fn = -1;
    for (int id = 0; acts.get(id); ++id) {   // stops at the first missing key, like the reference
      const std::vector<int>& a = *acts.get(id);
      if ((int)a.size() < 3) { *err = "action row too short"; return false; }
      int i = 2;
      for (; i < (int)a.size(); ++i) if (a[i] == 0 && i + 1 < (int)a.size()) { ++i; break; }
      for (; i < (int)a.size(); ++i) { if (a[i] < 0) { *err = "negative function id"; return false; } max_fn = std::max(max_fn, a[i]); }
    }
    T->fn_ini.clear();
    if (state_r >= 0 && max_fn >= 0) {
      if (max_fn > kMaxTag) { *err = "too many functions"; return false; }
      T->fn_ini.assign((size_t)max_fn + 1, kNoState);
      T->fn_ini[0] = (uint32_t)a_initial;   // renumbered below
      for (int f = 1; f <= max_fn; ++f) {
        const uint32_t c = cls(f);
        const int d = c < (uint32_t)NC ? A.dest(state_r, (int)c) : -1;
        T->fn_ini[f] = d >= 0 ? (uint32_t)d : kNoState;   // -1 and DEAD both make the function unusable
      }
    }
  }

  // ---- renumber: non-finals first, explicit dead sink, finals last ----
  const int n = A.num_states();
  std::vector<uint32_t> newid((size_t)n);
  uint32_t next_id = 0;
  for (int s = 0; s < n; ++s) if (!A.is_final[s]) newid[s] = next_id++;
  T->dead = next_id++;
  T->first_final = next_id;
  for (int s = 0; s < n; ++s) if (A.is_final[s]) newid[s] = next_id++;
  T->NS = (int)next_id;
  T->initial = newid[a_initial];
  for (auto& f : T->fn_ini) if (f != kNoState) f = newid[f];
  T->wide_states = T->NS >= 65535;

  // ---- dense table with the IW_ANY fallback folded in ----
  const bool dense = dense_wide || !T->wide_states;
  T->dense_on_host = dense;
  const size_t W = (size_t)NC + 1;
  if (dense) {
    const size_t cells = (size_t)T->NS * W;
    if (cells > ((size_t)1 << 33)) { *err = "dense transition table too large"; return false; }
    // The table starts as "no transition" everywhere and only the arcs are written: a row is touched in
    // full only when its state has an IW_ANY arc (the 9.3 GB table of bert_multi_cased has 0.03 % of its
    // cells set).
    auto fill_parallel = [&](auto* p, auto none) {
      unsigned nt = cells < ((size_t)1 << 26) ? 1u : std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
      std::vector<std::thread> th;
      const size_t per = (cells + nt - 1) / nt;
      for (unsigned t = 0; t < nt; ++t) {
        const size_t lo = (size_t)t * per, hi = std::min(cells, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { std::fill(p + lo, p + hi, none); });
      }
      for (auto& x : th) x.join();
    };
    if (T->wide_states) { T->trans32.resize(cells); fill_parallel(T->trans32.data(), (uint32_t)kNoState); }
    else { T->trans16.resize(cells); fill_parallel(T->trans16.data(), (uint16_t)0xFFFF); }
    auto fill_rows = [&](auto* table, auto none) {
      using E = decltype(none);
      for (int s = 0; s < n; ++s) {
        E* row = table + (size_t)newid[s] * W;
        for (int64_t k = A.arc_begin[s]; k < A.arc_begin[s + 1]; ++k) {
          const Arc& a = A.arcs[k];
          if (a.label < 0 || a.label >= NC) continue;   // a class no input can produce
          row[a.label] = (E)(a.dst == kDeadState ? T->dead : newid[a.dst]);
        }
        if (cls_any < (uint32_t)NC && row[cls_any] != none) {
          const E any = row[cls_any];
          for (size_t c = 0; c < W; ++c) if (row[c] == none) row[c] = any;
        }
      }
    };
    if (T->wide_states) fill_rows(T->trans32.data(), (uint32_t)kNoState); else fill_rows(T->trans16.data(), (uint16_t)0xFFFF);
  }

  // ---- the stored arcs in the new numbering ----
  {
    std::vector<int32_t> old_of(T->NS, -1);
    for (int s = 0; s < n; ++s) old_of[newid[s]] = s;
    T->arc_begin.assign((size_t)T->NS + 1, 0);
    T->arc_label.clear(); T->arc_dst.clear();
    for (int ns = 0; ns < T->NS; ++ns) {
      T->arc_begin[ns] = (int64_t)T->arc_label.size();
      const int s = old_of[ns];
      if (s < 0) continue;
      for (int64_t k = A.arc_begin[s]; k < A.arc_begin[s + 1]; ++k) {
        const Arc& a = A.arcs[k];
        if (a.label < 0 || a.label >= NC) continue;
        T->arc_label.push_back((uint32_t)a.label);
        T->arc_dst.push_back(a.dst == kDeadState ? T->dead : newid[a.dst]);
      }
    }
    T->arc_begin[T->NS] = (int64_t)T->arc_label.size();
    // IW_ANY: the fallback of every class the state has no arc for (FALexTools_t.h:266-270)
    T->any_dst.assign((size_t)T->NS, kNoState);
    if (cls_any < (uint32_t)NC)
      for (int ns = 0; ns < T->NS; ++ns)
        for (int64_t k = T->arc_begin[ns]; k < T->arc_begin[(size_t)ns + 1]; ++k)
          if (T->arc_label[k] == cls_any) T->any_dst[ns] = T->arc_dst[k];
  }

  // ---- rule ids, actions ----
  T->ow_of_state.assign((size_t)T->NS, -1);
  T->orig_offset.assign((size_t)T->NS, -1);
  for (int s = 0; s < n; ++s) T->orig_offset[newid[s]] = A.orig[s];
  T->tag_of_state.assign((size_t)T->NS, 0);
  int num_acts = (int)acts.rows.size();
  T->act_begin.assign((size_t)num_acts + 1, 0);
  T->act_data.clear();
  for (int k = 0; k < num_acts; ++k) {
    T->act_begin[k] = (int32_t)T->act_data.size();
    if (acts.present[k]) T->act_data.insert(T->act_data.end(), acts.rows[k].begin(), acts.rows[k].end());
  }
  T->act_begin[num_acts] = (int32_t)T->act_data.size();
  for (int k = 0; k < num_acts; ++k) {
    if (!acts.present[k]) continue;
    const std::vector<int>& a = acts.rows[k];
    const int fs = action_fn_start(a);
    if (fs < 0) { *err = "malformed action row"; return false; }
    for (int i = fs; i < (int)a.size(); ++i)
      if (a[i] < 0 || (size_t)a[i] >= T->fn_ini.size() || T->fn_ini[a[i]] == kNoState) { *err = "action calls an undefined function"; return false; }
  }
  for (int s = 0; s < n; ++s) {
    if (!A.is_final[s]) continue;
    const int ow = A.moore_ow[s];
    if (!acts.get(ow)) { *err = "final state without a rule/action"; return false; }
    T->ow_of_state[newid[s]] = ow;
    T->tag_of_state[newid[s]] = (*acts.get(ow))[2];
  }

  // ---- class maps ----
  const size_t max_iw = A.remap ? A.class_of_iw.size() : (size_t)NC;
  T->cls_of_iw.assign(max_iw, (uint16_t)NC);
  for (size_t iw = 0; iw < max_iw; ++iw) T->cls_of_iw[iw] = (uint16_t)cls((int)iw);

  T->cls_words_of_cp.assign((size_t)kMaxCodePoint + 1, (uint16_t)NC);
  for (int cp = 0; cp <= kMaxCodePoint; ++cp) {
    int x = cp == 0 ? 0x20 : cp;                                          // blingfiretokdll.cpp:482
    if (x < kIwEpsilon) x = kIwEpsilon;                                   // FALexTools_t.h:259-261
    T->cls_words_of_cp[cp] = (uint16_t)cls(x);
  }

  T->charmap_one_to_one = true;
  if (T->has_charmap) {
    int tmp[16];
    for (int cp = charmap.min_key; cp <= charmap.max_key && cp <= kMaxCodePoint; ++cp) {
      const int c = charmap.get(cp, tmp, 10);
      if (c != -1 && c != 1) { T->charmap_one_to_one = false; break; }
    }
  }
  if (T->charmap_one_to_one) {
    T->cls_of_cp.assign((size_t)kMaxCodePoint + 1, (uint16_t)NC);
    int tmp[16];
    for (int cp = 0; cp <= kMaxCodePoint; ++cp) {
      int x = cp;
      if (T->has_charmap && charmap.get(cp, tmp, 10) == 1) x = tmp[0];   // FANormalize (FAUtils_cl.h:311-369)
      if (x < kIwEpsilon) x = kIwEpsilon;                                   // FALexTools_t.h:259-261
      T->cls_of_cp[cp] = (uint16_t)cls(x);
    }
  } else {
    T->norm_count.assign((size_t)kMaxCodePoint + 1, 0xFF);
    T->norm_first.assign((size_t)kMaxCodePoint + 1, 0);
    int tmp[16];
    for (int cp = charmap.min_key; cp <= charmap.max_key && cp <= kMaxCodePoint; ++cp) {
      const int c = charmap.get(cp, tmp, 10);
      if (c == -1) continue;
      // counts outside 1..10 delete the character (no FANormalize branch fires)
      const int eff = (c >= 1 && c <= 10) ? c : 0;
      T->norm_count[cp] = (uint8_t)eff;
      T->norm_first[cp] = (uint32_t)T->norm_values.size();
      for (int i = 0; i < eff; ++i) T->norm_values.push_back(tmp[i]);
    }
  }

  // ---- flat two-level WordPiece shape? ----
  FastPath& F = T->fast;
  F = FastPath{};
  auto fail = [&](const char* why) { F.ok = false; F.why_not = why; return true; };

  // classes some data symbol (Iw >= 3) can produce, plus "unmapped"
  std::vector<uint8_t> is_data(W, 0);
  is_data[NC] = 1;
  for (size_t iw = kIwEpsilon; iw < max_iw; ++iw) is_data[T->cls_of_iw[iw]] = 1;

  // closure over the sparse arcs (renumbered ids); a superset of the dense-table closure is fine
  std::vector<int32_t> old_of_new((size_t)T->NS, -1);
  for (int s = 0; s < n; ++s) old_of_new[newid[s]] = s;
  auto closure = [&](const std::vector<uint32_t>& seeds, size_t limit, std::vector<uint32_t>* out) {
    std::vector<uint8_t> seen((size_t)T->NS, 0);
    std::vector<uint32_t> q(seeds.begin(), seeds.end());
    out->assign(seeds.begin(), seeds.end());
    for (uint32_t s : seeds) seen[s] = 1;
    while (!q.empty()) {
      const uint32_t s = q.back(); q.pop_back();
      const int os = old_of_new[s];
      if (os < 0) continue;   // the dead sink has no arcs
      for (int64_t k = A.arc_begin[os]; k < A.arc_begin[os + 1]; ++k) {
        const Arc& a = A.arcs[k];
        if (a.label < 0 || a.label >= NC) continue;
        if (!is_data[a.label] && (uint32_t)a.label != cls_any) continue;
        const uint32_t d = a.dst == kDeadState ? T->dead : newid[a.dst];
        if (seen[d]) continue;
        seen[d] = 1; out->push_back(d); q.push_back(d);
        if (out->size() > limit) return false;
      }
    }
    return true;
  };

  std::vector<uint32_t> walk;   // top-level states a walk can be in (global ids); walk[0] = initial
  {
    std::vector<uint32_t> seeds{T->initial};
    const uint32_t c0 = T->next(T->initial, T->cls_caret);
    if (c0 != kNoState && c0 != T->initial) seeds.push_back(c0);
    if (!closure(seeds, 200, &walk)) return fail("top-level automaton is not small (closure > 200 states)");
  }
  // states only reachable through the right anchor: finality/action matter, their rows do not
  std::vector<uint32_t> top = walk;
  std::map<uint32_t, int> local;
  for (size_t i = 0; i < top.size(); ++i) local[top[i]] = (int)i;
  const int K_walk = (int)walk.size();
  for (int i = 0; i < K_walk; ++i) {
    const uint32_t d = T->next(walk[i], T->cls_dollar);
    if (d != kNoState && !local.count(d)) { local[d] = (int)top.size(); top.push_back(d); }
  }
  if (top.size() > 250) return fail("top-level automaton is not small");
  F.K = (int)top.size();

  // columns -> top-level classes
  std::map<std::vector<uint8_t>, int> col_id;
  F.tc_of_class.assign(W, 0);
  std::vector<std::vector<uint8_t>> cols;
  for (size_t c = 0; c < W; ++c) {
    std::vector<uint8_t> col((size_t)K_walk, 0xFF);
    for (int i = 0; i < K_walk; ++i) {
      if (!is_data[c] && c != T->cls_dollar && !(c == T->cls_caret && i == 0)) continue;
      const uint32_t d = T->next(walk[i], (uint32_t)c);
      if (d == kNoState) continue;
      auto it = local.find(d);
      if (it == local.end()) { if (is_data[c]) return fail("internal: closure not closed"); continue; }
      col[i] = (uint8_t)it->second;
    }
    auto it = col_id.find(col);
    if (it == col_id.end()) { it = col_id.emplace(col, (int)cols.size()).first; cols.push_back(col); }
    if (cols.size() > 64) return fail("more than 64 top-level character classes");
    F.tc_of_class[c] = (uint8_t)it->second;
  }
  F.NT = (int)cols.size();
  F.tc_caret = F.tc_of_class[T->cls_caret];
  F.tc_dollar = F.tc_of_class[T->cls_dollar];
  F.tc_none = F.tc_of_class[NC];
  F.ttop.assign((size_t)F.K * F.NT, 0xFF);
  for (int t = 0; t < F.NT; ++t) for (int i = 0; i < K_walk; ++i) F.ttop[(size_t)i * F.NT + t] = cols[t][i];

  // which adjacent class pairs a single top-level walk can consume
  std::vector<uint8_t> tc_is_data((size_t)F.NT, 0);
  for (size_t c = 0; c < W; ++c) if (is_data[c]) tc_is_data[F.tc_of_class[c]] = 1;
  F.cross.assign((size_t)F.NT, 0);
  for (int i = 0; i < K_walk; ++i)
    for (int t1 = 0; t1 < F.NT; ++t1) {
      if (!tc_is_data[t1]) continue;
      const uint8_t d = F.ttop[(size_t)i * F.NT + t1];
      if (d == 0xFF || d >= K_walk) continue;
      for (int t2 = 0; t2 < F.NT; ++t2)
        if (tc_is_data[t2] && F.ttop[(size_t)d * F.NT + t2] != 0xFF) F.cross[t1] |= (1ull << t2);
    }

  // top-level actions: [0,0,Tag] or [0,0,Tag,0,fn] with 1 <= Tag <= 4
  F.top_final.assign((size_t)F.K, 0);
  F.top_tag.assign((size_t)F.K, 0);
  F.top_fn_root.assign((size_t)F.K, kNoState);
  F.top_fn_caret.assign((size_t)F.K, kNoState);
  std::set<int> called;
  for (int i = 0; i < F.K; ++i) {
    const uint32_t g = top[i];
    if (!T->is_final(g)) continue;
    const int32_t* a = T->act_data.data() + T->act_begin[T->ow_of_state[g]];
    const int an = T->act_begin[T->ow_of_state[g] + 1] - T->act_begin[T->ow_of_state[g]];
    if (a[0] != 0 || a[1] != 0) return fail("top-level action with left/right context");
    if (a[2] < 1 || a[2] > 4) return fail("top-level action tag outside 1..4");
    if (!(an == 3 || an == 5)) return fail("top-level action calls more than one function");
    F.top_final[i] = 1;
    F.top_tag[i] = a[2];
    if (an == 5) {
      const int fn = a[4];
      if (fn == 0) return fail("top-level action re-enters the main function");
      if (T->max_depth >= 2) {   // deeper calls return 0 tokens (FALexTools_t.h:222-224)
        F.top_fn_root[i] = T->fn_ini[fn];
        F.top_fn_caret[i] = T->next(T->fn_ini[fn], T->cls_caret);
        called.insert(fn);
      }
    }
  }
  // called functions: every reachable final is a plain [0,0,id] with id > 4
  for (int fn : called) {
    std::vector<uint32_t> sub;
    std::vector<uint32_t> seeds{T->fn_ini[fn]};
    const uint32_t c0 = T->next(T->fn_ini[fn], T->cls_caret);
    if (c0 != kNoState && c0 != T->fn_ini[fn]) seeds.push_back(c0);
    if (!closure(seeds, (size_t)T->NS + 1, &sub)) return fail("internal: function closure");
    const size_t nsub = sub.size();
    for (size_t i = 0; i < nsub; ++i) {
      const uint32_t d = T->next(sub[i], T->cls_dollar);
      if (d != kNoState) sub.push_back(d);
    }
    for (uint32_t g : sub) {
      if (!T->is_final(g)) continue;
      const int ow = T->ow_of_state[g];
      const int32_t* a = T->act_data.data() + T->act_begin[ow];
      const int an = T->act_begin[ow + 1] - T->act_begin[ow];
      if (an != 3 || a[0] != 0 || a[1] != 0) return fail("function action is not a plain [0,0,id]");
      if (a[2] <= 4) return fail("function action id <= 4");
    }
  }
  F.ok = true;
  if (T->charmap_one_to_one) {
    T->clsx_of_cp.resize(T->cls_of_cp.size());
    for (size_t cp = 0; cp < T->cls_of_cp.size(); ++cp)
      T->clsx_of_cp[cp] = (uint32_t)T->cls_of_cp[cp] | ((uint32_t)F.tc_of_class[T->cls_of_cp[cp]] << 16);
  }
  return true;
}

}  // namespace bfb200
