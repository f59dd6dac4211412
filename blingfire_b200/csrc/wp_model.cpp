// wp_model.cpp -- see wp_model.h.  Host-only, runs once per LoadModel.
#include "wp_model.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <array>

namespace bfb200 {

// Host-only view of the automaton through LexerTables::next() (stored arcs when the dense table is not on the host):
// wp_word runs over it at load time exactly as it runs over the dense table on the device.
struct HostArcs {};
template <>
struct WpGlobal<HostArcs> {
  const LexerTables* T;
  const int32_t* tag_of_state;
  uint32_t NC1, first_final, cls_caret, cls_dollar;
  int max_token_length;
};
inline uint32_t wp_step(const WpGlobal<HostArcs>& g, uint32_t q, uint32_t c) {
  const uint32_t d = g.T->next(q, c);
  return d == kNoState ? kNone32 : d;
}

namespace {

struct TopTok {
  int fq, f2, t2;
};

// The top-level loop of wp_chunk (wp_core.cuh; FALexTools_t.h:229-397 at depth 1) on a run of L positions of
// top-level class t, in isolation: a chunk ends where no walk can go on (the next position is a sync start) or
// at the end of the document, where the right anchor is fed (`dollar`).
void top_run(const FastPath& F, int t, int L, bool caret, bool dollar, int max_tok, std::vector<TopTok>* out) {
  out->clear();
  const int NT = F.NT;
  for (int from = caret ? -1 : 0; from < L; ++from) {
    int q = 0, j = from;
    const int bound = std::min(from + max_tok, L);
    if (j == -1) {
      const uint8_t d = F.ttop[F.tc_caret];
      if (d == 0xFF) continue;
      q = d; j = 0;
    }
    int fq = -1, fpos = -1;
    for (; j < bound; ++j) {
      const uint8_t d = F.ttop[(size_t)q * NT + t];
      if (d == 0xFF) break;
      q = d;
      if (F.top_final[q]) { fq = q; fpos = j; }
    }
    if (j == L && dollar) {
      const uint8_t d = F.ttop[(size_t)q * NT + F.tc_dollar];
      if (d != 0xFF && F.top_final[d]) { fq = d; fpos = j; }
    }
    if (fpos == -1) continue;
    out->push_back(TopTok{fq, std::max(from, 0), std::min(fpos, L - 1)});
    if (fpos > from) from = fpos;
  }
}

// Top-level classes are put into GROUPS such that a chunk whose positions all belong to one group has an outcome
// that does not depend on which of the group's classes stand where:
//   group 0            every class without such a property (chunks with one of them go through the loops)
//   group 1  DEAD      no walk state has a transition on the class: such a position matches nothing, starts nothing
//   closed(s)          delta(initial, t) = s and delta(s, t) = s for every class t of the group: whatever the mix, a run
//                      is ONE token whose final state is s
//   single(s)          delta(initial, t) = s and s has no transition at all: a run of length 1 is one token
// (bert_base_tok: white space and unmapped characters are DEAD, the eleven letter/digit classes the
// special-token patterns "[UNK]" ... split the alphabet into form one closed group, punctuation and CJK one single group,
// '[' stays in group 0.)
struct TopGroups {
  std::vector<int> group_of_tc;          // [NT]
  std::vector<std::vector<int>> members; // [group] classes
  std::vector<int> state_of_group;       // [group] s, -1 for groups 0 and 1
  std::vector<uint8_t> closed;           // [group]
};

TopGroups make_groups(const FastPath& F) {
  const int NT = F.NT;
  TopGroups G;
  G.group_of_tc.assign((size_t)NT, 0);
  G.members.assign(2, {});
  G.state_of_group.assign(2, -1);
  G.closed.assign(2, 0);
  std::map<std::pair<int, int>, int> ids;   // (s, closed) -> group
  for (int t = 0; t < NT; ++t) {
    // (an anchor's class is treated like any other column: the "unmapped" class often shares its column, all-none)
    bool dead = true;
    for (int q = 0; q < F.K; ++q) if (F.ttop[(size_t)q * NT + t] != 0xFF) dead = false;
    int grp = 0;
    if (dead) grp = 1;
    else {
      const uint8_t s = F.ttop[t];
      if (s != 0xFF && F.top_final[s]) {
        bool closed = F.ttop[(size_t)s * NT + t] == s;
        bool leaf = true;
        for (int t2 = 0; t2 < NT; ++t2) if (t2 != F.tc_dollar && F.ttop[(size_t)s * NT + t2] != 0xFF) leaf = false;
        if (closed || leaf) {
          auto key = std::make_pair((int)s, closed ? 1 : 0);
          auto it = ids.find(key);
          if (it == ids.end()) {
            it = ids.emplace(key, (int)G.members.size()).first;
            G.members.push_back({}); G.state_of_group.push_back(s); G.closed.push_back(closed ? 1 : 0);
          }
          grp = it->second;
        }
      }
    }
    G.group_of_tc[t] = grp;
    G.members[grp].push_back(t);
  }
  return G;
}

// No chunk made of classes of the group, of any length, at any place of the document, produces a WORD token: no
// state of the closure of {initial, delta(initial, ^)} under the group -- nor its successor on $ -- is a WORD final.
bool group_is_inert(const FastPath& F, const std::vector<int>& members) {
  const int NT = F.NT;
  std::vector<uint8_t> seen((size_t)F.K, 0);
  std::vector<int> q{0};
  seen[0] = 1;
  const uint8_t c0 = F.ttop[F.tc_caret];
  if (c0 != 0xFF && !seen[c0]) { seen[c0] = 1; q.push_back(c0); }
  for (size_t i = 0; i < q.size(); ++i)
    for (int t : members) {
      const uint8_t d = F.ttop[(size_t)q[i] * NT + t];
      if (d != 0xFF && !seen[d]) { seen[d] = 1; q.push_back(d); }
    }
  for (int s : q) {
    if (F.top_final[s] && F.top_tag[s] == 1) return false;
    const uint8_t d = F.ttop[(size_t)s * NT + F.tc_dollar];
    if (d != 0xFF && F.top_final[d] && F.top_tag[d] == 1) return false;
  }
  return true;
}

uint64_t splitmix(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

bool cuckoo_build(const std::vector<WpWordSlot>& keys, WpWords* W, std::vector<WpWordSlot>* slots) {
  // each half holds >= 16 x the vocabulary's words: the rest of the slots take the words learned at run time, and a new
  // word finds one of its two slots free as long as the table stays sparse (64 MB for bert_base_tok; what is touched
  // is a sector or two per distinct word of the corpus)
  uint32_t log2 = 12;
  while (((size_t)1 << log2) < keys.size() * 16 + 16) ++log2;
  uint64_t seed = 0x5EEDB200ull;
  for (; log2 <= 26; ++log2) {
    const uint32_t S = 1u << log2;
    for (int attempt = 0; attempt < 16; ++attempt) {
      for (int i = 0; i < 9; ++i) W->mul[i] = (uint32_t)splitmix(&seed) | 1u;
      W->log2_size = log2;
      slots->assign((size_t)2 * S, WpWordSlot{});
      bool ok = true;
      for (const WpWordSlot& k : keys) {
        WpWordSlot cur = k;
        int side = 0, kick = 0;
        for (; kick < 2000; ++kick) {
          uint32_t k8[8];
          for (int i = 0; i < 4; ++i) { k8[i] = cur.kw[i]; k8[4 + i] = cur.kw_hi[i]; }
          const uint32_t h = wp_key_hash(k8, W->mul, true);
          const uint32_t idx = side == 0 ? wp_slot1(*W, h) : wp_slot2(*W, h);
          WpWordSlot& dst = (*slots)[idx];
          if (dst.meta == 0) { dst = cur; break; }
          std::swap(cur, dst);
          side ^= 1;
        }
        if (kick == 2000) { ok = false; break; }
      }
      if (ok) return true;
    }
  }
  return false;
}

// Fills the whole-word table: every class sequence within one group of top-level classes (at most max_len classes) for which
// the function sub-grammar yields exactly one piece.  Candidates are the paths of the function automaton from
// its two entry states; each is decided by wp_word itself.
void build_words(const LexerTables& T, const TopGroups& G, const std::vector<uint32_t>& kind_of_group, WpBlob* out) {
  using TE = HostArcs;
  const FastPath& F = T.fast;
  WpWords& W = out->words;
  WpGlobal<TE> g{};
  g.T = &T; g.tag_of_state = T.tag_of_state.data();
  g.NC1 = (uint32_t)T.NC + 1; g.first_final = T.first_final; g.cls_caret = T.cls_caret; g.cls_dollar = T.cls_dollar;
  g.max_token_length = T.max_token_length;

  std::map<std::array<uint32_t, 8>, int32_t> found;
  const int max_len = (int)W.max_len;
  for (int grp = 2; grp < (int)G.members.size(); ++grp) {
    if (!(kind_of_group[grp] & (kKindWordRun | kKindWordOne))) continue;
    const uint32_t root = F.top_fn_root[G.state_of_group[grp]], caret = F.top_fn_caret[G.state_of_group[grp]];
    const uint32_t entries[2] = {caret, root};
    for (uint32_t entry : entries) {
      if (entry == kNoState) continue;
      // depth-first over the stored arcs whose class belongs to t
      struct Frame { uint32_t state; int64_t arc; };
      std::vector<Frame> st;
      uint16_t seq[kMaxFastLen + 1];
      int32_t ids[kMaxFastLen + 1];
      st.push_back(Frame{entry, T.arc_begin[entry]});
      while (!st.empty()) {
        Frame& f = st.back();
        if (f.arc >= T.arc_begin[(size_t)f.state + 1]) { st.pop_back(); continue; }
        const uint32_t c = T.arc_label[f.arc];
        const uint32_t d = T.arc_dst[f.arc];
        ++f.arc;
        if (G.group_of_tc[F.tc_of_class[c]] != grp || d == T.dead) continue;
        const int depth = (int)st.size();                   // length of the sequence with c appended
        seq[depth - 1] = (uint16_t)c;
        if (((kind_of_group[grp] & kKindWordRun) || depth == 1) &&
            (T.is_final(d) || T.is_final(T.next(d, T.cls_dollar)))) {
          for (int i = 0; i < depth; ++i) ids[i] = kNoPiece;
          const int n = wp_word<TE>(g, seq, 0, depth - 1, root, caret, ids);
          bool single = n == 1 && ids[0] != kNoPiece;
          for (int i = 1; i < depth && single; ++i) single = ids[i] == kNoPiece;
          if (single) {
            std::array<uint32_t, 8> kw;
            wp_pack_key_any(W.cpw, seq, depth, depth, W.cb, kw.data());
            found[kw] = ids[0];
          }
        }
        if (depth < max_len && (kind_of_group[grp] & kKindWordRun)) st.push_back(Frame{d, T.arc_begin[d]});
      }
    }
  }
  std::vector<WpWordSlot> keys;
  keys.reserve(found.size());
  for (const auto& kv : found) {
    WpWordSlot s{};
    for (int i = 0; i < 4; ++i) { s.kw[i] = kv.first[i]; s.kw_hi[i] = kv.first[4 + i]; }
    s.id[0] = kv.second;
    s.meta = kSlotValid | 1u;
    keys.push_back(s);
  }
  out->word_count = (int64_t)keys.size();
  if (!cuckoo_build(keys, &W, &out->word_slots)) {   // cannot happen below 2^26 slots; serve without the table
    W.max_len = 0; W.log2_size = 4;
    out->word_slots.assign(32, WpWordSlot{});
    out->word_count = 0;
  }
}

}  // namespace

void build_wp_blob(const LexerTables& T, WpBlob* out) {
  const FastPath& F = T.fast;
  WpBlobLayout& L = out->layout;
  L = WpBlobLayout{};
  L.K = F.K; L.NT = F.NT;
  L.tc_caret = F.tc_caret; L.tc_dollar = F.tc_dollar; L.tc_none = F.tc_none;

  // ---- chunk kinds: the top-level loop run at load time on runs of one class ----
  WpWords& W = out->words;
  W = WpWords{};
  W.cb = 1;
  while ((1u << W.cb) < (uint32_t)T.NC + 1) ++W.cb;
  W.cpw = std::max(1u, std::min(3u, 30u / W.cb));
  W.max_len = (uint32_t)std::max(0, std::min({kMaxFastLen, (int)(8 * W.cpw), T.max_token_length - 1}));
  const TopGroups G = make_groups(F);
  std::vector<uint32_t> kind_of_group(G.members.size(), 0);
  const uint8_t c0 = F.ttop[F.tc_caret];                      // delta(initial, ^)
  for (size_t grp = 1; grp < G.members.size(); ++grp) {
    uint32_t k = 0;
    if (grp == 1) k |= kKindDead;
    if (group_is_inert(F, G.members[grp])) k |= kKindInert;
    const int s = G.state_of_group[grp];
    if (s >= 0 && F.top_tag[s] == 1 && F.top_fn_root[s] != kNoState) {
      // one WORD token over the whole run, final state s whatever the mix of the group's classes
      k |= G.closed[grp] ? kKindWordRun : kKindWordOne;
      // at the start of the document the walk begins in delta(initial, ^) (FALexTools_t.h:244-252): same outcome iff that
      // state does not exist, or takes every class of the group to s as well
      bool caret_ok = true;
      if (c0 != 0xFF)
        for (int t : G.members[grp]) if (F.ttop[(size_t)c0 * F.NT + t] != s) caret_ok = false;
      if (caret_ok) k |= kKindCaretOk;
      // at the end of the document the right anchor is fed (:280-290): same outcome iff it leads nowhere final, or to a
      // final state with the same tag and function
      const uint8_t d = F.ttop[(size_t)s * F.NT + F.tc_dollar];
      if (d == 0xFF || !F.top_final[d] ||
          (F.top_tag[d] == F.top_tag[s] && F.top_fn_root[d] == F.top_fn_root[s] && F.top_fn_caret[d] == F.top_fn_caret[s]))
        k |= kKindDollarOk;
    }
    kind_of_group[grp] = k;
  }
  // cross-check of the structural argument: the top-level loop itself, run on runs of every single class
  {
    std::vector<TopTok> a;
    for (int t = 0; t < F.NT; ++t) {
      const uint32_t k = kind_of_group[G.group_of_tc[t]];
      for (int len = 1; len <= kMaxFastLen; ++len)
        for (int variant = 0; variant < 4; ++variant) {
          const bool caret = variant & 1, dollar = variant & 2;
          if ((caret && !(k & kKindCaretOk)) || (dollar && !(k & kKindDollarOk))) { if (!(k & kKindInert)) continue; }
          top_run(F, t, len, caret, dollar, T.max_token_length, &a);
          int words = 0;
          for (const TopTok& x : a) if (F.top_tag[x.fq] == 1) ++words;
          bool good = true;
          if (k & kKindInert) good = words == 0;
          else if ((k & kKindWordRun) || (len == 1 && (k & kKindWordOne))) {
            const int s = G.state_of_group[G.group_of_tc[t]];
            good = words == 1;
            for (const TopTok& x : a)
              if (F.top_tag[x.fq] == 1)
                good = good && x.f2 == 0 && x.t2 == len - 1 && F.top_fn_root[x.fq] == F.top_fn_root[s] && F.top_fn_caret[x.fq] == F.top_fn_caret[s];
          }
          if (!good) kind_of_group[G.group_of_tc[t]] = 0;      // never observed; the loops serve such a group
        }
    }
  }

  // ---- the whole-word table ----
  build_words(T, G, kind_of_group, out);
  if (W.max_len == 0)
    for (uint32_t& k : kind_of_group) k &= ~(kKindWordRun | kKindWordOne);
  std::vector<uint32_t> kind((size_t)F.NT, 0);
  for (int t = 0; t < F.NT; ++t) kind[t] = kind_of_group[G.group_of_tc[t]];

  // ---- the blob ----
  uint32_t off = 0;
  auto place = [&](uint32_t bytes, uint32_t align) { off = (off + align - 1) / align * align; const uint32_t o = off; off += bytes; return o; };
  L.off_ascii = place(128 * 4, 16);
  L.off_ttop = place((uint32_t)F.K * F.NT, 16);
  L.off_tag = place((uint32_t)F.K * 4, 16);
  L.off_root = place((uint32_t)F.K * 4, 16);
  L.off_caret = place((uint32_t)F.K * 4, 16);
  L.sync_shift = 1;
  while ((1 << L.sync_shift) < F.NT) ++L.sync_shift;
  L.off_sync = place(1u << (2 * L.sync_shift), 16);
  L.off_kind = place((uint32_t)F.NT * 4, 16);
  L.off_fn_root = place((uint32_t)F.NT * 4, 16);
  L.off_fn_caret = place((uint32_t)F.NT * 4, 16);
  L.total_bytes = (off + 15) / 16 * 16;

  out->bytes.assign(L.total_bytes, 0);
  uint8_t* bp = out->bytes.data();
  for (int cp = 0; cp < 128; ++cp) reinterpret_cast<uint32_t*>(bp + L.off_ascii)[cp] = T.clsx_of_cp[cp];
  // bit 7 of a top-level transition = "the destination is final" (wp_core.cuh kTopFinal); 0xFF stays "none"
  for (size_t i = 0; i < (size_t)F.K * F.NT; ++i) {
    const uint8_t d = F.ttop[i];
    bp[L.off_ttop + i] = (d != 0xFF && F.top_final[d]) ? (uint8_t)(d | 0x80) : d;
  }
  for (int t1 = 0; t1 < F.NT; ++t1)
    for (int t2 = 0; t2 < F.NT; ++t2)
      bp[L.off_sync + ((size_t)t1 << L.sync_shift) + t2] =
          (uint8_t)(((!((F.cross[(size_t)t1] >> t2) & 1ull) && F.ttop[(size_t)t2] != 0xFF) ? kSyncStart : 0) |
                    (G.group_of_tc[t1] != G.group_of_tc[t2] ? kSyncGroupChange : 0));
  std::memcpy(bp + L.off_tag, F.top_tag.data(), (size_t)F.K * 4);
  std::memcpy(bp + L.off_root, F.top_fn_root.data(), (size_t)F.K * 4);
  std::memcpy(bp + L.off_caret, F.top_fn_caret.data(), (size_t)F.K * 4);
  std::memcpy(bp + L.off_kind, kind.data(), (size_t)F.NT * 4);
  for (int t = 0; t < F.NT; ++t) {
    const int st = G.state_of_group[G.group_of_tc[t]];
    reinterpret_cast<uint32_t*>(bp + L.off_fn_root)[t] = st >= 0 ? F.top_fn_root[st] : kNoState;
    reinterpret_cast<uint32_t*>(bp + L.off_fn_caret)[t] = st >= 0 ? F.top_fn_caret[st] : kNoState;
  }
}

}  // namespace bfb200
