// sp_twin.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Sequential host driver over the product's own [pos-dict] table builder (seg_tables.cpp): the
// double-array Mealy automaton, the symbol map, the I2Info array and the flattened charmap.
// It runs the reference's algorithms in their sequential form over THOSE tables, so that
// tests/test_seg_tables.py can check the flattening against the oracle (which reads the packed
// image) on the CPU box.  The GPU kernels (sp_kernel.cu) are checked against the oracle
// directly in tests/test_gpu_parity_sp.py.  Never linked into the product.
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../blingfire_b200/csrc/ldb.h"
#include "../../blingfire_b200/csrc/seg_tables.h"

using namespace bfb200;

struct SpTwin {
  LdbImage ldb;
  SegTables S;
  std::string err;
};

static bool is_white(int c) {
  return c <= 0x20 || c == 0xa0 || (c >= 0x2000 && c <= 0x200f) || c == 0x202f || c == 0x205f || c == 0x2060 ||
         c == 0x2420 || c == 0x2424 || c == 0x3000 || c == 0xfeff;
}

static bool step(const SegTables& S, uint32_t* q, int c, int* ow, bool* fin) {
  if ((unsigned)c > 0x10FFFFu) return false;
  return S.step(q, S.sym_of_cp[c], ow, fin);
}

extern "C" {

void* sptwin_load(const char* path) {
  SpTwin* t = new SpTwin();
  if (!t->ldb.load_file(path)) { t->err = t->ldb.error(); return t; }
  if (!build_seg_tables(t->ldb, &t->S, &t->err) && t->err.empty()) t->err = "build failed";
  return t;
}
void sptwin_free(void* h) { delete (SpTwin*)h; }
const char* sptwin_error(void* h) { return ((SpTwin*)h)->err.c_str(); }
int sptwin_info(void* h, int what) {
  const SegTables& S = ((SpTwin*)h)->S;
  switch (what) {
    case 0: return (int)S.da.size();
    case 1: return S.alphabet;
    case 2: return (int)S.info.size();
    case 3: return S.max_arc_len;
    case 4: return S.delim_inside_tokens ? 1 : 0;
    case 5: return S.tok_algo;
    case 6: return S.id_offset;
    case 7: return S.use_raw_bytes ? 1 : 0;
    case 8: return S.has_charmap ? 1 : 0;
    case 9: return S.bpe_ord_ok ? 1 : 0;
    case 10: return S.bpe_singles_first ? 1 : 0;
    case 11: return (int)S.bpe_id_of_ord.size();
    case 12: return S.delim_is_token ? 1 : 0;
    default: return -1;
  }
}

// The BPE arc order as one integer (seg_tables.h): for every pair of usable keys sampled with
// `stride`, ord[a] < ord[b] must agree with the reference comparator on (rank desc when merges, id)
// (FATokenSegmentationTools_1best_bpe_t.h:238-255, ..._with_merges_t.h:242-262), and id_of_ord must
// invert it.  Returns the number of violations, -1 if the table is absent.
int sptwin_check_bpe_order(void* h, int stride) {
  const SegTables& S = ((SpTwin*)h)->S;
  if (S.bpe_ord.size() != S.info.size() || S.bpe_ord.empty()) return -1;
  const bool merges = S.tok_algo == kTokenizeBpeOptWithMerges;
  int bad = 0;
  std::vector<int> keys;
  for (size_t k = 0; k < S.info.size(); k += (size_t)stride) if (S.bpe_ord[k] >= 0) keys.push_back((int)k);
  for (int a : keys) {
    if (S.bpe_id_of_ord[(size_t)S.bpe_ord[a]] != S.info[a].id) ++bad;
    for (int b : keys) {
      const SegInfo& x = S.info[a]; const SegInfo& y = S.info[b];
      int cmp = 0;                                  // -1: a first
      if (merges && x.score != y.score) cmp = x.score > y.score ? -1 : 1;
      else if (x.id != y.id) cmp = x.id < y.id ? -1 : 1;
      const int oc = S.bpe_ord[a] < S.bpe_ord[b] ? -1 : (S.bpe_ord[a] > S.bpe_ord[b] ? 1 : 0);
      if (cmp != oc) ++bad;
    }
  }
  return bad;
}

}  // extern "C" (helpers below are C++)

namespace {

// blingfiretokdll.cpp:1372-1496: dummy prefix, decode, charmap, whitespace collapse.  1 = ok, 0 = the call returns 0.
int front_end(const SegTables& S, const char* s, int n, std::vector<int>* out) {
  const uint8_t* b = (const uint8_t*)s;
  int lo = 0;
  if (n >= 3 && b[0] == 0xEF && b[1] == 0xBB && b[2] == 0xBF) lo = 3;
  std::vector<int> buf;
  if (!S.no_dummy_prefix) buf.push_back(kSpDelim);
  const size_t off = buf.size();
  if (S.use_raw_bytes) {
    for (int p = lo; p < n; ++p) buf.push_back(b[p]);
  } else {
    int p = lo;
    while (p < n) {
      const int c = b[p];
      int len, cp;
      if (c < 0x80) { len = 1; cp = c; }
      else if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
      else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
      else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
      else return 0;
      if (p + len > n) return 0;
      for (int k = 1; k < len; ++k) { if ((b[p + k] & 0xC0) != 0x80) return 0; cp = (cp << 6) | (b[p + k] & 0x3F); }
      const int need = cp <= 0x7F ? 1 : cp <= 0x7FF ? 2 : cp <= 0xFFFF ? 3 : cp <= 0x10FFFF ? 4 : 0;
      if (need != len || (cp & 0xFFFFF800) == 0xD800) return 0;
      buf.push_back(cp);
      p += len;
    }
  }
  if (buf.size() == off) return 0;
  if (S.has_charmap) {
    std::vector<int> nb;
    for (int cp : buf) {
      const uint8_t c = ((unsigned)cp <= 0x10FFFFu) ? S.norm_count[cp] : 0xFF;
      if (c == 0xFF) nb.push_back(cp);
      else for (int k = 0; k < c; ++k) nb.push_back(S.norm_values[S.norm_first[cp] + k]);
    }
    if (nb.empty() || (int64_t)nb.size() > 2ll * (n + 1)) return 0;
    buf.swap(nb);
  }
  size_t j = 0;
  for (size_t i = 0; i < buf.size(); ++i) {
    const int c = buf[i];
    if (!is_white(c)) buf[j++] = c;
    else if (j == 0 || buf[j - 1] != kSpDelim) buf[j++] = kSpDelim;
  }
  if (j > 1 && buf[j - 1] == kSpDelim) --j;
  buf.resize(j);
  out->swap(buf);
  return out->empty() ? 0 : 1;
}

// FATokenSegmentationTools_1best_t.h:174-279 over buf[lo..hi) with the best score of position lo-1 given
// (0 for the empty prefix); appends the token ids.  false = an unusable key.
bool unigram_range(const SegTables& S, const std::vector<int>& buf, int lo, int hi, double* carry, int unk, std::vector<int>* res) {
  struct A { int begin, id; double score; };
  const int N = hi - lo;
  std::vector<A> best((size_t)N, A{-1, -1, -(double)FLT_MAX});
  for (int start = 0; start < N; ++start) {
    uint32_t q = S.root; int sum = 0; bool unknown = true;
    const double prev = start > 0 ? best[(size_t)start - 1].score : *carry;
    for (int i = start; i < N; ++i) {
      int ow; bool fin;
      if (!step(S, &q, buf[(size_t)lo + i], &ow, &fin)) break;
      sum += ow;
      if (fin) {
        if (sum < 0 || sum >= (int)S.info.size()) return false;
        const SegInfo si = S.info[sum];
        if (best[i].score < si.score + prev) best[i] = A{start, si.id, si.score + prev};
        unknown = false;
      }
    }
    if (unknown && best[start].score < -100000.0f + prev) {
      best[start] = A{start, -1, -100000.0f + prev};
      if (start > 0 && best[(size_t)start - 1].id == -1) best[start].begin = best[(size_t)start - 1].begin;
    }
  }
  std::vector<int> rev;
  int end = N - 1;
  while (end >= 0) { const A& a = best[end]; rev.push_back(a.id != -1 ? a.id : unk); end = a.begin - 1; if (a.begin < 0) break; }
  for (auto it = rev.rbegin(); it != rev.rend(); ++it) res->push_back(*it);
  *carry = best[(size_t)N - 1].score;
  return true;
}

// FATokenSegmentationTools_1best_bpe[_with_merges]_t.h:125-316 restricted to the starts of buf[a..b): arcs
// end before b; "token start" and "boundary" are read off the whole buffer, as the reference does.
bool bpe_range(const SegTables& S, const std::vector<int>& buf, int a, int b, int unk, std::vector<int>* res) {
  const int N = (int)buf.size();
  const bool merges = S.tok_algo == kTokenizeBpeOptWithMerges;
  const bool fast = merges || S.tok_algo == kTokenizeBpeOpt;
  struct Arc { int start, end, id; float rank; };
  std::vector<Arc> arcs;
  for (int start = a; start < b; ++start) {
    uint32_t q = S.root; int sum = 0; bool unknown = true;
    const bool tok_start = buf[start] == kSpDelim;
    const size_t cnt0 = arcs.size(); int ff = start;
    for (int i = start; i < b; ++i) {
      int ow; bool fin;
      if (!step(S, &q, buf[i], &ow, &fin)) break;
      sum += ow;
      if (fin) {
        if (sum < 0 || sum >= (int)S.info.size()) return false;
        const SegInfo si = S.info[sum];
        const bool opt = fast && tok_start && ((i < N - 1) ? buf[i + 1] == kSpDelim : true) && cnt0 < arcs.size();
        const Arc arc{start, i, si.id, merges ? si.score : 0.0f};
        if (!opt) arcs.push_back(arc); else { arcs[cnt0] = arc; arcs.resize(cnt0 + 1); ff = i; }
        unknown = false;
      }
    }
    if (unknown) {
      if (!arcs.empty() && arcs.back().id == unk) arcs.back().end = start;
      else arcs.push_back(Arc{start, start, unk, 0.0f});
    }
    if (fast) start = ff;
  }
  std::sort(arcs.begin(), arcs.end(), [&](const Arc& x, const Arc& y) {
    if (merges) { if (x.rank > y.rank) return true; if (x.rank < y.rank) return false; }
    if (x.id != y.id) return x.id < y.id;
    return x.start < y.start;
  });
  const int L = b - a;
  std::vector<int> tos((size_t)L), tid((size_t)L, unk);
  std::vector<uint8_t> inter((size_t)L + 1, 0);
  for (int i = 0; i < L; ++i) tos[i] = i;
  for (const Arc& x : arcs)
    if (!inter[x.start - a] && (x.end + 1 == b || !inter[x.end + 1 - a])) {
      tos[x.start - a] = x.end - a; tid[x.start - a] = x.id;
      for (int j = x.start + 1; j <= x.end; ++j) inter[j - a] = 1;
    }
  for (int s2 = 0; s2 < L; ++s2) { res->push_back(tid[s2]); s2 = tos[s2]; }
  return true;
}

int emit(const SegTables& S, const std::vector<int>& res, int32_t* ids, int max_ids) {
  int out = 0;
  for (size_t k = 0; k < res.size() && out < max_ids; ++k) ids[out++] = res[k] + S.id_offset;
  return out;
}

}  // namespace

extern "C" {

// The reference's whole-document algorithm over the flattened tables.
int sptwin_text_to_ids(void* h, const char* s, int n, int32_t* ids, int max_ids, int unk) {
  SpTwin* t = (SpTwin*)h;
  if (!t->err.empty()) return -2;
  const SegTables& S = t->S;
  if (n <= 0 || n > 1000000000 || !s) return 0;
  std::vector<int> buf, res;
  if (!front_end(S, s, n, &buf)) return 0;
  const bool bpe = S.tok_algo == kTokenizeBpe || S.tok_algo == kTokenizeBpeOpt || S.tok_algo == kTokenizeBpeOptWithMerges;
  double carry = 0.0;
  if (!(bpe ? bpe_range(S, buf, 0, (int)buf.size(), unk, &res) : unigram_range(S, buf, 0, (int)buf.size(), &carry, unk, &res))) return -3;
  return emit(S, res, ids, max_ids);
}

// The decompositions the streaming kernels rely on, run sequentially (sp_kernel.cu):
//   Unigram: the document in windows of `window` symbols cut at the last U+2581, the best score of the
//            last position carried over (a U+2581 is a forced token boundary when "U+2581" is a token
//            and no token contains it past its first symbol);
//   BPE:     every U+2581-delimited segment on its own, and a segment without unknown symbols split at
//            every position no token spans.
// -4: a window without a cut, -5: the model does not allow the decomposition.
int sptwin_text_to_ids_streamed(void* h, const char* s, int n, int32_t* ids, int max_ids, int unk, int window) {
  SpTwin* t = (SpTwin*)h;
  if (!t->err.empty()) return -2;
  const SegTables& S = t->S;
  if (n <= 0 || n > 1000000000 || !s) return 0;
  if (S.delim_inside_tokens) return -5;
  std::vector<int> buf, res;
  if (!front_end(S, s, n, &buf)) return 0;
  const int N = (int)buf.size();
  const bool bpe = S.tok_algo == kTokenizeBpe || S.tok_algo == kTokenizeBpeOpt || S.tok_algo == kTokenizeBpeOptWithMerges;
  if (!bpe) {
    if (!S.delim_is_token) return -5;
    double carry = 0.0;
    for (int pos = 0; pos < N;) {
      int cut = N;
      if (pos + window < N) {
        cut = -1;
        const int end = pos + window;
        for (int p = end - 1; p > pos; --p) if (buf[p] == kSpDelim) { cut = p; break; }
        if (cut < 0) {
          // no U+2581 in the window: the last position no token spans, not inside an unknown run, with
          // every earlier start walked to its end inside the window (sp_kernel.cu unigram_streamed)
          int reach = -1; bool prev_unknown = false;
          for (int p = pos; p < end; ++p) {
            uint32_t q = S.root; int far = -1; bool open = false;
            int i = p;
            for (; i < end; ++i) { int ow; bool fin; if (!step(S, &q, buf[i], &ow, &fin)) break; if (fin) far = i; if (q == 0) break; }
            if (i == end) open = true;
            const bool unknown = far < 0;
            if (p > pos && reach < p && !(unknown && prev_unknown)) cut = p;
            if (open) break;
            reach = std::max(reach, unknown ? p : far);
            prev_unknown = unknown;
          }
          if (cut < 0) return -4;
        }
      }
      if (!unigram_range(S, buf, pos, cut, &carry, unk, &res)) return -3;
      pos = cut;
    }
  } else {
    for (int a = 0; a < N;) {
      int b = a + 1;
      while (b < N && buf[b] != kSpDelim) ++b;
      // farthest token end from every start; a start without a token = an unknown symbol
      std::vector<int> fe((size_t)(b - a));
      bool unknown = false;
      for (int st = a; st < b; ++st) {
        uint32_t q = S.root; int far = -1;
        for (int i = st; i < b; ++i) { int ow; bool fin; if (!step(S, &q, buf[i], &ow, &fin)) break; if (fin) far = i; }
        if (far < 0) unknown = true;
        fe[(size_t)(st - a)] = far;
      }
      if (unknown) { if (!bpe_range(S, buf, a, b, unk, &res)) return -3; }
      else {
        int piece = a, reach = -1;
        for (int p = a; p < b; ++p) {
          if (p > a && reach < p) { if (!bpe_range(S, buf, piece, p, unk, &res)) return -3; piece = p; }
          reach = std::max(reach, fe[(size_t)(p - a)]);
        }
        if (!bpe_range(S, buf, piece, b, unk, &res)) return -3;
      }
      a = b;
    }
  }
  return emit(S, res, ids, max_ids);
}

}  // extern "C"
