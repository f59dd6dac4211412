// sp_kernel.cu -- TextToIdsWithOffsets_sp (blingfiretokdll.cpp:1349-1535) on the GPU for the
// [pos-dict] models: the shared front end (dummy prefix, UTF-8 decode or raw bytes, FANormalize,
// whitespace -> U+2581 collapse), then
//   * Unigram-LM best path  FATokenSegmentationTools_1best_t::Process        (.h:174-279)
//   * BPE / BPE-opt / BPE-opt-with-merges  FATokenSegmentationTools_1best_bpe[_with_merges]_t::Process (.h:125-316)
// over the Mealy MPH automaton laid out as a double-array (seg_tables.h): one 16-byte gather per
// GetDestOw.  Integer work plus fp64 adds in the reference's operand order (no FMA: there is no
// multiply to contract).
//
// One document per warp, persistent grid, documents handed out by an atomic counter.  Three paths
// (DESIGN.md section 4), each behind the other:
//   sp_unigram_fast  Unigram models with tokens <= 16 symbols: a 576-symbol window in shared memory (the
//                    whole document when it fits, else streamed and cut at U+2581 with the best score
//                    carried over), the relaxation in a register window (lane = position)
//   sp_bpe_fast      byte-level BPE models: a 512-symbol window slides over the document, cut at
//                    U+2581 (segments are independent); lanes re-dealt per phase (per segment, per
//                    start, per arc); long segments split at the positions no token spans
//   sp_doc_generic   everything else (any length, any token length, byte offsets): the document in
//                    the warp's global arena; Unigram with an arc tile in start order, BPE with the
//                    reference's arc vector, bitonic sort and sequential claim
#include "sp_kernel.cuh"

#include <cfloat>

#include "utf8_warp.cuh"

namespace bfb200 {

namespace {

constexpr int kTileArcs = 1024;           // arc slots of one tile of start positions (Unigram)
constexpr int kArcsPerSym = 8;            // warp-private BPE arc scratch, per symbol of capacity

struct Arc3 { int start, end, id; float rank; };   // 16 B

// per-warp workspace (shared memory or arena) for documents of up to `cap` symbols
struct Work {
  int32_t* sym;        // [cap+2] symbols (code points or bytes); later reused for ids
  int32_t* tmp;        // [cap+2] staging; Unigram: begin[]; BPE: ids_at[]
  double* score;       // [cap]   Unigram best score; BPE: segment-start list (int32 view)
  int32_t* bid;        // [cap]   Unigram best id; BPE: first-arc index, then tos[]
  uint8_t* flag;       // [cap+4] token-start marks / BPE intermediate[]
  int2* tile;          // [kTileArcs] Unigram arc tile {end, key}
  int32_t* boff_a;     // [cap+2] byte offset of every symbol (offsets requested; arena only)
  int32_t* boff_b;     // [cap+2] its staging twin; later the end position of the token that starts here
  int cap;
};

__host__ __device__ inline int64_t align16(int64_t v) { return (v + 15) & ~(int64_t)15; }
__host__ __device__ inline int64_t work_bytes(int cap) {
  return align16(4ll * (cap + 2)) * 2 + align16(8ll * cap) + align16(4ll * cap) + align16(cap + 4) + align16(8ll * kTileArcs);
}
// the arena variant also carries the two offset arrays
__host__ __device__ inline int64_t work_bytes_arena(int cap) { return work_bytes(cap) + 2 * align16(4ll * (cap + 2)); }
__device__ inline Work make_work(uint8_t* base, int cap, bool with_offsets) {
  Work w; int64_t o = 0;
  w.sym = (int32_t*)(base + o); o += align16(4ll * (cap + 2));
  w.tmp = (int32_t*)(base + o); o += align16(4ll * (cap + 2));
  w.score = (double*)(base + o); o += align16(8ll * cap);
  w.bid = (int32_t*)(base + o); o += align16(4ll * cap);
  w.flag = base + o; o += align16(cap + 4);
  w.tile = (int2*)(base + o); o += align16(8ll * kTileArcs);
  w.boff_a = with_offsets ? (int32_t*)(base + o) : nullptr; o += align16(4ll * (cap + 2));
  w.boff_b = with_offsets ? (int32_t*)(base + o) : nullptr;
  w.cap = cap;
  return w;
}

__device__ __forceinline__ bool sp_is_white(int c) {   // blingfiretokdll.h:17-21
  return c <= 0x20 || c == 0xa0 || (c >= 0x2000 && c <= 0x200f) || c == 0x202f || c == 0x205f || c == 0x2060 ||
         c == 0x2420 || c == 0x2424 || c == 0x3000 || c == 0xfeff;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// GetDestOw on the double-array (FAMealyDfa_pack_triv.cpp:69-244); q is the state's base
__device__ __forceinline__ bool da_step(const SpModelDev& m, uint32_t& q, int c, int& ow, bool& fin) {
  if ((unsigned)c > 0x10FFFFu) return false;
  const uint16_t s = __ldg(m.sym_of_cp + c);
  if (s == kNoSym) return false;
  const uint4 e = __ldg(reinterpret_cast<const uint4*>(m.da) + ((size_t)q + s));
  if (e.x != q) return false;
  ow = (int)e.z; fin = (e.y & kDaFinalBit) != 0; q = e.y & ~kDaFinalBit;
  return true;
}

// I2Info row (FAMultiMap_pack_fixed.cpp:140-162); an unusable key yields {unk, 0}
__device__ __forceinline__ void sp_info(const SpModelDev& m, int key, int unk, int& id, float& score) {
  id = unk; score = 0.0f;
  if (key >= 0 && key < m.info_count) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(m.info) + key);
    id = v.x; score = __int_as_float(v.y);
  }
}

// ---- front end: counts (store=false) or writes (store=true) the raw symbol stream -----------
// Returns the number of raw symbols incl. the dummy prefix, or -1 on invalid UTF-8 / no symbols.
// blingfiretokdll.cpp:1372-1412
__device__ int sp_raw_symbols(const SpModelDev& m, const uint8_t* text, int64_t lo0, int64_t hi, int64_t padded_bytes,
                              int32_t* out, int32_t* boff, bool store, int lane) {
  int64_t lo = lo0;
  if (hi - lo >= 3) {
    const uint32_t b0 = __ldg(text + lo), b1 = __ldg(text + lo + 1), b2 = __ldg(text + lo + 2);
    if (b0 == 0xEF && b1 == 0xBB && b2 == 0xBF) lo += 3;   // both decoders skip the BOM
  }
  const int off = m.no_dummy_prefix ? 0 : 1;
  if (store && off && lane == 0) { out[0] = kSpDelim; if (boff) boff[0] = -1; }   // :1372,:1387
  int cnt = 0;
  if (m.use_raw_bytes) {                                    // FAStrUtf8AsBytesToArray
    if (store) for (int64_t p = lo + lane; p < hi; p += 32) { out[off + (p - lo)] = (int)__ldg(text + p); if (boff) boff[off + (p - lo)] = (int)(p - lo0); }
    cnt = (int)(hi - lo);
  } else {                                                  // FAStrUtf8ToArray
    const uint32_t* text32 = reinterpret_cast<const uint32_t*>(text);
    unsigned bad = 0, sumlen = 0;
    for (int64_t bpos = lo; bpos < hi;) {
      const int64_t bs = bpos & ~(int64_t)3;
      const int64_t pos0 = bs + lane * 4;
      uint32_t w0, w1;
      utf8_load_words(text32, pos0, padded_bytes, &w0, &w1);
      const Utf8Lane d = utf8_decode_lane(w0, w1, pos0, bpos, hi);
      bad |= d.bad; sumlen += d.sumlen;
      const int c = __popc(d.start_mask);
      const int incl = warp_incl_scan(c, lane);
      if (store) {
        int idx = off + cnt + incl - c;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (d.start_mask & (1u << k)) { if (boff) boff[idx] = (int)(pos0 + k - lo0); out[idx++] = (int)d.cp[k]; }
      }
      cnt += __shfl_sync(0xffffffffu, incl, 31);
      bpos = bs + 128;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sumlen += __shfl_xor_sync(0xffffffffu, sumlen, o);
    if (__any_sync(0xffffffffu, bad != 0) || (int64_t)sumlen != hi - lo) return -1;
  }
  if (cnt <= 0) return -1;                                  // BuffSize <= 0 (:1409)
  return cnt + off;
}

// FANormalize over src[0..n) (FAUtils_cl.h:311-369): returns the normalized length; writes dst when given.
__device__ int sp_normalize(const SpModelDev& m, const int32_t* src, int n, int32_t* dst, int lane,
                            const int32_t* boff_src = nullptr, int32_t* boff_dst = nullptr) {
  int total = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    int c = 0, cp = 0; unsigned nc = 0xFF;
    if (i < n) {
      cp = src[i];
      nc = ((unsigned)cp <= 0x10FFFFu) ? (unsigned)__ldg(m.norm_count + cp) : 0xFFu;
      c = nc == 0xFF ? 1 : (int)nc;
    }
    const int incl = warp_incl_scan(c, lane);
    if (dst && i < n) {
      const int o = total + incl - c;
      if (nc == 0xFF) dst[o] = cp;
      else { const uint32_t f = __ldg(m.norm_first + cp); for (int k = 0; k < c; ++k) dst[o + k] = __ldg(m.norm_values + f + k); }
      if (boff_dst) for (int k = 0; k < c; ++k) boff_dst[o + k] = boff_src[i];   // pNormOffsets composed with pOffsets
    }
    total += __shfl_sync(0xffffffffu, incl, 31);
  }
  return total;
}

// whitespace runs -> one U+2581, one trailing U+2581 dropped (blingfiretokdll.cpp:1462-1496).
// A white symbol is kept iff the previous OUTPUT symbol is not U+2581, which is equivalent to
//   i == 0  ||  (src[i-1] is not white && src[i-1] != U+2581).
__device__ int sp_collapse(const int32_t* src, int n, int32_t* dst, int lane, const int32_t* boff_src = nullptr,
                           int32_t* boff_dst = nullptr) {
  int total = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    bool keep = false; int c = 0;
    if (i < n) {
      c = src[i];
      const bool w = sp_is_white(c);
      if (!w || i == 0) keep = true;
      else { const int p = src[i - 1]; keep = !sp_is_white(p) && p != kSpDelim; }
      if (w) c = kSpDelim;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) { const int o = total + __popc(bal & bf_lanemask_lt()); dst[o] = c; if (boff_dst) boff_dst[o] = boff_src[i]; }
    total += __popc(bal);
  }
  __syncwarp();
  if (total > 1 && dst[total - 1] == kSpDelim) --total;     // :1491-1493
  return total;
}

// FAUtf8Size of a lead byte (FAUtf8Utils.cpp:23-42)
__device__ __forceinline__ int sp_utf8_size_of_lead(unsigned ch) {
  if ((ch & 0x80) == 0) return 1;
  if ((ch & 0xE0) == 0xC0) return 2;
  if ((ch & 0xF0) == 0xE0) return 3;
  if ((ch & 0xF8) == 0xF0) return 4;
  return 0;
}

// where the offsets of one document go (blingfiretokdll.cpp:1519-1529); boff == nullptr: ids only
struct OffsetsOut {
  const int32_t* boff;   // byte offset (from the document start) of every final symbol, -1 = dummy prefix
  const uint8_t* doc;    // first byte of the document
  int32_t* starts;       // rows parallel to the ids row
  int32_t* ends;
};

// ordered emission of the tokens marked in w.flag (bit 1) with ids taken from idsrc[]; with offsets,
// the token that starts at q ends at symbol w.bid[q]
__device__ int sp_emit(const Work& w, const int32_t* idsrc, int N, int32_t* row, int max_ids, int unk, int id_offset,
                       bool map_unknown, int lane, const OffsetsOut& oo) {
  int out = 0;
  for (int p0 = 0; p0 < N; p0 += 32) {
    const int q = p0 + lane;
    const bool f = q < N && (w.flag[q] & 2);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int rank = out + __popc(bal & bf_lanemask_lt());
    if (f && rank < max_ids) {
      int id = idsrc[q];
      if (map_unknown && id == -1) id = unk;
      row[rank] = id + id_offset;                           // ids[k] = id + IdOffset, UNK included (:1516)
      if (oo.boff) {
        oo.starts[rank] = oo.boff[q];
        const int to_off = oo.boff[w.bid[q]];
        // a token that is only the dummy prefix has to_off == -1: the reference then sizes the byte
        // BEFORE the input (:1527, out of bounds); pinned to size 0 like the oracle does
        const int cs = to_off < 0 ? 0 : sp_utf8_size_of_lead(oo.doc[to_off]);
        oo.ends[rank] = to_off + (cs > 0 ? cs - 1 : 0);
      }
    }
    out += __popc(bal);
  }
  return out < max_ids ? out : max_ids;
}

// =====================================================================================
// Unigram-LM best path (FATokenSegmentationTools_1best_t.h:174-279)
// =====================================================================================
__device__ int sp_unigram(const SpModelDev& m, Work& w, int N, int32_t* row, int max_ids, int unk, int lane,
                          const OffsetsOut& oo) {
  int32_t* begin = w.tmp;
  for (int i = lane; i < N; i += 32) { w.score[i] = -(double)FLT_MAX; w.bid[i] = -1; begin[i] = -1; w.flag[i] = 0; }
  __syncwarp();
  int per = m.max_arc_len < 1 ? 1 : m.max_arc_len;          // an upper bound of the arcs of one start
  if (per > kTileArcs) per = kTileArcs;
  int S = kTileArcs / per; if (S > 32) S = 32;
  for (int t0 = 0; t0 < N; t0 += S) {
    // ---- phase A: lane l enumerates the arcs of start t0 + l (:196-224) ----
    const int start = t0 + lane;
    int narc = 0;
    if (lane < S && start < N) {
      uint32_t q = m.root; int sum = 0;
      for (int i = start; i < N; ++i) {
        int ow; bool fin;
        if (!da_step(m, q, w.sym[i], ow, fin)) break;
        sum += ow;
        if (fin && narc < per) { w.tile[lane * per + narc] = make_int2(i, sum); ++narc; }
        if (q == 0) break;                                   // a leaf: every further step fails
      }
    }
    __syncwarp();
    // ---- phase B: relax in start order; the arcs of one start end at distinct positions ----
    const int ns = min(S, N - t0);
    for (int l = 0; l < ns; ++l) {
      const int st = t0 + l;
      const int cnt = __shfl_sync(0xffffffffu, narc, l);
      const double prev = st > 0 ? w.score[st - 1] : 0.0;
      if (cnt > 0) {
        for (int k = lane; k < cnt; k += 32) {               // AddArc (:118-142)
          const int2 a = w.tile[l * per + k];
          int id; float sc;
          sp_info(m, a.y, -1, id, sc);
          const double cand = (double)sc + prev;
          if (w.score[a.x] < cand) { begin[a.x] = st; w.bid[a.x] = id; w.score[a.x] = cand; }
        }
      } else if (lane == 0) {                                // AddUnknownArc (:145-171)
        const double cand = (double)(-100000.0f) + prev;
        if (w.score[st] < cand) {
          begin[st] = st; w.bid[st] = -1; w.score[st] = cand;
          if (st > 0 && w.bid[st - 1] == -1) begin[st] = begin[st - 1];
        }
      }
      __syncwarp();
    }
  }
  // ---- back-trace (:227-257): mark token starts, move each token's id to its start slot ----
  if (lane == 0) {
    int end = N - 1;
    while (end >= 0) {
      const int b = begin[end];
      const int id = w.bid[end];
      if (b < 0) { w.flag[0] |= 2; w.sym[0] = id; w.bid[0] = end; break; }   // never-set arc: the reference emits it first and stops
      w.flag[b] |= 2;
      w.sym[b] = id;                                         // symbols before `end` are not read again
      w.bid[b] = end;                                        // ... nor are the ids at or before b: keep the token's end
      end = b - 1;
    }
  }
  __syncwarp();
  return sp_emit(w, w.sym, N, row, max_ids, unk, m.id_offset, true, lane, oo);
}

// =====================================================================================
// BPE family (FATokenSegmentationTools_1best_bpe_t.h:125-316, ..._with_merges_t.h)
// =====================================================================================
__device__ __forceinline__ bool arc_less(const Arc3& a, const Arc3& b, bool merges) {
  if (merges) {                                              // ..._with_merges_t.h:242-262: bigger ranks first
    if (a.rank > b.rank) return true;
    if (a.rank < b.rank) return false;
  }
  if (a.id != b.id) return a.id < b.id;                      // ..._bpe_t.h:238-255
  return a.start < b.start;
}

__device__ __forceinline__ int count_arcs_from(const SpModelDev& m, const Work& w, int s, int b) {
  uint32_t q = m.root; int cnt = 0;
  for (int i = s; i < b; ++i) {
    int ow; bool fin;
    if (!da_step(m, q, w.sym[i], ow, fin)) break;
    if (fin) ++cnt;
    if (q == 0) break;
  }
  return cnt;
}

// One segment [a, b), warp-cooperatively.  `arcs` is warp-private scratch of arc_cap entries.
// Tokens are written position-indexed: ids_at[start], w.flag[start] |= 2.  false = scratch overflow.
struct ArcScratch { Arc3* priv; int64_t priv_cap; Arc3* ovf; int64_t ovf_cap; int* lock; };

__device__ bool bpe_segment(const SpModelDev& m, Work& w, int N, int a, int b, int unk, const ArcScratch& scratch,
                            int32_t* ids_at, int lane, bool fast, bool merges) {
  // ---- arcs of every start, grouped by start (count -> scan -> write) ----
  int total = 0;
  for (int s0 = a; s0 < b; s0 += 32) {
    const int s = s0 + lane;
    const int cnt = s < b ? count_arcs_from(m, w, s, b) : 0;
    const int incl = warp_incl_scan(cnt, lane);
    if (s < b) w.bid[s] = total + incl - cnt;
    total += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  const int L = b - a;
  // raw arcs + the reference's arc vector (<= total + L entries) padded to a power of two
  const int64_t need = (int64_t)total + 2ll * ((int64_t)total + L) + 4;
  Arc3* arcs = scratch.priv;
  bool locked = false;
  if (need > scratch.priv_cap) {
    if (need > scratch.ovf_cap) return false;
    if (lane == 0) { while (atomicCAS(scratch.lock, 0, 1) != 0) __nanosleep(200); __threadfence(); }
    __syncwarp();
    arcs = scratch.ovf;
    locked = true;
  }
  for (int s0 = a; s0 < b; s0 += 32) {
    const int s = s0 + lane;
    if (s < b) {
      int wr = w.bid[s];
      uint32_t q = m.root; int sum = 0;
      for (int i = s; i < b; ++i) {
        int ow; bool fin;
        if (!da_step(m, q, w.sym[i], ow, fin)) break;
        sum += ow;
        if (fin) {
          Arc3 A; A.start = s; A.end = i;
          sp_info(m, sum, unk, A.id, A.rank);
          if (!merges) A.rank = 0.0f;
          arcs[wr++] = A;
        }
        if (q == 0) break;
      }
    }
  }
  __syncwarp();
  // ---- the reference's arc vector for this segment: bpe-opt at the start, unknown runs (:188-230) ----
  Arc3* fin_arcs = arcs + total;
  int nfin = 0;
  if (lane == 0) {
    for (int s = a; s < b; ++s) {
      const int first = w.bid[s];
      const int cnt = ((s + 1 < b) ? w.bid[s + 1] : total) - first;
      const bool tok_start = w.sym[s] == kSpDelim;
      const int cnt0 = nfin;
      int ff = s;
      for (int k = 0; k < cnt; ++k) {
        const Arc3 A = arcs[first + k];
        const bool boundary = (A.end < N - 1) ? (w.sym[A.end + 1] == kSpDelim) : true;
        if (fast && tok_start && boundary && cnt0 < nfin) { fin_arcs[cnt0] = A; nfin = cnt0 + 1; ff = A.end; }
        else fin_arcs[nfin++] = A;
      }
      if (cnt == 0) {
        if (nfin > 0 && fin_arcs[nfin - 1].id == unk) fin_arcs[nfin - 1].end = s;   // compares ids (:219-225)
        else { Arc3 U; U.start = s; U.end = s; U.id = unk; U.rank = 0.0f; fin_arcs[nfin++] = U; }
      }
      if (fast) s = ff;
    }
  }
  nfin = __shfl_sync(0xffffffffu, nfin, 0);
  __syncwarp();
  // ---- sort (:238-262): bitonic network over a power-of-two padded copy ----
  int P = 1; while (P < nfin) P <<= 1;
  for (int i = nfin + lane; i < P; i += 32) { Arc3 Z; Z.start = 0x7fffffff; Z.end = 0; Z.id = 0x7fffffff; Z.rank = -FLT_MAX; fin_arcs[i] = Z; }
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const Arc3 x = fin_arcs[i], y = fin_arcs[l];
          const bool up = (i & k) == 0;
          if (up ? arc_less(y, x, merges) : arc_less(x, y, merges)) { fin_arcs[i] = y; fin_arcs[l] = x; }
        }
      }
      __syncwarp();
    }
  }
  // ---- greedy claim in sorted order (:264-296); intermediate[] = bit 0 of w.flag, tos[] = w.bid ----
  for (int i = a + lane; i < b; i += 32) { w.flag[i] = 0; w.bid[i] = i; ids_at[i] = unk; }
  __syncwarp();
  if (lane == 0) {
    for (int k = 0; k < nfin; ++k) {
      const Arc3 A = fin_arcs[k];
      // position b starts the next segment: no arc crosses it, so it is never an intermediate
      const bool end_free = (A.end + 1 >= b) || (w.flag[A.end + 1] & 1) == 0;
      if ((w.flag[A.start] & 1) == 0 && end_free) {
        w.bid[A.start] = A.end; ids_at[A.start] = A.id;
        for (int j = A.start + 1; j <= A.end; ++j) w.flag[j] |= 1;
      }
    }
    // tokens: follow tos[] from the segment start (:299-313).  (tos[] starts as the identity; the
    // reference's zero-initialised tos[] would loop forever on an unclaimed start, which a
    // vocabulary with all single symbols never produces.)
    for (int s = a; s < b; ++s) { w.flag[s] |= 2; s = w.bid[s]; }
  }
  __syncwarp();
  if (locked && lane == 0) { __threadfence(); atomicExch(scratch.lock, 0); }
  return true;
}

__device__ int sp_bpe(const SpModelDev& m, Work& w, int N, int32_t* row, int max_ids, int unk, const ArcScratch& scratch,
                      int lane, bool* overflow, const OffsetsOut& oo) {
  const bool merges = m.tok_algo == kTokenizeBpeOptWithMerges;
  const bool fast = merges || m.tok_algo == kTokenizeBpeOpt;
  int32_t* ids_at = w.tmp;
  int32_t* seg = reinterpret_cast<int32_t*>(w.score);       // segment starts (w.score is unused by BPE)
  int nseg = 0;
  for (int p0 = 0; p0 < N; p0 += 32) {
    const int p = p0 + lane;
    const bool f = p < N && (p == 0 || (!m.delim_inside_tokens && w.sym[p] == kSpDelim));
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (f) seg[nseg + __popc(bal & bf_lanemask_lt())] = p;
    nseg += __popc(bal);
    if (p < N) w.flag[p] = 0;
  }
  __syncwarp();
  for (int g0 = 0; g0 < nseg; g0 += 32) {
    // ---- one lane per segment: the bpe-opt whole-word shortcut ----
    // Walking from the segment start, an arc that ends exactly at the segment end after a shorter
    // arc was already seen makes the reference keep ONLY that arc and skip the interior starts
    // (:188-206,:228-230); a one-symbol segment with an arc is a single arc as well.
    const int g = g0 + lane;
    bool hard = false; int a = 0, b = 0;
    if (g < nseg) {
      a = seg[g]; b = g + 1 < nseg ? seg[g + 1] : N;
      uint32_t q = m.root; int sum = 0, narcs = 0, whole_key = -1; bool whole = false;
      for (int i = a; i < b; ++i) {
        int ow; bool fin;
        if (!da_step(m, q, w.sym[i], ow, fin)) break;
        sum += ow;
        if (fin) { if (i == b - 1 && (narcs > 0 || b - a == 1)) { whole = true; whole_key = sum; } ++narcs; }
        if (q == 0) break;
      }
      const bool tok_start = w.sym[a] == kSpDelim;
      if (whole && ((fast && tok_start) || b - a == 1)) {
        int id; float r;
        sp_info(m, whole_key, unk, id, r);
        ids_at[a] = id; w.flag[a] = 2; w.bid[a] = b - 1;     // tos[a]
      } else hard = true;
    }
    // ---- the other segments of this round, one at a time, warp-cooperatively ----
    unsigned hb = __ballot_sync(0xffffffffu, hard);
    while (hb) {
      const int l = __ffs(hb) - 1; hb &= hb - 1;
      const int sa = __shfl_sync(0xffffffffu, a, l), sb = __shfl_sync(0xffffffffu, b, l);
      if (!bpe_segment(m, w, N, sa, sb, unk, scratch, ids_at, lane, fast, merges)) { *overflow = true; return 0; }
    }
  }
  __syncwarp();
  return sp_emit(w, ids_at, N, row, max_ids, unk, m.id_offset, false, lane, oo);
}

// One document through the general path: any length, any token length, offsets if wanted.  The
// document lives in the warp's arena workspace `wa`.
template <bool kBpe>
__device__ int sp_doc_generic(const SpLaunch& p, const SpModelDev& m, Work& wa, const ArcScratch& scratch, int64_t doc,
                              int64_t lo, int64_t hi, int64_t padded_bytes, int lane, int* error_flag) {
  const bool want_offsets = p.starts != nullptr;              // offsets ride in the arena workspace only
  const int64_t n = hi - lo;
  int result = 0;
  {
    {
      const int nraw = sp_raw_symbols(m, p.text, lo, hi, padded_bytes, nullptr, nullptr, false, lane);
      bool ok = nraw > 0;
      const int64_t need = (m.norm_count ? 2 * (n + 1) : (int64_t)nraw) + 2;   // staging bound (:1423)
      if (ok && need > (int64_t)p.arena_cap) { ok = false; if (lane == 0) atomicExch(error_flag, 2); }
      if (ok) {
        Work& w = wa;
        int32_t* boff = w.boff_a; int32_t* boff_other = w.boff_b;           // nullptr unless offsets are wanted
        sp_raw_symbols(m, p.text, lo, hi, padded_bytes, w.sym, boff, true, lane);
        __syncwarp();
        int N = nraw;
        int32_t* cur = w.sym; int32_t* other = w.tmp;
        if (m.norm_count) {
          const int nn = sp_normalize(m, cur, N, nullptr, lane);
          if (nn <= 0 || (int64_t)nn > 2 * (n + 1)) ok = false;          // :1442-1446
          else {
            sp_normalize(m, cur, N, other, lane, boff, boff_other);
            __syncwarp();
            N = nn;
            int32_t* t = cur; cur = other; other = t;
            t = boff; boff = boff_other; boff_other = t;
          }
        }
        if (ok) {
          N = sp_collapse(cur, N, other, lane, boff, boff_other);
          __syncwarp();
          if (other != w.sym) { for (int i = lane; i < N; i += 32) w.sym[i] = other[i]; __syncwarp(); }
          if (N > 0) {
            int32_t* row = p.ids + doc * (int64_t)p.max_ids;
            OffsetsOut oo;
            oo.boff = want_offsets ? boff_other : nullptr;                 // sp_collapse wrote the final offsets there
            oo.doc = p.text + lo;
            oo.starts = want_offsets ? p.starts + doc * (int64_t)p.max_ids : nullptr;
            oo.ends = want_offsets ? p.ends + doc * (int64_t)p.max_ids : nullptr;
            if (kBpe) {
              bool overflow = false;
              result = sp_bpe(m, w, N, row, p.max_ids, p.unk_id, scratch, lane, &overflow, oo);
              if (overflow) { result = 0; if (lane == 0) atomicExch(error_flag, 3); }
            } else {
              result = sp_unigram(m, w, N, row, p.max_ids, p.unk_id, lane, oo);
            }
          }
        }
      }
    }
  }
  return result;
}

__device__ __forceinline__ ArcScratch make_scratch(const SpLaunch& p, uint8_t* my_arena, int* error_flag) {
  ArcScratch scratch;
  scratch.priv = reinterpret_cast<Arc3*>(my_arena + work_bytes_arena(p.arena_cap));
  scratch.priv_cap = (int64_t)p.arena_cap * kArcsPerSym + 4096;
  scratch.ovf = reinterpret_cast<Arc3*>(p.overflow);
  scratch.ovf_cap = p.overflow_cap;
  scratch.lock = error_flag + 1;
  return scratch;
}

constexpr int kUFallback = -2;   // the document does not fit a fast path: the caller takes sp_doc_generic

// =====================================================================================
// BPE streaming fast path (byte-level models: gpt2, roberta).  Tokens never contain U+2581 past
// their first symbol, so the U+2581-delimited segments (words) are independent, and the document
// never has to be resident: a 512-symbol window slides over it, cut at the last U+2581.
//   front end   128 bytes per step: whitespace -> U+2581, runs collapsed, alphabet indices
//   easy pass   one lane per segment: the bpe-opt whole-word shortcut (:188-206)
//   hard pass   one lane per remaining segment (<= 64 symbols, <= 64 arcs): arcs are inserted into a
//               lane-private sorted list as they are found -- the sort key (ordinal of (rank, id),
//               start, end) is ONE 32-bit integer (seg_tables.h) -- then the reference's greedy
//               claim (:264-296) with the intermediate[] marks in a 64-bit register
//   coop        bigger segments (<= 1024 arcs): the warp together, bitonic sort in shared memory
// A segment longer than the window, more arcs than that, or a symbol outside the alphabet sends the
// whole document to sp_doc_generic.
// =====================================================================================
constexpr int kBWarps = 8;                 // per CTA
constexpr int kBCtasPerSm = 2;
constexpr int kBWin = 512;                 // symbols in the window
constexpr int kBLaneArcs = 48;             // lane-serial segments: at most this many listed arcs ...
constexpr int kBMaxLen = 64;               // ... and symbols (intermediate[] is one 64-bit register)
constexpr int kBSplitLen = 24;             // longer segments are first split at the positions no token spans
constexpr int kBCoopArcs = 512;            // warp-cooperative segments: arcs sorted in the window's scratch (64-bit keys)
constexpr unsigned kBUnclaimed = 0xFFFFFu; // ordinal of "no arc claimed from this start"

struct BWork {
  uint32_t* scratch;    // [32][kBLaneArcs] lane-interleaved sorted keys; or kBCoopArcs 64-bit keys
  int32_t* ids_at;      // [kBWin] claim state {ordinal, tos}, then the token id, at token starts
  uint32_t* mark;       // [kBWin/32] bit p: a token starts at p
  uint16_t* sym;        // [kBWin] alphabet indices
  uint16_t* seg;        // [kBWin/2+8] segment starts, window-relative, and the end sentinel
  uint16_t* hard_a;     // [kBWin] the pieces the easy pass left: first symbol ...
  uint16_t* hard_b;     // [kBWin] ... and one past the last
  uint32_t* cnt;        // [32] arcs listed per slot; [32] = slots that cannot be served lane-serially
  uint8_t* order;       // [kBLaneArcs][32] list index of the arc of rank r
};
constexpr int kBWorkBytes = 4 * 32 * kBLaneArcs + 4 * kBWin + 4 * (kBWin / 32) + 2 * kBWin + 2 * (kBWin / 2 + 8) + 4 * kBWin + 4 * 36 + 32 * kBLaneArcs;
static_assert(kBWorkBytes % 16 == 0 && 8 * kBCoopArcs <= 4 * 32 * kBLaneArcs && kBLaneArcs <= 64 && kBMaxLen <= 64 && kBWin <= 1024, "workspace layout");

__device__ inline BWork make_bwork(uint8_t* b) {
  BWork w;
  w.scratch = (uint32_t*)b; b += 4 * 32 * kBLaneArcs;
  w.ids_at = (int32_t*)b; b += 4 * kBWin;
  w.mark = (uint32_t*)b; b += 4 * (kBWin / 32);
  w.sym = (uint16_t*)b; b += 2 * kBWin;
  w.seg = (uint16_t*)b; b += 2 * (kBWin / 2 + 8);
  w.hard_a = (uint16_t*)b; b += 2 * kBWin;
  w.hard_b = (uint16_t*)b; b += 2 * kBWin;
  w.cnt = (uint32_t*)b; b += 4 * 36;
  w.order = b;
  return w;
}

// GetDestOw on alphabet indices
__device__ __forceinline__ bool b_step(const uint4* da, uint32_t& q, uint16_t s, int& sum, bool& fin) {
  if (s == kNoSym) return false;
  const uint4 e = __ldg(da + ((size_t)q + s));
  if (e.x != q) return false;
  sum += (int)e.z; fin = (e.y & kDaFinalBit) != 0; q = e.y & ~kDaFinalBit;
  return true;
}

__device__ __forceinline__ int b_ord(const SpModelDev& m, int key) {
  return (key >= 0 && key < m.info_count) ? __ldg(m.bpe_ord + key) : -1;
}

// One segment [a, b) of the window, the warp together.  false: it does not fit (general path).
__device__ bool bpe_coop(const SpModelDev& m, const BWork& w, const ArcScratch& scratch, int a, int b, int unk, int lane) {
  const unsigned full = 0xffffffffu;
  const uint4* da = reinterpret_cast<const uint4*>(m.da);
  const int L = b - a;
  const bool sf = m.bpe_singles_first;
  // arcs of every start, grouped by start (count -> scan -> write); one-symbol arcs are not listed
  // when they sort first (see bpe_window)
  int total = 0; bool bad = false;
  for (int s0 = 0; s0 < L; s0 += 32) {
    const int s = s0 + lane; int cnt = 0;
    if (s < L) {
      uint32_t q = m.root; int sum = 0; bool any = false;
      for (int i = a + s; i < b; ++i) {
        bool fin;
        if (!b_step(da, q, w.sym[i], sum, fin)) break;
        if (fin) { any = true; if (!(sf && i == a + s)) ++cnt; }
        if (q == 0) break;
      }
      if (!any) bad = true;                                    // an unknown symbol run (:208-227)
    }
    const int incl = warp_incl_scan(cnt, lane);
    if (s < L) w.ids_at[a + s] = total + incl - cnt;
    total += __shfl_sync(full, incl, 31);
  }
  if (__any_sync(full, bad)) return false;
  // keys in the window's scratch, or -- the rare big case -- in the warp's arena
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(w.scratch);
  int P = 1; while (P < total) P <<= 1;
  if (P > kBCoopArcs) {                                        // (P is a power of two)
    if ((int64_t)P > 2 * scratch.priv_cap) return false;
    keys = reinterpret_cast<unsigned long long*>(scratch.priv);
  }
  __syncwarp();
  for (int s0 = 0; s0 < L; s0 += 32) {
    const int s = s0 + lane;
    if (s < L) {
      int wr = w.ids_at[a + s];
      unsigned init = (kBUnclaimed << 10) | (unsigned)s;
      uint32_t q = m.root; int sum = 0;
      for (int i = a + s; i < b; ++i) {
        bool fin;
        if (!b_step(da, q, w.sym[i], sum, fin)) break;
        if (fin) {
          const int ord = b_ord(m, sum);
          if (ord < 0) bad = true;
          if (sf && i == a + s) init = ((unsigned)ord << 10) | (unsigned)s;
          else keys[wr++] = ((unsigned long long)(unsigned)ord << 20) | ((unsigned long long)s << 10) | (unsigned long long)(i - a);
        }
        if (q == 0) break;
      }
      w.ids_at[a + s] = (int)init;
    }
  }
  if (__any_sync(full, bad)) return false;
  for (int i = total + lane; i < P; i += 32) keys[i] = ~0ull;
  __syncwarp();
  for (int k = 2; k <= P; k <<= 1) {                           // (:238-262) as a bitonic network
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < P; i += 32) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = keys[i], y = keys[l];
          const bool up = (i & k) == 0;
          if (up ? (y < x) : (x < y)) { keys[i] = y; keys[l] = x; }
        }
      }
      __syncwarp();
    }
  }
  // greedy claim in sorted order (:264-296); lane i keeps bits 32i..32i+31 of intermediate[]
  unsigned inter = 0;
  for (int k = 0; k < total; ++k) {
    const unsigned long long key = keys[k];
    const int st = (int)(key >> 10) & 1023, en = (int)key & 1023;
    const unsigned ws = __shfl_sync(full, inter, st >> 5), we = __shfl_sync(full, inter, ((en + 1) >> 5) & 31);
    const bool end_free = (en + 1 >= L) || ((we >> ((en + 1) & 31)) & 1u) == 0;
    if (((ws >> (st & 31)) & 1u) == 0 && end_free) {
      if (lane == 0) w.ids_at[a + st] = (int)(((unsigned)(key >> 20) << 10) | (unsigned)en);
      const int lo = max(st + 1, lane * 32), hi = min(en, lane * 32 + 31);
      if (lo <= hi) inter |= ((2u << (hi & 31)) - 1u) & ~((1u << (lo & 31)) - 1u);
    }
  }
  __syncwarp();
  if (lane == 0) {                                             // tokens: follow tos[] (:299-313)
    for (int s = 0; s < L;) {
      const unsigned v = (unsigned)w.ids_at[a + s];
      const unsigned ord = v >> 10;
      w.ids_at[a + s] = ord == kBUnclaimed ? unk : __ldg(m.bpe_id_of_ord + ord);
      w.mark[(a + s) >> 5] |= 1u << ((a + s) & 31);
      s = (int)(v & 1023u) + 1;
    }
  }
  __syncwarp();
  return true;
}

// Farthest end (window position) of a token that starts at p and lies inside [p, limit); -1: none.
// *open: the walk ran into `limit` while it could still continue.
__device__ __forceinline__ int b_farthest(const SpModelDev& m, const uint4* da, const uint16_t* sym, int p, int limit, bool* open) {
  uint32_t q = m.root; int sum = 0, fe = -1;
  *open = false;
  int i = p;
  for (; i < limit; ++i) {
    bool fin;
    if (!b_step(da, q, sym[i], sum, fin)) break;
    if (fin) fe = i;
    if (q == 0) break;
  }
  if (i == limit) *open = true;
  return fe;
}

// The segments of sym[0..cut): tokens appended to row[out..).  Returns the new out (it may pass
// max_ids; nothing is written past it) or kUFallback.  open_ended: the last segment does not end at a
// U+2581 (nor at the end of the document) but at a position no token can span.
__device__ int bpe_window(const SpModelDev& m, const BWork& w, const ArcScratch& scratch, int cut, uint16_t delim, int32_t* row,
                          int out, int max_ids, int unk, bool fast, bool open_ended, int lane) {
  const unsigned full = 0xffffffffu;
  const uint4* da = reinterpret_cast<const uint4*>(m.da);
  int nseg = 0;
  for (int p0 = 0; p0 < cut; p0 += 32) {
    const int p = p0 + lane;
    const bool f = p < cut && (p == 0 || w.sym[p] == delim);
    const unsigned bal = __ballot_sync(full, f);
    if (f) w.seg[nseg + __popc(bal & bf_lanemask_lt())] = (uint16_t)p;
    nseg += __popc(bal);
  }
  if (lane == 0) w.seg[nseg] = (uint16_t)cut;
  for (int i = lane; i < kBWin / 32; i += 32) w.mark[i] = 0;
  __syncwarp();
  // ---- easy pass: the bpe-opt whole-word shortcut, one lane per segment ----
  // Walking from the segment start, an arc that ends exactly at the segment end after a shorter
  // arc was already seen makes the reference keep ONLY that arc and skip the interior starts
  // (:188-206,:228-230); a one-symbol segment with an arc is a single arc as well.
  int nhard = 0;
  for (int g0 = 0; g0 < nseg; g0 += 32) {
    const int g = g0 + lane;
    bool hard = false;
    int a = 0, b = 0;
    if (g < nseg && open_ended && g == nseg - 1) {             // its true end is not in the window: no shortcut
      a = w.seg[g]; b = w.seg[g + 1]; hard = true;
    } else if (g < nseg) {
      a = w.seg[g]; b = w.seg[g + 1];
      uint32_t q = m.root; int sum = 0, narcs = 0, whole_key = -1; bool whole = false;
      for (int i = a; i < b; ++i) {
        bool fin;
        if (!b_step(da, q, w.sym[i], sum, fin)) break;
        if (fin) { if (i == b - 1 && (narcs > 0 || b - a == 1)) { whole = true; whole_key = sum; } ++narcs; }
        if (q == 0) break;
      }
      if (whole && ((fast && w.sym[a] == delim) || b - a == 1)) {
        int id; float r;
        sp_info(m, whole_key, unk, id, r);
        w.ids_at[a] = id;
        atomicOr(&w.mark[a >> 5], 1u << (a & 31));
      } else hard = true;
    }
    const unsigned hb = __ballot_sync(full, hard);
    if (hard) { const int k = nhard + __popc(hb & bf_lanemask_lt()); w.hard_a[k] = (uint16_t)a; w.hard_b[k] = (uint16_t)b; }
    nhard += __popc(hb);
  }
  __syncwarp();
  // ---- split pass: a long segment falls apart at every position no token spans ----
  // No arc crosses such a position, so it is never marked intermediate and the claim tests on
  // either side never see the other side: the pieces are independent and each sorts only its own
  // arcs.  (Long whitespace-free runs are URLs and CJK/Thai text; they split every few bytes.)
  {
    const int nhard0 = nhard;
    for (int h0 = 0; h0 < nhard0; h0 += 32) {
      const int h = h0 + lane;
      unsigned lb = __ballot_sync(full, h < nhard0 && (int)w.hard_b[h] - (int)w.hard_a[h] > kBSplitLen);
      while (lb) {
        const int l = __ffs(lb) - 1; lb &= lb - 1;
        const int hh = h0 + l;
        const int sa = w.hard_a[hh], sb = w.hard_b[hh];
        const int n0 = nhard;
        int carry = -1; bool bad = false;
        for (int p0 = sa; p0 < sb; p0 += 32) {
          const int p = p0 + lane;
          int fe = -1;
          if (p < sb) { bool open; fe = b_farthest(m, da, w.sym, p, sb, &open); if (fe < 0) bad = true; }
          int incl = fe;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(full, incl, o); if (lane >= o) incl = max(incl, t); }
          int excl = __shfl_up_sync(full, incl, 1);
          excl = lane ? max(excl, carry) : carry;
          const bool piece = p < sb && p > sa && excl < p;     // (the segment start is piece 0: entry hh itself)
          const unsigned pb = __ballot_sync(full, piece);
          if (piece) w.hard_a[nhard + __popc(pb & bf_lanemask_lt())] = (uint16_t)p;
          nhard += __popc(pb);
          carry = max(carry, __shfl_sync(full, incl, 31));
        }
        if (__any_sync(full, bad)) return kUFallback;          // a symbol no token starts with: the general path
        __syncwarp();
        for (int k = n0 + lane; k < nhard; k += 32) w.hard_b[k] = k + 1 < nhard ? w.hard_a[k + 1] : (uint16_t)sb;
        if (lane == 0 && nhard > n0) w.hard_b[hh] = w.hard_a[n0];
        __syncwarp();
      }
    }
  }
  // ---- hard pass, 32 segments (slots) at a time; the lanes are re-dealt for every phase ----
  const bool sf = m.bpe_singles_first;
  for (int h0 = 0; h0 < nhard; h0 += 32) {
    const int h = h0 + lane;
    int a = 0, b = 0, L = 0; bool lane_ok = false;
    if (h < nhard) { a = w.hard_a[h]; b = w.hard_b[h]; L = b - a; lane_ok = L <= kBMaxLen; }
    const int preL = warp_incl_scan(lane_ok ? L : 0, lane);
    const int T = __shfl_sync(full, preL, 31);
    w.cnt[lane] = 0;
    if (lane == 0) w.cnt[32] = 0;                              // bit s: slot s cannot be served here
    __syncwarp();
    // phase 1, one lane per (slot, start): every arc of that start (:188-230) goes to the slot's
    // list.  ids_at[] becomes the claim state {ordinal, tos}: unclaimed, or -- when one-symbol
    // tokens sort first -- the one-symbol arc, which the claim loop would take before anything is
    // marked intermediate.
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      int slot = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) { const int v = __shfl_sync(full, preL, slot + step - 1); if (v <= t) slot += step; }
      const int sa = __shfl_sync(full, a, slot), sb = __shfl_sync(full, b, slot), sp = __shfl_sync(full, preL, slot);
      if (t < T) {
        const int s = t - (sp - (sb - sa));
        uint32_t q = m.root; int sum = 0, cnt = 0;
        unsigned init = (kBUnclaimed << 6) | (unsigned)s;
        for (int i = sa + s; i < sb; ++i) {
          bool fin;
          if (!b_step(da, q, w.sym[i], sum, fin)) break;
          if (fin) {
            const int ord = b_ord(m, sum);
            ++cnt;
            if (ord < 0) { cnt = 0; break; }
            if (sf && i == sa + s) init = ((unsigned)ord << 6) | (unsigned)s;
            else {
              const unsigned idx = atomicAdd(&w.cnt[slot], 1u);
              if (idx < (unsigned)kBLaneArcs) w.scratch[idx * 32 + slot] = ((uint32_t)ord << 12) | ((uint32_t)s << 6) | (uint32_t)(i - sa);
            }
          }
          if (q == 0) break;
        }
        if (cnt == 0) atomicOr(&w.cnt[32], 1u << slot);        // an unknown symbol run (or an unusable key): not here
        w.ids_at[sa + s] = (int)init;
      }
    }
    __syncwarp();
    // phase 2, one lane per (slot, arc): its rank in the order (:238-262) -- ordinal of (rank, id),
    // then start: one integer compare; keys are distinct, so the ranks are a permutation
    const int A = (int)w.cnt[lane];
    const bool ok = lane_ok && A <= kBLaneArcs && ((w.cnt[32] >> lane) & 1u) == 0;
    const int Aw = ok ? A : 0;
    const int preA = warp_incl_scan(Aw, lane);
    const int U = __shfl_sync(full, preA, 31);
    for (int u0 = 0; u0 < U; u0 += 32) {
      const int u = u0 + lane;
      int slot = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) { const int v = __shfl_sync(full, preA, slot + step - 1); if (v <= u) slot += step; }
      const int sA = __shfl_sync(full, Aw, slot), spA = __shfl_sync(full, preA, slot);
      if (u < U) {
        const int idx = u - (spA - sA);
        const uint32_t key = w.scratch[idx * 32 + slot];
        int rank = 0;
        for (int k = 0; k < sA; ++k) rank += w.scratch[k * 32 + slot] < key;
        w.order[rank * 32 + slot] = (uint8_t)idx;
      }
    }
    __syncwarp();
    // phases 3 and 4, one lane per slot: greedy claim in that order (:264-296) with intermediate[]
    // in a register, then the tokens by following tos[] (:299-313)
    if (ok) {
      unsigned long long inter = 0;
      for (int r = 0; r < A; ++r) {
        const uint32_t key = w.scratch[(int)w.order[r * 32 + lane] * 32 + lane];
        const int st = (int)(key >> 6) & 63, en = (int)key & 63;
        const bool end_free = (en + 1 >= L) || ((inter >> (en + 1)) & 1ull) == 0;
        if (((inter >> st) & 1ull) == 0 && end_free) {
          w.ids_at[a + st] = (int)(((key >> 12) << 6) | (unsigned)en);
          inter |= ((2ull << en) - 1ull) & ~((2ull << st) - 1ull);
        }
      }
      for (int s = 0; s < L;) {
        const unsigned v = (unsigned)w.ids_at[a + s];
        const unsigned ord = v >> 6;
        w.ids_at[a + s] = ord == kBUnclaimed ? unk : __ldg(m.bpe_id_of_ord + ord);
        atomicOr(&w.mark[(a + s) >> 5], 1u << ((a + s) & 31));
        s = (int)(v & 63u) + 1;
      }
    }
    unsigned cb = __ballot_sync(full, h < nhard && !ok);
    while (cb) {
      const int l = __ffs(cb) - 1; cb &= cb - 1;
      const int sa = __shfl_sync(full, a, l), sb = __shfl_sync(full, b, l);
      if (!bpe_coop(m, w, scratch, sa, sb, unk, lane)) return kUFallback;
    }
  }
  __syncwarp();
  // ---- ordered emission ----
  for (int p0 = 0; p0 < cut && out < max_ids; p0 += 32) {
    const uint32_t word = w.mark[p0 >> 5];
    const int rank = out + __popc(word & bf_lanemask_lt());
    if (((word >> lane) & 1u) && rank < max_ids) row[rank] = w.ids_at[p0 + lane] + m.id_offset;   // (:1516)
    out += __popc(word);
  }
  return out;
}

__device__ int sp_bpe_fast(const SpModelDev& m, const BWork& w, const ArcScratch& scratch, const uint8_t* text, int64_t lo0, int64_t hi,
                           int64_t padded_bytes, int32_t* row, int max_ids, int unk, uint16_t delim, int lane) {
  const unsigned full = 0xffffffffu;
  const bool fast = m.tok_algo == kTokenizeBpeOpt || m.tok_algo == kTokenizeBpeOptWithMerges;
  int64_t lo = lo0;
  if (hi - lo >= 3) {
    const uint32_t b0 = __ldg(text + lo), b1 = __ldg(text + lo + 1), b2 = __ldg(text + lo + 2);
    if (b0 == 0xEF && b1 == 0xBB && b2 == 0xBF) lo += 3;      // FAStrUtf8AsBytesToArray skips the BOM
  }
  if (hi <= lo) return 0;                                      // no symbols (:1409)
  const uint32_t* text32 = reinterpret_cast<const uint32_t*>(text);
  int fill = 0, out = 0, last_delim = 0;
  bool prior = false;                                          // an earlier window has been emitted
  unsigned carry = 0;                                          // the previous raw symbol is white, or the dummy prefix
  if (!m.no_dummy_prefix) { if (lane == 0) w.sym[0] = delim; fill = 1; carry = 1; }   // (:1372,:1387)
  int64_t bpos = lo;
  for (;;) {
    // ---- fill: whitespace -> U+2581, a white symbol survives iff its predecessor is neither (:1462-1496) ----
    while (bpos < hi && fill + 128 <= kBWin) {
      const int64_t bs = bpos & ~(int64_t)3;
      const int64_t pos0 = bs + lane * 4;
      const uint32_t word = pos0 < padded_bytes ? __ldg(text32 + (pos0 >> 2)) : 0u;
      unsigned white = 0, valid = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t pos = pos0 + k;
        const unsigned c = (word >> (8 * k)) & 0xFFu;
        if (pos < bpos) white |= carry << k;                   // filler before the first byte passes the carry on
        else if (pos < hi) { valid |= 1u << k; if (c <= 0x20u || c == 0xa0u) white |= 1u << k; }
      }
      const unsigned up = __shfl_up_sync(full, white >> 3, 1) & 1u;
      const unsigned prevw = ((white << 1) | (lane ? up : carry)) & 0xFu;
      const unsigned keep = valid & ~(white & prevw);
      const int c = __popc(keep);
      const int incl = warp_incl_scan(c, lane);
      int o = fill + incl - c, my_last = -1;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((keep >> k) & 1u) {
          const bool wh = (white >> k) & 1u;
          w.sym[o] = wh ? delim : __ldg(m.sym_of_cp + ((word >> (8 * k)) & 0xFFu));
          if (wh) my_last = o;
          ++o;
        }
      }
      last_delim = max(last_delim, __reduce_max_sync(full, my_last));   // a U+2581 at 0 is no cut point
      fill += __shfl_sync(full, incl, 31);
      const int last = (int)(min(hi, bs + 128) - 1 - bs);     // the step's last byte
      carry = (__shfl_sync(full, white, last >> 2) >> (last & 3)) & 1u;
      bpos = bs + 128;
    }
    __syncwarp();
    const bool at_end = bpos >= hi;
    int cut; bool open_ended = false;
    if (at_end) {
      if ((prior || fill > 1) && fill > 0 && w.sym[fill - 1] == delim) --fill;   // one trailing U+2581 goes (:1491-1493)
      cut = fill;
    } else if (last_delim > 0) {
      cut = last_delim;
    } else {
      // One segment fills the window.  Cut it at the last position p that no token spans (see the
      // split pass), provided every start before p has been walked to its end inside the window.
      const uint4* da = reinterpret_cast<const uint4*>(m.da);
      int carry = -1, best = 0; bool closed = true;
      for (int p0 = 0; p0 < fill && closed; p0 += 32) {
        const int p = p0 + lane;
        int fe = -1; bool open = false;
        if (p < fill) fe = b_farthest(m, da, w.sym, p, fill, &open);
        int incl = fe;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(full, incl, o); if (lane >= o) incl = max(incl, t); }
        int excl = __shfl_up_sync(full, incl, 1);
        excl = lane ? max(excl, carry) : carry;
        const unsigned ob = __ballot_sync(full, open);
        const int first_open = ob ? p0 + __ffs(ob) - 1 : fill;   // starts at or after it are not fully known
        const unsigned cb = __ballot_sync(full, p < fill && p > 0 && excl < p && p <= first_open);
        if (cb) best = p0 + 31 - __clz(cb);
        if (ob) closed = false;
        carry = max(carry, __shfl_sync(full, incl, 31));
      }
      if (best <= 0) return kUFallback;                        // no such position: the general path
      cut = best; open_ended = true;
    }
    if (cut > 0) {
      out = bpe_window(m, w, scratch, cut, delim, row, out, max_ids, unk, fast, open_ended, lane);
      if (out == kUFallback) return kUFallback;
      if (out >= max_ids) return max_ids;
    }
    if (at_end) break;
    // ---- slide: the unfinished segment moves to the front ----
    const int rest = fill - cut;
    for (int i0 = 0; i0 < rest; i0 += 32) {
      const int i = i0 + lane;
      const uint16_t v = i < rest ? w.sym[cut + i] : (uint16_t)0;
      __syncwarp();
      if (i < rest) w.sym[i] = v;
      __syncwarp();
    }
    fill = rest; last_delim = 0; prior = true;
  }
  return out;
}

// BPE family: one document per warp
__global__ void __launch_bounds__(kBWarps * 32, kBCtasPerSm) sp_bpe_kernel(const SpLaunch p, const SpModelDev m, int* error_flag) {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  uint8_t* smem = simt::shared_base();
#else
  extern __shared__ __align__(16) uint8_t smem[];
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kBWarps + warp;
  const BWork w = make_bwork(smem + (size_t)warp * kBWorkBytes);
  uint8_t* my_arena = p.arena + (size_t)gwarp * p.arena_stride;
  Work wa = make_work(my_arena, p.arena_cap, p.starts != nullptr);
  const ArcScratch scratch = make_scratch(p, my_arena, error_flag);
  const int64_t padded_bytes = (p.text_bytes + 3) & ~(int64_t)3;
  const uint16_t delim = __ldg(m.sym_of_cp + kSpDelim);
  const bool fast_model = p.starts == nullptr && m.use_raw_bytes && m.norm_count == nullptr && !m.delim_inside_tokens &&
                          m.bpe_ord != nullptr && delim != kNoSym;
  for (;;) {
    unsigned long long d64 = 0;
    if (lane == 0) d64 = atomicAdd(p.work_counter, 1ull);
    d64 = __shfl_sync(0xffffffffu, d64, 0);
    if ((int64_t)d64 >= p.ndocs) break;
    const int64_t doc = (int64_t)d64;
    const int64_t lo = __ldg(p.offsets + doc), hi = __ldg(p.offsets + doc + 1);
    const int64_t n = hi - lo;
    int result = 0;
    if (n > 0 && n <= 1000000000) {                                       // :1362
      result = kUFallback;
      if (fast_model)
        result = sp_bpe_fast(m, w, scratch, p.text, lo, hi, padded_bytes, p.ids + doc * (int64_t)p.max_ids, p.max_ids, p.unk_id, delim, lane);
      if (result == kUFallback)
        result = sp_doc_generic<true>(p, m, wa, scratch, doc, lo, hi, padded_bytes, lane, error_flag);
    }
    if (lane == 0) p.counts[doc] = result;
    __syncwarp();
  }
}

// =====================================================================================
// Unigram-LM fast path: tokens of up to kUMaxLen symbols, documents of any length streamed through a
// window of kUCap symbols cut at U+2581, ~8 KB of shared memory per warp (24 warps per SM).  Same arithmetic and the same
// visiting order as sp_unigram; what changes is where things live:
//   * one fused pass decodes UTF-8 and applies the charmap, one pass collapses whitespace and
//     maps code points to alphabet indices (so a walk step is ONE 16-byte gather);
//   * lanes walk 32 consecutive starts at a time and fetch {id, score} of every arc they find
//     right there (the I2Info gather overlaps the next step's gather), so the serial relaxation
//     touches shared memory only;
//   * a token spans <= 16 symbols, so the relaxation keeps the scores in REGISTERS: during a
//     half-tile of 16 starts lane j owns position t0-1+j (lane 0: the finished position before the
//     half-tile, lanes 1..31: everything its starts can reach).  Start st's turn: every lane takes
//     score[st-1] by shuffle, the lane whose position is st+k looks up arc (st, k) in the tile and
//     relaxes its own registers.  No shared-memory traffic but the arc fetch, no barriers.
// Anything that does not fit (a run of kUCap symbols without U+2581, offsets, raw bytes, longer
// tokens) takes sp_doc_generic in the warp's arena.
// =====================================================================================
constexpr int kUWarps = 8;                 // per CTA
constexpr int kUCtasPerSm = 3;
constexpr int kUCap = kSpUnigramFastCap;   // symbols
constexpr int kUMaxLen = 16;               // longest token (symbols): lanes 1..31 cover 16 starts + 15 more positions
constexpr int kUNoBegin = 0xFFFF;

struct UWork {
  int2* arc;           // [32][kUMaxLen] {id, score bits} of arc (start, length-1) for the tile's 32 starts
  int32_t* stage;      // [kUCap] normalised code points; then bid[]: id of the best arc ending at p
  uint32_t* mark;      // [kUCap/32] bit p: a token starts at p
  uint16_t* sym;       // [kUCap] alphabet indices after whitespace collapsing
  uint16_t* begin;     // [kUCap] start of the best arc ending at p
};
constexpr int kUWorkBytes = 8 * 32 * kUMaxLen + 4 * kUCap + 4 * (kUCap / 32) + 2 * kUCap + 2 * kUCap;
static_assert(kUWorkBytes % 8 == 0 && kUCap % 32 == 0 && kUCap < kUNoBegin, "workspace layout");

__device__ inline UWork make_uwork(uint8_t* b) {
  UWork w;
  w.arc = (int2*)b; b += 8 * 32 * kUMaxLen;
  w.stage = (int32_t*)b; b += 4 * kUCap;
  w.mark = (uint32_t*)b; b += 4 * (kUCap / 32);
  w.sym = (uint16_t*)b; b += 2 * kUCap;
  w.begin = (uint16_t*)b;
  return w;
}


// Best path over the window's symbols sym[0..N) (FATokenSegmentationTools_1best_t.h:174-279); the ids
// of its tokens are appended to row[out..).  *carry is the best score of the position before the
// window on entry and of position N-1 on exit.  Returns the new out.
__device__ int unigram_window(const SpModelDev& m, const UWork& w, int N, double* carry, int32_t* row, int out, int max_ids,
                              int unk, int lane) {
  const unsigned full = 0xffffffffu;
  int32_t* bid = w.stage;
  for (int i = lane; i < kUCap / 32; i += 32) w.mark[i] = 0;
  const uint4* da = reinterpret_cast<const uint4*>(m.da);
  // lane j owns position t0-1+j: best score, start and id of the best arc ending there
  double sc = lane == 0 ? *carry : -(double)FLT_MAX;          // lane 0: the position before the window
  int bg = kUNoBegin, bi = lane == 0 ? 0 : -1;
  for (int tA = 0; tA < N; tA += 32) {
    __syncwarp();                                              // the previous tile's arcs have been consumed
    // phase A: lane l finds the arcs of start tA + l (:196-224): bit k of amask = an arc of k+1 symbols
    const int start = tA + lane;
    unsigned amask = 0;
    if (start < N) {
      uint32_t q = m.root; int sum = 0;
      const int lim = min(kUMaxLen, N - start);
      for (int k = 0; k < lim; ++k) {
        const uint16_t s = w.sym[start + k];
        if (s == kNoSym) break;
        const uint4 e = __ldg(da + ((size_t)q + s));
        if (e.x != q) break;
        sum += (int)e.z;
        q = e.y & ~kDaFinalBit;
        if (e.y & kDaFinalBit) {
          int id; float score;
          sp_info(m, sum, -1, id, score);
          w.arc[lane * kUMaxLen + k] = make_int2(id, __float_as_int(score));
          amask |= 1u << k;
        }
        if (q == 0) break;                                     // a leaf: every further step fails
      }
    }
    __syncwarp();
    // phase B: relax in start order (ties keep the earlier start), 16 starts per register window
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int t0 = tA + 16 * h;
      if (t0 >= N) break;
      // off the serial chain: the score of the arc that reaches this lane's position from each of the
      // 16 starts (if there is one), and which starts have no arc at all
      float in_sc[16];
      unsigned have = 0, none = 0;
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        const unsigned M = __shfl_sync(full, amask, 16 * h + l);
        const int k = lane - 1 - l;
        in_sc[l] = 0.0f;
        if ((unsigned)k < (unsigned)kUMaxLen && ((M >> k) & 1u)) {
          in_sc[l] = __int_as_float(w.arc[(16 * h + l) * kUMaxLen + k].y);
          have |= 1u << l;
        }
        if (M == 0) none |= 1u << l;
      }
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        const int st = t0 + l;
        if (st >= N) break;
        const double prev = __shfl_sync(full, sc, l);          // score[st-1], final by now
        if ((none >> l) & 1u) {                                // AddUnknownArc (:145-171)
          const int pid = __shfl_sync(full, bi, l), pbg = __shfl_sync(full, bg, l);
          if (lane == l + 1) {
            const double cand = (double)(-100000.0f) + prev;
            if (sc < cand) { sc = cand; bi = -1; bg = (st > 0 && pid == -1) ? pbg : st; }
          }
        } else {                                               // AddArc (:118-142)
          const double cand = (double)in_sc[l] + prev;
          if (((have >> l) & 1u) && sc < cand) { sc = cand; bg = st; bi = w.arc[(16 * h + l) * kUMaxLen + lane - 1 - l].x; }
        }
      }
      // positions t0 .. t0+15 (lanes 1..16) are final: park them for the back-trace, slide the window
      if (lane >= 1 && lane <= 16 && t0 - 1 + lane < N) { w.begin[t0 - 1 + lane] = (uint16_t)bg; bid[t0 - 1 + lane] = bi; }
      if (N - 1 >= t0 && N - 1 <= t0 + 15) *carry = __shfl_sync(full, sc, N - t0);   // score of the window's last position
      sc = __shfl_down_sync(full, sc, 16); bg = __shfl_down_sync(full, bg, 16); bi = __shfl_down_sync(full, bi, 16);
      if (lane >= 16) { sc = -(double)FLT_MAX; bg = kUNoBegin; bi = -1; }
    }
  }
  __syncwarp();
  // ---- back-trace (:227-257): mark the token ENDS (same order as the starts; the id already sits there) ----
  if (lane == 0) {
    int end = N - 1;
    while (end >= 0) {
      w.mark[end >> 5] |= 1u << (end & 31);
      const int b = w.begin[end];
      if (b == kUNoBegin) break;                               // never-set arc: the reference emits it first and stops
      end = b - 1;
    }
  }
  __syncwarp();
  for (int p0 = 0; p0 < N && out < max_ids; p0 += 32) {
    const uint32_t word = w.mark[p0 >> 5];
    const int rank = out + __popc(word & bf_lanemask_lt());
    if (((word >> lane) & 1u) && rank < max_ids) {
      int id = bid[p0 + lane];
      if (id == -1) id = unk;
      row[rank] = id + m.id_offset;                            // ids[k] = id + IdOffset, UNK included (:1516)
    }
    out += __popc(word);
  }
  return out;
}

// The whole document in one window: one fused decode + charmap pass into stage[], one collapse pass.
// kUFallback when it has more than kUCap symbols after the charmap (the streamed form takes over).
__device__ int unigram_whole(const SpModelDev& m, const UWork& w, const uint8_t* text, int64_t lo0, int64_t hi,
                               int64_t padded_bytes, int32_t* row, int max_ids, int unk, int lane) {
  const unsigned full = 0xffffffffu;
  int64_t lo = lo0;
  if (hi - lo >= 3) {
    const uint32_t b0 = __ldg(text + lo), b1 = __ldg(text + lo + 1), b2 = __ldg(text + lo + 2);
    if (b0 == 0xEF && b1 == 0xBB && b2 == 0xBF) lo += 3;      // FAStrUtf8ToArray skips the BOM
  }
  if (hi <= lo) return 0;                                      // no symbols (:1409)
  const int64_t n = hi - lo0;
  const bool cm = m.norm_count != nullptr;
  // ---- pass 1: decode + charmap (FANormalize, FAUtils_cl.h:311-369), dummy prefix included (:1372,:1432) ----
  int total = 0;
  if (!m.no_dummy_prefix) {
    const unsigned nc = cm ? (unsigned)__ldg(m.norm_count + kSpDelim) : 0xFFu;
    if (nc == 0xFFu) { if (lane == 0) w.stage[0] = kSpDelim; total = 1; }
    else {
      const uint32_t f = __ldg(m.norm_first + kSpDelim);
      if (lane == 0) for (unsigned k = 0; k < nc; ++k) w.stage[k] = __ldg(m.norm_values + f + k);
      total = (int)nc;
    }
  }
  {
    const uint32_t* text32 = reinterpret_cast<const uint32_t*>(text);
    unsigned bad = 0, sumlen = 0;
    for (int64_t bpos = lo; bpos < hi;) {
      const int64_t bs = bpos & ~(int64_t)3;
      const int64_t pos0 = bs + lane * 4;
      uint32_t w0, w1;
      utf8_load_words(text32, pos0, padded_bytes, &w0, &w1);
      const Utf8Lane d = utf8_decode_lane(w0, w1, pos0, bpos, hi);
      bad |= d.bad; sumlen += d.sumlen;
      unsigned nck[4]; int c = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        nck[k] = 0xFFu;
        if (d.start_mask & (1u << k)) {
          if (cm) nck[k] = (unsigned)__ldg(m.norm_count + d.cp[k]);
          c += nck[k] == 0xFFu ? 1 : (int)nck[k];
        }
      }
      const int incl = warp_incl_scan(c, lane);
      const int wt = __shfl_sync(full, incl, 31);
      if (total + wt > kUCap) return kUFallback;
      int o = total + incl - c;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (d.start_mask & (1u << k)) {
          if (nck[k] == 0xFFu) w.stage[o++] = (int)d.cp[k];
          else {
            const uint32_t f = __ldg(m.norm_first + d.cp[k]);
            for (unsigned j = 0; j < nck[k]; ++j) w.stage[o++] = __ldg(m.norm_values + f + j);
          }
        }
      }
      total += wt;
      bpos = bs + 128;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sumlen += __shfl_xor_sync(full, sumlen, o);
    if (__any_sync(full, bad != 0) || (int64_t)sumlen != hi - lo) return 0;
  }
  if (cm && (total <= 0 || (int64_t)total > 2 * (n + 1))) return 0;      // :1442-1446
  __syncwarp();
  // ---- pass 2: whitespace runs -> one U+2581, trailing one dropped (:1462-1496); alphabet indices ----
  int N = 0, last_c = 0;
  for (int base = 0; base < total; base += 32) {
    const int i = base + lane;
    bool keep = false; int c = 0;
    if (i < total) {
      c = w.stage[i];
      const bool white = sp_is_white(c);
      if (!white || i == 0) keep = true;
      else { const int q = w.stage[i - 1]; keep = !sp_is_white(q) && q != kSpDelim; }
      if (white) c = kSpDelim;
    }
    const unsigned bal = __ballot_sync(full, keep);
    if (keep) w.sym[N + __popc(bal & bf_lanemask_lt())] = (unsigned)c <= 0x10FFFFu ? __ldg(m.sym_of_cp + c) : kNoSym;
    if (bal) last_c = __shfl_sync(full, c, 31 - __clz(bal));
    N += __popc(bal);
  }
  if (N > 1 && last_c == kSpDelim) --N;
  if (N <= 0) return 0;
  __syncwarp();
  double carry = 0.0;                                          // "position -1": the empty prefix
  const int out = unigram_window(m, w, N, &carry, row, 0, max_ids, unk, lane);
  return out < max_ids ? out : max_ids;
}

// Streams one document through the window.  A U+2581 is a forced token boundary: no token contains it
// past its first symbol, and "U+2581" itself is a token (both checked at load), so its start always
// has an arc and no unknown run merges across it.  Hence the best path up to the last U+2581 of the
// window is final: it is traced back and emitted, the rest slides to the front, and the best score of
// the last position carries over (the scores are absolute, as in the reference).
__device__ int unigram_streamed(const SpModelDev& m, const UWork& w, const uint8_t* text, int64_t lo0, int64_t hi,
                                int64_t padded_bytes, int32_t* row, int max_ids, int unk, int lane) {
  const unsigned full = 0xffffffffu;
  int64_t lo = lo0;
  if (hi - lo >= 3) {
    const uint32_t b0 = __ldg(text + lo), b1 = __ldg(text + lo + 1), b2 = __ldg(text + lo + 2);
    if (b0 == 0xEF && b1 == 0xBB && b2 == 0xBF) lo += 3;      // FAStrUtf8ToArray skips the BOM
  }
  if (hi <= lo) return 0;                                      // no symbols (:1409)
  const int64_t n = hi - lo0;
  const bool cm = m.norm_count != nullptr;
  const uint32_t* text32 = reinterpret_cast<const uint32_t*>(text);
  int32_t* blk = reinterpret_cast<int32_t*>(w.arc);            // normalised code points of one step (the tile is idle while filling)
  constexpr int kBlkCap = 8 * 32 * kUMaxLen / 4;               // 1024 code points
  int fill = 0, out = 0, last_delim = 0, prev_c = 0;
  int64_t stream = 0;                                          // normalised symbols so far (:1442-1446)
  bool prior = false, first_sym = true, last_is_delim = false;
  double carry = 0.0;                                          // "position -1": the empty prefix
  unsigned bad = 0, sumlen = 0;
  bool dummy_pending = !m.no_dummy_prefix;
  int64_t bpos = lo;
  for (;;) {
    // ---- fill: decode + charmap (FANormalize, FAUtils_cl.h:311-369) into blk, then whitespace -> U+2581
    // with runs collapsed (:1462-1496) and alphabet indices into the window ----
    while (dummy_pending || bpos < hi) {
      int wt = 0;
      int64_t next_bpos = bpos;
      if (dummy_pending) {                                     // the dummy prefix goes through the charmap too (:1372,:1432)
        const unsigned nc = cm ? (unsigned)__ldg(m.norm_count + kSpDelim) : 0xFFu;
        if (nc == 0xFFu) { if (lane == 0) blk[0] = kSpDelim; wt = 1; }
        else {
          const uint32_t f = __ldg(m.norm_first + kSpDelim);
          if (lane == 0) for (unsigned k = 0; k < nc; ++k) blk[k] = __ldg(m.norm_values + f + k);
          wt = (int)nc;
        }
      } else {
        const int64_t bs = bpos & ~(int64_t)3;
        const int64_t pos0 = bs + lane * 4;
        uint32_t w0, w1;
        utf8_load_words(text32, pos0, padded_bytes, &w0, &w1);
        const Utf8Lane d = utf8_decode_lane(w0, w1, pos0, bpos, hi);
        unsigned nck[4]; int c = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          nck[k] = 0xFFu;
          if (d.start_mask & (1u << k)) {
            if (cm) nck[k] = (unsigned)__ldg(m.norm_count + d.cp[k]);
            c += nck[k] == 0xFFu ? 1 : (int)nck[k];
          }
        }
        const int incl = warp_incl_scan(c, lane);
        wt = __shfl_sync(full, incl, 31);
        if (wt > kBlkCap) return kUFallback;
        if (fill + wt > kUCap) break;                          // the window is full: this step is decoded again later
        bad |= d.bad; sumlen += d.sumlen;
        int o = incl - c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (d.start_mask & (1u << k)) {
            if (nck[k] == 0xFFu) blk[o++] = (int)d.cp[k];
            else {
              const uint32_t f = __ldg(m.norm_first + d.cp[k]);
              for (unsigned j = 0; j < nck[k]; ++j) blk[o++] = __ldg(m.norm_values + f + j);
            }
          }
        }
        next_bpos = bs + 128;
      }
      if (fill + wt > kUCap) break;
      __syncwarp();
      // collapse blk[0..wt) behind the stream: a white symbol is kept iff it is the first symbol of all
      // or its predecessor is neither white nor U+2581
      int kept_total = 0;
      for (int base = 0; base < wt; base += 32) {
        const int i = base + lane;
        bool keep = false; int c = 0;
        if (i < wt) {
          c = blk[i];
          const bool white = sp_is_white(c);
          if (!white || (first_sym && i == 0)) keep = true;
          else { const int q = i > 0 ? blk[i - 1] : prev_c; keep = !sp_is_white(q) && q != kSpDelim; }
          if (white) c = kSpDelim;
        }
        const unsigned bal = __ballot_sync(full, keep);
        const int o = fill + kept_total + __popc(bal & bf_lanemask_lt());
        if (keep) w.sym[o] = (unsigned)c <= 0x10FFFFu ? __ldg(m.sym_of_cp + c) : kNoSym;
        const unsigned db = __ballot_sync(full, keep && c == kSpDelim);
        if (db) last_delim = max(last_delim, fill + kept_total + __popc(bal & ((2u << (31 - __clz(db))) - 1u)) - 1);
        if (bal) last_is_delim = (db >> (31 - __clz(bal))) & 1u;
        kept_total += __popc(bal);
      }
      if (wt > 0) { prev_c = blk[wt - 1]; first_sym = false; }
      __syncwarp();
      fill += kept_total;
      stream += wt;
      if (dummy_pending) dummy_pending = false; else bpos = next_bpos;
    }
    const bool at_end = !dummy_pending && bpos >= hi;
    if (__any_sync(full, bad != 0)) return 0;                  // invalid UTF-8 anywhere zeroes the document
    int cut;
    if (at_end) {
      unsigned tot = sumlen;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(full, tot, o);
      if ((int64_t)tot != hi - lo) return 0;
      if (cm && (stream <= 0 || stream > 2 * (n + 1))) return 0;                  // :1442-1446
      if ((prior || fill > 1) && fill > 0 && last_is_delim) --fill;               // one trailing U+2581 goes (:1491-1493)
      cut = fill;
    } else if (last_delim > 0) {
      cut = last_delim;
    } else {
      // A run without U+2581 fills the window (a URL, CJK text).  Any position p that no token spans is a
      // forced boundary as well, unless p-1 and p are both unknown symbols (an unknown run is one token,
      // :145-171); every start before p must have been walked to its end inside the window.
      const uint4* da = reinterpret_cast<const uint4*>(m.da);
      int reach = -1, best = 0, prev_unknown = 0; bool closed = true;
      for (int p0 = 0; p0 < fill && closed; p0 += 32) {
        const int p = p0 + lane;
        int fe = -1; bool open = false;
        if (p < fill) fe = b_farthest(m, da, w.sym, p, fill, &open);
        const bool unknown = p < fill && fe < 0;
        int incl = unknown ? p : fe;                           // an unknown symbol covers itself
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(full, incl, o); if (lane >= o) incl = max(incl, t); }
        int excl = __shfl_up_sync(full, incl, 1);
        excl = lane ? max(excl, reach) : reach;
        const unsigned ub = __ballot_sync(full, unknown);
        const bool before_unknown = lane ? ((ub >> (lane - 1)) & 1u) : (prev_unknown != 0);
        const unsigned ob = __ballot_sync(full, open);
        const int first_open = ob ? p0 + __ffs(ob) - 1 : fill;   // starts at or after it are not fully known
        const unsigned cb = __ballot_sync(full, p < fill && p > 0 && excl < p && p <= first_open && !(unknown && before_unknown));
        if (cb) best = p0 + 31 - __clz(cb);
        if (ob) closed = false;
        reach = max(reach, __shfl_sync(full, incl, 31));
        prev_unknown = (int)(ub >> 31);
      }
      if (best <= 0) return kUFallback;                        // no such position: the general path
      cut = best;
    }
    if (cut > 0) {
      out = unigram_window(m, w, cut, &carry, row, out, max_ids, unk, lane);
      if (out >= max_ids) {
        // the ids are complete, but an invalid byte or a charmap overflow later in the document must
        // still yield 0 (:1409, :1442-1446)
        if (!at_end) {
          unsigned more = 0;
          for (; bpos < hi; bpos = (bpos & ~(int64_t)3) + 128) {
            const int64_t bs = bpos & ~(int64_t)3, pos0 = bs + lane * 4;
            uint32_t w0, w1;
            utf8_load_words(text32, pos0, padded_bytes, &w0, &w1);
            const Utf8Lane d = utf8_decode_lane(w0, w1, pos0, bpos, hi);
            bad |= d.bad; sumlen += d.sumlen;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (d.start_mask & (1u << k)) {
                const unsigned nc = cm ? (unsigned)__ldg(m.norm_count + d.cp[k]) : 0xFFu;
                more += nc == 0xFFu ? 1u : nc;
              }
            }
          }
          unsigned tot = sumlen;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { tot += __shfl_xor_sync(full, tot, o); more += __shfl_xor_sync(full, more, o); }
          if (__any_sync(full, bad != 0) || (int64_t)tot != hi - lo) return 0;
          if (cm && stream + (int64_t)more > 2 * (n + 1)) return 0;
        }
        return max_ids;
      }
    }
    if (at_end) break;
    // ---- slide: the unfinished segment moves to the front ----
    const int rest = fill - cut;
    for (int i0 = 0; i0 < rest; i0 += 32) {
      const int i = i0 + lane;
      const uint16_t v = i < rest ? w.sym[cut + i] : (uint16_t)0;
      __syncwarp();
      if (i < rest) w.sym[i] = v;
      __syncwarp();
    }
    fill = rest; last_delim = 0; prior = true;
  }
  return out;
}

// Short documents take the one-window form (fewer passes); anything longer is streamed.
__device__ int sp_unigram_fast(const SpModelDev& m, const UWork& w, const uint8_t* text, int64_t lo0, int64_t hi,
                               int64_t padded_bytes, int32_t* row, int max_ids, int unk, int lane) {
  if (hi - lo0 <= 4ll * kUCap) {                               // a code point takes at most 4 bytes
    const int r = unigram_whole(m, w, text, lo0, hi, padded_bytes, row, max_ids, unk, lane);
    if (r != kUFallback) return r;
  }
  return unigram_streamed(m, w, text, lo0, hi, padded_bytes, row, max_ids, unk, lane);
}

__global__ void __launch_bounds__(kUWarps * 32, kUCtasPerSm) sp_unigram_kernel(const SpLaunch p, const SpModelDev m, int* error_flag) {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  uint8_t* smem = simt::shared_base();
#else
  extern __shared__ __align__(16) uint8_t smem[];
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kUWarps + warp;
  const UWork w = make_uwork(smem + (size_t)warp * kUWorkBytes);
  uint8_t* my_arena = p.arena + (size_t)gwarp * p.arena_stride;
  Work wa = make_work(my_arena, p.arena_cap, p.starts != nullptr);
  const ArcScratch scratch = make_scratch(p, my_arena, error_flag);
  const int64_t padded_bytes = (p.text_bytes + 3) & ~(int64_t)3;
  const bool fast_model = p.starts == nullptr && !m.use_raw_bytes && m.max_arc_len <= kUMaxLen && !m.delim_inside_tokens &&
                          m.delim_is_token;
  for (;;) {
    unsigned long long d64 = 0;
    if (lane == 0) d64 = atomicAdd(p.work_counter, 1ull);
    d64 = __shfl_sync(0xffffffffu, d64, 0);
    if ((int64_t)d64 >= p.ndocs) break;
    const int64_t doc = (int64_t)d64;
    const int64_t lo = __ldg(p.offsets + doc), hi = __ldg(p.offsets + doc + 1);
    const int64_t n = hi - lo;
    int result = 0;
    if (n > 0 && n <= 1000000000) {                                       // :1362
      result = kUFallback;
      if (fast_model)
        result = sp_unigram_fast(m, w, p.text, lo, hi, padded_bytes, p.ids + doc * (int64_t)p.max_ids, p.max_ids, p.unk_id, lane);
      if (result == kUFallback)
        result = sp_doc_generic<false>(p, m, wa, scratch, doc, lo, hi, padded_bytes, lane, error_flag);
    }
    if (lane == 0) p.counts[doc] = result;
    __syncwarp();
  }
}

}  // namespace

int64_t sp_arena_bytes_per_warp(int cap, int) {
  return align16(work_bytes_arena(cap) + 16ll * ((int64_t)cap * kArcsPerSym + 4096) + 256);
}

int64_t sp_overflow_entries(int cap, int max_arc_len) {
  // a start has at most min(max_arc_len, cap) arcs; see `need` in bpe_segment
  const int64_t per = max_arc_len < cap ? max_arc_len : cap;
  const int64_t total = (int64_t)cap * per;
  const int64_t need = total + 2 * (total + cap) + 64;
  const int64_t limit = (2ll << 30) / 16;          // 2 GB
  return need < limit ? need : limit;
}

static bool is_bpe_algo(int tok_algo) {
  return tok_algo == kTokenizeBpe || tok_algo == kTokenizeBpeOpt || tok_algo == kTokenizeBpeOptWithMerges;
}

#ifndef BF_SIMT_HOST
int sp_preferred_warps(int tok_algo) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return is_bpe_algo(tok_algo) ? sms * kBWarps * kBCtasPerSm : sms * kUWarps * kUCtasPerSm;
}

#endif

int sp_fast_cap(int tok_algo, int max_arc_len, bool use_raw_bytes) {
  return (!is_bpe_algo(tok_algo) && !use_raw_bytes && max_arc_len <= kUMaxLen) ? kUCap : 0;
}

#ifndef BF_SIMT_HOST
cudaError_t sp_tokenize_launch(const SpLaunch& p, const SpModelDev& m, cudaStream_t stream, int* launches) {
  if (p.ndocs <= 0) return cudaSuccess;
  const bool bpe = is_bpe_algo(m.tok_algo);
  const int cta_warps = bpe ? kBWarps : kUWarps;
  const size_t smem = bpe ? (size_t)kBWarps * kBWorkBytes : (size_t)kUWarps * kUWorkBytes;
  auto kern = bpe ? sp_bpe_kernel : sp_unigram_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = p.grid_warps / cta_warps;
  if (grid < 1) return cudaErrorInvalidValue;
  const int64_t needed = (p.ndocs + cta_warps - 1) / cta_warps;
  if (needed < grid) grid = (int)needed;
  e = cudaMemsetAsync(p.work_counter, 0, 16, stream);   // counter (8 B) + error flag (4 B)
  if (e != cudaSuccess) return e;
  int* err = reinterpret_cast<int*>(p.work_counter + 1);
  kern<<<grid, cta_warps * 32, smem, stream>>>(p, m, err);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

#endif

}  // namespace bfb200
