// sp_generic.cuh -- the general [pos-dict] path (sp_doc_generic): any document length (the warp's global arena), any
// token length, byte offsets; Unigram with an arc tile in start order, BPE with the reference's arc vector.
// Included by sp_kernel.cu only (one translation unit; everything lives in its anonymous namespace).
#pragma once

namespace bfb200 {
namespace {

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

}  // namespace
}  // namespace bfb200
