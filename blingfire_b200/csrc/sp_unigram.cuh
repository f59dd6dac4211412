// sp_unigram.cuh -- the Unigram-LM fast path (sp_unigram_fast): one window or streamed.
// Included by sp_kernel.cu only (one translation unit; everything lives in its anonymous namespace).
#pragma once

namespace bfb200 {
namespace {

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
// Anything that does not fit (a run of kUCap symbols without U+2581, raw bytes, longer tokens) takes sp_doc_generic in the
// warp's arena.  With offsets (sp_unigram_offsets_kernel) only the one-window form is fast: a document of more than kUCap
// symbols takes sp_doc_generic.
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
template <bool kOff>
__device__ int unigram_window(const SpModelDev& m, const UWork& w, int N, double* carry, int32_t* row, int out, int max_ids,
                              int unk, int lane, const WinOffsets& uo) {
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
      if constexpr (kOff) {                                    // the token ends at symbol p0 + lane and starts at begin[] of it
        int b = w.begin[p0 + lane];
        if (b == kUNoBegin) b = 0;                             // never-set arc: emitted as the token from symbol 0
        uo.starts[rank] = uo.boff[b];
        uo.ends[rank] = sp_end_offset(uo.doc, uo.boff[p0 + lane]);
      }
    }
    out += __popc(word);
  }
  return out;
}

// The whole document in one window: one fused decode + charmap pass into stage[], one collapse pass.
// kUFallback when it has more than kUCap symbols after the charmap (the streamed form takes over).
template <bool kOff>
__device__ int unigram_whole(const SpModelDev& m, const UWork& w, const uint8_t* text, int64_t lo0, int64_t hi,
                             int64_t padded_bytes, int32_t* row, int max_ids, int unk, int lane, const WinOffsets& uo) {
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
    if constexpr (kOff) { if (lane < total) uo.boff[lane] = -1; }          // (:1387; a charmap row has at most a few symbols)
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
          if (nck[k] == 0xFFu) {
            if constexpr (kOff) uo.boff[o] = (int)(pos0 + k - lo0);
            w.stage[o++] = (int)d.cp[k];
          } else {
            const uint32_t f = __ldg(m.norm_first + d.cp[k]);
            for (unsigned j = 0; j < nck[k]; ++j) {
              if constexpr (kOff) uo.boff[o] = (int)(pos0 + k - lo0);      // every symbol of an expansion: its source character
              w.stage[o++] = __ldg(m.norm_values + f + j);
            }
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
    int bv = 0;
    if constexpr (kOff) {                                      // compacted in place: a kept symbol never moves up
      if (i < total) bv = uo.boff[i];
      __syncwarp();
    }
    if (i < total) {
      c = w.stage[i];
      const bool white = sp_is_white(c);
      if (!white || i == 0) keep = true;
      else { const int q = w.stage[i - 1]; keep = !sp_is_white(q) && q != kSpDelim; }
      if (white) c = kSpDelim;
    }
    const unsigned bal = __ballot_sync(full, keep);
    if constexpr (kOff) { if (keep) uo.boff[N + __popc(bal & bf_lanemask_lt())] = bv; }
    if (keep) w.sym[N + __popc(bal & bf_lanemask_lt())] = (unsigned)c <= 0x10FFFFu ? __ldg(m.sym_of_cp + c) : kNoSym;
    if (bal) last_c = __shfl_sync(full, c, 31 - __clz(bal));
    N += __popc(bal);
  }
  if (N > 1 && last_c == kSpDelim) --N;
  if (N <= 0) return 0;
  __syncwarp();
  double carry = 0.0;                                          // "position -1": the empty prefix
  const int out = unigram_window<kOff>(m, w, N, &carry, row, 0, max_ids, unk, lane, uo);
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
      out = unigram_window<false>(m, w, cut, &carry, row, out, max_ids, unk, lane, WinOffsets{});
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
    const int r = unigram_whole<false>(m, w, text, lo0, hi, padded_bytes, row, max_ids, unk, lane, WinOffsets{});
    if (r != kUFallback) return r;
  }
  return unigram_streamed(m, w, text, lo0, hi, padded_bytes, row, max_ids, unk, lane);
}

}  // namespace
}  // namespace bfb200
