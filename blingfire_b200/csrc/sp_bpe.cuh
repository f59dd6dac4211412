// sp_bpe.cuh -- the streaming BPE fast path (sp_bpe_fast) for byte-level models.
// Included by sp_kernel.cu only (one translation unit; everything lives in its anonymous namespace).
#pragma once

namespace bfb200 {
namespace {

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
constexpr int kBCtasPerSm = 4;
constexpr int kBWin = 512;                 // symbols in the window
constexpr int kBLaneArcs = 48;             // lane-serial segments: at most this many listed arcs ...
constexpr int kBMaxLen = 64;               // ... and symbols (intermediate[] is one 64-bit register)
constexpr int kBSplitLen = 24;             // longer segments are first split at the positions no token spans
constexpr int kBCoopArcs = 512;            // warp-cooperative segments: arcs sorted in the window's scratch (64-bit keys)
constexpr unsigned kBUnclaimed = 0xFFFFFu; // ordinal of "no arc claimed from this start"

struct BWork {
  uint32_t* scratch;    // [32][kBLaneArcs] lane-interleaved sorted keys; or kBCoopArcs 64-bit keys.  GLOBAL memory (the warp's
                        // arena), like `order`: with the segment memo the hard pass runs for a fraction of the segments, and
                        // the 7.5 KB they took per warp in shared memory kept the kernel at 16 warps per SM
  int32_t* ids_at;      // [kBWin] claim state {ordinal, tos}, then the token id, at token starts
  uint32_t* mark;       // [kBWin/32] bit p: a token starts at p
  uint16_t* sym;        // [kBWin] alphabet indices
  uint16_t* seg;        // [kBWin/2+8] segment starts, window-relative, and the end sentinel
  uint16_t* hard_a;     // [kBWin] the pieces the easy pass left: first symbol ...
  uint16_t* hard_b;     // [kBWin] ... and one past the last
  uint32_t* cnt;        // [32] arcs listed per slot; [32] = slots that cannot be served lane-serially
  uint8_t* order;       // [kBLaneArcs][32] list index of the arc of rank r
};
constexpr int kBWorkBytes = 4 * kBWin + 4 * (kBWin / 32) + 2 * kBWin + 2 * (kBWin / 2 + 8) + 4 * kBWin + 4 * 36;
constexpr int kBGlobalBytes = 4 * 32 * kBLaneArcs + 32 * kBLaneArcs;   // scratch + order, at the start of the warp's private arc region
static_assert(kBWorkBytes % 16 == 0 && 8 * kBCoopArcs <= 4 * 32 * kBLaneArcs && kBLaneArcs <= 64 && kBMaxLen <= 64 && kBWin <= 1024, "workspace layout");

__device__ inline BWork make_bwork(uint8_t* b, uint8_t* g) {
  BWork w;
  w.scratch = (uint32_t*)g;
  w.order = g + 4 * 32 * kBLaneArcs;
  w.ids_at = (int32_t*)b; b += 4 * kBWin;
  w.mark = (uint32_t*)b; b += 4 * (kBWin / 32);
  w.sym = (uint16_t*)b; b += 2 * kBWin;
  w.seg = (uint16_t*)b; b += 2 * (kBWin / 2 + 8);
  w.hard_a = (uint16_t*)b; b += 2 * kBWin;
  w.hard_b = (uint16_t*)b; b += 2 * kBWin;
  w.cnt = (uint32_t*)b;
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
    if ((int64_t)P > 2 * (scratch.priv_cap - kBGlobalBytes / 16)) return false;
    keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(scratch.priv) + kBGlobalBytes);
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
template <bool kOff>
__device__ int bpe_window(const SpModelDev& m, const BWork& w, const ArcScratch& scratch, int cut, uint16_t delim, int32_t* row,
                          int out, int max_ids, int unk, bool fast, bool open_ended, int lane, const WinOffsets& wo) {
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
  // ---- memo pass: a segment's ids depend on its symbols alone (no token crosses a U+2581, the whole-word shortcut looks
  // only at the segment's own end) -- segments seen before come out of the table, one lane per segment.  A segment found
  // there is flagged in seg[] (bit 15) and skipped by the passes below; the others are added once those have resolved them.
  const WpWords& memo = m.seg_memo;
  // (not with offsets: a segment served from the table has its ids parked on its first symbols, not on the tokens' starts)
  const bool use_memo = !kOff && memo.max_len > 0;
  if (use_memo) {
    for (int g0 = 0; g0 < nseg; g0 += 32) {
      const int g = g0 + lane;
      int a = 0, L = 0;
      if (g < nseg && !(open_ended && g == nseg - 1)) { a = w.seg[g]; L = (int)w.seg[g + 1] - a; if (L > (int)memo.max_len) L = 0; }
      const int lcap = __reduce_max_sync(full, L);
      if (lcap == 0) continue;
      const bool wide = lcap > (int)(4 * memo.cpw);
      uint32_t kw[8];
      wp_pack_key_any(memo.cpw, w.sym + a, L, lcap, memo.cb, kw);
      const WpWordHit hit = wp_words_find(memo, kw, wide);
      if (L > 0 && hit.meta != 0) {
        const int n = (int)(hit.meta & 7u);                   // 1..6 tokens, at most one per symbol
        for (int k = 0; k < n; ++k) {
          w.ids_at[a + k] = hit.id[k];
          atomicOr(&w.mark[(a + k) >> 5], 1u << ((a + k) & 31));
        }
      }
      __syncwarp();
      if (L > 0 && hit.meta != 0) w.seg[g] = (uint16_t)(a | 0x8000);
    }
    __syncwarp();
  }
  // ---- easy pass: the bpe-opt whole-word shortcut, one lane per segment ----
  // Walking from the segment start, an arc that ends exactly at the segment end after a shorter
  // arc was already seen makes the reference keep ONLY that arc and skip the interior starts
  // (:188-206,:228-230); a one-symbol segment with an arc is a single arc as well.
  int nhard = 0;
  for (int g0 = 0; g0 < nseg; g0 += 32) {
    const int g = g0 + lane;
    bool hard = false;
    int a = 0, b = 0;
    if (g < nseg && (w.seg[g] & 0x8000u)) {
      // served by the memo pass
    } else if (g < nseg && open_ended && g == nseg - 1) {      // its true end is not in the window: no shortcut
      a = w.seg[g]; b = w.seg[g + 1] & 0x7FFF; hard = true;
    } else if (g < nseg) {
      a = w.seg[g]; b = w.seg[g + 1] & 0x7FFF;
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
  // ---- learn pass: the segments the table did not hold, now resolved, go into it (not those with a start nothing
  // claimed: that id is the caller's UnkId) ----
  if (use_memo) {
    for (int g0 = 0; g0 < nseg; g0 += 32) {
      const int g = g0 + lane;
      if (g < nseg && !(w.seg[g] & 0x8000u) && !(open_ended && g == nseg - 1)) {
        const int a = w.seg[g], b = w.seg[g + 1] & 0x7FFF, L = b - a;
        if (L <= (int)memo.max_len) {
          int n = 0, offs[kMaxLearnPieces];
          int32_t ids[kMaxLearnPieces];
          bool clean = true;
          for (int p = a; p < b; ++p)
            if ((w.mark[p >> 5] >> (p & 31)) & 1u) {
              const int32_t id = w.ids_at[p];
              if (id == unk) clean = false;
              if (n < kMaxLearnPieces) { ids[n] = id; offs[n] = n; }
              ++n;
            }
          if (clean && n >= 1 && n <= kMaxLearnPieces) {
            uint32_t kw[8];
            wp_pack_key_any(memo.cpw, w.sym + a, L, L, memo.cb, kw);
            wp_words_insert(memo, kw, L > (int)(4 * memo.cpw), n, ids, offs);
          }
        }
      }
    }
  }
  // ---- ordered emission ----
  for (int p0 = 0; p0 < cut && out < max_ids; p0 += 32) {
    const uint32_t word = w.mark[p0 >> 5];
    const int rank = out + __popc(word & bf_lanemask_lt());
    if (((word >> lane) & 1u) && rank < max_ids) {
      row[rank] = w.ids_at[p0 + lane] + m.id_offset;           // (:1516)
      if constexpr (kOff) {
        // the tokens tile the window: this one ends before the next start, the last one at cut - 1
        int e = cut - 1;
        const uint32_t rest = lane < 31 ? word >> (lane + 1) : 0u;
        if (rest) e = p0 + lane + __ffs(rest) - 1;
        else {
          for (int q = p0 + 32; q < cut; q += 32) {
            const uint32_t nw = w.mark[q >> 5];
            if (nw) { e = q + __ffs(nw) - 2; break; }
          }
        }
        wo.starts[rank] = wo.boff[p0 + lane];
        wo.ends[rank] = sp_end_offset(wo.doc, wo.boff[e]);
      }
    }
    out += __popc(word);
  }
  return out;
}

template <bool kOff>
__device__ int sp_bpe_fast(const SpModelDev& m, const BWork& w, const ArcScratch& scratch, const uint8_t* text, int64_t lo0, int64_t hi,
                           int64_t padded_bytes, int32_t* row, int max_ids, int unk, uint16_t delim, int lane, const WinOffsets& wo) {
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
  if (!m.no_dummy_prefix) {                                    // (:1372,:1387)
    if (lane == 0) { w.sym[0] = delim; if constexpr (kOff) wo.boff[0] = -1; }
    fill = 1; carry = 1;
  }
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
          if constexpr (kOff) wo.boff[o] = (int)(pos0 + k - lo0);
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
      out = bpe_window<kOff>(m, w, scratch, cut, delim, row, out, max_ids, unk, fast, open_ended, lane, wo);
      if (out == kUFallback) return kUFallback;
      if (out >= max_ids) return max_ids;
    }
    if (at_end) break;
    // ---- slide: the unfinished segment moves to the front ----
    const int rest = fill - cut;
    for (int i0 = 0; i0 < rest; i0 += 32) {
      const int i = i0 + lane;
      const uint16_t v = i < rest ? w.sym[cut + i] : (uint16_t)0;
      int bv = 0;
      if constexpr (kOff) { if (i < rest) bv = wo.boff[cut + i]; }
      __syncwarp();
      if (i < rest) w.sym[i] = v;
      if constexpr (kOff) { if (i < rest) wo.boff[i] = bv; }
      __syncwarp();
    }
    fill = rest; last_delim = 0; prior = true;
  }
  return out;
}

}  // namespace
}  // namespace bfb200
