// wp_kernel.cu -- fused TextToIds kernel for FastPath (flat two-level WordPiece) lexer models,
// hand-written for sm_100a.  Integer / indexing work only: no tensor cores.
//
// Replaces, per document, the reference's whole TextToIdsWithOffsets_wp pipeline
// (blingfiretokdll.cpp:1108-1314): FAStrUtf8ToArray -> FANormalize -> FALexTools_t::Process
// (with the nested FnTokWord call) -> tiling/UNK post-pass -> MaxIdsArrLength truncation.
//
// Mapping (DESIGN.md 3 has the derivation and the exactness argument):
//   * persistent CTAs, one document per WARP at a time, documents handed out by an atomic
//     counter (ragged lengths balance themselves);
//   * the small hot part of the model (top-level automaton, ASCII class table, the pair table of sync points and
//     class groups, the per-class memo kinds) is staged ONCE per CTA into shared memory by a bulk async copy
//     (cp.async.bulk + mbarrier, the 1-D TMA path; SASS: UBLKCP);
//   * bytes are read as coalesced 32-bit words (uchar4 per lane, 128 B per warp step).  An all-ASCII block is
//     classified from shared memory and its EVENTS (positions where a chunk may start: no top-level match can cross,
//     or where the group of top-level classes changes) are made in registers, four positions per lane; other blocks
//     are decoded and validated strictly (FAUtf8ToInt), compacted with a warp scan, classified from the L2-resident
//     class map;
//   * one lane per event turns events into chunks; the load-time memo of the lexer loops (wp_core.cuh, wp_model.cpp)
//     says per chunk: emits nothing / one word for the table / run the loops;
//   * one lane per word looks the packed class sequence up in the two-choice word table in HBM (vocabulary words placed
//     at load time, multi-piece words added at run time); what it does not hold runs the reference's loops verbatim
//     (gathers from the dense state x class table), one lane per chunk, and is added to the table;
//   * ids are written position-indexed in shared memory and compacted in order (four positions per lane, warp scan)
//     into the output row.
#include "wp_kernel.cuh"

namespace bfb200 {

namespace {

constexpr int kWarpsPerCta = 16;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kWin = 576;            // window capacity in code points (per warp)
constexpr int kBlockBytes = 128;     // bytes consumed per decode step (32 lanes x uchar4)
constexpr int kPad = 4;              // the window starts (offset & 3) entries into its arrays, so that a lane's 4 ASCII bytes land
                                     // on a 4-entry boundary and go out as one vector store per array
constexpr int kRing = 64;            // entries per work ring (a round consumes 32; at most 63 are ever queued)
// per-warp shared memory: ids_at i32 | cls u16 (+ slack for the key reads past a chunk) | top class u8 | events u16 |
// table ring u32 | loops ring u32
constexpr int kOffCls = (kWin + kPad) * 4;
constexpr int kOffMeta = kOffCls + (kWin + kPad + kMaxFastLen + 4) * 2;
constexpr int kOffEv = (kOffMeta + kWin + kPad + 15) & ~15;
constexpr int kOffFastq = kOffEv + (kWin + 8) * 2;
constexpr int kOffSlowq = kOffFastq + kRing * 4;
constexpr int kWarpSmem = kOffSlowq + kRing * 4;

static_assert(kWin % 32 == 0, "window must be a multiple of the warp size");
static_assert(kOffCls % 16 == 0 && kOffMeta % 4 == 0 && kOffFastq % 4 == 0 && kWarpSmem % 16 == 0, "vector stores need aligned arrays");
static_assert(kWin + kPad < 0x8000, "event entries keep the position in 15 bits");

__device__ __forceinline__ unsigned lanemask_lt() {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  return (1u << simt::tl.lane) - 1u;
#else
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
#endif
}

#ifndef BF_SIMT_HOST
// ---- bulk async copy (TMA 1-D) of the model blob into shared memory ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

#endif  // BF_SIMT_HOST

// ---- UTF-8: decode the (up to 4) sequences that START in this lane's word ----
// Strictness follows FAUtf8ToInt (FAUtf8Utils.cpp:121-196): shortest form, no surrogates,
// <= U+10FFFF, continuation bytes 10xxxxxx, no truncation at the end of the document.
// Offsets are relative to the 4-byte-aligned word the document starts in: the document is [ulo, uhi).
struct LaneDecode {
  uint32_t cp[4];
  unsigned start_mask;   // bit k: a sequence starts at byte k of the lane's word
  unsigned bad;          // nonzero: some sequence starting here is invalid
  unsigned sumlen;       // total length of the sequences that start here
};

__device__ __forceinline__ LaneDecode decode_lane(uint32_t w0, uint32_t w1, int u0, int ulo, int uhi) {
  LaneDecode r;
  r.start_mask = 0; r.bad = 0; r.sumlen = 0;
  const uint64_t x = ((uint64_t)w1 << 32) | w0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int pos = u0 + k;
    const uint32_t y = (uint32_t)(x >> (8 * k));
    const uint32_t b0 = y & 0xFF, b1 = (y >> 8) & 0xFF, b2 = (y >> 16) & 0xFF, b3 = y >> 24;
    r.cp[k] = 0;
    if (pos >= ulo && pos < uhi && (b0 & 0xC0) != 0x80) {
      uint32_t cp, len, bad = 0;
      if (b0 < 0x80) { cp = b0; len = 1; }
      else if ((b0 & 0xE0) == 0xC0) {
        len = 2; cp = ((b0 & 0x1F) << 6) | (b1 & 0x3F);
        bad = ((b1 & 0xC0) != 0x80) | (cp < 0x80);
      } else if ((b0 & 0xF0) == 0xE0) {
        len = 3; cp = ((b0 & 0x0F) << 12) | ((b1 & 0x3F) << 6) | (b2 & 0x3F);
        bad = ((b1 & 0xC0) != 0x80) | ((b2 & 0xC0) != 0x80) | (cp < 0x800) | ((cp & 0xFFFFF800u) == 0xD800u);
      } else if ((b0 & 0xF8) == 0xF0) {
        len = 4; cp = ((b0 & 0x07) << 18) | ((b1 & 0x3F) << 12) | ((b2 & 0x3F) << 6) | (b3 & 0x3F);
        bad = ((b1 & 0xC0) != 0x80) | ((b2 & 0xC0) != 0x80) | ((b3 & 0xC0) != 0x80) | (cp < 0x10000) | (cp > 0x10FFFF);
      } else { cp = 0; len = 1; bad = 1; }
      bad |= (pos + (int)len > uhi);
      r.bad |= bad;
      r.sumlen += len;
      r.cp[k] = bad ? 0u : cp;
      r.start_mask |= 1u << k;
    }
  }
  return r;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// The per-warp working set: one window of the current document, its events, and the two work rings.
struct WarpWin {
  int32_t* ids_at;     // [kWin] id of the piece starting at a position, kNoPiece elsewhere
  uint16_t* cls;       // [kWin] class of every position
  uint8_t* meta;       // [kWin] top-level class of every position
  uint16_t* ev;        // events in position order: position | 0x8000 when a chunk may start there (else: only the group
                       // of top-level classes changes)
  uint32_t* fastq;     // ring of word runs for the table: start | length << 16
  uint32_t* slowq;     // ring of chunks for the lexer loops: start | end << 16; bit 15: a word run the table did not hold
  int nev, nfast, fh, nslow, sh;   // counts and ring heads (warp-uniform)
  int carry;           // per lane: largest position known to be a finished chunk boundary
};

// Events of the positions [pb, pe), one position per lane (the all-ASCII blocks make theirs in registers).
__device__ __forceinline__ void gen_events(WarpWin& w, const WpTop& top, int pb, int pe, int lane) {
  for (int p0 = pb; p0 < pe; p0 += 32) {
    const int q = p0 + lane;
    unsigned sv = 0;
    if (q < pe) sv = q == 0 ? (unsigned)kSyncStart : (unsigned)top.sync_start[((unsigned)w.meta[q - 1] << top.sync_shift) | (unsigned)w.meta[q]];
    const unsigned bal = __ballot_sync(0xffffffffu, sv != 0);
    if (sv) w.ev[w.nev + __popc(bal & lanemask_lt())] = (uint16_t)((unsigned)q | ((sv & kSyncStart) ? 0x8000u : 0u));
    w.nev += __popc(bal);
  }
}

// Events -> chunks: every lane takes one event and its two successors.  A chunk start followed by another chunk
// start is a run inside ONE group of top-level classes (a change of group would be an event of its own); a chunk
// start, a change to the DEAD group, and the next chunk start is such a run followed by positions that match
// nothing ("word" + white space in bert_*).  The per-group memo says whether the run emits nothing, is one word for
// the table, or needs the loops -- as does every chunk of another shape.
__device__ __forceinline__ void classify_round(WarpWin& w, const WpTop& top, const WpWords& words, int i0, int cnt, int lane, int m,
                                               bool at_end, int limit, bool first) {
  // straight-line on purpose: every lane reads three (clamped) events and two kinds, the decisions are selects
  const int i = i0 + (lane < cnt ? lane : cnt - 1);
  const unsigned e0 = w.ev[i], e1 = w.ev[i + 1], e2 = w.ev[i + 2];
  const int s = (int)(e0 & 0x7FFFu), p1 = (int)(e1 & 0x7FFFu), p2 = (int)(e2 & 0x7FFFu);
  const bool b0 = lane < cnt && (e0 & 0x8000u), b1 = (e1 & 0x8000u) != 0, b2 = (e2 & 0x8000u) != 0;
  const uint32_t k0 = top.kind_of_tc[w.meta[s]];
  const uint32_t k1 = top.kind_of_tc[w.meta[p1 < m ? p1 : m - 1]];
  const bool shaped = b1 || (b2 && (k1 & kKindDead));      // run [s, p1) of one group, then nothing or positions that match nothing
  const int we = p1;
  int e = b1 ? p1 : p2;
  if (__any_sync(0xffffffffu, b0 && !shaped)) {            // rare: some other mix of groups -- find where the chunk really ends
    if (b0 && !shaped) {
      int p = we + 1;
      while (p < m && !(top.sync_start[((unsigned)w.meta[p - 1] << top.sync_shift) | (unsigned)w.meta[p]] & kSyncStart)) ++p;
      e = p;
    }
  }
  const int len = we - s;
  const bool inert = (k0 & kKindInert) != 0;
  const bool anchors_ok = !(first && s == 0 && !(k0 & kKindCaretOk)) && !(at_end && we == m && !(k0 & kKindDollarOk));
  const bool word = len <= (int)words.max_len && ((k0 & kKindWordRun) || (len == 1 && (k0 & kKindWordOne)));
  // (a chunk that is not final in this window goes to the loops, which clip it)
  const bool memo = shaped && (at_end || e <= limit) && (inert || (anchors_ok && word));
  const bool to_table = b0 && memo && !inert, to_loops = b0 && !memo;
  if (b0 && memo) w.carry = max(w.carry, e);
  const unsigned lt = lanemask_lt();
  const unsigned bf = __ballot_sync(0xffffffffu, to_table), bs = __ballot_sync(0xffffffffu, to_loops);
  if (to_table) w.fastq[(w.fh + w.nfast + __popc(bf & lt)) & (kRing - 1)] = (uint32_t)s | ((uint32_t)len << 16);
  if (to_loops) w.slowq[(w.sh + w.nslow + __popc(bs & lt)) & (kRing - 1)] = (uint32_t)s | ((uint32_t)e << 16);
  w.nfast += __popc(bf); w.nslow += __popc(bs);
  __syncwarp();
}

// One word per lane through the table; what it does not hold goes to the loops (and from there into the table).
__device__ __forceinline__ void fast_round(WarpWin& w, const WpWords& words, int cnt, int lane, int unk_id) {
  int s = 0, L = 0;
  if (lane < cnt) {
    const uint32_t ent = w.fastq[(w.fh + lane) & (kRing - 1)];
    s = (int)(ent & 0xFFFFu); L = (int)(ent >> 16);
  }
  const int lcap = __reduce_max_sync(0xffffffffu, L);
  const bool wide = lcap > (int)(4 * words.cpw);     // some word of the round needs the upper half of the key
  uint32_t kw[8];
  wp_pack_key_any(words.cpw, w.cls + s, L, lcap, words.cb, kw);
  const WpWordHit hit = wp_words_find(words, kw, wide);
  const bool found = lane < cnt && hit.meta != 0;
  if (found) wp_apply_hit(hit, s, unk_id, w.ids_at);
  const bool miss = lane < cnt && !found;
  const unsigned bm = __ballot_sync(0xffffffffu, miss);
  if (miss) w.slowq[(w.sh + w.nslow + __popc(bm & lanemask_lt())) & (kRing - 1)] = (uint32_t)s | 0x8000u | ((uint32_t)(s + L) << 16);
  w.nslow += __popc(bm);
  w.fh = (w.fh + cnt) & (kRing - 1); w.nfast -= cnt;
  __syncwarp();
}

// One chunk per lane through the reference's loops (wp_core.cuh).  A word run the table did not hold is resolved by
// the function sub-grammar alone -- its top-level outcome is the memo's -- and then added to the table.
template <typename TE>
__device__ __forceinline__ void slow_round(WarpWin& w, const WpTop& top, const WpGlobal<TE>& g, const WpWords& words, int cnt, int lane,
                                           int m, bool at_end, int limit, bool first, int unk_id) {
  if (lane < cnt) {
    const uint32_t ent = w.slowq[(w.sh + lane) & (kRing - 1)];
    const int s = (int)(ent & 0x7FFFu);
    int fe = (int)(ent >> 16);
    if (ent & 0x8000u) {
      const unsigned tc = w.meta[s];
      const int tiled = wp_word<TE>(g, w.cls, s, fe - 1, top.fn_root_of_tc[tc], top.fn_caret_of_tc[tc], w.ids_at);
      int n = 0, offs[kMaxLearnPieces];
      int32_t ids[kMaxLearnPieces];
      if (!tiled) {                      // not covered without gaps -> one UnkId (blingfiretokdll.cpp:1282-1301)
        for (int p = s + 1; p < fe; ++p) w.ids_at[p] = kNoPiece;
        w.ids_at[s] = unk_id;
      } else {
        for (int p = s; p < fe; ++p) {
          const int32_t id = w.ids_at[p];
          if (id != kNoPiece) { if (n < kMaxLearnPieces) { ids[n] = id; offs[n] = p - s; } ++n; }
        }
      }
      if (n <= kMaxLearnPieces) {
        uint32_t kw[8];
        wp_pack_key_any(words.cpw, w.cls + s, fe - s, fe - s, words.cb, kw);
        wp_words_insert(words, kw, fe - s > (int)(4 * words.cpw), n, ids, offs);
      }
    } else {
      if (fe > limit) fe = limit;
      const int fb = (s == 0 && first) ? -1 : s;
      if (fb < fe) w.carry = max(w.carry, wp_chunk<TE>(top, g, w.cls, m, at_end, fb, fe, unk_id, w.ids_at, w.meta));
    }
  }
  w.sh = (w.sh + cnt) & (kRing - 1); w.nslow -= cnt;
  __syncwarp();
}

// Runs full rounds (32 entries) of the two rings; with `all`, empties them.  A ring holds at most 63 entries: the
// loops' ring is brought below 32 before a table round can add its misses.
template <typename TE>
__device__ __forceinline__ void drain_rounds(WarpWin& w, const WpTop& top, const WpGlobal<TE>& g, const WpWords& words, int lane, int m,
                                             bool at_end, int limit, bool first, int unk_id, bool all) {
  for (;;) {
    if (w.nslow >= 32 || (all && w.nslow > 0 && w.nfast == 0))
      slow_round<TE>(w, top, g, words, w.nslow < 32 ? w.nslow : 32, lane, m, at_end, limit, first, unk_id);
    else if (w.nfast >= 32 || (all && w.nfast > 0))
      fast_round(w, words, w.nfast < 32 ? w.nfast : 32, lane, unk_id);
    else
      break;
  }
}

template <typename TE>
__global__ void __launch_bounds__(kThreads, 2) wp_tokenize_kernel(const WpLaunch p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned blob_bytes = (p.layout.total_bytes + 127u) & ~127u;
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  uint8_t* smem = simt::shared_base();
  if (threadIdx.x == 0) std::memcpy(smem, p.blob, p.layout.total_bytes);
  __syncthreads();
#else
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t blob_bar;

  // stage the model blob: one bulk async copy per CTA, completion through an mbarrier
  if (threadIdx.x == 0) {
    mbar_init(&blob_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&blob_bar, p.layout.total_bytes);
    bulk_copy_g2s(smem, p.blob, p.layout.total_bytes, &blob_bar);
  }
  mbar_wait(&blob_bar, 0);
#endif

  const WpTop top = make_wp_top(smem, p.layout);
  WpGlobal<TE> g;
  g.trans = reinterpret_cast<const TE*>(p.trans);
  g.tag_of_state = p.tag_of_state;
  g.NC1 = p.NC1; g.first_final = p.first_final; g.cls_caret = p.cls_caret; g.cls_dollar = p.cls_dollar;
  g.max_token_length = p.max_token_length;
  const WpWords words = p.words;

  uint8_t* wbase = smem + blob_bytes + (size_t)warp * kWarpSmem;
  WarpWin w;
  w.ev = reinterpret_cast<uint16_t*>(wbase + kOffEv);
  w.fastq = reinterpret_cast<uint32_t*>(wbase + kOffFastq);
  w.slowq = reinterpret_cast<uint32_t*>(wbase + kOffSlowq);

  const uint32_t* text32 = reinterpret_cast<const uint32_t*>(p.text);
  const int max_tok = p.max_token_length;
  const unsigned sh = top.sync_shift;

  for (;;) {
    unsigned long long d64 = 0;
    if (lane == 0) d64 = atomicAdd(p.work_counter, 1ull);
    d64 = __shfl_sync(0xffffffffu, d64, 0);
    if ((int64_t)d64 >= p.ndocs) break;
    const int64_t doc = (int64_t)d64;
    const int64_t lo = __ldg(p.offsets + doc);
    const int64_t n64 = __ldg(p.offsets + doc + 1) - lo;
    int result = 0;

    // parameter validation (blingfiretokdll.cpp:1121).  From here on offsets are relative to the aligned word the
    // document starts in: the document is the bytes [a, uend) of the words tw[0..].
    bool run = n64 > 0 && n64 <= 1000000000;
    const uint32_t* tw = text32 + (lo >> 2);
    int a = (int)(lo & 3);
    const int uend = a + (int)(run ? n64 : 0);
    if (run && n64 >= 3) {                 // the BOM (FAUtf8Utils.cpp:247-252)
      const uint32_t f0 = __ldg(tw), f1 = a + 3 > 4 ? __ldg(tw + 1) : 0u;
      const uint32_t b = (uint32_t)((((uint64_t)f1 << 32) | f0) >> (8 * a)) & 0xFFFFFFu;
      if (b == 0xBFBBEFu) a += 3;
    }
    run = run && a < uend;

    // Documents that may not fit one window are validated up front, because ids are
    // emitted window by window and an invalid byte anywhere must yield 0 ids.
    // a document of up to kWin-3 bytes always fits one window (its code points <= its bytes)
    const bool multi = run && (uend - a) > (kWin - 4);
    if (multi) {
      unsigned bad = 0, sumlen = 0;
      for (int ub = 0; ub < uend; ub += kBlockBytes) {
        const int u0 = ub + lane * 4;
        const uint32_t w0 = u0 < uend ? __ldg(tw + (u0 >> 2)) : 0u;
        const uint32_t w1 = u0 + 4 < uend ? __ldg(tw + (u0 >> 2) + 1) : 0u;
        const LaneDecode dcd = decode_lane(w0, w1, u0, a, uend);
        bad |= dcd.bad; sumlen += dcd.sumlen;
      }
      if (__any_sync(0xffffffffu, bad != 0) || (int)__reduce_add_sync(0xffffffffu, sumlen) != uend - a) run = false;
    }

    if (run) {
      // the window starts (a & 3) entries into its arrays: a lane's 4 bytes of ASCII text then land on a
      // 4-entry boundary, as long as everything before them was ASCII too
      const int pad = a & 3;
      w.ids_at = reinterpret_cast<int32_t*>(wbase) + pad;
      w.cls = reinterpret_cast<uint16_t*>(wbase + kOffCls) + pad;
      w.meta = wbase + kOffMeta + pad;
      int m = 0, out = 0, ub = 0;
      bool first = true, ok = true;
      unsigned sumlen = 0;          // per lane: bytes covered by the sequences decoded here
      int ulen = 0;                 // uniform: bytes of the all-ASCII blocks
      int32_t* row = p.ids + doc * (int64_t)p.max_ids;
      for (;;) {
        // ---- fill the window: decode, validate, classify, and list the events ----
        unsigned bad = 0;
        unsigned prev_tc = 0;
        w.nev = 0;
        if (m > 0) {                // the tail kept from the previous window
          __syncwarp();
          gen_events(w, top, 0, m, lane);
          prev_tc = w.meta[m - 1];
        }
        while (ub < uend) {
          const int blo = ub > a ? ub : a, bhi = ub + kBlockBytes < uend ? ub + kBlockBytes : uend;
          const int nvalid = bhi - blo;        // the block yields at most that many code points
          if (m + nvalid > kWin) break;
          const int u0 = ub + lane * 4;
          const bool full = ub >= a && ub + kBlockBytes <= uend;
          const uint32_t w0 = (full || (u0 < uend && u0 + 4 > a)) ? __ldg(tw + (u0 >> 2)) : 0u;
          uint32_t vmask = 0xFFFFFFFFu;        // this lane's bytes that belong to the document
          if (!full) {
            const int vfirst = a - u0, vlast = uend - u0;
            vmask = vlast >= 4 ? 0xFFFFFFFFu : (vlast <= 0 ? 0u : (1u << (8 * vlast)) - 1u);
            if (vfirst > 0) vmask = vfirst >= 4 ? 0u : (vmask & ~((1u << (8 * vfirst)) - 1u));
          }
          if (!__any_sync(0xffffffffu, (w0 & 0x80808080u & vmask) != 0)) {
            // all ASCII: position = byte offset; classes from the shared-memory table; events in registers
            const int p0 = m + (u0 - blo);     // window position of this lane's first byte
            const uint32_t x0 = top.ascii_clsx[w0 & 0x7F], x1 = top.ascii_clsx[(w0 >> 8) & 0x7F];
            const uint32_t x2 = top.ascii_clsx[(w0 >> 16) & 0x7F], x3 = top.ascii_clsx[(w0 >> 24) & 0x7F];
            const unsigned t0 = x0 >> 16, t1 = x1 >> 16, t2 = x2 >> 16, t3 = x3 >> 16;
            unsigned tp = __shfl_up_sync(0xffffffffu, t3, 1);
            if (lane == 0) tp = prev_tc;
            unsigned s0 = top.sync_start[(tp << sh) | t0], s1 = top.sync_start[(t0 << sh) | t1];
            unsigned s2 = top.sync_start[(t1 << sh) | t2], s3 = top.sync_start[(t2 << sh) | t3];
            if (m == 0) {                      // the first position of a window is a chunk start
              const int k = blo - u0;
              if (k == 0) s0 = kSyncStart; else if (k == 1) s1 = kSyncStart; else if (k == 2) s2 = kSyncStart; else if (k == 3) s3 = kSyncStart;
            }
            if (full && ((pad + m) & 3) == 0) {
              *reinterpret_cast<uint2*>(w.cls + p0) = make_uint2(__byte_perm(x0, x1, 0x5410), __byte_perm(x2, x3, 0x5410));
              *reinterpret_cast<uint32_t*>(w.meta + p0) = __byte_perm(__byte_perm(x0, x1, 0x0062), __byte_perm(x2, x3, 0x0062), 0x5410);
              *reinterpret_cast<int4*>(w.ids_at + p0) = make_int4(kNoPiece, kNoPiece, kNoPiece, kNoPiece);
            } else if (full) {
              w.cls[p0] = (uint16_t)x0; w.cls[p0 + 1] = (uint16_t)x1; w.cls[p0 + 2] = (uint16_t)x2; w.cls[p0 + 3] = (uint16_t)x3;
              w.meta[p0] = (uint8_t)t0; w.meta[p0 + 1] = (uint8_t)t1; w.meta[p0 + 2] = (uint8_t)t2; w.meta[p0 + 3] = (uint8_t)t3;
              w.ids_at[p0] = kNoPiece; w.ids_at[p0 + 1] = kNoPiece; w.ids_at[p0 + 2] = kNoPiece; w.ids_at[p0 + 3] = kNoPiece;
            } else {
              if (vmask & 0xFFu) { w.cls[p0] = (uint16_t)x0; w.meta[p0] = (uint8_t)t0; w.ids_at[p0] = kNoPiece; } else s0 = 0;
              if (vmask & 0xFF00u) { w.cls[p0 + 1] = (uint16_t)x1; w.meta[p0 + 1] = (uint8_t)t1; w.ids_at[p0 + 1] = kNoPiece; } else s1 = 0;
              if (vmask & 0xFF0000u) { w.cls[p0 + 2] = (uint16_t)x2; w.meta[p0 + 2] = (uint8_t)t2; w.ids_at[p0 + 2] = kNoPiece; } else s2 = 0;
              if (vmask & 0xFF000000u) { w.cls[p0 + 3] = (uint16_t)x3; w.meta[p0 + 3] = (uint8_t)t3; w.ids_at[p0 + 3] = kNoPiece; } else s3 = 0;
            }
            const unsigned sx = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
            const int cnt = __popc((sx | (sx >> 1)) & 0x01010101u);
            const int incl = warp_incl_scan(cnt, lane);
            int ei = w.nev + incl - cnt;
            if (s0) w.ev[ei++] = (uint16_t)((unsigned)p0 | ((s0 & kSyncStart) << 15));
            if (s1) w.ev[ei++] = (uint16_t)((unsigned)(p0 + 1) | ((s1 & kSyncStart) << 15));
            if (s2) w.ev[ei++] = (uint16_t)((unsigned)(p0 + 2) | ((s2 & kSyncStart) << 15));
            if (s3) w.ev[ei++] = (uint16_t)((unsigned)(p0 + 3) | ((s3 & kSyncStart) << 15));
            w.nev += __shfl_sync(0xffffffffu, incl, 31);
            prev_tc = __shfl_sync(0xffffffffu, t3, 31);   // (a block that ends before its lane 31 is the document's last)
            m += nvalid;
            ulen += nvalid;
          } else {
            const uint32_t w1 = u0 + 4 < uend ? __ldg(tw + (u0 >> 2) + 1) : 0u;
            const LaneDecode dcd = decode_lane(w0, w1, u0, blo, uend);
            bad |= dcd.bad; sumlen += dcd.sumlen;
            const int cnt = __popc(dcd.start_mask);
            const int incl = warp_incl_scan(cnt, lane);
            int idx = m + incl - cnt;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dcd.start_mask & (1u << k)) {
                const uint32_t cp = dcd.cp[k];
                const uint32_t x = cp < 128 ? top.ascii_clsx[cp] : __ldg(p.clsx_of_cp + cp);
                w.cls[idx] = (uint16_t)x;
                w.meta[idx] = (uint8_t)(x >> 16);
                w.ids_at[idx] = kNoPiece;
                ++idx;
              }
            }
            const int m1 = m + __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            gen_events(w, top, m, m1, lane);
            if (m1 > 0) prev_tc = w.meta[m1 - 1];
            m = m1;
          }
          ub += kBlockBytes;
        }
        const bool at_end = ub >= uend;
        if (!multi) {
          // single-window document: validity is known only now, before anything is emitted
          if (__any_sync(0xffffffffu, bad != 0) || (at_end && (int)__reduce_add_sync(0xffffffffu, sumlen) + ulen != uend - a)) { ok = false; break; }
        }
        if (m == 0) break;
        // the end of the window closes the last chunk (twice: every event looks at two successors)
        if (lane < 2) w.ev[w.nev + lane] = (uint16_t)((unsigned)m | 0x8000u);
        __syncwarp();

        // ---- events -> chunks -> ids ----
        const int limit = at_end ? m : m - max_tok;
        w.nfast = w.fh = w.nslow = w.sh = 0;
        w.carry = 0;
        for (int i0 = 0; i0 < w.nev; i0 += 32) {
          const int cnt = w.nev - i0 < 32 ? w.nev - i0 : 32;
          classify_round(w, top, words, i0, cnt, lane, m, at_end, limit, first);
          drain_rounds<TE>(w, top, g, words, lane, m, at_end, limit, first, p.unk_id, i0 + 32 >= w.nev);
        }
        int carry = at_end ? m : __reduce_max_sync(0xffffffffu, w.carry);
        __syncwarp();

        // ---- ordered compaction of the ids of positions [0, carry), four (array-aligned) entries per lane ----
        const int32_t* ids_base = w.ids_at - pad;
        for (int pp = 0; pp < pad + carry; pp += 128) {
          const int ph = pp + 4 * lane;
          int4 v = make_int4(kNoPiece, kNoPiece, kNoPiece, kNoPiece);
          if (ph < pad + carry) v = *reinterpret_cast<const int4*>(ids_base + ph);
          const int q = ph - pad;                // window position of v.x
          const bool f0 = v.x != kNoPiece && q >= 0 && q < carry, f1 = v.y != kNoPiece && q + 1 >= 0 && q + 1 < carry;
          const bool f2 = v.z != kNoPiece && q + 2 >= 0 && q + 2 < carry, f3 = v.w != kNoPiece && q + 3 < carry;
          const int cnt = (int)f0 + (int)f1 + (int)f2 + (int)f3;
          const int incl = warp_incl_scan(cnt, lane);
          const int r0 = out + incl - cnt, r1 = r0 + (int)f0, r2 = r1 + (int)f1, r3 = r2 + (int)f2;
          if (f0 && r0 < p.max_ids) row[r0] = v.x;
          if (f1 && r1 < p.max_ids) row[r1] = v.y;
          if (f2 && r2 < p.max_ids) row[r2] = v.z;
          if (f3 && r3 < p.max_ids) row[r3] = v.w;
          out += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (at_end) break;

        // ---- keep the unprocessed tail [carry, m) and refill ----
        const int rest = m - carry;
        for (int k0 = 0; k0 < rest; k0 += 32) {
          const int k = k0 + lane;
          const uint16_t v = k < rest ? w.cls[carry + k] : (uint16_t)0;
          const uint8_t mt = k < rest ? w.meta[carry + k] : (uint8_t)0;
          __syncwarp();
          if (k < rest) { w.cls[k] = v; w.meta[k] = mt; w.ids_at[k] = kNoPiece; }
        }
        __syncwarp();
        m = rest;
        first = false;
      }
      if (ok) result = out < p.max_ids ? out : p.max_ids;
    }
    if (lane == 0) p.counts[doc] = result;
    __syncwarp();
  }
}

#ifndef BF_SIMT_HOST
template <typename OutT>
__global__ void wp_compact_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ counts,
                                  const int64_t* __restrict__ row_off, int64_t ndocs, int max_ids,
                                  OutT* __restrict__ csr) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = warp; d < ndocs; d += nwarps) {
    const int c = counts[d];
    const int32_t* src = ids + d * (int64_t)max_ids;
    OutT* dst = csr + row_off[d];
    for (int k = lane; k < c; k += 32) dst[k] = (OutT)src[k];
  }
}

// Exclusive prefix sum of per-document counts (int32 -> int64), one CTA: the input is a few ten thousand items per
// chunk and sits between two much longer kernels.  16 consecutive items per thread (two int4 loads), a warp scan and a
// scan of the warp totals per 16 384 items.
constexpr int kScanPerThread = 16;
__global__ void __launch_bounds__(1024) wp_scan_kernel(const int32_t* __restrict__ counts, int64_t* __restrict__ row_off, int64_t n) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024 * kScanPerThread) {
    const int64_t i0 = base + (int64_t)threadIdx.x * kScanPerThread;
    int32_t v[kScanPerThread];
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) v[k] = i0 + k < n ? counts[i0 + k] : 0;
    int64_t mine = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) mine += v[k];
    int64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int64_t run = carry_s + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - mine;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
      if (i0 + k < n) row_off[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_off[n] = carry_s;
}

#endif  // BF_SIMT_HOST

}  // namespace

#ifndef BF_SIMT_HOST
cudaError_t wp_tokenize_launch(const WpLaunch& p, cudaStream_t stream, WpLaunchInfo* info) {
  const size_t blob_bytes = ((size_t)p.layout.total_bytes + 127) & ~(size_t)127;
  const size_t smem = blob_bytes + (size_t)kWarpsPerCta * kWarpSmem;
  auto kern = p.wide ? wp_tokenize_kernel<uint32_t> : wp_tokenize_kernel<uint16_t>;
  // the launch geometry depends on (device, kernel variant, smem size) only: resolve it once
  struct Geo { int dev = -1; size_t smem = 0; int grid = 0; };
  static thread_local Geo cache[2];
  Geo& geo = cache[p.wide ? 1 : 0];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (geo.dev != dev || geo.smem != smem) {
    int sms = 0, per_sm = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    geo.dev = dev; geo.smem = smem; geo.grid = sms * per_sm;
  }
  // persistent grid: a whole number of CTAs per SM
  int grid = geo.grid;
  const int64_t needed = (p.ndocs + kWarpsPerCta - 1) / kWarpsPerCta;
  if (needed < grid) grid = (int)(needed > 0 ? needed : 1);
  e = cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  kern<<<grid, kThreads, smem, stream>>>(p);
  if (info) { info->grid = grid; info->block = kThreads; info->smem_bytes = smem; info->launches = 1; }
  return cudaGetLastError();
}

template <typename OutT>
static cudaError_t compact_launch(const int32_t* ids, const int32_t* counts, const int64_t* row_off, int64_t ndocs,
                                  int max_ids, OutT* csr, cudaStream_t stream) {
  if (ndocs <= 0) return cudaSuccess;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int block = 256;
  int64_t grid = (ndocs * 32 + block - 1) / block;
  if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
  wp_compact_kernel<OutT><<<(int)grid, block, 0, stream>>>(ids, counts, row_off, ndocs, max_ids, csr);
  return cudaGetLastError();
}
cudaError_t wp_compact_launch(const int32_t* ids, const int32_t* counts, const int64_t* row_off, int64_t ndocs,
                              int max_ids, int32_t* csr, cudaStream_t stream) {
  return compact_launch<int32_t>(ids, counts, row_off, ndocs, max_ids, csr, stream);
}
cudaError_t wp_compact_launch_u16(const int32_t* ids, const int32_t* counts, const int64_t* row_off, int64_t ndocs,
                                  int max_ids, uint16_t* csr, cudaStream_t stream) {
  return compact_launch<uint16_t>(ids, counts, row_off, ndocs, max_ids, csr, stream);
}

cudaError_t wp_scan_counts(const int32_t* counts, int64_t* row_off, int64_t ndocs, cudaStream_t stream) {
  wp_scan_kernel<<<1, 1024, 0, stream>>>(counts, row_off, ndocs);
  return cudaGetLastError();
}

#endif  // BF_SIMT_HOST

}  // namespace bfb200
