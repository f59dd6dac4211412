// wp_kernel.cu -- fused TextToIds kernel for FastPath (flat two-level WordPiece) lexer models,
// hand-written for sm_100a.  Integer / indexing work only: no tensor cores.
//
// Replaces, per document, the reference's whole TextToIdsWithOffsets_wp pipeline
// (blingfiretokdll.cpp:1108-1314): FAStrUtf8ToArray -> FANormalize -> FALexTools_t::Process
// (with the nested FnTokWord call) -> tiling/UNK post-pass -> MaxIdsArrLength truncation.
//
// Mapping (DESIGN.md has the derivation and the exactness argument):
//   * persistent CTAs, one document per WARP at a time, documents handed out by an atomic
//     counter (ragged lengths balance themselves);
//   * the small hot part of the model (top-level automaton, class tables, the two hottest
//     transition rows) is staged ONCE per CTA into shared memory by a bulk async copy
//     (cp.async.bulk + mbarrier, the 1-D TMA path; SASS: UBLKCP);
//   * bytes are read as coalesced 32-bit words (uchar4 per lane, 128 B per warp step), UTF-8
//     is decoded and validated in registers, code points are compacted with a warp scan and
//     mapped to classes (ASCII from shared memory, the rest from the L2-resident class map);
//   * a ballot over the class pairs marks "sync points" no top-level match can cross; each
//     lane then owns one chunk between sync points and runs the reference's loops verbatim
//     (wp_core.cuh), gathering from the dense state x class table in HBM;
//   * ids are written position-indexed in shared memory and compacted in order with
//     ballot + popc into the output row.
#include "wp_kernel.cuh"

namespace bfb200 {

namespace {

constexpr int kWarpsPerCta = 16;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kWin = 576;            // window capacity in code points (per warp)
constexpr int kBlockBytes = 128;     // bytes consumed per decode step (32 lanes x uchar4)
constexpr int kPad = 4;              // the window starts (lo & 3) entries into its arrays, so that a lane's 4 ASCII bytes land on
                                     // a 4-entry boundary and go out as one vector store per array
constexpr int kRing = 64;            // entries per work ring (a round consumes 32; at most 63 are ever queued)
constexpr int kEvRing = 128;         // the event ring: a round needs 34 queued (every event looks at its two successors)
// per-warp shared memory: ids_at i32 | cls u16 (+ slack for the key reads past a chunk) | top class u8 | event ring u16 |
// fast ring u32 | slow ring u32
constexpr int kOffCls = (kWin + kPad) * 4;
constexpr int kOffMeta = kOffCls + (kWin + kPad + kMaxFastLen + 4) * 2;
constexpr int kOffEvq = (kOffMeta + kWin + kPad + 15) & ~15;
constexpr int kOffFastq = kOffEvq + kEvRing * 2;
constexpr int kOffSlowq = kOffFastq + kRing * 4;
constexpr int kWarpSmem = kOffSlowq + kRing * 4;

static_assert(kWin % 32 == 0, "window must be a multiple of the warp size");
static_assert(kOffCls % 16 == 0 && kOffMeta % 4 == 0 && kWarpSmem % 16 == 0, "vector stores need aligned arrays");
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
struct LaneDecode {
  uint32_t cp[4];
  unsigned start_mask;   // bit k: a sequence starts at byte k of the lane's word
  unsigned bad;          // nonzero: some sequence starting here is invalid
  unsigned sumlen;       // total length of the sequences that start here
};

__device__ __forceinline__ LaneDecode decode_lane(uint32_t w0, uint32_t w1, int64_t pos0, int64_t bpos, int64_t hi) {
  LaneDecode r;
  r.start_mask = 0; r.bad = 0; r.sumlen = 0;
  const uint64_t x = ((uint64_t)w1 << 32) | w0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t pos = pos0 + k;
    const uint32_t y = (uint32_t)(x >> (8 * k));
    const uint32_t b0 = y & 0xFF, b1 = (y >> 8) & 0xFF, b2 = (y >> 16) & 0xFF, b3 = y >> 24;
    r.cp[k] = 0;
    if (pos >= bpos && pos < hi && (b0 & 0xC0) != 0x80) {
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
      bad |= (pos + len > hi);
      r.bad |= bad;
      r.sumlen += len;
      r.cp[k] = bad ? 0u : cp;
      r.start_mask |= 1u << k;
    }
  }
  return r;
}

// The per-warp working set: one window of the current document and the three work rings.
struct WarpWin {
  int32_t* ids_at;     // [kWin] id of the piece starting at a position, kNoPiece elsewhere
  uint16_t* cls;       // [kWin] class of every position
  uint8_t* meta;       // [kWin] top-level class of every position
  uint16_t* evq;       // ring of events: position | 0x8000 when a chunk may start there (else: only the class changes)
  uint32_t* fastq;     // ring of word runs for the whole-word table: start | length << 16
  uint32_t* slowq;     // ring of chunks for the lexer loops: start | end << 16
  int nev, evh, nfast, fh, nslow, sh;   // ring counts and heads (warp-uniform)
  int carry;           // per lane: largest position known to be a finished chunk boundary
};

// Events -> chunks: every lane takes one event and its two successors.  A chunk start followed by another chunk
// start is a run inside ONE group of top-level classes (a change of group would be an event of its own); a chunk
// start, a change to the DEAD group, and the next chunk start is such a run followed by positions that match
// nothing ("word" + white space in bert_*).  The per-group memo says whether the run emits nothing, is one word for
// the table, or needs the loops -- as does every chunk of another shape.
template <typename TE>
__device__ __forceinline__ void classify_round(WarpWin& w, const WpTop& top, const WpWords& words, int cnt, int lane, int m,
                                               bool at_end, int limit, bool first) {
  int action = 0;      // 1 = table, 2 = loops
  int s = 0, we = 0, e = 0;
  if (lane < cnt) {
    const unsigned e0 = w.evq[(w.evh + lane) & (kEvRing - 1)], e1 = w.evq[(w.evh + lane + 1) & (kEvRing - 1)];
    s = (int)(e0 & 0x7FFFu);
    if (e0 & 0x8000u) {
      bool shaped = true;
      we = e = (int)(e1 & 0x7FFFu);
      if (!(e1 & 0x8000u)) {
        const unsigned e2 = w.evq[(w.evh + lane + 2) & (kEvRing - 1)];
        e = (int)(e2 & 0x7FFFu);
        shaped = (e2 & 0x8000u) && (top.kind_of_tc[w.meta[we]] & kKindDead);
        if (!shaped) {
          // some other mix of groups: find where the chunk really ends
          int p = we + 1;
          while (p < m && !(top.sync_start[((unsigned)w.meta[p - 1] << top.sync_shift) | (unsigned)w.meta[p]] & kSyncStart)) ++p;
          e = p;
        }
      }
      if (!shaped || (!at_end && e > limit)) {
        action = 2;                                  // (a chunk that is not final in this window: the loops clip it)
      } else {
        const int r = wp_classify_run(top.kind_of_tc[w.meta[s]], we - s, words.max_len, first && s == 0, at_end && we == m);
        if (r == 0) action = 2;
        else { action = r == 2 ? 1 : 0; w.carry = max(w.carry, e); }
      }
    }
  }
  const unsigned lt = lanemask_lt();
  const unsigned bf = __ballot_sync(0xffffffffu, action == 1), bs = __ballot_sync(0xffffffffu, action == 2);
  if (action == 1) w.fastq[(w.fh + w.nfast + __popc(bf & lt)) & (kRing - 1)] = (uint32_t)s | ((uint32_t)(we - s) << 16);
  if (action == 2) w.slowq[(w.sh + w.nslow + __popc(bs & lt)) & (kRing - 1)] = (uint32_t)s | ((uint32_t)e << 16);
  w.nfast += __popc(bf); w.nslow += __popc(bs);
  w.evh = (w.evh + cnt) & (kEvRing - 1); w.nev -= cnt;
  __syncwarp();
}

// One word per lane through the whole-word table; what the table does not hold goes to the loops.
__device__ __forceinline__ void fast_round(WarpWin& w, const WpWords& words, int cnt, int lane) {
  int s = 0, L = 0;
  if (lane < cnt) {
    const uint32_t ent = w.fastq[(w.fh + lane) & (kRing - 1)];
    s = (int)(ent & 0xFFFFu); L = (int)(ent >> 16);
  }
  const int lcap = __reduce_max_sync(0xffffffffu, L);
  uint32_t kw[4];
  wp_pack_key_any(words.cpw, w.cls + s, L, lcap, words.cb, kw);
  const int32_t id = wp_words_find(words, kw);
  const bool hit = lane < cnt && id != kNoPiece;
  if (hit) w.ids_at[s] = id;
  const bool miss = lane < cnt && !hit;
  const unsigned bm = __ballot_sync(0xffffffffu, miss);
  if (miss) w.slowq[(w.sh + w.nslow + __popc(bm & lanemask_lt())) & (kRing - 1)] = (uint32_t)s | ((uint32_t)(s + L) << 16);
  w.nslow += __popc(bm);
  w.fh = (w.fh + cnt) & (kRing - 1); w.nfast -= cnt;
  __syncwarp();
}

// One chunk per lane through the reference's loops (wp_core.cuh).
template <typename TE>
__device__ __forceinline__ void slow_round(WarpWin& w, const WpTop& top, const WpGlobal<TE>& g, int cnt, int lane, int m, bool at_end,
                                           int limit, bool first, int unk_id) {
  if (lane < cnt) {
    const uint32_t ent = w.slowq[(w.sh + lane) & (kRing - 1)];
    const int s = (int)(ent & 0xFFFFu);
    int fe = (int)(ent >> 16);
    if (fe > limit) fe = limit;
    const int fb = (s == 0 && first) ? -1 : s;
    if (fb < fe) w.carry = max(w.carry, wp_chunk<TE>(top, g, w.cls, m, at_end, fb, fe, unk_id, w.ids_at, w.meta));
  }
  w.sh = (w.sh + cnt) & (kRing - 1); w.nslow -= cnt;
  __syncwarp();
}

template <typename TE>
__device__ __forceinline__ void drain_rounds(WarpWin& w, const WpTop& top, const WpGlobal<TE>& g, const WpWords& words, int lane, int m,
                                             bool at_end, int limit, bool first, int unk_id, bool all) {
  // a ring holds at most 63 entries: the loops' ring is emptied before the table round can add 32 more
  if (w.nslow >= 32) slow_round<TE>(w, top, g, 32, lane, m, at_end, limit, first, unk_id);
  if (w.nfast >= 32) fast_round(w, words, 32, lane);
  if (w.nslow >= 32) slow_round<TE>(w, top, g, 32, lane, m, at_end, limit, first, unk_id);
  if (all) {
    if (w.nfast > 0) fast_round(w, words, w.nfast, lane);
    while (w.nslow > 0) slow_round<TE>(w, top, g, w.nslow < 32 ? w.nslow : 32, lane, m, at_end, limit, first, unk_id);
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
  w.evq = reinterpret_cast<uint16_t*>(wbase + kOffEvq);
  w.fastq = reinterpret_cast<uint32_t*>(wbase + kOffFastq);
  w.slowq = reinterpret_cast<uint32_t*>(wbase + kOffSlowq);

  const uint32_t* text32 = reinterpret_cast<const uint32_t*>(p.text);
  const int64_t padded_bytes = (p.text_bytes + 3) & ~(int64_t)3;
  const int max_tok = p.max_token_length;

  for (;;) {
    unsigned long long d64 = 0;
    if (lane == 0) d64 = atomicAdd(p.work_counter, 1ull);
    d64 = __shfl_sync(0xffffffffu, d64, 0);
    if ((int64_t)d64 >= p.ndocs) break;
    const int64_t doc = (int64_t)d64;
    int64_t lo = __ldg(p.offsets + doc);
    const int64_t hi = __ldg(p.offsets + doc + 1);
    const int64_t n = hi - lo;
    int result = 0;

    // parameter validation (blingfiretokdll.cpp:1121) and the BOM (FAUtf8Utils.cpp:247-252)
    bool run = n > 0 && n <= 1000000000;
    if (run && n >= 3) {
      const uint32_t b0 = __ldg(p.text + lo), b1 = __ldg(p.text + lo + 1), b2 = __ldg(p.text + lo + 2);
      if (b0 == 0xEF && b1 == 0xBB && b2 == 0xBF) lo += 3;
    }
    run = run && lo < hi;
    const int64_t lo0 = lo;

    // Documents that may not fit one window are validated up front, because ids are
    // emitted window by window and an invalid byte anywhere must yield 0 ids.
    // a document of up to kWin-3 bytes always fits one window (its code points <= its bytes; the
    // first decode block may start up to 3 bytes before the document)
    const bool multi = run && (hi - lo0) > (kWin - 4);
    if (multi) {
      unsigned bad = 0, sumlen = 0;
      for (int64_t bpos = lo0; bpos < hi;) {
        const int64_t bs = bpos & ~(int64_t)3;
        const int64_t pos0 = bs + lane * 4;
        const uint32_t w0 = pos0 < padded_bytes ? __ldg(text32 + (pos0 >> 2)) : 0u;
        const uint32_t w1 = pos0 + 4 < padded_bytes ? __ldg(text32 + (pos0 >> 2) + 1) : 0u;
        const LaneDecode dcd = decode_lane(w0, w1, pos0, bpos, hi);
        bad |= dcd.bad; sumlen += dcd.sumlen;
        bpos = bs + kBlockBytes;
      }
      bad = __any_sync(0xffffffffu, bad != 0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sumlen += __shfl_xor_sync(0xffffffffu, sumlen, o);
      if (bad || (int64_t)sumlen != hi - lo0) run = false;
    }

    if (run) {
      // the window starts (lo0 & 3) entries into its arrays: a lane's 4 bytes of ASCII text then land on a
      // 4-entry boundary, as long as everything before them was ASCII too
      const int pad = (int)(lo0 & 3);
      w.ids_at = reinterpret_cast<int32_t*>(wbase) + pad;
      w.cls = reinterpret_cast<uint16_t*>(wbase + kOffCls) + pad;
      w.meta = wbase + kOffMeta + pad;
      int m = 0, out = 0;
      int64_t bpos = lo0;
      bool first = true, ok = true;
      unsigned sumlen = 0;          // per lane: bytes covered by the sequences decoded here
      int ulen = 0;                 // uniform: bytes of the all-ASCII blocks
      int32_t* row = p.ids + doc * (int64_t)p.max_ids;
      for (;;) {
        // ---- fill the window: decode, validate, classify ----
        unsigned bad = 0;
        while (bpos < hi) {
          const int64_t bs = bpos & ~(int64_t)3;
          // the block [bpos, min(bs+128, hi)) yields at most that many code points
          const int64_t blk_end = bs + kBlockBytes < hi ? bs + kBlockBytes : hi;
          if (m + (int)(blk_end - bpos) > kWin) break;
          const int64_t pos0 = bs + lane * 4;
          const uint32_t w0 = pos0 < padded_bytes ? __ldg(text32 + (pos0 >> 2)) : 0u;
          const int vfirst = (int)(bpos - pos0), vlast = (int)(blk_end - pos0);   // this lane's bytes [vfirst, vlast) belong to the block
          uint32_t vmask = vlast >= 4 ? 0xFFFFFFFFu : (vlast <= 0 ? 0u : (1u << (8 * vlast)) - 1u);
          if (vfirst > 0) vmask &= ~((1u << (8 * vfirst)) - 1u);
          if (!__any_sync(0xffffffffu, (w0 & 0x80808080u & vmask) != 0)) {
            // all ASCII: position = byte offset, classes from the shared-memory table
            const int idx = m + (int)(pos0 - bpos);
            const uint32_t x0 = top.ascii_clsx[w0 & 0x7F], x1 = top.ascii_clsx[(w0 >> 8) & 0x7F];
            const uint32_t x2 = top.ascii_clsx[(w0 >> 16) & 0x7F], x3 = top.ascii_clsx[(w0 >> 24) & 0x7F];
            if (vmask == 0xFFFFFFFFu && ((pad + idx) & 3) == 0) {
              *reinterpret_cast<uint2*>(w.cls + idx) = make_uint2(__byte_perm(x0, x1, 0x5410), __byte_perm(x2, x3, 0x5410));
              *reinterpret_cast<uint32_t*>(w.meta + idx) = __byte_perm(__byte_perm(x0, x1, 0x0062), __byte_perm(x2, x3, 0x0062), 0x5410);
              *reinterpret_cast<int4*>(w.ids_at + idx) = make_int4(kNoPiece, kNoPiece, kNoPiece, kNoPiece);
            } else {
              if (vmask & 0xFFu) { w.cls[idx] = (uint16_t)x0; w.meta[idx] = (uint8_t)(x0 >> 16); w.ids_at[idx] = kNoPiece; }
              if (vmask & 0xFF00u) { w.cls[idx + 1] = (uint16_t)x1; w.meta[idx + 1] = (uint8_t)(x1 >> 16); w.ids_at[idx + 1] = kNoPiece; }
              if (vmask & 0xFF0000u) { w.cls[idx + 2] = (uint16_t)x2; w.meta[idx + 2] = (uint8_t)(x2 >> 16); w.ids_at[idx + 2] = kNoPiece; }
              if (vmask & 0xFF000000u) { w.cls[idx + 3] = (uint16_t)x3; w.meta[idx + 3] = (uint8_t)(x3 >> 16); w.ids_at[idx + 3] = kNoPiece; }
            }
            m += (int)(blk_end - bpos);
            ulen += (int)(blk_end - bpos);
          } else {
            const uint32_t w1 = pos0 + 4 < padded_bytes ? __ldg(text32 + (pos0 >> 2) + 1) : 0u;
            const LaneDecode dcd = decode_lane(w0, w1, pos0, bpos, hi);
            bad |= dcd.bad; sumlen += dcd.sumlen;
            const int cnt = __popc(dcd.start_mask);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int v = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += v;
            }
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
            m += __shfl_sync(0xffffffffu, incl, 31);
          }
          bpos = bs + kBlockBytes;
        }
        const bool at_end = bpos >= hi;
        if (!multi) {
          // single-window document: validity is known only now, before anything is emitted
          unsigned tot = sumlen;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
          if (__any_sync(0xffffffffu, bad != 0) || (at_end && (int64_t)tot + ulen != hi - lo0)) { ok = false; break; }
        }
        if (m == 0) break;
        __syncwarp();

        // ---- events: positions where a chunk may start (sync points) or the top-level class changes ----
        const int limit = at_end ? m : m - max_tok;
        w.nev = w.evh = w.nfast = w.fh = w.nslow = w.sh = 0;
        w.carry = 0;
        for (int p0 = 0; p0 < m; p0 += 32) {
          const int q = p0 + lane;
          bool ev = false;
          unsigned entry = 0;
          if (q < m) {
            const unsigned tc = w.meta[q];
            const unsigned tp = q > 0 ? (unsigned)w.meta[q - 1] : tc;
            const unsigned sv = q == 0 ? (unsigned)kSyncStart : (unsigned)top.sync_start[(tp << top.sync_shift) | tc];
            ev = sv != 0;
            entry = (unsigned)q | ((sv & kSyncStart) ? 0x8000u : 0u);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, ev);
          if (ev) w.evq[(w.evh + w.nev + __popc(bal & lanemask_lt())) & (kEvRing - 1)] = (uint16_t)entry;
          w.nev += __popc(bal);
          __syncwarp();
          if (w.nev >= 34) {         // 32 events with their successors
            classify_round<TE>(w, top, words, 32, lane, m, at_end, limit, first);
            drain_rounds<TE>(w, top, g, words, lane, m, at_end, limit, first, p.unk_id, false);
          }
        }
        // the end of the window closes the last chunk (twice: every event looks at two successors)
        if (lane < 2) w.evq[(w.evh + w.nev + lane) & (kEvRing - 1)] = (uint16_t)((unsigned)m | 0x8000u);
        w.nev += 2;
        __syncwarp();
        while (w.nev >= 3) {
          classify_round<TE>(w, top, words, w.nev - 2 < 32 ? w.nev - 2 : 32, lane, m, at_end, limit, first);
          drain_rounds<TE>(w, top, g, words, lane, m, at_end, limit, first, p.unk_id, w.nev < 3);
        }
        int carry = at_end ? m : w.carry;
        if (!at_end) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) carry = max(carry, __shfl_xor_sync(0xffffffffu, carry, o));
        }
        __syncwarp();

        // ---- ordered compaction of the ids of positions [0, carry) ----
        for (int p0 = 0; p0 < carry; p0 += 32) {
          const int q = p0 + lane;
          const int32_t id = q < carry ? w.ids_at[q] : kNoPiece;
          const bool f = id != kNoPiece;
          const unsigned bal = __ballot_sync(0xffffffffu, f);
          const int rank = out + __popc(bal & lanemask_lt());
          if (f && rank < p.max_ids) row[rank] = id;
          out += __popc(bal);
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

// Exclusive prefix sum of per-document counts (int32 -> int64), one CTA: the input is a few
// hundred thousand items per chunk and sits between two much longer kernels.
__global__ void __launch_bounds__(1024) wp_scan_kernel(const int32_t* __restrict__ counts, int64_t* __restrict__ row_off, int64_t n) {
  __shared__ int64_t warp_sums[32];
  __shared__ int64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n ? (int64_t)counts[i] : 0;
    int64_t incl = v;
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
    const int64_t carry = carry_s;
    const int64_t prefix = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - v;
    if (i < n) row_off[i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = prefix + v;
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
