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
constexpr int kWarpSmem = kWin * (2 + 4 + 1 + 2 + 2);   // cls u16, ids_at i32, top class u8, starts u16, order u16

static_assert(kWin % 32 == 0, "window must be a multiple of the warp size");

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

__device__ __forceinline__ void load_words(const uint32_t* text32, int64_t pos0, int64_t padded_bytes, uint32_t* w0, uint32_t* w1) {
  *w0 = pos0 < padded_bytes ? __ldg(text32 + (pos0 >> 2)) : 0u;
  *w1 = pos0 + 4 < padded_bytes ? __ldg(text32 + (pos0 >> 2) + 1) : 0u;
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
  g.cls_of_cp = p.cls_of_cp;
  g.NC1 = p.NC1; g.first_final = p.first_final; g.cls_caret = p.cls_caret; g.cls_dollar = p.cls_dollar;
  g.max_token_length = p.max_token_length;

  uint8_t* wbase = smem + blob_bytes + (size_t)warp * kWarpSmem;
  int32_t* ids_at = reinterpret_cast<int32_t*>(wbase);
  uint16_t* cls = reinterpret_cast<uint16_t*>(wbase + kWin * 4);
  uint16_t* starts = reinterpret_cast<uint16_t*>(wbase + kWin * 6);
  uint8_t* meta = wbase + kWin * 8;          // top-level class of every position
  uint16_t* order = reinterpret_cast<uint16_t*>(wbase + kWin * 9 + (kWin & 1));

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
        uint32_t w0, w1;
        load_words(text32, pos0, padded_bytes, &w0, &w1);
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
      int m = 0, out = 0;
      int64_t bpos = lo0;
      bool first = true, ok = true;
      unsigned sumlen = 0;
      int32_t* row = p.ids + doc * (int64_t)p.max_ids;
      for (;;) {
        // ---- fill the window: decode, validate, classify, compact ----
        unsigned bad = 0;
        while (bpos < hi) {
          const int64_t bs = bpos & ~(int64_t)3;
          // the block [bpos, min(bs+128, hi)) yields at most that many code points
          const int64_t blk_end = bs + kBlockBytes < hi ? bs + kBlockBytes : hi;
          if (m + (int)(blk_end - bpos) > kWin) break;
          const int64_t pos0 = bs + lane * 4;
          uint32_t w0, w1;
          load_words(text32, pos0, padded_bytes, &w0, &w1);
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
              const uint16_t c = cp < 128 ? top.ascii_cls[cp] : __ldg(g.cls_of_cp + cp);
              cls[idx] = c;
              meta[idx] = top.tc_of_class[c];
              ids_at[idx] = kNoPiece;
              ++idx;
            }
          }
          m += __shfl_sync(0xffffffffu, incl, 31);
          bpos = bs + kBlockBytes;
        }
        const bool at_end = bpos >= hi;
        if (!multi) {
          // single-window document: validity is known only now, before anything is emitted
          unsigned tot = sumlen;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
          if (__any_sync(0xffffffffu, bad != 0) || (at_end && (int64_t)tot != hi - lo0)) { ok = false; break; }
        }
        if (m == 0) break;
        __syncwarp();

        // ---- sync points -> chunk starts ----
        int nst = 0;
        for (int p0 = 0; p0 < m; p0 += 32) {
          const int q = p0 + lane;
          bool flag = false;
          if (q < m) {
            flag = q == 0 || top.sync_start[((unsigned)meta[q - 1] << top.sync_shift) | (unsigned)meta[q]] != 0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, flag);
          if (flag) starts[nst + __popc(bal & lanemask_lt())] = (uint16_t)q;
          nst += __popc(bal);
        }
        __syncwarp();

        // ---- order the chunks by length class (counting sort over the chunk list), so that the 32
        // chunks of a round cost about the same and the lanes of the warp stay together ----
        const int limit = at_end ? m : m - max_tok;
        int carry = at_end ? m : 0;
        {
          int cnt0 = 0, cnt1 = 0, cnt2 = 0;     // chunks of length <= 2, 3..4, 5..7 (the rest is class 3)
          for (int base = 0; base < nst; base += 32) {
            const int i = base + lane;
            int cl = 4;
            if (i < nst) { const int len = (i + 1 < nst ? (int)starts[i + 1] : m) - (int)starts[i]; cl = len <= 2 ? 0 : len <= 4 ? 1 : len <= 7 ? 2 : 3; }
            cnt0 += __popc(__ballot_sync(0xffffffffu, cl == 0));
            cnt1 += __popc(__ballot_sync(0xffffffffu, cl == 1));
            cnt2 += __popc(__ballot_sync(0xffffffffu, cl == 2));
          }
          int o0 = 0, o1 = cnt0, o2 = cnt0 + cnt1, o3 = cnt0 + cnt1 + cnt2;
          for (int base = 0; base < nst; base += 32) {
            const int i = base + lane;
            int cl = 4;
            if (i < nst) { const int len = (i + 1 < nst ? (int)starts[i + 1] : m) - (int)starts[i]; cl = len <= 2 ? 0 : len <= 4 ? 1 : len <= 7 ? 2 : 3; }
            const unsigned b0 = __ballot_sync(0xffffffffu, cl == 0), b1 = __ballot_sync(0xffffffffu, cl == 1);
            const unsigned b2 = __ballot_sync(0xffffffffu, cl == 2), b3 = __ballot_sync(0xffffffffu, cl == 3);
            const unsigned lt = lanemask_lt();
            if (cl == 0) order[o0 + __popc(b0 & lt)] = (uint16_t)i;
            else if (cl == 1) order[o1 + __popc(b1 & lt)] = (uint16_t)i;
            else if (cl == 2) order[o2 + __popc(b2 & lt)] = (uint16_t)i;
            else if (cl == 3) order[o3 + __popc(b3 & lt)] = (uint16_t)i;
            o0 += __popc(b0); o1 += __popc(b1); o2 += __popc(b2); o3 += __popc(b3);
          }
        }
        __syncwarp();

        // ---- one lane per chunk: the reference's loops, verbatim in structure ----
        for (int base = 0; base < nst; base += 32) {
          const int k = base + lane;
          if (k < nst) {
            const int i = order[k];
            int fb = starts[i];
            int fe = i + 1 < nst ? (int)starts[i + 1] : m;
            if (fe > limit) fe = limit;
            if (i == 0 && first) fb = -1;
            if (fb < fe) {
              const int r = wp_chunk<TE>(top, g, cls, m, at_end, fb, fe, p.unk_id, ids_at, meta);
              carry = max(carry, r);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) carry = max(carry, __shfl_xor_sync(0xffffffffu, carry, o));
        __syncwarp();

        // ---- ordered compaction of the ids of positions [0, carry) ----
        for (int p0 = 0; p0 < carry; p0 += 32) {
          const int q = p0 + lane;
          const int32_t id = q < carry ? ids_at[q] : kNoPiece;
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
          const uint16_t v = k < rest ? cls[carry + k] : (uint16_t)0;
          const uint8_t mt = k < rest ? meta[carry + k] : (uint8_t)0;
          __syncwarp();
          if (k < rest) { cls[k] = v; meta[k] = mt; ids_at[k] = kNoPiece; }
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
