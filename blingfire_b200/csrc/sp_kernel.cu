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

#include "sp_common.cuh"
#include "sp_generic.cuh"
#include "sp_bpe.cuh"       // (b_step / b_farthest are shared with the Unigram cuts)
#include "sp_unigram.cuh"

namespace bfb200 {

namespace {

// BPE family: one document per warp
__global__ void __launch_bounds__(kBWarps * 32, kBCtasPerSm) sp_bpe_kernel(const SpLaunch p, const SpModelDev m, int* error_flag) {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  uint8_t* smem = simt::shared_base();
#else
  extern __shared__ __align__(16) uint8_t smem[];
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kBWarps + warp;
  uint8_t* my_arena = p.arena + (size_t)gwarp * p.arena_stride;
  Work wa = make_work(my_arena, p.arena_cap, p.starts != nullptr);
  const ArcScratch scratch = make_scratch(p, my_arena, error_flag);
  const BWork w = make_bwork(smem + (size_t)warp * kBWorkBytes, reinterpret_cast<uint8_t*>(scratch.priv));
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
        result = sp_bpe_fast<false>(m, w, scratch, p.text, lo, hi, padded_bytes, p.ids + doc * (int64_t)p.max_ids, p.max_ids, p.unk_id, delim, lane,
                                    WinOffsets{});
      if (result == kUFallback)
        result = sp_doc_generic<true>(p, m, wa, scratch, doc, lo, hi, padded_bytes, lane, error_flag);
    }
    if (lane == 0) p.counts[doc] = result;
    __syncwarp();
  }
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

static_assert(4 * kBWin <= kSpOffsetsTailBytes && 4 * kUCap <= kSpOffsetsTailBytes, "the window's byte offsets live in the arena tail");

// TextToIdsWithOffsets_sp for the BPE family: sp_bpe_kernel with the byte offsets of the window's symbols carried along (in the
// tail of the warp's arena).  A kernel of its own, so that sp_bpe_kernel stays what it is.
__global__ void __launch_bounds__(kBWarps * 32, kBCtasPerSm) sp_bpe_offsets_kernel(const SpLaunch p, const SpModelDev m, int* error_flag) {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  uint8_t* smem = simt::shared_base();
#else
  extern __shared__ __align__(16) uint8_t smem[];
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kBWarps + warp;
  uint8_t* my_arena = p.arena + (size_t)gwarp * p.arena_stride;
  Work wa = make_work(my_arena, p.arena_cap, true);
  const ArcScratch scratch = make_scratch(p, my_arena, error_flag);
  const BWork w = make_bwork(smem + (size_t)warp * kBWorkBytes, reinterpret_cast<uint8_t*>(scratch.priv));
  const int64_t padded_bytes = (p.text_bytes + 3) & ~(int64_t)3;
  const uint16_t delim = __ldg(m.sym_of_cp + kSpDelim);
  const bool fast_model = m.use_raw_bytes && m.norm_count == nullptr && !m.delim_inside_tokens && m.bpe_ord != nullptr && delim != kNoSym;
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
      if (fast_model) {
        WinOffsets wo;
        wo.boff = reinterpret_cast<int32_t*>(my_arena + p.arena_stride - kSpOffsetsTailBytes);
        wo.doc = p.text + lo;
        wo.starts = p.starts + doc * (int64_t)p.max_ids;
        wo.ends = p.ends + doc * (int64_t)p.max_ids;
        result = sp_bpe_fast<true>(m, w, scratch, p.text, lo, hi, padded_bytes, p.ids + doc * (int64_t)p.max_ids, p.max_ids, p.unk_id, delim, lane, wo);
        __syncwarp();
      }
      if (result == kUFallback)
        result = sp_doc_generic<true>(p, m, wa, scratch, doc, lo, hi, padded_bytes, lane, error_flag);
    }
    if (lane == 0) p.counts[doc] = result;
    __syncwarp();
  }
}

// TextToIdsWithOffsets_sp for the Unigram family: the one-window fast path with the byte offsets carried along (in the tail of
// the warp's arena), sp_doc_generic for what does not fit it.  A kernel of its own, so that sp_unigram_kernel stays what it is.
__global__ void __launch_bounds__(kUWarps * 32, kUCtasPerSm) sp_unigram_offsets_kernel(const SpLaunch p, const SpModelDev m, int* error_flag) {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  uint8_t* smem = simt::shared_base();
#else
  extern __shared__ __align__(16) uint8_t smem[];
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * kUWarps + warp;
  const UWork w = make_uwork(smem + (size_t)warp * kUWorkBytes);
  uint8_t* my_arena = p.arena + (size_t)gwarp * p.arena_stride;
  Work wa = make_work(my_arena, p.arena_cap, true);
  const ArcScratch scratch = make_scratch(p, my_arena, error_flag);
  const int64_t padded_bytes = (p.text_bytes + 3) & ~(int64_t)3;
  const bool fast_model = !m.use_raw_bytes && m.max_arc_len <= kUMaxLen && !m.delim_inside_tokens && m.delim_is_token;
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
      if (fast_model && n <= 4ll * kUCap) {                               // a code point takes at most 4 bytes
        WinOffsets uo;
        uo.boff = reinterpret_cast<int32_t*>(my_arena + p.arena_stride - kSpOffsetsTailBytes);
        uo.doc = p.text + lo;
        uo.starts = p.starts + doc * (int64_t)p.max_ids;
        uo.ends = p.ends + doc * (int64_t)p.max_ids;
        result = unigram_whole<true>(m, w, p.text, lo, hi, padded_bytes, p.ids + doc * (int64_t)p.max_ids, p.max_ids, p.unk_id, lane, uo);
        __syncwarp();
      }
      if (result == kUFallback)
        result = sp_doc_generic<false>(p, m, wa, scratch, doc, lo, hi, padded_bytes, lane, error_flag);
    }
    if (lane == 0) p.counts[doc] = result;
    __syncwarp();
  }
}

}  // namespace

int64_t sp_arena_bytes_per_warp(int cap, int) {
  return align16(work_bytes_arena(cap) + 16ll * ((int64_t)cap * kArcsPerSym + 4096) + 256) + kSpOffsetsTailBytes;
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
  // (per thread and device: the per-document calls come here once per call)
  static thread_local int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) { cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev); cached_dev = dev; }
  const int sms = cached_sms;
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
  const int which = (bpe ? 1 : 0) + (p.starts != nullptr ? 2 : 0);
  auto kern = which == 0 ? sp_unigram_kernel : which == 1 ? sp_bpe_kernel : which == 2 ? sp_unigram_offsets_kernel : sp_bpe_offsets_kernel;
  static thread_local int attr_dev[4] = {-1, -1, -1, -1};       // the attribute is set once per (thread, device, kernel)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (attr_dev[which] != dev) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_dev[which] = dev;
  }
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
