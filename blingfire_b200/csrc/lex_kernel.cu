// lex_kernel.cu -- the generic lexer engine on the GPU: any [wbd] grammar (contexts, IW_ANY,
// tag-less actions, nested calls), used for TextToWords (wbd.bin) and for TextToIds on lexer
// models outside the FastPath shape.  Correctness-first: the grammar-agnostic path does not
// attempt the chunk parallelism of wp_kernel.cu.
//
//   lex_decode_kernel   warp per document: strict UTF-8 decode + class lookup, compacted into
//                       a per-document class array (FAStrUtf8ToArray + [FANormalize] + GetNewIw)
//   lex_run_kernel      THREAD per document: FALexTools_t::Process_int with an explicit frame
//                       stack (lex_core.cuh) -> (Tag, From, To) triples
//   lex_wp_kernel       thread per document: TextToIdsWithOffsets_wp's post-pass over the triples
#include "lex_kernel.cuh"

#include "utf8_warp.cuh"

namespace bfb200 {

namespace {

constexpr int kDecodeWarps = 8;

__global__ void __launch_bounds__(kDecodeWarps * 32) lex_decode_kernel(const LexLaunch p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t* text32 = reinterpret_cast<const uint32_t*>(p.text);
  const int64_t padded_bytes = (p.text_bytes + 3) & ~(int64_t)3;
  for (int64_t doc = warp; doc < p.ndocs; doc += nwarps) {
    int64_t lo = __ldg(p.offsets + doc);
    const int64_t hi = __ldg(p.offsets + doc + 1);
    const int64_t n = hi - lo;
    int result = 0;   // number of code points, -1 = invalid input, 0 = nothing to do
    if (n > 0 && n <= 1000000000) {
      if (n >= 3) {   // BOM (FAUtf8Utils.cpp:247-252)
        const uint32_t b0 = __ldg(p.text + lo), b1 = __ldg(p.text + lo + 1), b2 = __ldg(p.text + lo + 2);
        if (b0 == 0xEF && b1 == 0xBB && b2 == 0xBF) lo += 3;
      }
      const int64_t doc_lo = __ldg(p.offsets + doc);
      uint16_t* cls = p.cls_buf + (doc_lo - p.base_offset);
      int32_t* boff = p.boff_buf ? p.boff_buf + (doc_lo - p.base_offset) : nullptr;
      int m = 0;
      unsigned bad = 0, sumlen = 0;
      for (int64_t bpos = lo; bpos < hi;) {
        const int64_t bs = bpos & ~(int64_t)3;
        const int64_t pos0 = bs + lane * 4;
        uint32_t w0, w1;
        utf8_load_words(text32, pos0, padded_bytes, &w0, &w1);
        const Utf8Lane d = utf8_decode_lane(w0, w1, pos0, bpos, hi);
        bad |= d.bad; sumlen += d.sumlen;
        const int cnt = __popc(d.start_mask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int idx = m + incl - cnt;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (d.start_mask & (1u << k)) {
            if (boff) boff[idx] = (int32_t)(pos0 + k - doc_lo);   // offsets count from the document start, BOM included
            cls[idx++] = __ldg(p.cls_of_cp + d.cp[k]);
          }
        m += __shfl_sync(0xffffffffu, incl, 31);
        bpos = bs + 128;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sumlen += __shfl_xor_sync(0xffffffffu, sumlen, o);
      const bool invalid = __any_sync(0xffffffffu, bad != 0) || (int64_t)sumlen != hi - lo;
      result = invalid ? -1 : m;
    } else if (n != 0) {
      result = -1;
    }
    if (lane == 0) p.ncps[doc] = result;
  }
}

template <typename TE>
__global__ void __launch_bounds__(128) lex_run_kernel(const LexLaunch p, const LexGlobal<TE> g) {
  const int64_t doc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (doc >= p.ndocs) return;
  const int n = p.ncps[doc];
  int count = 0;
  if (n > 0) {
    const int64_t rel = __ldg(p.offsets + doc) - p.base_offset;
    count = lex_process<TE>(g, p.cls_buf + rel, n, p.tri_buf + 3 * (int64_t)p.tri_mul * rel, 3 * p.tri_mul * n);
  }
  p.tri_count[doc] = count;
}

__global__ void __launch_bounds__(128) lex_wp_kernel(const LexLaunch p, int32_t* ids, int32_t* counts, int max_ids, int unk) {
  const int64_t doc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (doc >= p.ndocs) return;
  int c = 0;
  if (p.ncps[doc] > 0) {
    const int64_t rel = __ldg(p.offsets + doc) - p.base_offset;
    c = wp_postpass(p.tri_buf + 3 * (int64_t)p.tri_mul * rel, p.tri_count[doc], ids + doc * (int64_t)max_ids, max_ids, unk);
  }
  counts[doc] = c;
}

// FAUtf8Size of a lead byte (FAUtf8Utils.cpp:23-42)
__device__ __forceinline__ int utf8_size_of_lead(unsigned ch) {
  if ((ch & 0x80) == 0x00) return 1;
  if ((ch & 0xE0) == 0xC0) return 2;
  if ((ch & 0xF0) == 0xE0) return 3;
  if ((ch & 0xF8) == 0xF0) return 4;
  return 0;
}

// TextToIdsWithOffsets_wp's post-pass including offsets (blingfiretokdll.cpp:1207-1313).  The
// charmap is 1 -> 1 for every model served here, so pNormOffsets is the identity.
__global__ void __launch_bounds__(128) lex_wp_offsets_kernel(const LexLaunch p, int32_t* ids, int32_t* starts, int32_t* ends,
                                                             int32_t* counts, int max_ids, int unk) {
  const int64_t doc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (doc >= p.ndocs) return;
  int out = 0;
  if (p.ncps[doc] > 0) {
    const int64_t doc_lo = __ldg(p.offsets + doc);
    const int64_t rel = doc_lo - p.base_offset;
    const int32_t* res = p.tri_buf + 3 * (int64_t)p.tri_mul * rel;
    const int32_t* boff = p.boff_buf + rel;
    const uint8_t* text = p.text + doc_lo;
    const int rn = p.tri_count[doc];
    int32_t* oi = ids + doc * (int64_t)max_ids;
    int32_t* os = starts + doc * (int64_t)max_ids;
    int32_t* oe = ends + doc * (int64_t)max_ids;
    auto emit = [&](int id, int from, int to) {
      oi[out] = id;
      os[out] = boff[from];
      const int to_off = boff[to];
      const int cs = utf8_size_of_lead(text[to_off]);
      oe[out] = to_off + (cs > 0 ? cs - 1 : 0);
      ++out;
    };
    for (int i = 0; i < rn; i += 3) {
      const int tag = res[i];
      if (tag == 4) continue;
      if (tag == 1) {
        const int tfrom = res[i + 1], tto = res[i + 2];
        int j = i + 3, nsub = 0;
        bool covered = false;
        if (j < rn) {
          int expect = tfrom, stag = res[j], sfrom = res[j + 1], sto = res[j + 2];
          while (j <= rn && stag > 4 && expect == sfrom) {
            expect = sto + 1; ++nsub; j += 3;
            if (j < rn) { stag = res[j]; sfrom = res[j + 1]; sto = res[j + 2]; }
          }
          if (nsub > 0 && expect - 1 == tto) {
            for (int k = 0; k < nsub && out < max_ids; ++k) { const int ti = (k + 1) * 3 + i; emit(res[ti], res[ti + 1], res[ti + 2]); }
            covered = true;
          }
        }
        if (!covered && out < max_ids) emit(unk, tfrom, tto);
        i = j - 3;
      }
      if (out >= max_ids) break;
    }
  }
  counts[doc] = out;
}

// ---- TextToWords for a batch: the output strings are put together on the device ----
// blingfiretokdll.cpp:507-555: every token that is not IGNORE, ' ' inside a token -> '_', U+0000 -> ' ' (:482), tokens joined
// by one ' ', a trailing NUL.  One thread per document walks its triples twice: lengths first, bytes after the scan.
struct WordsDoc {
  const int32_t* tri; const int32_t* boff; const uint8_t* text;
  int rn, ncps, nbytes;
};
__device__ __forceinline__ int words_doc(const LexLaunch& p, int64_t doc, WordsDoc* d) {   // 1 ok, 0 empty input, -1 error
  const int64_t lo = __ldg(p.offsets + doc), n = __ldg(p.offsets + doc + 1) - lo;
  if (n == 0) return 0;                                             // :446-448
  if (n < 0 || n > 1000000000) return -1;                           // :449-454
  const int ncps = p.ncps[doc];
  if (ncps <= 0) return -1;                                         // invalid UTF-8, or nothing decoded (:475-478)
  const int rn = p.tri_count[doc];
  if (rn < 0 || rn > 3 * ncps || rn % 3 != 0) return -1;            // :500-502
  const int64_t rel = lo - p.base_offset;
  d->tri = p.tri_buf + 3 * (int64_t)p.tri_mul * rel; d->boff = p.boff_buf + rel; d->text = p.text + lo;
  d->rn = rn; d->ncps = ncps; d->nbytes = (int)n;
  return 1;
}
__device__ __forceinline__ int words_cp_off(const WordsDoc& d, int i) { return i >= d.ncps ? d.nbytes : d.boff[i]; }

// __FAIsWhiteSpace__ (blingfiretokdll.h:17-21) of the (valid) code point at byte b; U+0000 was replaced by U+0020 before
// the lexer ran (:237)
__device__ __forceinline__ bool words_white_at(const uint8_t* s, int b) {
  const uint32_t c0 = s[b];
  uint32_t cp;
  if (c0 < 0x80) cp = c0;
  else if ((c0 & 0xE0) == 0xC0) cp = ((c0 & 0x1Fu) << 6) | (s[b + 1] & 0x3Fu);
  else if ((c0 & 0xF0) == 0xE0) cp = ((c0 & 0x0Fu) << 12) | ((s[b + 1] & 0x3Fu) << 6) | (s[b + 2] & 0x3Fu);
  else cp = ((c0 & 0x07u) << 18) | ((s[b + 1] & 0x3Fu) << 12) | ((s[b + 2] & 0x3Fu) << 6) | (s[b + 3] & 0x3Fu);
  if (cp == 0) cp = 0x20;
  return cp <= 0x20 || cp == 0xa0 || (cp >= 0x2000 && cp <= 0x200f) || cp == 0x202f || cp == 0x205f || cp == 0x2060 ||
         cp == 0x2420 || cp == 0x2424 || cp == 0x3000 || cp == 0xfeff;
}

// The pieces of a document's output string, in order: f(b0, b1) for the bytes [b0, b1) of each.  false: malformed triples.
//   words      every token that is not IGNORE (blingfiretokdll.cpp:507-552)
//   sentences  one per triple, from right after the previous one's end to this one's To, leading white space dropped
//              (FAGetFirstNonWhiteSpace, :138-150), empty ones skipped; what follows the last boundary is the last sentence
//              (:257-338)
template <bool kSentences, typename F>
__device__ __forceinline__ bool words_pieces(const WordsDoc& d, F f) {
  if (!kSentences) {
    for (int i = 0; i < d.rn; i += 3) {
      if (d.tri[i] == 4) continue;                                  // WBD_IGNORE_TAG (:511-514)
      const int from = d.tri[i + 1], to = d.tri[i + 2];
      if (from < 0 || from > d.ncps || to >= d.ncps || to < -1) return false;
      const int b0 = words_cp_off(d, from), b1 = words_cp_off(d, to + 1);
      f(b0, b1 > b0 ? b1 : b0);
    }
    return true;
  }
  int prev_end = -1;
  auto sentence = [&](int from, int to) {
    int first = from;
    while (first <= to && words_white_at(d.text, words_cp_off(d, first))) ++first;
    if (first > to) return;
    f(words_cp_off(d, first), words_cp_off(d, to + 1));
  };
  for (int i = 0; i < d.rn; i += 3) {
    const int to = d.tri[i + 2];
    if (to < -1 || to >= d.ncps) return false;
    sentence(prev_end + 1, to);
    prev_end = to;
  }
  if (prev_end + 1 < d.ncps) sentence(prev_end + 1, d.ncps - 1);
  return true;
}

template <bool kSentences>
__global__ void __launch_bounds__(128) lex_words_len_kernel(const LexLaunch p, int32_t* lens, int32_t* results) {
  const int64_t doc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (doc >= p.ndocs) return;
  WordsDoc d;
  int r = words_doc(p, doc, &d), total = 0;
  if (r == 1) {
    int pieces = 0;
    if (!words_pieces<kSentences>(d, [&](int b0, int b1) { total += (b1 - b0) + (pieces > 0 ? 1 : 0); ++pieces; })) r = -1;
    total += 1;                                                     // the NUL (:555, :343)
  }
  lens[doc] = r == 1 ? total : 0;
  results[doc] = r == 1 ? total : r;
}

template <bool kSentences>
__global__ void __launch_bounds__(128) lex_words_write_kernel(const LexLaunch p, const int64_t* out_off, const int32_t* results, char* out) {
  const int64_t doc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (doc >= p.ndocs || results[doc] <= 0) return;
  WordsDoc d;
  if (words_doc(p, doc, &d) != 1) return;
  char* o = out + out_off[doc];
  int pieces = 0;
  const char sep = kSentences ? '\n' : ' ', repl = kSentences ? ' ' : '_';
  words_pieces<kSentences>(d, [&](int b0, int b1) {
    if (pieces > 0) *o++ = sep;
    for (int b = b0; b < b1; ++b) {
      char c = (char)d.text[b];
      if (c == 0) c = 0x20;                                         // U+0000 -> U+0020 (:482, :237)
      if (c == sep) c = repl;                                       // the delimiter inside a piece (:546, :292)
      *o++ = c;
    }
    ++pieces;
  });
  *o = 0;
}

}  // namespace

#ifndef BF_SIMT_HOST                       // tests/simt compiles the kernels above for the host
cudaError_t lex_words_len_launch(const LexLaunch& p, bool sentences, int32_t* lens, int32_t* results, cudaStream_t stream) {
  if (p.ndocs <= 0) return cudaSuccess;
  if (!p.boff_buf) return cudaErrorInvalidValue;
  const int grid = (int)((p.ndocs + 127) / 128);
  if (sentences) lex_words_len_kernel<true><<<grid, 128, 0, stream>>>(p, lens, results);
  else lex_words_len_kernel<false><<<grid, 128, 0, stream>>>(p, lens, results);
  return cudaGetLastError();
}
cudaError_t lex_words_write_launch(const LexLaunch& p, bool sentences, const int64_t* out_off, const int32_t* results, char* out,
                                   cudaStream_t stream) {
  if (p.ndocs <= 0) return cudaSuccess;
  const int grid = (int)((p.ndocs + 127) / 128);
  if (sentences) lex_words_write_kernel<true><<<grid, 128, 0, stream>>>(p, out_off, results, out);
  else lex_words_write_kernel<false><<<grid, 128, 0, stream>>>(p, out_off, results, out);
  return cudaGetLastError();
}

cudaError_t lex_wp_offsets_launch(const LexLaunch& p, int32_t* ids, int32_t* starts, int32_t* ends, int32_t* counts,
                                  int max_ids, int unk, cudaStream_t stream, int* launches) {
  if (p.ndocs <= 0) return cudaSuccess;
  if (!p.boff_buf) return cudaErrorInvalidValue;
  lex_wp_offsets_kernel<<<(int)((p.ndocs + 127) / 128), 128, 0, stream>>>(p, ids, starts, ends, counts, max_ids, unk);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t lex_launch(const LexLaunch& p, const LexModelDev& m, cudaStream_t stream, int* launches) {
  if (p.ndocs <= 0) return cudaSuccess;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (p.ndocs + kDecodeWarps - 1) / kDecodeWarps;
  if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
  lex_decode_kernel<<<(int)grid, kDecodeWarps * 32, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int64_t g2 = (p.ndocs + 127) / 128;
  if (m.wide) {
    LexGlobal<uint32_t> g{};
    g.trans = reinterpret_cast<const uint32_t*>(m.trans);
    g.ow_of_state = m.ow_of_state; g.act_begin = m.act_begin; g.act_data = m.act_data; g.fn_ini = m.fn_ini; g.fn_count = m.fn_count;
    g.NC1 = m.NC1; g.first_final = m.first_final; g.cls_caret = m.cls_caret; g.cls_dollar = m.cls_dollar; g.initial = m.initial;
    g.max_depth = m.max_depth; g.max_token_length = m.max_token_length;
    lex_run_kernel<uint32_t><<<(int)g2, 128, 0, stream>>>(p, g);
  } else {
    LexGlobal<uint16_t> g{};
    g.trans = reinterpret_cast<const uint16_t*>(m.trans);
    g.ow_of_state = m.ow_of_state; g.act_begin = m.act_begin; g.act_data = m.act_data; g.fn_ini = m.fn_ini; g.fn_count = m.fn_count;
    g.NC1 = m.NC1; g.first_final = m.first_final; g.cls_caret = m.cls_caret; g.cls_dollar = m.cls_dollar; g.initial = m.initial;
    g.max_depth = m.max_depth; g.max_token_length = m.max_token_length;
    lex_run_kernel<uint16_t><<<(int)g2, 128, 0, stream>>>(p, g);
  }
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t lex_wp_launch(const LexLaunch& p, int32_t* ids, int32_t* counts, int max_ids, int unk, cudaStream_t stream, int* launches) {
  if (p.ndocs <= 0) return cudaSuccess;
  lex_wp_kernel<<<(int)((p.ndocs + 127) / 128), 128, 0, stream>>>(p, ids, counts, max_ids, unk);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

#endif  // BF_SIMT_HOST

}  // namespace bfb200
