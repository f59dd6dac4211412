// sp_common.cuh -- what the [pos-dict] paths share: the per-warp workspace, the double-array step, I2Info,
// the front end of the general path (raw symbols, FANormalize, whitespace collapse), ordered emission.
// Included by sp_kernel.cu only (one translation unit; everything lives in its anonymous namespace).
#pragma once

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

// TextToIdsWithOffsets_sp (blingfiretokdll.cpp:1519-1529) on the fast paths: where the byte offsets of the window's symbols
// live and where those of the tokens go.  Only the kOff = true instantiations touch it.
struct WinOffsets {
  int32_t* boff;       // byte offset (from the document start) of every symbol of the window, -1 = dummy prefix; the warp's own
                       // global scratch (the tail of its arena), so that the shared-memory layouts stay what they are
  const uint8_t* doc;  // first byte of the document
  int32_t* starts;     // rows parallel to the ids row
  int32_t* ends;
};
// the end offset of a token whose last symbol sits at byte to_off (:1527): the last byte of that character.  A token that is
// only the dummy prefix has to_off == -1 (the reference reads the byte before the input): size 0, like sp_emit
__device__ __forceinline__ int sp_end_offset(const uint8_t* doc, int to_off) {
  const int cs = to_off < 0 ? 0 : sp_utf8_size_of_lead(doc[to_off]);
  return to_off + (cs > 0 ? cs - 1 : 0);
}

constexpr int kUFallback = -2;   // the document does not fit a fast path: the caller takes sp_doc_generic

}  // namespace
}  // namespace bfb200
