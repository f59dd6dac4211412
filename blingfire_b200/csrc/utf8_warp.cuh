// utf8_warp.cuh -- warp-cooperative strict UTF-8 decoding (device only).
//
// Each lane owns one 32-bit word (uchar4) of a 128-byte step and decodes the sequences that
// START in its word, looking into the next word for sequences that straddle the lane boundary.
// Strictness follows FAUtf8ToInt (FAUtf8Utils.cpp:121-196): shortest form only, no surrogates,
// nothing above U+10FFFF, continuation bytes must be 10xxxxxx, no truncation at the document end.
// "Every continuation byte belongs to some lead" is checked by the caller through
// sum(len of sequences) == number of bytes.
#pragma once

#include <cstdint>

namespace bfb200 {

__device__ __forceinline__ unsigned bf_lanemask_lt() {
#ifdef BF_SIMT_HOST                        // tests/simt: the kernel source on the CPU
  return (1u << simt::tl.lane) - 1u;
#else
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
#endif
}

struct Utf8Lane {
  uint32_t cp[4];
  unsigned start_mask;   // bit k: a sequence starts at byte k of the lane's word
  unsigned bad;          // nonzero: some sequence starting here is invalid
  unsigned sumlen;       // total length of the sequences that start here
};

__device__ __forceinline__ Utf8Lane utf8_decode_lane(uint32_t w0, uint32_t w1, int64_t pos0, int64_t bpos, int64_t hi) {
  Utf8Lane r;
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

// The lane's word and the following one; reads never go past the 4-byte-padded end of the buffer.
__device__ __forceinline__ void utf8_load_words(const uint32_t* text32, int64_t pos0, int64_t padded_bytes, uint32_t* w0, uint32_t* w1) {
  *w0 = pos0 < padded_bytes ? __ldg(text32 + (pos0 >> 2)) : 0u;
  *w1 = pos0 + 4 < padded_bytes ? __ldg(text32 + (pos0 >> 2) + 1) : 0u;
}

}  // namespace bfb200
