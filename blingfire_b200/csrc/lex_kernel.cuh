// lex_kernel.cuh -- launch interface of the generic lexer engine (lex_kernel.cu).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "lex_core.cuh"

namespace bfb200 {

// device-resident generic lexer model
struct LexModelDev {
  const void* trans;
  bool wide;
  const int32_t* ow_of_state;
  const int32_t* act_begin;
  const int32_t* act_data;
  const uint32_t* fn_ini;
  int fn_count;
  uint32_t NC1, first_final, cls_caret, cls_dollar, initial;
  int max_depth, max_token_length;
};

struct LexLaunch {
  const uint8_t* text;        // biased so that absolute offsets index it (see capi.cu)
  const int64_t* offsets;     // [ndocs+1] absolute
  int64_t ndocs;
  int64_t text_bytes;         // offsets[ndocs]
  int64_t base_offset;        // offsets[0] of the chunk: scratch arrays are indexed by offset - base_offset
  const uint16_t* cls_of_cp;  // [0x110000] the symbol view to use (TextToIds or TextToWords)
  uint16_t* cls_buf;          // [chunk bytes] classes of document d at cls_buf + offsets[d] - base_offset
  int32_t* ncps;              // [ndocs] code points per document, -1 = invalid UTF-8
  int32_t* tri_buf;           // [3 * tri_mul * chunk bytes] triples of document d at 3*tri_mul*(offsets[d]-base_offset)
  int32_t* tri_count;         // [ndocs] ints written (3 per triple)
  int tri_mul;                // capacity in triples per code point: 1 (TextToWords, :492) or 2 (TextToIds_wp, :1194)
  int32_t* boff_buf;          // optional [chunk bytes]: byte offset (from the document start) of every code point
};

cudaError_t lex_launch(const LexLaunch& p, const LexModelDev& m, cudaStream_t stream, int* launches);
cudaError_t lex_wp_launch(const LexLaunch& p, int32_t* ids, int32_t* counts, int max_ids, int unk, cudaStream_t stream, int* launches);
// the post-pass with offsets (blingfiretokdll.cpp:1263-1273,1289-1297); needs p.boff_buf
cudaError_t lex_wp_offsets_launch(const LexLaunch& p, int32_t* ids, int32_t* starts, int32_t* ends, int32_t* counts,
                                  int max_ids, int unk, cudaStream_t stream, int* launches);

// TextToWords for a batch (needs p.boff_buf): bytes of every document's output string incl. its NUL (0 when the document has
// none) and what TextToWords returns for it; then the strings themselves at out + out_off[doc]
// (sentences: TextToSentences' string instead, blingfiretokdll.cpp:257-338)
cudaError_t lex_words_len_launch(const LexLaunch& p, bool sentences, int32_t* lens, int32_t* results, cudaStream_t stream);
cudaError_t lex_words_write_launch(const LexLaunch& p, bool sentences, const int64_t* out_off, const int32_t* results, char* out,
                                   cudaStream_t stream);

}  // namespace bfb200
