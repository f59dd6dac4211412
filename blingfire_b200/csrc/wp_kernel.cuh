// wp_kernel.cuh -- launch interface of the fused WordPiece TextToIds kernel (wp_kernel.cu).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "wp_core.cuh"
#include "wp_model.h"

namespace bfb200 {

struct WpLaunch {
  // batch, all device pointers
  const uint8_t* text;         // concatenated documents (4-byte aligned allocation, 8 bytes of slack readable)
  const int64_t* offsets;      // [ndocs+1] document i = text[offsets[i] .. offsets[i+1])
  int64_t ndocs;
  int64_t text_bytes;          // offsets[ndocs]
  int32_t* ids;                // [ndocs][max_ids]; only the first counts[i] entries of a row are written
  int32_t* counts;             // [ndocs] what TextToIds would have returned for the document
  int max_ids;
  int unk_id;
  // model, device pointers
  const uint8_t* blob;         // WpBlob bytes in global memory (16-byte aligned)
  WpBlobLayout layout;
  const void* trans;           // dense table, uint16_t or uint32_t entries
  bool wide;
  const int32_t* tag_of_state;
  const uint32_t* clsx_of_cp;  // [0x110000] class | top-level class << 16
  WpWords words;               // whole-word table (slots in global memory)
  uint32_t NC1, first_final, cls_caret, cls_dollar;
  int max_token_length;
  unsigned long long* work_counter;   // device scalar, zeroed by the launcher
};

struct WpLaunchInfo {
  int grid, block;
  size_t smem_bytes;
  int launches;                // kernels launched by this call (for bench.py's gpu_launches)
};

// Enqueues the tokenization of one batch on `stream`.  Returns cudaSuccess or the launch error.
cudaError_t wp_tokenize_launch(const WpLaunch& p, cudaStream_t stream, WpLaunchInfo* info);

// Compacts row-major ids into CSR order: csr[row_off[i] + k] = ids[i*max_ids + k], k < counts[i].
cudaError_t wp_compact_launch(const int32_t* ids, const int32_t* counts, const int64_t* row_off,
                              int64_t ndocs, int max_ids, int32_t* csr, cudaStream_t stream);

// Same, 16-bit ids (the caller guarantees every id fits): halves the device->host bytes of a host batch call.
cudaError_t wp_compact_launch_u16(const int32_t* ids, const int32_t* counts, const int64_t* row_off,
                                  int64_t ndocs, int max_ids, uint16_t* csr, cudaStream_t stream);

// Exclusive prefix sum of counts[0..ndocs) into int64 row offsets row_off[0..ndocs].
cudaError_t wp_scan_counts(const int32_t* counts, int64_t* row_off, int64_t ndocs, cudaStream_t stream);

}  // namespace bfb200
