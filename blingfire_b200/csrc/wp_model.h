// wp_model.h -- packs the small, hot part of a FastPath lexer model into one contiguous
// blob so a CTA can stage it into shared memory with a single bulk async copy, and rebuilds
// the WpTop view over any base address (shared memory on the device, host memory in the
// test twin).
#pragma once

#include <cstdint>
#include <vector>

#include "lexer_tables.h"
#include "wp_core.cuh"

namespace bfb200 {

constexpr int kMaxStagedRows = 4;
constexpr size_t kMaxStagedBytes = 24 * 1024;   // shared-memory budget for staged transition rows

// Byte offsets inside the blob.  Plain data: passed to the kernel by value.
struct WpBlobLayout {
  uint32_t total_bytes;       // multiple of 16 (cp.async.bulk granularity)
  uint32_t off_ascii, off_tc, off_ttop, off_cross, off_final, off_tag, off_root, off_caret;
  uint32_t off_row_root, off_row_caret, off_rows, off_sync;
  int32_t K, NT, num_rows;
  uint32_t row_bytes;         // (NC+1) * sizeof(table entry)
  uint8_t tc_caret, tc_dollar, tc_none;
  uint8_t sync_shift;         // sync_start is [1 << sync_shift][1 << sync_shift]
};

struct WpBlob {
  WpBlobLayout layout{};
  std::vector<uint8_t> bytes;
};

// Builds the blob from flattened tables (requires T.fast.ok && T.charmap_one_to_one).
void build_wp_blob(const LexerTables& T, WpBlob* out);

// View over a blob located at `base` (16-byte aligned).
BF_HD WpTop make_wp_top(const uint8_t* base, const WpBlobLayout& L) {
  WpTop t;
  t.ascii_cls = reinterpret_cast<const uint16_t*>(base + L.off_ascii);
  t.tc_of_class = base + L.off_tc;
  t.ttop = base + L.off_ttop;
  t.cross = reinterpret_cast<const unsigned long long*>(base + L.off_cross);
  t.top_final = base + L.off_final;
  t.top_tag = reinterpret_cast<const int32_t*>(base + L.off_tag);
  t.top_fn_root = reinterpret_cast<const uint32_t*>(base + L.off_root);
  t.top_fn_caret = reinterpret_cast<const uint32_t*>(base + L.off_caret);
  t.top_row_root = reinterpret_cast<const int8_t*>(base + L.off_row_root);
  t.top_row_caret = reinterpret_cast<const int8_t*>(base + L.off_row_caret);
  t.staged_rows = base + L.off_rows;
  t.sync_start = base + L.off_sync; t.sync_shift = L.sync_shift;
  t.K = L.K; t.NT = L.NT;
  t.tc_caret = L.tc_caret; t.tc_dollar = L.tc_dollar; t.tc_none = L.tc_none;
  return t;
}

}  // namespace bfb200
