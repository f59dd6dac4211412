// wp_model.h -- packs the small, hot part of a FastPath lexer model into one contiguous
// blob so a CTA can stage it into shared memory with a single bulk async copy, rebuilds
// the WpTop view over any base address (shared memory on the device, host memory in the
// test twin), and builds the load-time memo of the lexer loops (wp_core.cuh): the chunk kind of
// every top-level class and the whole-word table.
#pragma once

#include <cstdint>
#include <vector>

#include "lexer_tables.h"
#include "wp_core.cuh"

namespace bfb200 {

// Byte offsets inside the blob.  Plain data: passed to the kernel by value.
struct WpBlobLayout {
  uint32_t total_bytes;       // multiple of 16 (cp.async.bulk granularity)
  uint32_t off_ascii, off_ttop, off_tag, off_root, off_caret, off_sync, off_kind, off_fn_root, off_fn_caret;
  int32_t K, NT;
  uint8_t tc_caret, tc_dollar, tc_none;
  uint8_t sync_shift;         // sync_start is [1 << sync_shift][1 << sync_shift]
};

struct WpBlob {
  WpBlobLayout layout{};
  std::vector<uint8_t> bytes;
  // whole-word table (uploaded to global memory; `words.slots` is left null here)
  std::vector<WpWordSlot> word_slots;
  WpWords words{};
  int64_t word_count = 0;     // keys in the table
};

// Builds the blob and the memo from flattened tables (requires T.fast.ok && T.charmap_one_to_one and
// the dense table still resident on the host).
void build_wp_blob(const LexerTables& T, WpBlob* out);

// View over a blob located at `base` (16-byte aligned).
BF_HD WpTop make_wp_top(const uint8_t* base, const WpBlobLayout& L) {
  WpTop t;
  t.ascii_clsx = reinterpret_cast<const uint32_t*>(base + L.off_ascii);
  t.ttop = base + L.off_ttop;
  t.top_tag = reinterpret_cast<const int32_t*>(base + L.off_tag);
  t.top_fn_root = reinterpret_cast<const uint32_t*>(base + L.off_root);
  t.top_fn_caret = reinterpret_cast<const uint32_t*>(base + L.off_caret);
  t.sync_start = base + L.off_sync; t.sync_shift = L.sync_shift;
  t.kind_of_tc = reinterpret_cast<const uint32_t*>(base + L.off_kind);
  t.fn_root_of_tc = reinterpret_cast<const uint32_t*>(base + L.off_fn_root);
  t.fn_caret_of_tc = reinterpret_cast<const uint32_t*>(base + L.off_fn_caret);
  t.K = L.K; t.NT = L.NT;
  t.tc_caret = L.tc_caret; t.tc_dollar = L.tc_dollar; t.tc_none = L.tc_none;
  return t;
}

}  // namespace bfb200
