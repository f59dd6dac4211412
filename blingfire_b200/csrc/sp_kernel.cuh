// sp_kernel.cuh -- launch interface of the [pos-dict] (SentencePiece-style) engines: Unigram-LM
// best path and BPE over the Mealy MPH automaton (sp_kernel.cu): a shared-memory fast path per family,
// the general path in a per-warp global arena behind it.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "seg_tables.h"
#include "wp_core.cuh"   // WpWords: the run-time memo table (segment -> ids) of the streaming BPE path

namespace bfb200 {

struct SpModelDev {
  const DaEntry* da;             // double-array automaton
  uint32_t root;
  const uint16_t* sym_of_cp;     // [0x110000]
  const SegInfo* info;           // [info_count]
  int info_count;
  const uint8_t* norm_count;     // [0x110000] or nullptr (no charmap)
  const uint32_t* norm_first;
  const int32_t* norm_values;
  int tok_algo, id_offset;
  bool use_raw_bytes, no_dummy_prefix, delim_inside_tokens;
  bool delim_is_token;           // "U+2581" alone is a token (seg_tables.h)
  int max_arc_len;
  // BPE family: ordinal of a key in the arc sort order and its inverse (seg_tables.h); nullptr when
  // the ordinals do not fit a sort key (the streaming BPE path is then off)
  const int32_t* bpe_ord;        // [info_count]
  const int32_t* bpe_id_of_ord;
  bool bpe_singles_first;        // one-symbol tokens sort before all others (seg_tables.h)
  // streaming BPE path: U+2581-delimited segments are independent subproblems, so a segment's ids are a function of its
  // symbols alone; segments resolved once are kept in this table (filled at run time only; max_len 0 = off)
  WpWords seg_memo;
};

// Parameters of the BPE segment memo for an alphabet of `alphabet` symbols (slots = nullptr: the caller allocates
// (2 << log2_size) zeroed WpWordSlot and sets the pointer).  max_len 0: the model's symbols do not fit a key.
inline WpWords sp_seg_memo_params(int alphabet, uint32_t log2_size) {
  WpWords W{};
  W.log2_size = log2_size;
  W.cb = 1;
  while ((1u << W.cb) < (uint32_t)alphabet) ++W.cb;
  W.cpw = W.cb > 30 ? 0u : (30u / W.cb > 3u ? 3u : 30u / W.cb);
  W.max_len = W.cpw == 0 ? 0u : (8u * W.cpw < (uint32_t)kMaxFastLen ? 8u * W.cpw : (uint32_t)kMaxFastLen);
  const uint32_t mul[9] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u, 0x2545F491u};
  for (int i = 0; i < 9; ++i) W.mul[i] = mul[i];
  return W;
}

struct SpLaunch {
  const uint8_t* text;           // biased: absolute offsets index it
  const int64_t* offsets;        // [ndocs+1]
  int64_t ndocs;
  int64_t text_bytes;
  int32_t* ids;                  // [ndocs][max_ids]
  int32_t* counts;               // [ndocs]
  // optional (both or neither): byte offsets of the first byte of every token's first character and
  // of the last byte of its last character, relative to the document start, [ndocs][max_ids].
  // Offsets ride in the arena workspace: arena_cap must then cover every document of the launch.
  int32_t* starts;
  int32_t* ends;
  int max_ids, unk_id;
  unsigned long long* work_counter;
  // per-warp global scratch for the documents (or segments) the fast paths do not hold
  uint8_t* arena;                // [grid_warps][arena_stride] bytes
  int64_t arena_stride;          // bytes per warp
  int arena_cap;                 // symbols a warp's arena region can hold
  int grid_warps;                // warps the launch may use (arena rows)
  // BPE: a segment whose arcs do not fit the warp's private scratch takes this grid-wide region
  // under a spin lock (rare: long runs of one character have dozens of vocabulary matches per start)
  uint8_t* overflow;             // [overflow_cap] 16-byte arc entries
  int64_t overflow_cap;
};

// arc entries the shared overflow region must hold for documents of up to `cap` symbols
int64_t sp_overflow_entries(int cap, int max_arc_len);

// bytes of arena one warp needs to process documents of up to `cap` symbols; its last kSpOffsetsTailBytes hold the byte
// offsets of the fast paths' window when offsets are asked for (sp_*_offsets_kernel)
constexpr int kSpOffsetsTailBytes = 4096;
int64_t sp_arena_bytes_per_warp(int cap, int max_arc_len);
// preferred number of warps in the grid on the current device
int sp_preferred_warps(int tok_algo);
// symbols (after the charmap) a document may have and still be served without the arena: the
// Unigram fast path's shared-memory capacity; 0 for the models that always use the arena
constexpr int kSpUnigramFastCap = 576;
int sp_fast_cap(int tok_algo, int max_arc_len, bool use_raw_bytes);

cudaError_t sp_tokenize_launch(const SpLaunch& p, const SpModelDev& m, cudaStream_t stream, int* launches);

}  // namespace bfb200
