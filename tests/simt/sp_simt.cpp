// sp_simt.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the SOURCE of the segmentation kernels (blingfire_b200/csrc/sp_kernel.cu) for the host with the
// SIMT shim (simt.h: one OS thread per lane, warp intrinsics as rendezvous) and runs sp_unigram_kernel /
// sp_bpe_kernel over a batch of documents: the kernels' warp-level logic -- shuffles, ballots, the
// register window, the streaming windows, the lane re-dealing -- is exercised on the CPU box against the
// oracle (tests/test_simt.py).  The tables are the product's own (seg_tables.cpp), in host memory.
// Never linked into the product.
#include "simt.h"

#include <cstdio>
#include <string>
#include <vector>

#define BF_SIMT_HOST 1
#include "../../blingfire_b200/csrc/sp_kernel.cu"

using namespace bfb200;

namespace {

struct SpSim {
  LdbImage ldb;
  SegTables S;
  std::string err;
  std::vector<WpWordSlot> memo_slots;
};

SpModelDev model_view(const SegTables& S) {
  SpModelDev d{};
  d.da = S.da.data(); d.root = S.root; d.sym_of_cp = S.sym_of_cp.data(); d.info = S.info.data(); d.info_count = (int)S.info.size();
  d.norm_count = S.has_charmap ? S.norm_count.data() : nullptr; d.norm_first = S.norm_first.data(); d.norm_values = S.norm_values.data();
  d.tok_algo = S.tok_algo; d.id_offset = S.id_offset; d.use_raw_bytes = S.use_raw_bytes; d.no_dummy_prefix = S.no_dummy_prefix;
  d.delim_inside_tokens = S.delim_inside_tokens; d.delim_is_token = S.delim_is_token; d.max_arc_len = S.max_arc_len;
  d.bpe_ord = S.bpe_ord_ok ? S.bpe_ord.data() : nullptr; d.bpe_id_of_ord = S.bpe_id_of_ord.data();
  d.bpe_singles_first = S.bpe_singles_first;
  d.seg_memo = WpWords{};
  return d;
}

}  // namespace

extern "C" {

void* spsim_load(const char* path) {
  SpSim* t = new SpSim();
  if (!t->ldb.load_file(path)) { t->err = t->ldb.error(); return t; }
  if (!build_seg_tables(t->ldb, &t->S, &t->err) && t->err.empty()) t->err = "build failed";
  return t;
}
void spsim_free(void* h) { delete (SpSim*)h; }
const char* spsim_error(void* h) { return ((SpSim*)h)->err.c_str(); }

// The batch through the kernel source: ids [ndocs][max_ids], counts [ndocs]; with starts/ends the offsets too
// (sp_bpe_offsets_kernel / sp_unigram_offsets_kernel).  `warps` lane groups pull documents from the shared counter, like a (tiny)
// persistent grid of one CTA.  Returns the kernel's error flag (0 = fine), -1 on bad arguments.
int spsim_batch(void* h, const char* text, const int64_t* offsets, int64_t ndocs, int32_t* ids, int32_t* counts,
                int32_t* starts, int32_t* ends, int max_ids, int unk, int warps) {
  SpSim* t = (SpSim*)h;
  if (!t->err.empty() || ndocs <= 0 || warps < 1) return -1;
  const SegTables& S = t->S;
  const bool bpe = S.tok_algo == kTokenizeBpe || S.tok_algo == kTokenizeBpeOpt || S.tok_algo == kTokenizeBpeOptWithMerges;
  const int cta_warps = bpe ? kBWarps : kUWarps;
  if (warps > cta_warps) warps = cta_warps;
  int64_t max_len = 0;
  for (int64_t d = 0; d < ndocs; ++d) max_len = std::max(max_len, offsets[d + 1] - offsets[d]);
  const int cap = (int)((S.has_charmap ? 2 * (max_len + 1) : max_len + 1) + 2);       // capi.cu launch_segmentation
  const int64_t per_warp = sp_arena_bytes_per_warp(cap, S.max_arc_len);
  std::vector<uint8_t> arena((size_t)per_warp * (size_t)cta_warps + 64, 0xCD);   // cudaMalloc'ed scratch is not zeroed
  const int64_t ovf_entries = bpe ? sp_overflow_entries(cap, S.max_arc_len) : 1;
  std::vector<uint8_t> overflow((size_t)ovf_entries * 16 + 64, 0xCD);
  // a padded copy of the text (the kernels read whole 32-bit words)
  const int64_t total = offsets[ndocs];
  std::vector<uint8_t> padded((size_t)total + 64, 0);
  std::memcpy(padded.data(), text, (size_t)total);
  alignas(16) unsigned long long counter[2] = {0, 0};
  SpLaunch X{};
  X.text = padded.data(); X.offsets = offsets; X.ndocs = ndocs; X.text_bytes = total;
  X.ids = ids; X.counts = counts; X.starts = starts; X.ends = ends; X.max_ids = max_ids; X.unk_id = unk;
  X.work_counter = counter; X.arena = arena.data(); X.arena_stride = per_warp; X.arena_cap = cap; X.grid_warps = cta_warps;
  X.overflow = overflow.data(); X.overflow_cap = ovf_entries;
  SpModelDev m = model_view(S);
  if (bpe && S.bpe_ord_ok) {      // the segment memo persists across the batches of one handle, like the device table
    if (t->memo_slots.empty()) t->memo_slots.assign((size_t)2 << 12, WpWordSlot{});
    m.seg_memo = sp_seg_memo_params(S.alphabet, 12);
    m.seg_memo.slots = t->memo_slots.data();
  }
  int* err = reinterpret_cast<int*>(counter + 1);
  blockDim.x = (unsigned)cta_warps * 32; gridDim.x = 1;
  const size_t smem = bpe ? (size_t)kBWarps * kBWorkBytes : (size_t)kUWarps * kUWorkBytes;
  simt::run_cta(warps, smem, [&] {
    if (bpe) { if (starts) sp_bpe_offsets_kernel(X, m, err); else sp_bpe_kernel(X, m, err); }      // sp_tokenize_launch's choice
    else if (starts) sp_unigram_offsets_kernel(X, m, err);
    else sp_unigram_kernel(X, m, err);
  });
  return *err;
}

}  // extern "C"
