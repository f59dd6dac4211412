// wp_simt.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The SOURCE of the fused WordPiece kernel (blingfire_b200/csrc/wp_kernel.cu: wp_tokenize_kernel, with
// wp_core.cuh and utf8_warp-style decoding inside it) compiled for the host over the SIMT shim and run
// over a batch of documents against the oracle (tests/test_simt.py): decode + validation, the sync-point
// ballot pass, the counting sort of the chunks, lane-per-chunk walks, ordered emission, the multi-window
// carry.  The blob staging (cp.async.bulk + mbarrier on the device) is a memcpy here.
// Never linked into the product.
#include "simt.h"

#include <cstdio>
#include <string>
#include <vector>

#define BF_SIMT_HOST 1
#include "../../blingfire_b200/csrc/wp_kernel.cu"

using namespace bfb200;

namespace {
struct WpSim {
  LdbImage ldb;
  LexerTables T;
  WpBlob blob;
  std::string err;
};
}  // namespace

extern "C" {

void* wpsim_load(const char* path) {
  WpSim* t = new WpSim();
  if (!t->ldb.load_file(path)) { t->err = t->ldb.error(); return t; }
  if (!build_lexer_tables(t->ldb, &t->T, &t->err)) return t;
  if (!(t->T.fast.ok && t->T.charmap_one_to_one)) { t->err = "not a FastPath model"; return t; }
  build_wp_blob(t->T, &t->blob);
  return t;
}
void wpsim_free(void* h) { delete (WpSim*)h; }
const char* wpsim_error(void* h) { return ((WpSim*)h)->err.c_str(); }

// ids [ndocs][max_ids], counts [ndocs]; `warps` lane groups of one emulated CTA pull documents from the counter.
int wpsim_batch(void* h, const char* text, const int64_t* offsets, int64_t ndocs, int32_t* ids, int32_t* counts, int max_ids,
                int unk, int warps) {
  WpSim* t = (WpSim*)h;
  if (!t->err.empty() || ndocs <= 0 || warps < 1) return -1;
  if (warps > kWarpsPerCta) warps = kWarpsPerCta;
  const LexerTables& T = t->T;
  const int64_t total = offsets[ndocs];
  std::vector<uint8_t> padded((size_t)total + 64, 0);
  std::memcpy(padded.data(), text, (size_t)total);
  unsigned long long counter = 0;
  WpLaunch L{};
  L.wide = T.wide_states;
  L.trans = T.wide_states ? (const void*)T.trans32.data() : (const void*)T.trans16.data();
  L.tag_of_state = T.tag_of_state.data();
  L.clsx_of_cp = T.clsx_of_cp.data();
  L.words = t->blob.words;
  L.words.slots = t->blob.word_slots.data();
  L.blob = t->blob.bytes.data();
  L.layout = t->blob.layout;
  L.NC1 = (uint32_t)T.NC + 1; L.first_final = T.first_final; L.cls_caret = T.cls_caret; L.cls_dollar = T.cls_dollar;
  L.max_token_length = T.max_token_length;
  L.text = padded.data(); L.offsets = offsets; L.ndocs = ndocs; L.text_bytes = total;
  L.ids = ids; L.counts = counts; L.max_ids = max_ids; L.unk_id = unk; L.work_counter = &counter;
  blockDim.x = (unsigned)kThreads; gridDim.x = 1;
  const size_t blob_bytes = ((size_t)L.layout.total_bytes + 127) & ~(size_t)127;
  const size_t smem = blob_bytes + (size_t)kWarpsPerCta * kWarpSmem;
  simt::run_cta(warps, smem, [&] {
    if (L.wide) wp_tokenize_kernel<uint32_t>(L); else wp_tokenize_kernel<uint16_t>(L);
  });
  return 0;
}

}  // extern "C"
