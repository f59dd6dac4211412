// lex_simt.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The SOURCE of the generic lexer engine (blingfire_b200/csrc/lex_kernel.cu + lex_core.cuh: warp-per-document
// decode, thread-per-document Process_int with an explicit frame stack, the WordPiece post-pass with and
// without offsets) compiled for the host over the SIMT shim, against the oracle (tests/test_simt.py).
// Never linked into the product.
#include "simt.h"

#include <string>
#include <vector>

#define BF_SIMT_HOST 1
#include "../../blingfire_b200/csrc/lex_kernel.cu"
#include "../../blingfire_b200/csrc/lexer_tables.h"

using namespace bfb200;

namespace {

struct LexSim {
  LdbImage ldb;
  LexerTables T;
  std::string err;
};

template <typename TE>
LexGlobal<TE> global_view(const LexerTables& T, const TE* trans) {
  LexGlobal<TE> g{};
  g.trans = trans;
  g.ow_of_state = T.ow_of_state.data(); g.act_begin = T.act_begin.data(); g.act_data = T.act_data.data();
  g.fn_ini = T.fn_ini.data(); g.fn_count = (int)T.fn_ini.size();
  g.NC1 = (uint32_t)T.NC + 1; g.first_final = T.first_final; g.cls_caret = T.cls_caret; g.cls_dollar = T.cls_dollar; g.initial = T.initial;
  g.max_depth = T.max_depth; g.max_token_length = T.max_token_length;
  return g;
}

struct Scratch {
  std::vector<uint8_t> text;
  std::vector<uint16_t> cls;
  std::vector<int32_t> ncps, tri, tri_count, boff;
};

LexLaunch make_launch(const LexerTables& T, const char* text, const int64_t* offsets, int64_t ndocs, bool words, bool want_boff, Scratch* S) {
  const int64_t total = offsets[ndocs];
  S->text.assign((size_t)total + 64, 0);
  std::memcpy(S->text.data(), text, (size_t)total);
  const int tri_mul = words ? 1 : 2;
  // (device scratch is not zeroed: poison it)
  S->cls.assign((size_t)total + 8, 0xCDCD);
  S->ncps.assign((size_t)ndocs, (int32_t)0xCDCDCDCD);
  S->tri_count.assign((size_t)ndocs, (int32_t)0xCDCDCDCD);
  S->tri.assign((size_t)(3 * tri_mul) * (size_t)total + 8, (int32_t)0xCDCDCDCD);
  if (want_boff) S->boff.assign((size_t)total + 8, (int32_t)0xCDCDCDCD);
  LexLaunch X{};
  X.text = S->text.data(); X.offsets = offsets; X.ndocs = ndocs; X.text_bytes = total; X.base_offset = offsets[0];
  X.cls_of_cp = words ? T.cls_words_of_cp.data() : T.cls_of_cp.data();
  X.cls_buf = S->cls.data(); X.ncps = S->ncps.data(); X.tri_buf = S->tri.data(); X.tri_count = S->tri_count.data();
  X.tri_mul = tri_mul; X.boff_buf = want_boff ? S->boff.data() : nullptr;
  return X;
}

// every emulated thread runs the kernel once per block index, so that `blocks` blocks of blockDim threads are covered
template <class F>
void run_grid(int warps, int64_t blocks, F&& kernel) {
  blockDim.x = (unsigned)warps * 32; gridDim.x = (unsigned)blocks;
  simt::run_cta(warps, 64, [&] { for (int64_t b = 0; b < blocks; ++b) { blockIdx.x = (unsigned)b; kernel(); } });
}

void lex_and_run(const LexerTables& T, const LexLaunch& X, int warps) {
  // the decode kernel strides over the documents with whatever grid it gets
  blockDim.x = (unsigned)warps * 32; gridDim.x = 1;
  simt::run_cta(warps, 64, [&] { blockIdx.x = 0; lex_decode_kernel(X); });
  const int64_t blocks = (X.ndocs + warps * 32 - 1) / (warps * 32);
  if (T.wide_states) { const auto g = global_view<uint32_t>(T, T.trans32.data()); run_grid(warps, blocks, [&] { lex_run_kernel<uint32_t>(X, g); }); }
  else { const auto g = global_view<uint16_t>(T, T.trans16.data()); run_grid(warps, blocks, [&] { lex_run_kernel<uint16_t>(X, g); }); }
}

}  // namespace

extern "C" {

void* lexsim_load(const char* path) {
  LexSim* t = new LexSim();
  if (!t->ldb.load_file(path)) { t->err = t->ldb.error(); return t; }
  if (!build_lexer_tables(t->ldb, &t->T, &t->err)) return t;
  if (t->T.max_depth > kMaxLexDepth) t->err = "grammar nests deeper than the engine's frame stack";
  return t;
}
void lexsim_free(void* h) { delete (LexSim*)h; }
const char* lexsim_error(void* h) { return ((LexSim*)h)->err.c_str(); }

// TextToIds[WithOffsets]_wp through the generic engine: ids [ndocs][max_ids], counts; starts/ends may be NULL
int lexsim_ids(void* h, const char* text, const int64_t* offsets, int64_t ndocs, int32_t* ids, int32_t* counts, int32_t* starts,
               int32_t* ends, int max_ids, int unk, int warps) {
  LexSim* t = (LexSim*)h;
  if (!t->err.empty() || ndocs <= 0 || !t->T.charmap_one_to_one) return -1;
  Scratch S;
  const LexLaunch X = make_launch(t->T, text, offsets, ndocs, /*words=*/false, starts != nullptr, &S);
  lex_and_run(t->T, X, warps);
  const int64_t blocks = (ndocs + warps * 32 - 1) / (warps * 32);
  if (starts) run_grid(warps, blocks, [&] { lex_wp_offsets_kernel(X, ids, starts, ends, counts, max_ids, unk); });
  else run_grid(warps, blocks, [&] { lex_wp_kernel(X, ids, counts, max_ids, unk); });
  return 0;
}

// the TextToWords / TextToSentences view: code points per document (-1 = invalid UTF-8) and the (Tag, From, To) triples,
// document d's at tri[3 * (offsets[d] - offsets[0]) ...], tri_count[d] ints
int lexsim_triples(void* h, const char* text, const int64_t* offsets, int64_t ndocs, int32_t* ncps, int32_t* tri, int32_t* tri_count, int warps) {
  LexSim* t = (LexSim*)h;
  if (!t->err.empty() || ndocs <= 0) return -1;
  Scratch S;
  const LexLaunch X = make_launch(t->T, text, offsets, ndocs, /*words=*/true, false, &S);
  lex_and_run(t->T, X, warps);
  std::memcpy(ncps, S.ncps.data(), sizeof(int32_t) * (size_t)ndocs);
  std::memcpy(tri_count, S.tri_count.data(), sizeof(int32_t) * (size_t)ndocs);
  std::memcpy(tri, S.tri.data(), sizeof(int32_t) * 3 * (size_t)(offsets[ndocs] - offsets[0]));
  return 0;
}

}  // extern "C"
