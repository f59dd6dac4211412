// wp_model.cpp -- see wp_model.h.
#include "wp_model.h"

#include <cstring>

namespace bfb200 {

void build_wp_blob(const LexerTables& T, WpBlob* out) {
  const FastPath& F = T.fast;
  WpBlobLayout& L = out->layout;
  L = WpBlobLayout{};
  const uint32_t NC1 = (uint32_t)T.NC + 1;
  const uint32_t esz = T.wide_states ? 4u : 2u;
  L.K = F.K; L.NT = F.NT;
  L.tc_caret = F.tc_caret; L.tc_dollar = F.tc_dollar; L.tc_none = F.tc_none;
  L.row_bytes = NC1 * esz;

  // hot rows: FnIni and delta(FnIni, ^) of every called function, at most kMaxStagedRows
  std::vector<uint32_t> rows;
  std::vector<int8_t> row_root((size_t)F.K, -1), row_caret((size_t)F.K, -1);
  auto stage = [&](uint32_t s) -> int8_t {
    if (s == kNoState) return -1;
    for (size_t i = 0; i < rows.size(); ++i) if (rows[i] == s) return (int8_t)i;
    if ((int)rows.size() >= kMaxStagedRows) return -1;
    if ((rows.size() + 1) * (size_t)L.row_bytes > kMaxStagedBytes) return -1;   // keep the blob small
    rows.push_back(s);
    return (int8_t)(rows.size() - 1);
  };
  for (int i = 0; i < F.K; ++i) {
    if (!F.top_final[i] || F.top_fn_root[i] == kNoState) continue;
    row_caret[i] = stage(F.top_fn_caret[i]);
    row_root[i] = stage(F.top_fn_root[i]);
  }
  L.num_rows = (int32_t)rows.size();

  uint32_t off = 0;
  auto place = [&](uint32_t bytes, uint32_t align) { off = (off + align - 1) / align * align; const uint32_t o = off; off += bytes; return o; };
  L.off_ascii = place(128 * 2, 16);
  L.off_tc = place(NC1, 16);
  L.off_ttop = place((uint32_t)F.K * F.NT, 16);
  L.off_cross = place((uint32_t)F.NT * 8, 16);
  L.off_final = place((uint32_t)F.K, 16);
  L.off_tag = place((uint32_t)F.K * 4, 16);
  L.off_root = place((uint32_t)F.K * 4, 16);
  L.off_caret = place((uint32_t)F.K * 4, 16);
  L.off_row_root = place((uint32_t)F.K, 16);
  L.off_row_caret = place((uint32_t)F.K, 16);
  L.sync_shift = 1;
  while ((1 << L.sync_shift) < F.NT) ++L.sync_shift;
  L.off_sync = place(1u << (2 * L.sync_shift), 16);
  L.off_rows = place(L.row_bytes * (uint32_t)(rows.empty() ? 1 : rows.size()), 16);
  L.total_bytes = (off + 15) / 16 * 16;

  out->bytes.assign(L.total_bytes, 0);
  uint8_t* b = out->bytes.data();
  for (int cp = 0; cp < 128; ++cp) reinterpret_cast<uint16_t*>(b + L.off_ascii)[cp] = T.cls_of_cp[cp];
  std::memcpy(b + L.off_tc, F.tc_of_class.data(), NC1);
  // bit 7 of a top-level transition = "the destination is final" (wp_core.cuh kTopFinal); 0xFF stays "none"
  for (size_t i = 0; i < (size_t)F.K * F.NT; ++i) {
    const uint8_t d = F.ttop[i];
    b[L.off_ttop + i] = (d != 0xFF && F.top_final[d]) ? (uint8_t)(d | 0x80) : d;
  }
  std::memcpy(b + L.off_cross, F.cross.data(), (size_t)F.NT * 8);
  for (int t1 = 0; t1 < F.NT; ++t1)
    for (int t2 = 0; t2 < F.NT; ++t2)
      b[L.off_sync + ((size_t)t1 << L.sync_shift) + t2] = (!((F.cross[(size_t)t1] >> t2) & 1ull) && F.ttop[(size_t)t2] != 0xFF) ? 1 : 0;
  std::memcpy(b + L.off_final, F.top_final.data(), (size_t)F.K);
  std::memcpy(b + L.off_tag, F.top_tag.data(), (size_t)F.K * 4);
  std::memcpy(b + L.off_root, F.top_fn_root.data(), (size_t)F.K * 4);
  std::memcpy(b + L.off_caret, F.top_fn_caret.data(), (size_t)F.K * 4);
  std::memcpy(b + L.off_row_root, row_root.data(), (size_t)F.K);
  std::memcpy(b + L.off_row_caret, row_caret.data(), (size_t)F.K);
  for (size_t r = 0; r < rows.size(); ++r) {
    const uint8_t* src = T.wide_states ? reinterpret_cast<const uint8_t*>(T.trans32.data() + (size_t)rows[r] * NC1)
                                       : reinterpret_cast<const uint8_t*>(T.trans16.data() + (size_t)rows[r] * NC1);
    std::memcpy(b + L.off_rows + r * (size_t)L.row_bytes, src, L.row_bytes);
  }
}

}  // namespace bfb200
