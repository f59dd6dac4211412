// seg_tables.h -- flattened, HBM-ready form of a [pos-dict] segmentation model (the
// SentencePiece-style models: Unigram-LM xlm_roberta_base/xlnet/laser/uri, BPE gpt2/roberta).
//
// Built once at LoadModel time from the packed image.  Replaces the per-symbol packed-record
// decoding of FAMealyDfa_pack_triv::GetDestOw (FAMealyDfa_pack_triv.cpp:69-244),
// FARSDfa_pack_triv::IsFinal and FAMultiMap_pack_fixed::Get on I2Info
// (FAMultiMap_pack_fixed.cpp:140-162) with:
//
//   sym_of_cp[cp]     symbol -> dense alphabet index (0xFFFF: not in the alphabet, every
//                     transition on it fails); bytes index it directly in byte mode
//   da[]              the Mealy MPH automaton as a DOUBLE-ARRAY: a state is its `base`; the arc
//                     on symbol index s is da[base + s], valid iff da[base+s].check == base.
//                     One 16-byte load per GetDestOw, hit or miss, no search, no probing.
//   info[key]         MPH key (sum of arc weights on the path) -> {id, score/rank}
//   norm_*            the charmap (FANormalize, 1 -> 0..N code points)
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "ldb.h"

namespace bfb200 {

struct DaEntry {
  uint32_t check;   // base of the state that owns this slot; 0xFFFFFFFF = empty
  uint32_t dst;     // base of the destination (0 = no outgoing arcs) | 0x80000000 if it is final
  int32_t ow;       // arc output weight (MPH partial sum); -1 when the source state has none
  uint32_t pad;
};
constexpr uint32_t kDaFinalBit = 0x80000000u;
constexpr uint16_t kNoSym = 0xFFFF;
constexpr int kSpDelim = 0x2581;      // blingfiretokdll.h:11

struct SegInfo {
  int32_t id;
  float score;      // Unigram: log-prob; BPE with merges: rank.  Bit pattern of pValues[1]
};

struct SegTables {
  // configuration (FADictConfKeeper.cpp:57-228)
  int tok_algo = 0;            // 0/2 unigram, 3 bpe, 4 bpe-opt, 5 bpe-opt-with-merges
  int id_offset = 0;
  bool use_raw_bytes = false;
  bool no_dummy_prefix = false;
  bool has_charmap = false;

  int alphabet = 0;                     // number of distinct arc labels
  std::vector<uint16_t> sym_of_cp;      // [0x110000]
  uint32_t root = 0;                    // base of the initial state
  std::vector<DaEntry> da;
  std::vector<SegInfo> info;            // [max_key+1]; id == INT32_MIN marks an unusable row
  int max_arc_len = 0;                  // longest path from the root (symbols)
  bool delim_inside_tokens = false;     // some token has U+2581 past its first symbol
  bool delim_is_token = false;          // "U+2581" alone is a token: a start on U+2581 always has an arc

  // BPE family: dense ordinal of a key in the arc sort order (rank descending for with-merges,
  // then id); -1 for an unusable row.  bpe_id_of_ord inverts it.  bpe_ord_ok: ordinals fit 20 bits.
  std::vector<int32_t> bpe_ord;         // [max_key+1]
  std::vector<int32_t> bpe_id_of_ord;
  bool bpe_ord_ok = false;
  bool bpe_singles_first = false;       // every one-symbol token sorts before every longer one

  std::vector<uint8_t> norm_count;      // [0x110000] 0..10, 0xFF = unmapped (keep code point)
  std::vector<uint32_t> norm_first;     // [0x110000]
  std::vector<int32_t> norm_values;

  // GetDestOw: returns false on "no transition"
  bool step(uint32_t* q, uint16_t s, int* ow, bool* final) const {
    if (s == kNoSym) return false;
    const DaEntry& e = da[(size_t)*q + s];
    if (e.check != *q) return false;
    *ow = e.ow; *final = (e.dst & kDaFinalBit) != 0; *q = e.dst & ~kDaFinalBit;
    return true;
  }
};

bool build_seg_tables(const LdbImage& ldb, SegTables* out, std::string* err);

// shared with lexer_tables.cpp: flattens a charmap into count/first/values arrays
void flatten_charmap(const FixedMap& charmap, std::vector<uint8_t>* count, std::vector<uint32_t>* first,
                     std::vector<int32_t>* values);

}  // namespace bfb200
