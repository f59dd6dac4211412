// ldb.h -- host-side reader of BlingFire's compiled .bin model container (LDB).
//
// Replaces, for the TextToIds / TextToWords path only, the reference's model
// container and packed-image readers:
//   FALDB                    blingfireclient.library/src/FALDB.cpp:24-116
//   FAImageDump              blingfireclient.library/src/FAImageDump.cpp:62-118
//   FAMultiMap_pack          blingfireclient.library/src/FAMultiMap_pack.cpp:22-126
//   FAMultiMap_pack_fixed    blingfireclient.library/src/FAMultiMap_pack_fixed.cpp:25-162
//   FAIwMap_pack             blingfireclient.library/src/FAIwMap_pack.cpp:35-62, inc/FAIwMap_pack.h:55-109
//   FARSDfa_pack_triv        blingfireclient.library/src/FARSDfa_pack_triv.cpp:27-399
//   FAState2Ow_pack_triv     blingfireclient.library/src/FAState2Ow_pack_triv.cpp:34-130
//   FAMealyDfa_pack_triv     blingfireclient.library/src/FAMealyDfa_pack_triv.cpp:24-244
// Format: blingfirecompile.library/inc/FADfaPack_triv.h:27-88.
//
// Unlike the reference (which walks the packed image at run time, one virtual
// call per transition), this reader is used ONCE at LoadModel time: it expands
// every automaton into explicit (state, label, dest[, weight]) arcs which the
// table builders (lexer_tables.cpp, seg_tables.cpp) then lay out for HBM.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace bfb200 {

// Section / parameter ids: blingfireclient.library/inc/FAFsmConst.h:152-273.
enum : int {
  kFuncW2H = 11, kFuncPosDict = 12, kFuncWbd = 19, kFuncGlobal = 20, kFuncI2W = 35,
  kParamFsm = 2, kParamReverse = 10, kParamDirection = 11, kParamMapMode = 16, kParamNoTr = 18,
  kParamIgnoreCase = 22, kParamArray = 24, kParamMultiMap = 25, kParamFsmType = 26,
  kParamDictMode = 31, kParamNormalize = 35, kParamDoW2B = 37, kParamDepth = 38,
  kParamMaxTag = 39, kParamLogScale = 40, kParamWord = 42, kParamPunkt = 43, kParamEos = 44,
  kParamEop = 45, kParamUseNfst = 46, kParamCharmap = 47, kParamXWord = 51, kParamSeg = 52,
  kParamIgnore = 53, kParamActData = 68, kParamMaxLength = 69, kParamVerifyLdbBin = 70,
  kParamTokenizationType = 71, kParamIdOffset = 72, kParamUseByteEncoding = 73,
  kParamNoDummyPrefix = 74, kParamStringArray = 75, kParamTokenIdMin = 76, kParamTokenIdMax = 77,
  kTypeMooreDfa = 3, kTypeMealyDfa = 7,
  kModePackTriv = 1, kModePackMph = 2, kModePackFixed = 3,
  kTokenizeBpe = 3, kTokenizeBpeOpt = 4, kTokenizeBpeOptWithMerges = 5,
  kIwAny = 0, kIwLAnchor = 1, kIwRAnchor = 2, kIwEpsilon = 3,
  kDeadState = -2,
};

struct Span {
  const uint8_t* p = nullptr;
  size_t n = 0;
};

// Key -> int vector map (actions, configuration).  FAMultiMap_pack.
struct MultiMap {
  std::vector<std::vector<int>> rows;   // rows[key]; empty+present[key]==0 means "no entry"
  std::vector<uint8_t> present;
  bool ptr_interface_ok = false;        // values stored as ints (the lexer needs Get(key, &ptr))
  const std::vector<int>* get(int key) const {
    if (key < 0 || (size_t)key >= rows.size() || !present[key]) return nullptr;
    return &rows[key];
  }
};

// Fixed-row map (charmap, I2Info).  FAMultiMap_pack_fixed.  Row = {count, v0..v[max_count-1]}.
struct FixedMap {
  int size_of_value = 0, max_count = 0, min_key = 0, max_key = -1;
  const uint8_t* data = nullptr;
  // returns count (may be any int <= max_count), or -1 when there is no entry
  int get(int key, int* out, int max_out) const;
};

// One expanded automaton: states are renumbered 0..n-1 in discovery (BFS) order from the
// initial state; original ids are byte offsets into the dump (kept in `orig`).
struct Arc {
  int32_t label;   // class id when the automaton remaps input weights, else the raw Iw
  int32_t dst;     // renumbered destination, or kDeadState
  int32_t ow;      // Mealy per-arc output weight (0 for Moore automata)
};
struct Automaton {
  bool remap = false;                 // labels are classes (FAIwMap_pack applied first)
  int num_classes = 0;                // when remap: 1 + max class id
  std::vector<int32_t> class_of_iw;   // when remap: direct table over [0, max_iw]; -1 = unmapped
  std::vector<int64_t> arc_begin;     // CSR over states, size n+1
  std::vector<Arc> arcs;              // sorted by label within a state
  std::vector<uint8_t> is_final;
  std::vector<int32_t> moore_ow;      // Moore: State2Ow (rule id), -1 if none
  std::vector<int32_t> orig;          // original byte-offset ids
  int num_states() const { return (int)is_final.size(); }
  int class_of(int iw) const {
    if (!remap) return iw;
    if (iw < 0 || (size_t)iw >= class_of_iw.size()) return -1;
    return class_of_iw[iw];
  }
  // dest on label (class or raw iw): -1 none, kDeadState, or a renumbered id
  int dest(int state, int label, int* ow = nullptr) const;
};

class LdbImage {
 public:
  // Reads the file; returns false (with error()) on any structural problem instead of
  // throwing across the C ABI like the reference (FAImageDump.cpp:98 LogAssert).
  bool load_file(const char* path);
  bool set_image(const uint8_t* bytes, size_t size);   // blingfiretokdll.cpp:1055-1071 (SetModel)
  const std::string& error() const { return err_; }

  int dump_count() const { return (int)dumps_.size(); }
  Span dump(int i) const { return (i >= 0 && i < (int)dumps_.size()) ? dumps_[i] : Span{}; }
  const MultiMap& conf() const { return conf_; }
  // FALDB::GetValue (FALDB.cpp:143-190)
  bool get_value(int section, int param, int* value) const;

  // decoders for individual dumps
  static bool parse_multimap(Span d, MultiMap* out, std::string* err);
  static bool parse_fixedmap(Span d, FixedMap* out, std::string* err);
  // Expands the automaton dump.  `mealy` selects the FAMealyDfa_pack_triv reading of the
  // trailing per-state word (offset into the Ows chains) instead of the Moore rule id.
  static bool parse_automaton(Span d, bool mealy, Automaton* out, std::string* err);

 private:
  bool parse();
  std::vector<uint8_t> owned_;
  const uint8_t* base_ = nullptr;
  size_t size_ = 0;
  std::vector<Span> dumps_;
  std::vector<int> offsets_;
  MultiMap conf_;
  std::string err_;
};

uint32_t crc32_update(const uint8_t* buf, size_t n, uint32_t crc);  // FAUtils_cl.cpp:148-159

}  // namespace bfb200
