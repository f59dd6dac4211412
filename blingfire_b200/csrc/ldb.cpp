// ldb.cpp -- see ldb.h.  Host-only, runs once per LoadModel.
#include "ldb.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <unordered_map>

namespace bfb200 {

namespace {

inline int32_t rd_i32(const uint8_t* p) { int32_t v; std::memcpy(&v, p, 4); return v; }
inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }

// little-endian 1/2/4-byte field (FAEncodeUtils.h:292-310)
inline uint32_t rd_le(const uint8_t* p, int size) {
  if (size == 1) return p[0];
  if (size == 2) return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
  return rd_u32(p);
}
// big-endian 1..4-byte field (FAEncodeUtils.h:418-448)
inline uint32_t rd_be(const uint8_t* p, int size) {
  uint32_t v = 0;
  for (int i = 0; i < size; ++i) v = (v << 8) | p[i];
  return v;
}
inline int32_t rd_signed_le(const uint8_t* p, int size) {
  if (size == 1) return (int8_t)p[0];
  if (size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
  return rd_i32(p);
}

// Chains container (FAChains_pack_triv.h): {int SizeOfValue; int MaxCount; chains}
struct Chains {
  const uint8_t* img = nullptr;
  size_t size = 0;
  int size_of_value = 0;
  bool set(Span s) {
    if (s.n < 8) return false;
    img = s.p; size = s.n; size_of_value = rd_i32(s.p);
    return size_of_value == 1 || size_of_value == 2 || size_of_value == 4;
  }
  bool read(size_t off, std::vector<int>* out) const {
    if (off + (size_t)size_of_value > size) return false;
    int cnt = rd_signed_le(img + off, size_of_value);
    if (cnt < 0 || off + (size_t)size_of_value * (size_t)(cnt + 1) > size) return false;
    out->resize(cnt);
    for (int i = 0; i < cnt; ++i) (*out)[i] = rd_signed_le(img + off + (size_t)size_of_value * (i + 1), size_of_value);
    return true;
  }
  // value by index or -1 (FAChains_pack_triv.h:166-222)
  int at(size_t off, int idx) const {
    if (off + (size_t)size_of_value > size) return -1;
    int cnt = rd_signed_le(img + off, size_of_value);
    if (idx < 0 || idx >= cnt) return -1;
    size_t p = off + (size_t)size_of_value * (size_t)(idx + 1);
    if (p + (size_t)size_of_value > size) return -1;
    return rd_signed_le(img + p, size_of_value);
  }
};

}  // namespace

uint32_t crc32_update(const uint8_t* buf, size_t n, uint32_t crc) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  crc ^= ~0u;
  while (n--) crc = table[(crc ^ *buf++) & 0xFF] ^ (crc >> 8);
  return crc ^ ~0u;
}

int FixedMap::get(int key, int* out, int max_out) const {
  if (key < min_key || key > max_key) return -1;
  const uint8_t* arr = data + (size_t)(max_count + 1) * size_of_value * (size_t)(key - min_key);
  const int cnt = rd_signed_le(arr, size_of_value);
  if (cnt > max_count) return -1;
  if (out && max_out >= cnt)
    for (int i = 0; i < cnt; ++i) out[i] = rd_signed_le(arr + (size_t)size_of_value * (i + 1), size_of_value);
  return cnt;
}

int Automaton::dest(int state, int label, int* ow) const {
  if (state < 0 || state >= num_states()) return -1;
  const Arc* b = arcs.data() + arc_begin[state];
  const Arc* e = arcs.data() + arc_begin[state + 1];
  const Arc* it = std::lower_bound(b, e, label, [](const Arc& a, int l) { return a.label < l; });
  if (it == e || it->label != label) return -1;
  if (ow) *ow = it->ow;
  return it->dst;
}

bool LdbImage::load_file(const char* path) {
  if (!path) { err_ = "null path"; return false; }
  FILE* f = std::fopen(path, "rb");
  if (!f) { err_ = std::string("cannot open ") + path; return false; }
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 8) { std::fclose(f); err_ = "file too small"; return false; }
  owned_.resize((size_t)sz + 16);   // slack so unaligned tail reads stay in bounds
  size_t got = std::fread(owned_.data(), 1, (size_t)sz, f);
  std::fclose(f);
  if (got != (size_t)sz) { err_ = "short read"; return false; }
  base_ = owned_.data();
  size_ = (size_t)sz;
  return parse();
}

bool LdbImage::set_image(const uint8_t* bytes, size_t size) {
  if (!bytes || size < 8) { err_ = "empty image"; return false; }
  owned_.assign(bytes, bytes + size);
  owned_.resize(size + 16);
  base_ = owned_.data();
  size_ = size;
  return parse();
}

// FALDB::SetImage (FALDB.cpp:24-64) + IsValidBinary (:67-116)
bool LdbImage::parse() {
  const int count = rd_i32(base_);
  if (count <= 0 || count > 3 * 36 || 4 + 4 * (size_t)count > size_) { err_ = "bad dump count"; return false; }
  dumps_.clear(); offsets_.clear();
  for (int i = 0; i < count; ++i) {
    const int off = rd_i32(base_ + 4 + 4 * (size_t)i);
    if (off < 0 || (size_t)off >= size_) { err_ = "bad dump offset"; return false; }
    offsets_.push_back(off);
  }
  for (int i = 0; i < count; ++i) {
    // dumps are laid out back to back in increasing offset order in every shipped model;
    // the extent is only used for bounds checks, so fall back to "until end of file".
    size_t end = size_;
    if (i + 1 < count && offsets_[i + 1] > offsets_[i]) end = (size_t)offsets_[i + 1];
    dumps_.push_back(Span{base_ + offsets_[i], end - (size_t)offsets_[i]});
  }
  if (!parse_multimap(dumps_[0], &conf_, &err_)) { err_ = "bad configuration dump: " + err_; return false; }

  int verify = 0;
  get_value(kFuncGlobal, kParamVerifyLdbBin, &verify);
  if (verify) {
    if (count < 2) { err_ = "verify-ldb-bin without validation dump"; return false; }
    Span v = dumps_[count - 1];
    if (v.n < 12) { err_ = "short validation dump"; return false; }
    if (rd_u32(v.p) == 0) {
      const uint32_t exp_size = rd_u32(v.p + 4), exp_hash = rd_u32(v.p + 8);
      uint32_t size = 0, hash = 0;
      for (int i = 0; i < count - 1; ++i) {
        const int s = offsets_[i + 1] - offsets_[i];
        if (s < 0 || (size_t)offsets_[i] + (size_t)s > size_) { err_ = "bad dump extent"; return false; }
        size += (uint32_t)s;
        hash = crc32_update(base_ + offsets_[i], (size_t)s, hash);
      }
      if (size != exp_size || hash != exp_hash) { err_ = "LDB CRC32 validation failed"; return false; }
    }
  }
  return true;
}

static bool is_boolean_param(int p) {  // FALDB.cpp:130-141
  return p == kParamReverse || p == kParamNoTr || p == kParamIgnoreCase || p == kParamDictMode ||
         p == kParamNormalize || p == kParamLogScale || p == kParamUseNfst || p == kParamDoW2B ||
         p == kParamVerifyLdbBin;
}

bool LdbImage::get_value(int section, int param, int* value) const {
  *value = 0;
  const std::vector<int>* v = conf_.get(section);
  const int size = v ? (int)v->size() : -1;
  for (int i = 0; i < size; ++i) {
    const int np = (*v)[i];
    const bool isb = is_boolean_param(np);
    if (!isb) { ++i; if (i >= size) return false; }
    if (np == param) { *value = isb ? 1 : (*v)[i]; return true; }
  }
  return is_boolean_param(param);
}

// FAMultiMap_pack::SetImage (FAMultiMap_pack.cpp:22-53)
bool LdbImage::parse_multimap(Span d, MultiMap* out, std::string* err) {
  if (d.n < 8) { *err = "multimap too small"; return false; }
  const uint32_t max_key = rd_u32(d.p);
  const int so = (int)rd_u32(d.p + 4);
  if (so < 1 || so > 4 || max_key > (1u << 28)) { *err = "multimap header"; return false; }
  size_t off = 8;
  const uint8_t* offsets = d.p + off;
  off += (size_t)so * ((size_t)max_key + 1);
  if (off % 4) off += 4 - off % 4;
  if (off + 8 > d.n) { *err = "multimap extent"; return false; }
  Chains ch;
  if (!ch.set(Span{d.p + off, d.n - off})) { *err = "multimap chains"; return false; }
  out->ptr_interface_ok = ch.size_of_value == 4;
  out->rows.assign((size_t)max_key + 1, {});
  out->present.assign((size_t)max_key + 1, 0);
  for (uint32_t k = 0; k <= max_key; ++k) {
    const uint32_t vo = rd_be(offsets + (size_t)k * so, so);
    if (vo == 0) continue;
    if (!ch.read(vo - 1, &out->rows[k])) { *err = "multimap row"; return false; }
    out->present[k] = 1;
  }
  return true;
}

// FAMultiMap_pack_fixed::SetImage (FAMultiMap_pack_fixed.cpp:25-58)
bool LdbImage::parse_fixedmap(Span d, FixedMap* out, std::string* err) {
  if (d.n < 16) { *err = "fixed map too small"; return false; }
  out->size_of_value = (int)rd_u32(d.p);
  out->max_count = rd_i32(d.p + 4);
  out->min_key = rd_i32(d.p + 8);
  out->max_key = rd_i32(d.p + 12);
  out->data = d.p + 16;
  const int sv = out->size_of_value;
  if (!(sv == 1 || sv == 2 || sv == 4) || out->max_count <= 0 || out->min_key < 0 || out->max_key < out->min_key) {
    *err = "fixed map header"; return false;
  }
  const size_t need = 16 + (size_t)(out->max_count + 1) * sv * ((size_t)out->max_key - out->min_key + 1);
  if (need > d.n + 16) { *err = "fixed map extent"; return false; }
  return true;
}

// Expands FADfaPack_triv records (FADfaPack_triv.h:27-88) by BFS from the initial state.
bool LdbImage::parse_automaton(Span d, bool mealy, Automaton* out, std::string* err) {
  if (d.n < 12) { *err = "automaton too small"; return false; }
  int dst_size = rd_i32(d.p);
  if (dst_size < 1 || dst_size > 4) dst_size = 3;   // FARSDfa_pack_triv.cpp:38-40
  const int ows_offset = rd_i32(d.p + 4);
  uint32_t iwc = rd_u32(d.p + 8);
  const bool remap = (iwc & 0x80000000u) != 0;
  iwc &= 0x7fffffffu;
  size_t off = 12 + 4 * (size_t)iwc;
  if (off > d.n || iwc == 0 || (iwc % 2) != 0) { *err = "automaton alphabet"; return false; }

  out->remap = remap;
  out->num_classes = 0;
  out->class_of_iw.clear();
  Chains ows;
  if (mealy) {
    if (remap || ows_offset <= 0 || (size_t)ows_offset + 8 > d.n) { *err = "mealy automaton header"; return false; }
    if (!ows.set(Span{d.p + ows_offset, d.n - (size_t)ows_offset})) { *err = "mealy ows"; return false; }
  } else if (remap) {
    if (off + 4 > d.n) { *err = "iw map"; return false; }
    const int msz = rd_i32(d.p + off); off += 4;
    if (msz < 8 || off + (size_t)msz > d.n) { *err = "iw map size"; return false; }
    // FAIwMap_pack::SetImage (FAIwMap_pack.cpp:35-62)
    const uint8_t* m = d.p + off;
    const int sznew = rd_i32(m), ic = rd_i32(m + 4);
    if (sznew < 1 || sznew > 4 || ic <= 0 || 8 + 12 * (size_t)ic > (size_t)msz) { *err = "iw map header"; return false; }
    const uint8_t* from = m + 8;
    const uint8_t* pairs = m + 8 + 4 * (size_t)ic;
    const uint8_t* newiws = m + 8 + 12 * (size_t)ic;
    const size_t newiws_len = (size_t)msz - (8 + 12 * (size_t)ic);
    const int max_iw = rd_i32(pairs + 8 * (size_t)(ic - 1));
    if (max_iw < 0 || max_iw > (1 << 26)) { *err = "iw map range"; return false; }
    out->class_of_iw.assign((size_t)max_iw + 1, -1);
    int maxc = -1;
    for (int k = 0; k < ic; ++k) {
      const int f = rd_i32(from + 4 * (size_t)k), t = rd_i32(pairs + 8 * (size_t)k), io = rd_i32(pairs + 8 * (size_t)k + 4);
      if (f < 0 || t < f || t > max_iw || io < 0) { *err = "iw map interval"; return false; }
      if ((size_t)io + (size_t)(t - f + 1) * sznew > newiws_len) { *err = "iw map interval extent"; return false; }
      for (int iw = f; iw <= t; ++iw) {
        // GetNewIw takes the LAST interval whose From <= iw (FAFindEqualOrLess_log), then
        // rejects iw > To; intervals are disjoint and sorted in shipped models.
        const uint32_t v = rd_be(newiws + io + (size_t)(iw - f) * sznew, sznew);
        out->class_of_iw[iw] = v ? (int)v - 1 : -1;
        if (v && (int)v - 1 > maxc) maxc = (int)v - 1;
      }
    }
    out->num_classes = maxc + 1;
    off += (size_t)msz;
  }
  const int initial = (int)off;
  if ((size_t)initial >= d.n) { *err = "automaton initial"; return false; }

  std::unordered_map<int, int> id_of;      // byte offset -> renumbered id
  std::vector<int> order;                  // renumbered id -> byte offset
  std::deque<int> queue;
  auto intern = [&](int s) {
    auto it = id_of.find(s);
    if (it != id_of.end()) return it->second;
    const int id = (int)order.size();
    id_of.emplace(s, id); order.push_back(s); queue.push_back(s);
    return id;
  };
  intern(initial);

  struct RawArc { int label, dst_off, ow; };
  std::vector<std::vector<RawArc>> raw;     // per renumbered state
  std::vector<uint8_t> finals;
  std::vector<int32_t> mows;
  size_t total_arcs = 0;

  auto dec_dst = [&](const uint8_t* p, size_t idx) -> int {
    const uint32_t v = rd_be(p + idx * dst_size, dst_size);
    const uint32_t ones = dst_size == 4 ? 0xffffffffu : ((1u << (8 * dst_size)) - 1u);
    return v == ones ? kDeadState : (int)v;
  };

  while (!queue.empty()) {
    const int s = queue.front(); queue.pop_front();
    const int sid = id_of[s];
    if ((size_t)sid >= raw.size()) { raw.resize(sid + 1); finals.resize(sid + 1, 0); mows.resize(sid + 1, -1); }
    if (s < initial || (size_t)s >= d.n) { *err = "state offset out of range"; return false; }
    const uint8_t* p = d.p + s;
    const uint8_t* const end = d.p + d.n;
    const uint8_t info = *p++;
    const int iw_size = ((info & 0x18) >> 3) + 1;
    const int owc = (info & 0x60) >> 5;
    const int ow_size = owc == 3 ? 4 : owc;
    const int tr = info & 0x07;
    finals[sid] = (info & 0x80) ? 1 : 0;
    std::vector<RawArc>& arcs = raw[sid];
    const uint8_t* ows_ptr = nullptr;
    if (iw_size == 3) { *err = "iw size 3"; return false; }
    switch (tr) {
      case 4: {  // TRS_PARA
        if (p + iw_size > end) { *err = "record extent"; return false; }
        const size_t n = (size_t)rd_le(p, iw_size) + 1; p += iw_size;
        if (p + n * (size_t)(iw_size + dst_size) > end) { *err = "record extent"; return false; }
        const uint8_t* iws = p; const uint8_t* dsts = p + n * iw_size;
        for (size_t i = 0; i < n; ++i) arcs.push_back({(int)rd_le(iws + i * iw_size, iw_size), dec_dst(dsts, i), (int)i});
        p = dsts + n * dst_size;
        ows_ptr = p;
        break;
      }
      case 6: {  // TRS_IWIA
        if (p + 2 * iw_size > end) { *err = "record extent"; return false; }
        const uint32_t base = rd_le(p, iw_size); p += iw_size;
        const uint32_t mx = rd_le(p, iw_size); p += iw_size;
        if (mx < base || p + (size_t)(mx - base + 1) * dst_size > end) { *err = "record extent"; return false; }
        if (!mealy)   // FAMealyDfa_pack_triv.cpp:204-210: IWIA is not readable as Mealy
          for (uint32_t i = 0; i <= mx - base; ++i) {
            const int dd = dec_dst(p, i);
            if (dd != 0) arcs.push_back({(int)(base + i), dd, 0});
          }
        p += (size_t)(mx - base + 1) * dst_size;
        break;
      }
      case 1: {  // TRS_RANGE
        if (p + iw_size > end) { *err = "record extent"; return false; }
        const size_t n = (size_t)rd_le(p, iw_size) + 1; p += iw_size;
        if (p + n * (size_t)(2 * iw_size + dst_size) > end) { *err = "record extent"; return false; }
        const uint8_t* fr = p; const uint8_t* to = p + n * iw_size; const uint8_t* dsts = to + n * iw_size;
        if (!mealy)
          for (size_t i = 0; i < n; ++i) {
            // FAFindEqualOrLess_log picks the last range whose From <= Iw; later ranges shadow
            // earlier ones on overlap, so bound each range by the next From.
            const uint32_t f = rd_le(fr + i * iw_size, iw_size);
            uint32_t t = rd_le(to + i * iw_size, iw_size);
            if (i + 1 < n) { const uint32_t nf = rd_le(fr + (i + 1) * iw_size, iw_size); if (nf > f && t >= nf) t = nf - 1; }
            if (t >= f && (size_t)(t - f) > (1u << 22)) { *err = "range transition too wide to expand"; return false; }
            for (uint32_t l = f; l <= t && t >= f; ++l) { arcs.push_back({(int)l, dec_dst(dsts, i), 0}); if (l == 0xffffffffu) break; }
          }
        p = dsts + n * dst_size;
        break;
      }
      case 2: {  // TRS_IMPL: destination is the next record
        if (p + iw_size + ow_size > end) { *err = "record extent"; return false; }
        arcs.push_back({(int)rd_le(p, iw_size), s + 1 + iw_size + ow_size, 0});
        p += iw_size;
        ows_ptr = p;
        break;
      }
      case 0: break;
      default: { *err = "unknown transition encoding"; return false; }
    }
    if (owc) {
      if (p + ow_size > end) { *err = "record extent"; return false; }
      const int w = rd_signed_le(p, ow_size);
      if (!mealy) {
        mows[sid] = w;   // FAState2Ow_pack_triv.cpp:34-130
      } else {
        // FAMealyDfa_pack_triv.cpp:214-236: w is the offset of this state's Ows chain
        (void)ows_ptr;
        if (w < 0) { *err = "mealy ows offset"; return false; }
        for (auto& a : arcs) a.ow = ows.at((size_t)w, a.ow);
      }
    } else if (mealy) {
      for (auto& a : arcs) a.ow = -1;   // *pOw = -1 when the state carries no weights (:238-241)
    }
    if (!mealy) for (auto& a : arcs) a.ow = 0;
    total_arcs += arcs.size();
    if (total_arcs > (1u << 27)) { *err = "automaton too large"; return false; }
    for (auto& a : arcs) {
      if (a.dst_off == kDeadState) continue;
      if (a.dst_off < initial || (size_t)a.dst_off >= d.n) { *err = "destination out of range"; return false; }
      intern(a.dst_off);
    }
  }

  const int n = (int)order.size();
  raw.resize(n); finals.resize(n, 0); mows.resize(n, -1);
  out->orig.assign(order.begin(), order.end());
  out->is_final.assign(finals.begin(), finals.end());
  out->moore_ow.assign(mows.begin(), mows.end());
  out->arc_begin.assign((size_t)n + 1, 0);
  out->arcs.clear();
  out->arcs.reserve(total_arcs);
  for (int s = 0; s < n; ++s) {
    out->arc_begin[s] = (int64_t)out->arcs.size();
    std::vector<RawArc>& ra = raw[s];
    // PARA label arrays are sorted-unique in valid dumps (binary search in the reference);
    // a stable sort keeps the first occurrence first for lower_bound.
    std::stable_sort(ra.begin(), ra.end(), [](const RawArc& a, const RawArc& b) { return a.label < b.label; });
    for (size_t i = 0; i < ra.size(); ++i) {
      if (i > 0 && ra[i].label == ra[i - 1].label) continue;
      out->arcs.push_back(Arc{ra[i].label, ra[i].dst_off == kDeadState ? kDeadState : id_of[ra[i].dst_off], ra[i].ow});
    }
  }
  out->arc_begin[n] = (int64_t)out->arcs.size();
  if (!remap) {
    int maxl = -1;
    for (auto& a : out->arcs) maxl = std::max(maxl, a.label);
    out->num_classes = maxl + 1;
  }
  return true;
}

}  // namespace bfb200
