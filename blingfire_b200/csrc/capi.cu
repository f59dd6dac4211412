// capi.cu -- the C ABI (include/blingfiretokdll_b200.h): model lifecycle, device residency of
// the flattened tables, and the host-side batch pipeline around the kernels.
//
// Host code only orchestrates: every tokenizing entry point runs the kernels of wp_kernel.cu /
// lex_kernel.cu / sp_kernel.cu on the GPU.  There is deliberately no CPU path; if CUDA is unavailable
// the calls fail loudly.
//
// Concurrency: after LoadModel a model is immutable (SetNoDummyPrefix aside, like the reference,
// blingfiretokdll.cpp:1670-1679) and every entry point is re-entrant: a call leases a private
// working context (streams, device and pinned buffers) from the model's pool, so N caller threads on
// one handle run N pipelines side by side (README.md:105 "can be called from multiple threads").
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/blingfiretokdll_b200.h"
#include "copy_pool.h"
#include "ldb.h"
#include "lex_kernel.cuh"
#include "lexer_tables.h"
#include "seg_tables.h"
#include "sp_kernel.cuh"
#include "wp_kernel.cuh"
#include "wp_model.h"

using namespace bfb200;

namespace {

thread_local std::string g_last_error;
thread_local double g_last_kernel_ms = 0.0;   // device time of the tokenization kernels of the last host batch call
std::atomic<int64_t> g_launches{0};

void set_error(const std::string& e) { g_last_error = e; }
bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (!cuda_ok(cudaMalloc(&p, n * sizeof(T)), "cudaMalloc")) return false;
    cap = n;
    return true;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// CPUs of the current model's NUMA node (set by CtxLease for the duration of a call): pinned staging
// memory is allocated while the calling thread runs there, so that first touch places it next to the GPU
thread_local const cpu_set_t* g_numa_cpus = nullptr;

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cpu_set_t saved;
    const bool moved = g_numa_cpus && n * sizeof(T) >= (1u << 20) && sched_getaffinity(0, sizeof(saved), &saved) == 0 &&
                       sched_setaffinity(0, sizeof(cpu_set_t), g_numa_cpus) == 0;
    const bool ok = cuda_ok(cudaMallocHost(&p, n * sizeof(T)), "cudaMallocHost");
    if (moved) sched_setaffinity(0, sizeof(saved), &saved);
    if (!ok) return false;
    cap = n;
    return true;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// per-slot working set of the host batch pipeline
constexpr int kMaxSlots = 8;
// chunks in flight in the host pipeline / chunks enqueued behind the one whose counts the host waits for
// (BLINGFIRE_B200_SLOTS / BLINGFIRE_B200_AHEAD override the defaults: tuning knobs)
static const int kSlots = [] { const char* e = std::getenv("BLINGFIRE_B200_SLOTS"); const int v = e ? std::atoi(e) : 8; return v < 2 ? 2 : (v > kMaxSlots ? kMaxSlots : v); }();
static const int kAhead = [] { const char* e = std::getenv("BLINGFIRE_B200_AHEAD"); const int v = e ? std::atoi(e) : 4; return v < 1 ? 1 : (v >= kSlots ? kSlots - 1 : v); }();

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t counts_ready = nullptr;   // recorded after the chunk's row offsets were copied to the host
  cudaEvent_t k_begin = nullptr, k_end = nullptr;   // around the engine's kernels (device time of the tokenization alone)
  cudaEvent_t h2d_done = nullptr;       // the chunk's text and offsets have left the (staging) host buffers
  DevBuf<uint8_t> text;
  DevBuf<int64_t> offsets;
  DevBuf<int32_t> ids;
  DevBuf<int32_t> counts;      // ndocs + 1 (trailing zero for the scan)
  DevBuf<int64_t> row_off;     // ndocs + 1
  DevBuf<int32_t> csr;         // compact ids of the chunk (int32, or uint16 packed two per word)
  DevBuf<unsigned long long> counter;
  PinBuf<int64_t> h_row_off;
  PinBuf<int64_t> h_offsets;   // staging of the chunk's document offsets when the caller's array is pageable
  PinBuf<uint8_t> h_text;      // staging of the chunk's text when the caller's buffer is pageable
  PinBuf<int32_t> h_csr;       // staging of the chunk's ids: row-major API, or a pageable CSR destination
  DevBuf<uint8_t> sp_arena;    // [pos-dict] engine: per-warp scratch for documents beyond the smem window
  DevBuf<uint8_t> sp_overflow; // [pos-dict] BPE: grid-wide arc scratch for segments beyond the private one
  // generic lexer engine scratch
  DevBuf<uint16_t> lex_cls;
  DevBuf<int32_t> lex_ncps, lex_tri, lex_tri_count, lex_boff;
  // bookkeeping of the chunk in flight
  int64_t doc0 = 0, ndocs = 0;
  int64_t out_base = 0, out_n = 0;      // where the chunk's ids go in the caller's CSR buffer
  // pageable caller buffers: copies between them and the pinned staging buffers run on the copy pool, asynchronously
  CopyJob in_job, out_job;
  int64_t staged_doc0 = -1;             // the chunk whose text is (being) staged into h_text
  void release() {
    lex_cls.release(); lex_ncps.release(); lex_tri.release(); lex_tri_count.release(); lex_boff.release();
    sp_arena.release(); sp_overflow.release();
    text.release(); offsets.release(); ids.release(); counts.release(); row_off.release(); csr.release();
    counter.release(); h_row_off.release(); h_offsets.release(); h_text.release(); h_csr.release();
    if (counts_ready) cudaEventDestroy(counts_ready);
    counts_ready = nullptr;
    if (k_begin) cudaEventDestroy(k_begin);
    if (h2d_done) cudaEventDestroy(h2d_done);
    h2d_done = nullptr;
    if (k_end) cudaEventDestroy(k_end);
    k_begin = k_end = nullptr;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
};

// everything one host call needs; leased from the model for the duration of the call
struct Ctx {
  Slot slots[kMaxSlots];
  PinBuf<int32_t> h_words;     // single-document calls: [ncps, tri_count, triples...] / ids, starts, ends
  void release() {
    for (auto& s : slots) s.release();
    h_words.release();
  }
};

// FAStringArray_pack (FAStringArray_pack.cpp:22-71): count, count+1 offsets, bytes
struct I2w {
  bool present = false;
  std::vector<uint8_t> image;  // private copy of the dump
  int count = 0;
  const uint32_t* offsets = nullptr;
  const uint8_t* data = nullptr;
  int min_id = 0, max_id = 1000000000;     // FALimits::MaxArrSize (blingfiretokdll.cpp:999-1000)
};

struct Model {
  int device = 0;
  int engine = 0;              // 1 = fused WordPiece kernel
  bool has_wbd = false, has_seg = false;
  LexerTables T;               // big host vectors are dropped after upload
  WpBlob blob;
  // device-resident model
  void* d_trans = nullptr;
  int32_t* d_tag = nullptr;
  uint16_t* d_cls = nullptr;
  uint8_t* d_blob = nullptr;
  uint32_t* d_clsx = nullptr;
  WpWordSlot* d_words = nullptr;
  // generic lexer engine (any [wbd] model)
  bool lex_ok = false;
  int32_t* d_ow = nullptr;
  int32_t* d_act_begin = nullptr;
  int32_t* d_act_data = nullptr;
  uint32_t* d_fn_ini = nullptr;
  uint16_t* d_cls_words = nullptr;
  // [pos-dict] engine (Unigram-LM / BPE over the Mealy automaton)
  SegTables S;
  std::atomic<bool> no_dummy_prefix{false};   // SetNoDummyPrefix may flip it after load
  DaEntry* d_da = nullptr;
  uint16_t* d_sym = nullptr;
  SegInfo* d_info = nullptr;
  uint8_t* d_norm_count = nullptr;
  uint32_t* d_norm_first = nullptr;
  int32_t* d_norm_values = nullptr;
  int32_t* d_bpe_ord = nullptr;
  int32_t* d_bpe_id_of_ord = nullptr;
  WpWordSlot* d_seg_memo = nullptr;   // BPE: run-time memo of resolved segments (sp_bpe.cuh)
  WpWords seg_memo{};
  I2w i2w;
  int32_t max_tag = 0;         // largest id the model itself can emit (UnkId aside)

  // working contexts of the host-pointer entry points
  std::mutex pool_mu;
  std::vector<std::unique_ptr<Ctx>> all_ctx;
  std::vector<Ctx*> idle_ctx;
  // device-pointer entry points: scratch per caller stream (work is stream-ordered, so a stream can
  // reuse its scratch call after call), and a ring of work counters for the scratch-free engine
  std::mutex dev_mu;
  std::map<cudaStream_t, std::unique_ptr<Slot>> dev_slots;
  DevBuf<unsigned long long> dev_counters;
  std::atomic<uint32_t> dev_counter_next{0};
  // pageable caller buffers: copy threads on the GPU's NUMA node
  std::once_flag pool_once;
  std::unique_ptr<CopyPool> copy_pool;
  cpu_set_t numa_cpus;
  bool numa_known = false;

  ~Model() {
    cudaSetDevice(device);
    for (auto& c : all_ctx) c->release();
    for (auto& kv : dev_slots) kv.second->release();
    dev_counters.release();
    if (d_trans) cudaFree(d_trans);
    if (d_tag) cudaFree(d_tag);
    if (d_cls) cudaFree(d_cls);
    if (d_blob) cudaFree(d_blob);
    if (d_clsx) cudaFree(d_clsx);
    if (d_words) cudaFree(d_words);
    if (d_ow) cudaFree(d_ow);
    if (d_act_begin) cudaFree(d_act_begin);
    if (d_act_data) cudaFree(d_act_data);
    if (d_fn_ini) cudaFree(d_fn_ini);
    if (d_cls_words) cudaFree(d_cls_words);
    if (d_da) cudaFree(d_da);
    if (d_sym) cudaFree(d_sym);
    if (d_info) cudaFree(d_info);
    if (d_norm_count) cudaFree(d_norm_count);
    if (d_norm_first) cudaFree(d_norm_first);
    if (d_norm_values) cudaFree(d_norm_values);
    if (d_bpe_ord) cudaFree(d_bpe_ord);
    if (d_bpe_id_of_ord) cudaFree(d_bpe_id_of_ord);
    if (d_seg_memo) cudaFree(d_seg_memo);
  }
};

// RAII lease of a working context
struct CtxLease {
  Model* m;
  Ctx* c = nullptr;
  const cpu_set_t* saved_numa;
  explicit CtxLease(Model* model) : m(model), saved_numa(g_numa_cpus) {
    std::lock_guard<std::mutex> l(m->pool_mu);
    if (m->idle_ctx.empty()) {
      m->all_ctx.emplace_back(new Ctx());
      c = m->all_ctx.back().get();
    } else {
      c = m->idle_ctx.back();
      m->idle_ctx.pop_back();
    }
    g_numa_cpus = m->numa_known ? &m->numa_cpus : nullptr;
  }
  ~CtxLease() {
    g_numa_cpus = saved_numa;
    std::lock_guard<std::mutex> l(m->pool_mu);
    m->idle_ctx.push_back(c);
  }
};

CopyPool* copy_pool_of(Model* m) {
  std::call_once(m->pool_once, [m] {
    int threads = 12;
    if (const char* e = std::getenv("BLINGFIRE_B200_COPY_THREADS")) threads = std::max(0, std::atoi(e));
    if (m->numa_known) threads = std::min(threads, std::max(1, CPU_COUNT(&m->numa_cpus) - 1));
    m->copy_pool.reset(new CopyPool(threads, m->numa_known ? &m->numa_cpus : nullptr));
  });
  return m->copy_pool.get();
}

// true when the driver would stage a copy from/to this host pointer (plain malloc / numpy memory)
bool is_pageable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

template <typename T>
bool upload(T** dst, const T* src, size_t n, size_t slack_elems = 0) {
  if (!cuda_ok(cudaMalloc((void**)dst, (n + slack_elems) * sizeof(T)), "cudaMalloc(model)")) return false;
  return cuda_ok(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(model)");
}

// blingfiretokdll.cpp:997-1045: the [i2w] section (string array dump, regular id range)
bool load_i2w(const LdbImage& ldb, I2w* w) {
  const std::vector<int>* v = ldb.conf().get(kFuncI2W);
  if (!v) return true;
  const int n = (int)v->size();
  int dump = -1;
  for (int i = 0; i < n; ++i) {
    const int p = (*v)[i];
    if (p == kParamStringArray && i + 1 < n) dump = (*v)[++i];
    else if (p == kParamTokenIdMin && i + 1 < n) w->min_id = (*v)[++i];
    else if (p == kParamTokenIdMax && i + 1 < n) w->max_id = (*v)[++i];
  }
  if (dump < 0) return true;
  const Span d = ldb.dump(dump);
  if (!d.p || d.n < 4) { set_error("[i2w]: bad string-array dump"); return false; }
  w->image.assign(d.p, d.p + d.n);
  uint32_t count;
  std::memcpy(&count, w->image.data(), 4);
  if (count > 0x7fffffffu || 4 + 4 * ((size_t)count + 1) > w->image.size()) { set_error("[i2w]: truncated string array"); return false; }
  w->count = (int)count;
  w->offsets = reinterpret_cast<const uint32_t*>(w->image.data() + 4);
  w->data = w->image.data() + 4 + 4 * ((size_t)count + 1);
  const size_t data_bytes = w->image.size() - (4 + 4 * ((size_t)count + 1));
  for (int i = 0; i < w->count; ++i)
    if (w->offsets[i + 1] < w->offsets[i] || w->offsets[i + 1] > data_bytes) { set_error("[i2w]: bad offsets"); return false; }
  w->present = true;
  return true;
}

// The dense state x class table of a wide (32-bit entries) lexer model, built in HBM from the stored arcs: the table is
// 9.3 GB for bert_multi_cased with 0.03 % of its cells set -- staging it on the host took 12 s and 9.3 GB of RAM.
// One thread per state: its arcs, then -- for the rare state with an IW_ANY arc -- the fallback into every empty cell
// (FALexTools_t.h:266-270), exactly like lexer_tables.cpp does for the host copy.
__global__ void trans_fill_kernel(uint32_t* __restrict__ trans, const int64_t* __restrict__ arc_begin, const uint32_t* __restrict__ arc_label,
                                  const uint32_t* __restrict__ arc_dst, const uint32_t* __restrict__ any_dst, int ns, uint32_t W) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  uint32_t* row = trans + (size_t)s * W;
  for (int64_t k = arc_begin[s]; k < arc_begin[s + 1]; ++k) row[arc_label[k]] = arc_dst[k];
  const uint32_t any = any_dst[s];
  if (any != 0xFFFFFFFFu)
    for (uint32_t c = 0; c < W; ++c) if (row[c] == 0xFFFFFFFFu) row[c] = any;
}

bool build_trans_on_device(Model* m) {
  const LexerTables& T = m->T;
  const size_t W = (size_t)T.NC + 1, cells = (size_t)T.NS * W;
  uint32_t* d = nullptr;
  int64_t* d_begin = nullptr; uint32_t *d_label = nullptr, *d_dst = nullptr, *d_any = nullptr;
  bool ok = cuda_ok(cudaMalloc(&d, (cells + 16) * sizeof(uint32_t)), "cudaMalloc (transition table)") &&
            cuda_ok(cudaMemset(d, 0xFF, (cells + 16) * sizeof(uint32_t)), "cudaMemset") &&
            upload(&d_begin, T.arc_begin.data(), T.arc_begin.size()) && upload(&d_label, T.arc_label.data(), T.arc_label.size(), 1) &&
            upload(&d_dst, T.arc_dst.data(), T.arc_dst.size(), 1) && upload(&d_any, T.any_dst.data(), T.any_dst.size());
  if (ok) {
    trans_fill_kernel<<<(T.NS + 127) / 128, 128>>>(d, d_begin, d_label, d_dst, d_any, T.NS, (uint32_t)W);
    ok = cuda_ok(cudaGetLastError(), "table fill launch") && cuda_ok(cudaDeviceSynchronize(), "table fill");
  }
  cudaFree(d_begin); cudaFree(d_label); cudaFree(d_dst); cudaFree(d_any);
  if (!ok) { cudaFree(d); return false; }
  m->d_trans = d;
  return true;
}

Model* finish_model(std::unique_ptr<Model> m, const LdbImage& ldb) {
  m->has_wbd = ldb.conf().get(kFuncWbd) != nullptr;
  m->has_seg = ldb.conf().get(kFuncPosDict) != nullptr;
  if (!load_i2w(ldb, &m->i2w)) return nullptr;
  if (!m->has_wbd && !m->has_seg) {
    // an *.i2w file holds only the id -> text array (IdsToText); nothing to put on the GPU
    if (m->i2w.present) return m.release();
    set_error("model has neither a [wbd], a [pos-dict] nor an [i2w] section");
    return nullptr;
  }
  if (!cuda_ok(cudaGetDevice(&m->device), "cudaGetDevice")) return nullptr;
  {
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), m->device) == cudaSuccess && !std::getenv("BLINGFIRE_B200_NO_NUMA"))
      m->numa_known = numa_cpus_of_pci(bus, &m->numa_cpus);
    else
      cudaGetLastError();
  }
  if (m->has_wbd && !m->has_seg) {
    std::string err;
    // a table with 32-bit entries is built in HBM from the stored arcs, not staged on the host
    if (!build_lexer_tables(ldb, &m->T, &err, /*dense_wide=*/false)) { set_error("lexer model: " + err); return nullptr; }
    LexerTables& T = m->T;
    const size_t cells = (size_t)T.NS * ((size_t)T.NC + 1);
    if (!T.dense_on_host) { if (!build_trans_on_device(m.get())) return nullptr; }
    else if (T.wide_states) { if (!upload((uint32_t**)&m->d_trans, T.trans32.data(), cells, 16)) return nullptr; }
    else { if (!upload((uint16_t**)&m->d_trans, T.trans16.data(), cells, 16)) return nullptr; }
    // generic lexer engine: serves TextToWords for every [wbd] model, and TextToIds for the
    // grammars outside the FastPath shape
    if (T.max_depth <= kMaxLexDepth) {
      if (!upload(&m->d_ow, T.ow_of_state.data(), T.ow_of_state.size())) return nullptr;
      if (!upload(&m->d_act_begin, T.act_begin.data(), T.act_begin.size())) return nullptr;
      if (!upload(&m->d_act_data, T.act_data.data(), T.act_data.size(), 4)) return nullptr;
      if (!upload(&m->d_fn_ini, T.fn_ini.data(), T.fn_ini.size(), 4)) return nullptr;
      if (!upload(&m->d_cls_words, T.cls_words_of_cp.data(), T.cls_words_of_cp.size())) return nullptr;
      m->lex_ok = true;
    }
    if (T.charmap_one_to_one && !upload(&m->d_cls, T.cls_of_cp.data(), T.cls_of_cp.size())) return nullptr;
    if (T.fast.ok && T.charmap_one_to_one && T.max_token_length <= 400) {
      build_wp_blob(T, &m->blob);
      if (!upload(&m->d_tag, T.tag_of_state.data(), T.tag_of_state.size())) return nullptr;
      if (!upload(&m->d_blob, m->blob.bytes.data(), m->blob.bytes.size())) return nullptr;
      if (!upload(&m->d_clsx, T.clsx_of_cp.data(), T.clsx_of_cp.size())) return nullptr;
      if (!upload(&m->d_words, m->blob.word_slots.data(), m->blob.word_slots.size())) return nullptr;
      m->blob.words.slots = m->d_words;
      std::vector<WpWordSlot>().swap(m->blob.word_slots);
      std::vector<uint32_t>().swap(T.clsx_of_cp);
      m->engine = 1;
    } else if (m->lex_ok && T.charmap_one_to_one) {
      m->engine = 2;
    }
    for (int32_t t : T.tag_of_state) m->max_tag = std::max(m->max_tag, t);
    for (int32_t t : T.act_data) m->max_tag = std::max(m->max_tag, t);
    // the dense table is the bulk of the host footprint; the device copy is the one that serves
    DenseVec<uint16_t>().swap(T.trans16);
    DenseVec<uint32_t>().swap(T.trans32);
    T.dense_on_host = false;               // (next() answers from the stored arcs from here on)
  }
  if (m->has_seg) {
    // blingfiretokdll.cpp:1636-1645: a [pos-dict] model is served by the segmentation engine
    std::string err;
    SegTables& S = m->S;
    if (!build_seg_tables(ldb, &S, &err)) { set_error("segmentation model: " + err); return nullptr; }
    if (S.max_arc_len > 1024) { set_error("segmentation model: tokens longer than 1024 symbols are not served"); return nullptr; }
    if (!upload(&m->d_da, S.da.data(), S.da.size(), 8)) return nullptr;
    if (!upload(&m->d_sym, S.sym_of_cp.data(), S.sym_of_cp.size())) return nullptr;
    if (!upload(&m->d_info, S.info.data(), S.info.size(), 1)) return nullptr;
    if (S.has_charmap) {
      if (!upload(&m->d_norm_count, S.norm_count.data(), S.norm_count.size())) return nullptr;
      if (!upload(&m->d_norm_first, S.norm_first.data(), S.norm_first.size())) return nullptr;
      if (!upload(&m->d_norm_values, S.norm_values.data(), S.norm_values.size(), 16)) return nullptr;
    }
    if (S.bpe_ord_ok) {
      if (!upload(&m->d_bpe_ord, S.bpe_ord.data(), S.bpe_ord.size(), 1)) return nullptr;
      if (!upload(&m->d_bpe_id_of_ord, S.bpe_id_of_ord.data(), S.bpe_id_of_ord.size(), 1)) return nullptr;
      // the segment memo: 2 x 2^18 slots of 64 B (32 MB), empty at load
      m->seg_memo = sp_seg_memo_params(S.alphabet, 18);
      if (m->seg_memo.max_len > 0 && !std::getenv("BLINGFIRE_B200_NO_MEMO")) {
        const size_t bytes = ((size_t)2 << m->seg_memo.log2_size) * sizeof(WpWordSlot);
        if (!cuda_ok(cudaMalloc(&m->d_seg_memo, bytes), "cudaMalloc (segment memo)") ||
            !cuda_ok(cudaMemset(m->d_seg_memo, 0, bytes), "cudaMemset (segment memo)"))
          return nullptr;
        m->seg_memo.slots = m->d_seg_memo;
      } else {
        m->seg_memo.max_len = 0;
      }
    }
    m->no_dummy_prefix = S.no_dummy_prefix;
    for (const SegInfo& i : S.info) if (i.id != INT32_MIN) m->max_tag = std::max<int64_t>(m->max_tag, (int64_t)i.id + S.id_offset);
    m->engine = 3;
  }
  return m.release();
}

SpModelDev make_sp_model(const Model* m) {
  SpModelDev d{};
  const SegTables& S = m->S;
  d.da = m->d_da; d.root = S.root; d.sym_of_cp = m->d_sym; d.info = m->d_info; d.info_count = (int)S.info.size();
  d.norm_count = S.has_charmap ? m->d_norm_count : nullptr; d.norm_first = m->d_norm_first; d.norm_values = m->d_norm_values;
  d.tok_algo = S.tok_algo; d.id_offset = S.id_offset; d.use_raw_bytes = S.use_raw_bytes; d.no_dummy_prefix = m->no_dummy_prefix.load();
  d.delim_inside_tokens = S.delim_inside_tokens; d.delim_is_token = S.delim_is_token; d.max_arc_len = S.max_arc_len;
  d.bpe_ord = S.bpe_ord_ok ? m->d_bpe_ord : nullptr; d.bpe_id_of_ord = m->d_bpe_id_of_ord;
  d.bpe_singles_first = S.bpe_singles_first;
  d.seg_memo = m->seg_memo;
  return d;
}

WpLaunch make_launch(const Model* m) {
  WpLaunch L{};
  L.blob = m->d_blob;
  L.layout = m->blob.layout;
  L.trans = m->d_trans;
  L.wide = m->T.wide_states;
  L.tag_of_state = m->d_tag;
  L.clsx_of_cp = m->d_clsx;
  L.words = m->blob.words;
  L.NC1 = (uint32_t)m->T.NC + 1;
  L.first_final = m->T.first_final;
  L.cls_caret = m->T.cls_caret;
  L.cls_dollar = m->T.cls_dollar;
  L.max_token_length = m->T.max_token_length;
  return L;
}

LexModelDev make_lex_model(const Model* m) {
  LexModelDev d{};
  d.trans = m->d_trans; d.wide = m->T.wide_states;
  d.ow_of_state = m->d_ow; d.act_begin = m->d_act_begin; d.act_data = m->d_act_data; d.fn_ini = m->d_fn_ini;
  d.fn_count = (int)m->T.fn_ini.size();
  d.NC1 = (uint32_t)m->T.NC + 1; d.first_final = m->T.first_final; d.cls_caret = m->T.cls_caret; d.cls_dollar = m->T.cls_dollar;
  d.initial = m->T.initial; d.max_depth = m->T.max_depth; d.max_token_length = m->T.max_token_length;
  return d;
}

// b0 = 4-byte-aligned absolute offset the device text starts at, first_off = offsets[doc0]
LexLaunch make_lex_launch(Slot& s, const uint8_t* text_biased, const int64_t* d_offsets, int64_t first_off, int64_t b1,
                          int64_t ndocs, const uint16_t* cls_table, int tri_mul) {
  LexLaunch X{};
  X.text = text_biased;
  X.offsets = d_offsets;
  X.ndocs = ndocs;
  X.text_bytes = b1;
  X.base_offset = first_off;
  X.cls_of_cp = cls_table;
  X.cls_buf = s.lex_cls.p;
  X.ncps = s.lex_ncps.p;
  X.tri_buf = s.lex_tri.p;
  X.tri_count = s.lex_tri_count.p;
  X.tri_mul = tri_mul;
  return X;
}

bool ensure_stream(Slot& s) {
  if (s.stream) return true;
  if (!cuda_ok(cudaEventCreateWithFlags(&s.counts_ready, cudaEventDisableTiming), "cudaEventCreate")) return false;
  if (!cuda_ok(cudaEventCreate(&s.k_begin), "cudaEventCreate") || !cuda_ok(cudaEventCreate(&s.k_end), "cudaEventCreate")) return false;
  if (!cuda_ok(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming), "cudaEventCreate")) return false;
  return cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
}

// ids a document of `len` bytes can produce at most (before the MaxIdsPerDoc cap): one per byte for the
// lexer engines; the [pos-dict] front end prepends the dummy prefix and its charmap maps one code point
// to up to two symbols per input symbol on average (blingfiretokdll.cpp:1387-1425: buffers of 2(n+1))
inline int64_t max_ids_of_doc(const Model* m, int64_t len) {
  if (len <= 0) return 0;
  if (m->engine != 3) return len;
  return m->S.has_charmap ? 2 * (len + 1) + 2 : len + 1;
}

// ---- small device helpers of the host pipeline (the tokenization itself is in the *_kernel.cu files) ----

// UnkId == INT32_MIN collides with the fused kernel's "no piece starts here" marker: the kernel runs with
// a stand-in and the ids are patched afterwards (one pass over the written part of the rows)
__global__ void patch_unk_kernel(int32_t* ids, const int32_t* counts, int64_t ndocs, int max_ids, int32_t from, int32_t to) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t d = warp; d < ndocs; d += nwarps) {
    const int c = counts[d];
    int32_t* row = ids + d * (int64_t)max_ids;
    for (int k = lane; k < c; k += 32) if (row[k] == from) row[k] = to;
  }
}

// longest document of a device-resident batch (sizes the [pos-dict] engine's arena)
__global__ void max_doc_len_kernel(const int64_t* offsets, int64_t ndocs, unsigned long long* out) {
  unsigned long long mx = 0;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < ndocs; d += (int64_t)gridDim.x * blockDim.x) {
    const int64_t len = offsets[d + 1] - offsets[d];
    if (len > 0 && (unsigned long long)len > mx) mx = (unsigned long long)len;
  }
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, mx, o); if (v > mx) mx = v; }
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

// ---- engines ----

// [pos-dict] engine over documents already on the device.  Documents whose symbols exceed the
// shared-memory window use a per-warp arena sized for the longest document of the launch; the grid
// shrinks if the arena would not fit.  With starts/ends the offsets ride in the arena as well.
bool launch_segmentation(Model* m, Slot& s, const uint8_t* text_biased, const int64_t* d_offsets, int64_t b1, int64_t ndocs,
                         int64_t max_len, int32_t* d_ids, int32_t* d_counts, int32_t* d_starts, int32_t* d_ends, int max_ids,
                         int unk, cudaStream_t stream, int* launches) {
  int64_t cap64 = (m->S.has_charmap ? 2 * (max_len + 1) : max_len + 1) + 2;
  const bool is_bpe = m->S.tok_algo == kTokenizeBpe || m->S.tok_algo == kTokenizeBpeOpt || m->S.tok_algo == kTokenizeBpeOptWithMerges;
  // the Unigram fast path keeps everything in shared memory when even the worst-case staging fits
  // (the same predicate the kernel uses: sp_kernel.cu fast_model); BPE always keeps its arc scratch in the arena
  const bool unigram_fast = !is_bpe && !m->S.use_raw_bytes && !m->S.delim_inside_tokens && m->S.delim_is_token;
  if (unigram_fast && cap64 <= sp_fast_cap(m->S.tok_algo, m->S.max_arc_len, m->S.use_raw_bytes) && !d_starts) cap64 = 16;
  if (cap64 > (1ll << 28)) { set_error("document too large for the segmentation engine"); return false; }
  const int cap = (int)cap64;
  const int64_t per_warp = sp_arena_bytes_per_warp(cap, m->S.max_arc_len);
  int warps = sp_preferred_warps(m->S.tok_algo);
  const int64_t budget = 6ll << 30;
  if (per_warp * warps > budget) warps = (int)std::max<int64_t>(8, (budget / per_warp) / 8 * 8);
  if (per_warp * warps > (24ll << 30)) { set_error("document too large for the segmentation engine arena"); return false; }
  if (!s.sp_arena.reserve((size_t)(per_warp * warps)) || !s.counter.reserve(2)) return false;
  const int64_t ovf_entries = is_bpe ? sp_overflow_entries(cap, m->S.max_arc_len) : 1;
  if (!s.sp_overflow.reserve((size_t)ovf_entries * 16)) return false;
  SpLaunch X{};
  X.overflow = s.sp_overflow.p; X.overflow_cap = ovf_entries;
  X.text = text_biased; X.offsets = d_offsets; X.ndocs = ndocs; X.text_bytes = b1;
  X.ids = d_ids; X.counts = d_counts; X.starts = d_starts; X.ends = d_ends; X.max_ids = max_ids; X.unk_id = unk;
  X.work_counter = s.counter.p; X.arena = s.sp_arena.p; X.arena_stride = per_warp; X.arena_cap = cap; X.grid_warps = warps;
  return cuda_ok(sp_tokenize_launch(X, make_sp_model(m), stream, launches), "segmentation launch");
}

// generic lexer engine (any [wbd] grammar): decode/classify -> Process_int triples -> wp post-pass
bool launch_lexer_ids(Model* m, Slot& s, const uint8_t* text_biased, const int64_t* d_offsets, int64_t first_off, int64_t b1,
                      int64_t ndocs, int32_t* d_ids, int32_t* d_counts, int max_ids, int unk, cudaStream_t stream, int* launches) {
  const size_t span = (size_t)(b1 - first_off) + 8;
  if (!s.lex_cls.reserve(span) || !s.lex_ncps.reserve((size_t)ndocs) || !s.lex_tri_count.reserve((size_t)ndocs) ||
      !s.lex_tri.reserve(6 * span))
    return false;
  LexLaunch X = make_lex_launch(s, text_biased, d_offsets, first_off, b1, ndocs, m->d_cls, 2);
  if (!cuda_ok(lex_launch(X, make_lex_model(m), stream, launches), "lexer launch")) return false;
  return cuda_ok(lex_wp_launch(X, d_ids, d_counts, max_ids, unk, stream, launches), "post-pass launch");
}

// fused WordPiece kernel.  `counter` = a zeroable device word private to this launch.
bool launch_wordpiece(Model* m, const uint8_t* text_biased, const int64_t* d_offsets, int64_t b1, int64_t ndocs, int32_t* d_ids,
                      int32_t* d_counts, int max_ids, int unk, unsigned long long* counter, cudaStream_t stream, int* launches) {
  WpLaunch L = make_launch(m);
  L.text = text_biased;            // biased: absolute offsets index it directly
  L.offsets = d_offsets;
  L.ndocs = ndocs;
  L.text_bytes = b1;
  L.ids = d_ids;
  L.counts = d_counts;
  L.max_ids = max_ids;
  // the kernel marks positions without a piece with INT32_MIN; an UnkId of that value runs under a stand-in
  // no rule id of the model can equal (rule ids are the model's own, max_tag < INT32_MAX) and is patched back
  const bool patch = unk == INT32_MIN;
  L.unk_id = patch ? INT32_MAX : unk;
  L.work_counter = counter;
  WpLaunchInfo info{};
  if (!cuda_ok(wp_tokenize_launch(L, stream, &info), "tokenize launch")) return false;
  *launches += info.launches;
  if (patch && ndocs > 0) {
    patch_unk_kernel<<<(int)std::min<int64_t>((ndocs + 7) / 8, 148 * 16), 256, 0, stream>>>(d_ids, d_counts, ndocs, max_ids, INT32_MAX, INT32_MIN);
    if (!cuda_ok(cudaGetLastError(), "patch launch")) return false;
    *launches += 1;
  }
  return true;
}

// how the ids of a host batch travel back
enum class OutKind { kCsr32, kCsr16 };

struct HostBatch {
  const char* utf8; const int64_t* offsets; int64_t ndocs; int max_ids; int unk;
  bool stage_in = false;       // the caller's text/offsets are pageable: stage through pinned chunk buffers
  OutKind out = OutKind::kCsr32;
};

// Pageable caller text: start copying the chunk into the slot's pinned staging buffers (copy-pool threads).
bool stage_chunk(Model* m, Slot& s, const HostBatch& B, int64_t doc0, int64_t ndocs) {
  const int64_t b0 = B.offsets[doc0] & ~(int64_t)3, b1 = B.offsets[doc0 + ndocs];
  const size_t nbytes = (size_t)(b1 - b0);
  if (!s.h_text.reserve(nbytes + 64) || !s.h_offsets.reserve((size_t)ndocs + 1)) return false;
  CopyPool* pool = copy_pool_of(m);
  pool->submit(s.h_text.p, B.utf8 + b0, nbytes, &s.in_job);
  pool->submit(s.h_offsets.p, B.offsets + doc0, ((size_t)ndocs + 1) * sizeof(int64_t), &s.in_job);
  s.staged_doc0 = doc0;
  return true;
}

// Enqueue one chunk [doc0, doc0+ndocs) of a host CSR batch on slot s: H2D, tokenize, scan, compact,
// D2H of the row offsets.  Offsets stay absolute; the device text pointer is biased instead.
bool enqueue_chunk(Model* m, Slot& s, const HostBatch& B, int64_t doc0, int64_t ndocs) {
  if (!ensure_stream(s)) return false;
  const int64_t* offsets = B.offsets;
  const int max_ids = B.max_ids;
  const int64_t b0 = offsets[doc0] & ~(int64_t)3;       // keep 32-bit word alignment of absolute offsets
  const int64_t b1 = offsets[doc0 + ndocs];
  const size_t nbytes = (size_t)(b1 - b0);
  if (!s.text.reserve(nbytes + 64) || !s.offsets.reserve((size_t)ndocs + 1) || !s.counts.reserve((size_t)ndocs + 1) ||
      !s.row_off.reserve((size_t)ndocs + 1) || !s.ids.reserve((size_t)ndocs * (size_t)max_ids) ||
      !s.counter.reserve(2) || !s.h_row_off.reserve((size_t)ndocs + 2))
    return false;
  size_t csr_cap = 0;
  int64_t max_len = 0;
  for (int64_t d = doc0; d < doc0 + ndocs; ++d) {
    const int64_t len = offsets[d + 1] - offsets[d];
    max_len = std::max(max_len, len);
    csr_cap += (size_t)std::min<int64_t>(max_ids_of_doc(m, len), max_ids);
  }
  if (!s.csr.reserve(csr_cap + 2)) return false;

  const char* text_src = B.utf8 + b0;
  const int64_t* offs_src = offsets + doc0;
  if (B.stage_in) {
    if (s.staged_doc0 != doc0 && !stage_chunk(m, s, B, doc0, ndocs)) return false;   // (normally prefetched by run_pipeline)
    copy_pool_of(m)->wait(&s.in_job);
    s.staged_doc0 = -1;
    text_src = (const char*)s.h_text.p;
    offs_src = s.h_offsets.p;
  }
  if (nbytes && !cuda_ok(cudaMemcpyAsync(s.text.p, text_src, nbytes, cudaMemcpyHostToDevice, s.stream), "H2D text")) return false;
  if (!cuda_ok(cudaMemcpyAsync(s.offsets.p, offs_src, ((size_t)ndocs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream), "H2D offsets")) return false;

  int nl = 0;
  if (!cuda_ok(cudaEventRecord(s.h2d_done, s.stream), "event record")) return false;
  if (!cuda_ok(cudaEventRecord(s.k_begin, s.stream), "event record")) return false;
  const uint8_t* text_biased = s.text.p - b0;
  if (m->engine == 3) {
    if (!launch_segmentation(m, s, text_biased, s.offsets.p, b1, ndocs, max_len, s.ids.p, s.counts.p, nullptr, nullptr, max_ids, B.unk,
                             s.stream, &nl))
      return false;
  } else if (m->engine == 2) {
    if (!launch_lexer_ids(m, s, text_biased, s.offsets.p, offsets[doc0], b1, ndocs, s.ids.p, s.counts.p, max_ids, B.unk, s.stream, &nl))
      return false;
  } else {
    if (!launch_wordpiece(m, text_biased, s.offsets.p, b1, ndocs, s.ids.p, s.counts.p, max_ids, B.unk, s.counter.p, s.stream, &nl))
      return false;
  }
  if (!cuda_ok(cudaEventRecord(s.k_end, s.stream), "event record")) return false;
  if (!cuda_ok(wp_scan_counts(s.counts.p, s.row_off.p, ndocs, s.stream), "scan")) return false;
  if (B.out == OutKind::kCsr16) {
    if (!cuda_ok(wp_compact_launch_u16(s.ids.p, s.counts.p, s.row_off.p, ndocs, max_ids, reinterpret_cast<uint16_t*>(s.csr.p), s.stream), "compact"))
      return false;
  } else if (!cuda_ok(wp_compact_launch(s.ids.p, s.counts.p, s.row_off.p, ndocs, max_ids, s.csr.p, s.stream), "compact")) {
    return false;
  }
  g_launches += nl + 2;
  if (!cuda_ok(cudaMemcpyAsync(s.h_row_off.p, s.row_off.p, ((size_t)ndocs + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream), "D2H offsets")) return false;
  // the kernel's error word sits right after the work counter (only the [pos-dict] engine sets it)
  s.h_row_off.p[ndocs + 1] = 0;
  if (m->engine == 3 &&
      !cuda_ok(cudaMemcpyAsync(s.h_row_off.p + ndocs + 1, s.counter.p + 1, sizeof(int64_t), cudaMemcpyDeviceToHost, s.stream), "D2H flag"))
    return false;
  if (!cuda_ok(cudaEventRecord(s.counts_ready, s.stream), "event record")) return false;
  s.doc0 = doc0; s.ndocs = ndocs;
  return true;
}

// a kernel that could not serve a document exactly raises a flag instead of guessing
bool chunk_ok(Slot& s) {
  const int64_t flag = s.h_row_off.p[s.ndocs + 1] & 0xffffffffll;
  if (flag == 0) return true;
  set_error(flag == 2 ? "a document exceeds the segmentation engine's arena" :
            flag == 3 ? "BPE arc scratch overflow (more than 32 vocabulary matches per symbol on average)" :
                        "kernel reported an error");
  return false;
}

// how many documents go into the next chunk: bounded text bytes and bounded id-matrix size
int64_t chunk_docs(const int64_t* offsets, int64_t doc0, int64_t ndocs_total, int max_ids, int engine) {
  // the generic lexer keeps 26 scratch bytes per input byte (classes + triples)
  // chunk size of the host pipeline; BLINGFIRE_B200_CHUNK_MB overrides it (tuning knob)
  static const int64_t env_mb = [] { const char* e = std::getenv("BLINGFIRE_B200_CHUNK_MB"); return e ? std::atoll(e) : 0ll; }();
  // (measured, tools/e2e_sweep.py / tools/e2e_trace.sh: the fused WordPiece kernel is fastest end to end with 16 MB chunks, the
  // segmentation kernels -- longer, less even documents per warp -- with 32 MB)
  const int64_t kMaxBytes = engine == 2 ? (8ll << 20) : ((env_mb > 0 ? env_mb : (engine == 3 ? 32 : 16)) << 20);
  const int64_t kMaxIdsCells = 160ll << 20;     // 640 MB of int32 per slot
  const int64_t max_docs = std::max<int64_t>(1, kMaxIdsCells / std::max(1, max_ids));
  int64_t d = doc0;
  const int64_t start = offsets[doc0];
  while (d < ndocs_total && d - doc0 < max_docs && (offsets[d + 1] - start <= kMaxBytes || d == doc0)) ++d;
  return d - doc0;
}

// Runs the whole host batch through a pipeline of kSlots chunks so that the H2D copy of chunk c, the
// kernels of chunk c-1 and the D2H copy of chunk c-2 overlap (PCIe is full duplex).
//   issue(slot)   called in chunk order once the chunk's row offsets are on the host
//                 (slot.h_row_off); enqueues the asynchronous D2H of the ids on slot.stream
//   finish(slot)  called in chunk order once that D2H has completed (host-side post-processing)
template <typename Issue, typename Finish>
bool run_pipeline(Model* m, Ctx* ctx, const HostBatch& B, Issue issue, Finish finish) {
  if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return false;
  g_last_kernel_ms = 0.0;
  bool copying[kMaxSlots] = {};
  auto settle = [&](int si) -> bool {   // wait for the ids copy of the chunk that last used slot si
    if (!copying[si]) return true;
    copying[si] = false;
    Slot& s = ctx->slots[si];
    if (!cuda_ok(cudaStreamSynchronize(s.stream), "sync")) return false;
    return finish(s);
  };
  auto counts_ready = [&](int si) -> bool {   // kernels of the chunk in slot si are done
    Slot& s = ctx->slots[si];
    if (!cuda_ok(cudaEventSynchronize(s.counts_ready), "event sync")) return false;
    float kms = 0.0f;
    if (cudaEventElapsedTime(&kms, s.k_begin, s.k_end) == cudaSuccess) g_last_kernel_ms += kms;
    if (!chunk_ok(s) || !issue(s)) return false;
    copying[si] = true;
    return true;
  };
  int64_t d = 0;
  int c = 0;
  static const bool trace = std::getenv("BLINGFIRE_B200_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  double t_settle = 0, t_enqueue = 0, t_counts = 0;
  bool ok = true;
  while (ok && d < B.ndocs) {
    const int si = c % kSlots;
    double t0 = now();
    if (!settle(si)) { ok = false; break; }
    double t1 = now();
    const int64_t nd = chunk_docs(B.offsets, d, B.ndocs, B.max_ids, m->engine);
    if (!enqueue_chunk(m, ctx->slots[si], B, d, nd)) { ok = false; break; }
    if (B.stage_in && d + nd < B.ndocs) {
      // pageable text: the copy threads stage the next chunk while this one is in flight
      const int sj = (c + 1) % kSlots;
      if (!settle(sj)) { ok = false; break; }
      if (ctx->slots[sj].h2d_done && !cuda_ok(cudaEventSynchronize(ctx->slots[sj].h2d_done), "event sync")) { ok = false; break; }
      const int64_t nd2 = chunk_docs(B.offsets, d + nd, B.ndocs, B.max_ids, m->engine);
      if (!stage_chunk(m, ctx->slots[sj], B, d + nd, nd2)) { ok = false; break; }
    }
    double t2 = now();
    // chunk c-kAhead: its kernels had kAhead chunks of queued work behind them, so the copy
    // engines and the SMs never wait for the host
    if (c >= kAhead && !counts_ready((c - kAhead) % kSlots)) { ok = false; break; }
    double t3 = now();
    t_settle += t1 - t0; t_enqueue += t2 - t1; t_counts += t3 - t2;
    d += nd; ++c;
  }
  if (trace)
    std::fprintf(stderr, "[bfb200] pipeline: %d chunks, loop %.2f ms (settle %.2f, enqueue %.2f, wait-counts %.2f)\n", c,
                 now() - t_begin, t_settle, t_enqueue, t_counts);
  if (ok)
    for (int k = c >= kAhead ? c - kAhead : 0; k < c && ok; ++k)     // the chunks whose counts were not consumed yet
      ok = counts_ready(k % kSlots);
  if (ok)
    for (int k = c >= kSlots ? c - kSlots : 0; k < c && ok; ++k)     // the chunks still copying, oldest first
      ok = settle(k % kSlots);
  // the asynchronous host copies of this call (they read / write the caller's buffers)
  if (m->copy_pool)
    for (int k = 0; k < kMaxSlots; ++k) {
      m->copy_pool->wait(&ctx->slots[k].in_job);
      m->copy_pool->wait(&ctx->slots[k].out_job);
      ctx->slots[k].staged_doc0 = -1;
    }
  if (!ok) {
    // the context goes back to the pool: nothing of this call may still be in flight on its streams
    const std::string keep = g_last_error;
    for (auto& s : ctx->slots) if (s.stream) cudaStreamSynchronize(s.stream);
    cudaGetLastError();
    g_last_error = keep;
  }
  return ok;
}

bool check_batch_args(Model* m, const char* utf8, const int64_t* offsets, int64_t ndocs, int max_ids) {
  if (!m) { set_error("null model"); return false; }
  if (m->engine == 0) { set_error("no GPU engine for this model type"); return false; }
  if (ndocs < 0 || max_ids < 0 || (ndocs > 0 && (!utf8 || !offsets))) { set_error("bad batch arguments"); return false; }
  return true;
}

// the compact-output batch call behind TextToIdsBatchCsr / TextToIdsBatchCsrU16
template <typename OutT>
int64_t batch_csr(Model* m, const char* utf8, const int64_t* offsets, int64_t ndocs, OutT* ids_csr, int64_t capacity,
                  int64_t* id_offsets, int max_ids, int unk) {
  if (!check_batch_args(m, utf8, offsets, ndocs, max_ids)) return -1;
  if (!id_offsets || (capacity > 0 && !ids_csr)) { set_error("bad output arguments"); return -1; }
  if (sizeof(OutT) == 2 && (m->max_tag > 0xFFFF || unk < 0 || unk > 0xFFFF)) {
    set_error("16-bit ids need a model whose ids (and UnkId) fit 16 bits");
    return -1;
  }
  if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return -1;
  CtxLease lease(m);
  HostBatch B{utf8, offsets, ndocs, max_ids, unk};
  B.stage_in = ndocs > 0 && is_pageable(utf8);
  B.out = sizeof(OutT) == 2 ? OutKind::kCsr16 : OutKind::kCsr32;
  const bool stage_out = capacity > 0 && is_pageable(ids_csr);
  int64_t total = 0;
  bool overflow = false;
  id_offsets[0] = 0;
  auto issue = [&](Slot& s) -> bool {
    const int64_t n = s.h_row_off.p[s.ndocs];
    for (int64_t i = 0; i < s.ndocs; ++i) id_offsets[s.doc0 + i + 1] = total + s.h_row_off.p[i + 1];
    s.out_base = total; s.out_n = 0;
    if (total + n > capacity) overflow = true;
    else if (n > 0) {
      void* dst = ids_csr + total;
      if (stage_out) {
        copy_pool_of(m)->wait(&s.out_job);          // the slot's previous ids have left the staging buffer
        if (!s.h_csr.reserve(((size_t)n * sizeof(OutT) + 3) / 4 + 1)) return false;
        dst = s.h_csr.p;
        s.out_n = n;
      }
      if (!cuda_ok(cudaMemcpyAsync(dst, s.csr.p, (size_t)n * sizeof(OutT), cudaMemcpyDeviceToHost, s.stream), "D2H ids")) return false;
    }
    total += n;
    return true;
  };
  auto finish = [&](Slot& s) -> bool {
    if (s.out_n > 0) copy_pool_of(m)->submit(ids_csr + s.out_base, s.h_csr.p, (size_t)s.out_n * sizeof(OutT), &s.out_job);
    return true;
  };
  if (!run_pipeline(m, lease.c, B, issue, finish)) return -1;
  return overflow ? -total : total;
}

// per-stream scratch of the device-pointer entry points
Slot* dev_slot_of(Model* m, cudaStream_t stream) {
  std::lock_guard<std::mutex> l(m->dev_mu);
  auto& p = m->dev_slots[stream];
  if (!p) p.reset(new Slot());
  return p.get();
}

}  // namespace

extern "C" {

int GetBlingFireTokVersion(void) { return 18 * 1000 + 0; }

const char* BlingFireB200LastError(void) { return g_last_error.c_str(); }
int64_t BlingFireB200KernelLaunches(void) { return g_launches.load(); }
double BlingFireB200LastKernelMs(void) { return g_last_kernel_ms; }
int BlingFireB200ModelEngine(void* h) { return h ? ((Model*)h)->engine : 0; }

void* LoadModel(const char* path) {
  try {
    g_last_error.clear();
    LdbImage ldb;
    if (!ldb.load_file(path)) { set_error(ldb.error()); return nullptr; }
    return finish_model(std::unique_ptr<Model>(new Model()), ldb);
  } catch (const std::exception& e) { set_error(e.what()); return nullptr; }
}

void* SetModel(const unsigned char* bytes, int n) {
  try {
    g_last_error.clear();
    if (!bytes || n <= 0) { set_error("empty image"); return nullptr; }
    LdbImage ldb;
    if (!ldb.set_image(bytes, (size_t)n)) { set_error(ldb.error()); return nullptr; }
    return finish_model(std::unique_ptr<Model>(new Model()), ldb);
  } catch (const std::exception& e) { set_error(e.what()); return nullptr; }
}

int FreeModel(void* h) {
  if (!h) return 0;
  delete (Model*)h;
  return 1;
}

// blingfiretokdll.cpp:1669-1679
int SetNoDummyPrefix(void* h, bool fNoDummyPrefix) {
  if (!h) return 0;
  ((Model*)h)->no_dummy_prefix = fNoDummyPrefix;
  return 1;
}

// blingfiretokdll.cpp:1689-1745.  A table read on the host, like the reference's: there is no arithmetic
// to move to the GPU (the id -> text array stays in host memory next to the caller's buffers).
int IdsToText(void* h, const int32_t* ids, const int count, char* out, const int max_out, bool skip_special) {
  if (!h) return 0;
  if (count == 0 || !ids) return 0;
  const Model* m = (const Model*)h;
  if (!m->i2w.present) return 0;
  int len = 0;
  for (int i = 0; i < count; ++i) {
    const int id = ids[i];
    if (skip_special && (id < m->i2w.min_id || id > m->i2w.max_id)) continue;       // :1712
    if (id < 0 || id >= m->i2w.count) return 0;                                       // unknown id (:1719-1721)
    const uint8_t* tok = m->i2w.data + m->i2w.offsets[id];
    int tl = (int)(m->i2w.offsets[id + 1] - m->i2w.offsets[id]);
    if (len == 0 && tl > 0 && tok[0] == 0x20) { ++tok; --tl; }                       // no leading space (:1724-1727)
    if (tl > 0 && max_out - len >= tl) std::memcpy(out + len, tok, (size_t)tl);
    len += tl;
  }
  if (max_out > len) out[len] = 0;
  return len + 1;
}

int TextToIdsBatchDeviceSized(void* h, const char* d_utf8, const int64_t* d_offsets, int64_t ndocs, int64_t total_bytes,
                              int64_t max_doc_bytes, int32_t* d_ids, int32_t* d_counts, int max_ids, int unk, void* stream_) {
  try {
    g_last_error.clear();
    Model* m = (Model*)h;
    if (!m || m->engine == 0) { set_error(m ? "no GPU engine for this model type" : "null model"); return -1; }
    if (ndocs == 0) return 0;
    if (ndocs < 0 || !d_utf8 || !d_offsets || !d_ids || !d_counts || max_ids < 0) { set_error("bad arguments"); return -1; }
    if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return -1;
    cudaStream_t stream = (cudaStream_t)stream_;
    int nl = 0;
    if (m->engine == 1) {
      // every call takes the next word of a ring: calls on different streams never share a work counter
      constexpr uint32_t kRing = 1024;
      {
        std::lock_guard<std::mutex> l(m->dev_mu);
        if (!m->dev_counters.reserve(kRing)) return -1;
      }
      unsigned long long* counter = m->dev_counters.p + (m->dev_counter_next.fetch_add(1) % kRing);
      if (!launch_wordpiece(m, (const uint8_t*)d_utf8, d_offsets, total_bytes, ndocs, d_ids, d_counts, max_ids, unk, counter, stream, &nl))
        return -1;
    } else {
      Slot& s = *dev_slot_of(m, stream);      // stream-ordered reuse of the scratch
      if (m->engine == 3) {
        if (!s.counter.reserve(4)) return -1;
        if (max_doc_bytes <= 0) {
          // size the arena from the batch itself: one small kernel and a 8-byte read back (synchronises the stream)
          unsigned long long* dmax = s.counter.p + 2;
          if (!cuda_ok(cudaMemsetAsync(dmax, 0, 8, stream), "memset")) return -1;
          max_doc_len_kernel<<<(int)std::min<int64_t>((ndocs + 255) / 256, 1184), 256, 0, stream>>>(d_offsets, ndocs, dmax);
          unsigned long long hmax = 0;
          if (!cuda_ok(cudaMemcpyAsync(&hmax, dmax, 8, cudaMemcpyDeviceToHost, stream), "D2H") ||
              !cuda_ok(cudaStreamSynchronize(stream), "sync"))
            return -1;
          max_doc_bytes = (int64_t)hmax;
          ++nl;
        }
        if (!launch_segmentation(m, s, (const uint8_t*)d_utf8, d_offsets, total_bytes, ndocs, max_doc_bytes, d_ids, d_counts, nullptr,
                                 nullptr, max_ids, unk, stream, &nl))
          return -1;
      } else {
        if (!launch_lexer_ids(m, s, (const uint8_t*)d_utf8, d_offsets, 0, total_bytes, ndocs, d_ids, d_counts, max_ids, unk, stream, &nl))
          return -1;
      }
    }
    g_launches += nl;
    return 0;
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int TextToIdsBatchDevice(void* h, const char* d_utf8, const int64_t* d_offsets, int64_t ndocs, int64_t total_bytes,
                         int32_t* d_ids, int32_t* d_counts, int max_ids, int unk, void* stream) {
  return TextToIdsBatchDeviceSized(h, d_utf8, d_offsets, ndocs, total_bytes, 0, d_ids, d_counts, max_ids, unk, stream);
}

// the [pos-dict] kernels raise a flag instead of guessing when their scratch cannot hold a document
int BlingFireB200DeviceStatus(void* h, void* stream_) {
  try {
    g_last_error.clear();
    Model* m = (Model*)h;
    if (!m) { set_error("null model"); return -1; }
    if (m->engine != 3) return 0;
    if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return -1;
    cudaStream_t stream = (cudaStream_t)stream_;
    Slot& s = *dev_slot_of(m, stream);
    if (!s.counter.p) return 0;
    unsigned long long flag = 0;
    if (!cuda_ok(cudaMemcpyAsync(&flag, s.counter.p + 1, 8, cudaMemcpyDeviceToHost, stream), "D2H") ||
        !cuda_ok(cudaStreamSynchronize(stream), "sync"))
      return -1;
    const int f = (int)(flag & 0xffffffffu);
    if (f) set_error("segmentation engine: scratch exhausted (code " + std::to_string(f) + ")");
    return f;
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int64_t BlingFireB200CompactDevice(const int32_t* d_ids, const int32_t* d_counts, int64_t ndocs, int max_ids, int32_t* d_csr,
                                   int64_t* d_row_off, void* stream_) {
  g_last_error.clear();
  if (ndocs < 0 || !d_counts || !d_row_off || (ndocs > 0 && (!d_ids || !d_csr))) { set_error("bad arguments"); return -1; }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!cuda_ok(wp_scan_counts(d_counts, d_row_off, ndocs, stream), "scan")) return -1;
  if (!cuda_ok(wp_compact_launch(d_ids, d_counts, d_row_off, ndocs, max_ids, d_csr, stream), "compact")) return -1;
  g_launches += 2;
  return 0;
}

int64_t TextToIdsBatchCsr(void* h, const char* utf8, const int64_t* offsets, int64_t ndocs, int32_t* ids_csr,
                          int64_t capacity, int64_t* id_offsets, int max_ids, int unk) {
  try {
    g_last_error.clear();
    return batch_csr<int32_t>((Model*)h, utf8, offsets, ndocs, ids_csr, capacity, id_offsets, max_ids, unk);
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int64_t TextToIdsBatchCsrU16(void* h, const char* utf8, const int64_t* offsets, int64_t ndocs, uint16_t* ids_csr,
                             int64_t capacity, int64_t* id_offsets, int max_ids, int unk) {
  try {
    g_last_error.clear();
    return batch_csr<uint16_t>((Model*)h, utf8, offsets, ndocs, ids_csr, capacity, id_offsets, max_ids, unk);
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int64_t TextToIdsBatch(void* h, const char* utf8, const int64_t* offsets, int64_t ndocs, int32_t* ids, int32_t* counts,
                       int max_ids, int unk) {
  try {
    g_last_error.clear();
    Model* m = (Model*)h;
    if (!check_batch_args(m, utf8, offsets, ndocs, max_ids)) return -1;
    if (ndocs > 0 && (!ids || !counts)) { set_error("bad output arguments"); return -1; }
    if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return -1;
    CtxLease lease(m);
    HostBatch B{utf8, offsets, ndocs, max_ids, unk};
    B.stage_in = ndocs > 16 && is_pageable(utf8);
    int64_t total = 0;
    auto issue = [&](Slot& s) -> bool {
      const int64_t n = s.h_row_off.p[s.ndocs];
      if (!s.h_csr.reserve((size_t)n + 1)) return false;
      if (n > 0 && !cuda_ok(cudaMemcpyAsync(s.h_csr.p, s.csr.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream), "D2H ids"))
        return false;
      total += n;
      return true;
    };
    auto finish = [&](Slot& s) -> bool {
      // rows beyond their count stay untouched, as in the reference (blingfiretokdll.cpp:1098-1101)
      auto rows = [&](int64_t a0, int64_t a1) {
        for (int64_t i = a0; i < a1; ++i) {
          const int64_t a = s.h_row_off.p[i], b = s.h_row_off.p[i + 1];
          counts[s.doc0 + i] = (int32_t)(b - a);
          if (b > a) std::memcpy(ids + (s.doc0 + i) * (int64_t)max_ids, s.h_csr.p + a, (size_t)(b - a) * sizeof(int32_t));
        }
      };
      constexpr int64_t kPart = 4096;
      if (s.ndocs <= 2 * kPart) rows(0, s.ndocs);
      else copy_pool_of(m)->parallel_for((size_t)((s.ndocs + kPart - 1) / kPart),
                                         [&](size_t k) { rows((int64_t)k * kPart, std::min<int64_t>(s.ndocs, ((int64_t)k + 1) * kPart)); });
      return true;
    };
    if (!run_pipeline(m, lease.c, B, issue, finish)) return -1;
    return total;
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int TextToIds(void* h, const char* s, int n, int32_t* ids, const int max_ids, const int unk) {
  // blingfiretokdll.cpp:1619-1646 -> :1121: parameter validation happens before any work
  g_last_error.clear();
  if (!h || n <= 0 || n > 1000000000 || !s) return 0;
  Model* m = (Model*)h;
  if (m->engine == 0) { set_error("no GPU engine for this model type"); return 0; }
  if (max_ids <= 0 || !ids) return 0;
  if (n <= (32 << 10) && max_ids <= 8192) {
    // One short document: ONE copy in (offsets and text together, from the context's pinned buffer), the engine's kernels,
    // ONE copy out (the id row and its count together), one synchronisation.  The batch pipeline's chunking, scan,
    // compaction and two round trips are for batches; a call still costs a launch and two small copies.
    try {
      if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return 0;
      CtxLease lease(m);
      Slot& sl = lease.c->slots[0];
      if (!ensure_stream(sl)) return 0;
      const size_t tb = ((size_t)n + 3) & ~(size_t)3;
      const size_t in_words = 4 + (tb + 64) / 4, out_words = (size_t)max_ids + 4;
      if (!lease.c->h_words.reserve(in_words + out_words) || !sl.text.reserve(16 + tb + 64) || !sl.ids.reserve(out_words) || !sl.counter.reserve(2))
        return 0;
      int32_t* hw = lease.c->h_words.p;
      int64_t* ho = reinterpret_cast<int64_t*>(hw);
      ho[0] = 0; ho[1] = n;
      std::memcpy(hw + 4, s, (size_t)n);
      if (!cuda_ok(cudaMemcpyAsync(sl.text.p, hw, 16 + (size_t)n, cudaMemcpyHostToDevice, sl.stream), "H2D")) return 0;
      const uint8_t* d_text = sl.text.p + 16;
      const int64_t* d_offs = reinterpret_cast<const int64_t*>(sl.text.p);
      int32_t* d_ids = sl.ids.p;
      int32_t* d_counts = d_ids + max_ids;          // the count travels back with the row
      int nl = 0;
      bool ok;
      if (m->engine == 3) ok = launch_segmentation(m, sl, d_text, d_offs, n, 1, n, d_ids, d_counts, nullptr, nullptr, max_ids, unk, sl.stream, &nl);
      else if (m->engine == 2) ok = launch_lexer_ids(m, sl, d_text, d_offs, 0, n, 1, d_ids, d_counts, max_ids, unk, sl.stream, &nl);
      else ok = launch_wordpiece(m, d_text, d_offs, n, 1, d_ids, d_counts, max_ids, unk, sl.counter.p, sl.stream, &nl);
      if (!ok) return 0;
      g_launches += nl;
      int32_t* hr = hw + in_words;
      if (!cuda_ok(cudaMemcpyAsync(hr, d_ids, ((size_t)max_ids + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H")) return 0;
      hr[max_ids + 1] = 0;
      if (m->engine == 3 &&
          !cuda_ok(cudaMemcpyAsync(hr + max_ids + 1, reinterpret_cast<const int32_t*>(sl.counter.p + 1), 4, cudaMemcpyDeviceToHost, sl.stream), "D2H flag"))
        return 0;
      if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return 0;
      if (hr[max_ids + 1] != 0) { set_error("segmentation engine: scratch exhausted (code " + std::to_string(hr[max_ids + 1]) + ")"); return 0; }
      const int count = hr[max_ids];
      if (count <= 0 || count > max_ids) return 0;
      std::memcpy(ids, hr, (size_t)count * sizeof(int32_t));     // the rest of the caller's array stays untouched (:1098-1101)
      return count;
    } catch (const std::exception& e) { set_error(e.what()); return 0; }
  }
  const int64_t offsets[2] = {0, n};
  int32_t count = 0;
  const int64_t r = TextToIdsBatch(h, s, offsets, 1, ids, &count, max_ids, unk);
  return r < 0 ? 0 : (int)count;
}

// blingfiretokdll.cpp:1563-1609 -> TextToIdsWithOffsets_wp (:1108-1314) with offsets.  Served by the
// generic lexer engine (decode with byte offsets -> Process_int triples -> the exact post-pass);
// the fused kernel does not carry offsets.  [pos-dict] models: not served yet (returns 0).
int TextToIdsWithOffsets(void* h, const char* s, int n, int32_t* ids, int* starts, int* ends, const int max_ids, const int unk) {
  try {
    g_last_error.clear();
    if (!h || n <= 0 || n > 1000000000 || !s) return 0;              // :1121
    if (!starts || !ends) return TextToIds(h, s, n, ids, max_ids, unk);
    Model* m = (Model*)h;
    if (m->engine == 0) { set_error("no GPU engine for this model type"); return 0; }
    if (m->engine != 3 && (!m->has_wbd || !m->lex_ok || !m->d_cls)) { set_error("model has no lexer engine with a 1->1 charmap"); return 0; }
    if (max_ids <= 0 || !ids) return 0;
    if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return 0;
    CtxLease lease(m);
    Slot& sl = lease.c->slots[0];
    PinBuf<int32_t>& h_words = lease.c->h_words;
    if (!ensure_stream(sl)) return 0;
    const size_t nb = (size_t)n;
    if (m->engine == 3) {
      // TextToIdsWithOffsets_sp (:1349-1535): the segmentation kernel with the byte offsets of every
      // symbol carried through normalisation and whitespace collapsing
      if (!sl.text.reserve(nb + 64) || !sl.offsets.reserve(2) || !sl.ids.reserve(3 * (size_t)max_ids) || !sl.counts.reserve(2) ||
          !h_words.reserve(3 * (size_t)max_ids + 16))
        return 0;
      const int64_t offs[2] = {0, n};
      if (!cuda_ok(cudaMemcpyAsync(sl.text.p, s, nb, cudaMemcpyHostToDevice, sl.stream), "H2D text")) return 0;
      if (!cuda_ok(cudaMemcpyAsync(sl.offsets.p, offs, sizeof(offs), cudaMemcpyHostToDevice, sl.stream), "H2D offsets")) return 0;
      int32_t* d_ids = sl.ids.p;
      int nl = 0;
      if (!launch_segmentation(m, sl, sl.text.p, sl.offsets.p, n, 1, n, d_ids, sl.counts.p, d_ids + max_ids, d_ids + 2 * (size_t)max_ids,
                               max_ids, unk, sl.stream, &nl))
        return 0;
      g_launches += nl;
      int32_t* hw = h_words.p;
      if (!cuda_ok(cudaMemcpyAsync(hw, sl.counts.p, 4, cudaMemcpyDeviceToHost, sl.stream), "D2H")) return 0;
      if (!cuda_ok(cudaMemcpyAsync(hw + 2, sl.counter.p + 1, 4, cudaMemcpyDeviceToHost, sl.stream), "D2H flag")) return 0;
      if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return 0;
      if (hw[2] != 0) { set_error("segmentation engine: scratch exhausted (code " + std::to_string(hw[2]) + ")"); return 0; }
      const int count = hw[0];
      if (count <= 0 || count > max_ids) return 0;
      for (int k = 0; k < 3; ++k)
        if (!cuda_ok(cudaMemcpyAsync(hw + 4 + (size_t)k * max_ids, d_ids + (size_t)k * max_ids, (size_t)count * 4, cudaMemcpyDeviceToHost, sl.stream), "D2H rows"))
          return 0;
      if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return 0;
      std::memcpy(ids, hw + 4, (size_t)count * 4);                   // the rest of the arrays stays untouched
      std::memcpy(starts, hw + 4 + (size_t)max_ids, (size_t)count * 4);
      std::memcpy(ends, hw + 4 + 2 * (size_t)max_ids, (size_t)count * 4);
      return count;
    }
    if (!sl.text.reserve(nb + 64) || !sl.offsets.reserve(2) || !sl.lex_cls.reserve(nb + 8) || !sl.lex_ncps.reserve(1) ||
        !sl.lex_tri_count.reserve(1) || !sl.lex_tri.reserve(6 * nb + 8) || !sl.lex_boff.reserve(nb + 8) ||
        !sl.ids.reserve(3 * (size_t)max_ids) || !sl.counts.reserve(2) || !h_words.reserve(3 * (size_t)max_ids + 16))
      return 0;
    const int64_t offs[2] = {0, n};
    if (!cuda_ok(cudaMemcpyAsync(sl.text.p, s, nb, cudaMemcpyHostToDevice, sl.stream), "H2D text")) return 0;
    if (!cuda_ok(cudaMemcpyAsync(sl.offsets.p, offs, sizeof(offs), cudaMemcpyHostToDevice, sl.stream), "H2D offsets")) return 0;
    LexLaunch X = make_lex_launch(sl, sl.text.p, sl.offsets.p, 0, n, 1, m->d_cls, 2);
    X.boff_buf = sl.lex_boff.p;
    int nl = 0;
    if (!cuda_ok(lex_launch(X, make_lex_model(m), sl.stream, &nl), "lexer launch")) return 0;
    int32_t* d_ids = sl.ids.p;
    if (!cuda_ok(lex_wp_offsets_launch(X, d_ids, d_ids + max_ids, d_ids + 2 * (size_t)max_ids, sl.counts.p, max_ids, unk, sl.stream, &nl),
                 "post-pass launch"))
      return 0;
    g_launches += nl;
    int32_t* hw = h_words.p;
    if (!cuda_ok(cudaMemcpyAsync(hw, sl.counts.p, 4, cudaMemcpyDeviceToHost, sl.stream), "D2H")) return 0;
    if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return 0;
    const int count = hw[0];
    if (count <= 0 || count > max_ids) return 0;
    for (int k = 0; k < 3; ++k)
      if (!cuda_ok(cudaMemcpyAsync(hw + 4 + (size_t)k * max_ids, d_ids + (size_t)k * max_ids, (size_t)count * 4, cudaMemcpyDeviceToHost, sl.stream), "D2H rows"))
        return 0;
    if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return 0;
    std::memcpy(ids, hw + 4, (size_t)count * 4);                     // the rest of the arrays stays untouched
    std::memcpy(starts, hw + 4 + (size_t)max_ids, (size_t)count * 4);
    std::memcpy(ends, hw + 4 + 2 * (size_t)max_ids, (size_t)count * 4);
    return count;
  } catch (const std::exception& e) { set_error(e.what()); return 0; }
}

// Additive: TextToIdsWithOffsets for a batch.  The same kernels as the per-document call -- they take batches already --
// over chunks of a few MB; the three row-major device arrays of a chunk are compacted on the device (scan of the counts,
// three gathers), so that 12 bytes per id come back instead of 12 * max_ids per document.
//   row-major form (id_offsets == nullptr): ids / starts / ends are [ndocs][max_ids], counts [ndocs]; a row beyond its
//     count stays untouched, like the arrays of the per-document call.
//   CSR form: ids / starts / ends hold `capacity` entries each, id_offsets [ndocs + 1]; returns the total, negated if it
//     exceeds the capacity (id_offsets is complete then, the arrays hold the chunks that still fitted).
static int64_t offsets_batch(Model* m, const char* utf8, const int64_t* offsets, int64_t ndocs, int32_t* ids, int32_t* starts,
                             int32_t* ends, int32_t* counts, int64_t capacity, int64_t* id_offsets, int max_ids, int unk) {
  const bool csr = id_offsets != nullptr;
  if (!check_batch_args(m, utf8, offsets, ndocs, max_ids)) return -1;
  if (csr) {
    if (capacity < 0 || (capacity > 0 && (!ids || !starts || !ends))) { set_error("bad output arguments"); return -1; }
    id_offsets[0] = 0;
  } else if (ndocs > 0 && (!ids || !starts || !ends || !counts)) {
    set_error("bad output arguments");
    return -1;
  }
  if (max_ids <= 0) {
    for (int64_t i = 0; i < ndocs; ++i) {
      if (csr) id_offsets[i + 1] = 0; else counts[i] = 0;
    }
    return 0;
  }
  const bool seg = m->has_seg;
  if (!seg && !(m->has_wbd && m->lex_ok && m->T.charmap_one_to_one)) { set_error("offsets are not served for this lexer model"); return -1; }
  if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return -1;
  CtxLease lease(m);
  // Up to kDepth chunks are in flight, each on the stream of its own slot: the generic lexer engine runs one thread per document,
  // so a launch lasts as long as its longest document takes -- the launches of several chunks overlap instead of queueing.
  const int kDepth = std::min(4, kSlots);
  struct InFlight { int64_t d0 = 0, nd = 0; size_t cells = 0; bool busy = false; };
  InFlight fl[kMaxSlots];
  int64_t total = 0;
  const int64_t kChunkBytes = 8ll << 20, kMaxCells = 16ll << 20;

  // the kernels of chunk [d0, d1) and the copy of its row offsets, all on the slot's stream
  auto issue = [&](Slot& sl, InFlight& f, int64_t d0, int64_t d1, int64_t max_len) -> bool {
    if (!ensure_stream(sl)) return false;
    const int64_t nd = d1 - d0;
    const int64_t b0 = offsets[d0] & ~(int64_t)3, b1 = offsets[d1];
    const size_t nb = (size_t)(b1 - b0), span = (size_t)(b1 - offsets[d0]), cells = (size_t)nd * (size_t)max_ids;
    if (!sl.text.reserve(nb + 64) || !sl.offsets.reserve((size_t)nd + 1) || !sl.ids.reserve(3 * cells) || !sl.counts.reserve((size_t)nd + 1) ||
        !sl.counter.reserve(2) || !sl.row_off.reserve((size_t)nd + 1) || !sl.csr.reserve(3 * cells) || !sl.h_row_off.reserve((size_t)nd + 2))
      return false;
    if (nb && !cuda_ok(cudaMemcpyAsync(sl.text.p, utf8 + b0, nb, cudaMemcpyHostToDevice, sl.stream), "H2D text")) return false;
    if (!cuda_ok(cudaMemcpyAsync(sl.offsets.p, offsets + d0, ((size_t)nd + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, sl.stream), "H2D offsets")) return false;
    int32_t* d_ids = sl.ids.p;
    int nl = 0;
    if (seg) {
      if (!launch_segmentation(m, sl, sl.text.p - b0, sl.offsets.p, b1, nd, max_len, d_ids, sl.counts.p, d_ids + cells, d_ids + 2 * cells, max_ids, unk,
                               sl.stream, &nl))
        return false;
    } else {
      if (!sl.lex_cls.reserve(span + 8) || !sl.lex_ncps.reserve((size_t)nd) || !sl.lex_tri_count.reserve((size_t)nd) ||
          !sl.lex_tri.reserve(6 * span + 8) || !sl.lex_boff.reserve(span + 8))
        return false;
      LexLaunch X = make_lex_launch(sl, sl.text.p - b0, sl.offsets.p, offsets[d0], b1, nd, m->d_cls, 2);
      X.boff_buf = sl.lex_boff.p;
      if (!cuda_ok(lex_launch(X, make_lex_model(m), sl.stream, &nl), "lexer launch")) return false;
      if (!cuda_ok(lex_wp_offsets_launch(X, d_ids, d_ids + cells, d_ids + 2 * cells, sl.counts.p, max_ids, unk, sl.stream, &nl), "post-pass launch"))
        return false;
    }
    if (!cuda_ok(wp_scan_counts(sl.counts.p, sl.row_off.p, nd, sl.stream), "scan")) return false;
    for (int k = 0; k < 3; ++k)
      if (!cuda_ok(wp_compact_launch(d_ids + (size_t)k * cells, sl.counts.p, sl.row_off.p, nd, max_ids, sl.csr.p + (size_t)k * cells, sl.stream), "compact"))
        return false;
    g_launches += nl + 4;
    int64_t* hro = sl.h_row_off.p;                                   // [nd + 1] row offsets of the chunk, then the error word
    if (!cuda_ok(cudaMemcpyAsync(hro, sl.row_off.p, ((size_t)nd + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, sl.stream), "D2H offsets")) return false;
    hro[nd + 1] = 0;
    if (seg && !cuda_ok(cudaMemcpyAsync(hro + nd + 1, sl.counter.p + 1, 4, cudaMemcpyDeviceToHost, sl.stream), "D2H flag")) return false;
    f.d0 = d0; f.nd = nd; f.cells = cells; f.busy = true;
    return true;
  };

  // waits for the chunk, copies its compact ids / starts / ends out; chunks finish in document order
  auto finish = [&](Slot& sl, InFlight& f) -> bool {
    f.busy = false;
    const int64_t d0 = f.d0, nd = f.nd;
    const size_t cells = f.cells;
    if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return false;
    const int64_t* hro = sl.h_row_off.p;
    if (hro[nd + 1] != 0) { set_error("segmentation engine: scratch exhausted (code " + std::to_string((long long)hro[nd + 1]) + ")"); return false; }
    const int64_t nout = hro[nd];
    if (nout < 0 || (size_t)nout > cells) { set_error("offsets batch: inconsistent counts"); return false; }
    if (csr) {
      for (int64_t i = 0; i < nd; ++i) id_offsets[d0 + i + 1] = total + hro[i + 1];
      if (nout > 0 && total + nout <= capacity) {
        int32_t* dst[3] = {ids + total, starts + total, ends + total};
        for (int k = 0; k < 3; ++k)
          if (!cuda_ok(cudaMemcpyAsync(dst[k], sl.csr.p + (size_t)k * cells, (size_t)nout * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H ids"))
            return false;
        if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return false;
      }
    } else {
      if (!sl.h_csr.reserve(3 * (size_t)nout + 4)) return false;
      int32_t* hw = sl.h_csr.p;                                      // the chunk's compact ids, starts, ends
      if (nout > 0) {
        for (int k = 0; k < 3; ++k)
          if (!cuda_ok(cudaMemcpyAsync(hw + (size_t)k * nout, sl.csr.p + (size_t)k * cells, (size_t)nout * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H ids"))
            return false;
        if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return false;
      }
      for (int64_t i = 0; i < nd; ++i) {
        const int64_t r0 = hro[i], c = hro[i + 1] - r0;
        counts[d0 + i] = (int32_t)c;
        if (c > 0) {
          const size_t dst = (size_t)(d0 + i) * (size_t)max_ids;
          std::memcpy(ids + dst, hw + r0, (size_t)c * 4);
          std::memcpy(starts + dst, hw + nout + r0, (size_t)c * 4);
          std::memcpy(ends + dst, hw + 2 * nout + r0, (size_t)c * 4);
        }
      }
    }
    total += nout;
    return true;
  };

  bool ok = true;
  int64_t chunk = 0;
  for (int64_t d0 = 0; d0 < ndocs && ok; ++chunk) {
    int64_t d1 = d0, max_len = 0;
    while (d1 < ndocs && (d1 == d0 || (offsets[d1 + 1] - offsets[d0] <= kChunkBytes && (d1 - d0 + 1) * (int64_t)max_ids <= kMaxCells))) {
      max_len = std::max(max_len, offsets[d1 + 1] - offsets[d1]);
      ++d1;
    }
    const int si = (int)(chunk % kDepth);
    if (fl[si].busy) ok = finish(lease.c->slots[si], fl[si]);        // the oldest chunk in flight: the slot is its
    if (ok) ok = issue(lease.c->slots[si], fl[si], d0, d1, max_len);
    d0 = d1;
  }
  // the chunks still in flight, oldest first; after an error they are only waited for (their buffers are about to be reused)
  for (int64_t k = chunk >= kDepth ? chunk - kDepth : 0; k < chunk; ++k) {
    const int si = (int)(k % kDepth);
    if (!fl[si].busy) continue;
    if (ok) ok = finish(lease.c->slots[si], fl[si]);
    else { fl[si].busy = false; cudaStreamSynchronize(lease.c->slots[si].stream); }
  }
  if (!ok) return -1;
  return csr && total > capacity ? -total : total;
}

int64_t TextToIdsWithOffsetsBatch(void* h, const char* utf8, const int64_t* offsets, int64_t ndocs, int32_t* ids, int32_t* starts,
                                  int32_t* ends, int32_t* counts, int max_ids, int unk) {
  try {
    g_last_error.clear();
    return offsets_batch((Model*)h, utf8, offsets, ndocs, ids, starts, ends, counts, 0, nullptr, max_ids, unk);
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int64_t TextToIdsWithOffsetsBatchCsr(void* h, const char* utf8, const int64_t* offsets, int64_t ndocs, int32_t* ids_csr, int32_t* starts_csr,
                                     int32_t* ends_csr, int64_t capacity, int64_t* id_offsets, int max_ids, int unk) {
  try {
    g_last_error.clear();
    if (!id_offsets) { set_error("bad output arguments"); return -1; }
    return offsets_batch((Model*)h, utf8, offsets, ndocs, ids_csr, starts_csr, ends_csr, nullptr, capacity, id_offsets, max_ids, unk);
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}

int TextToIdsWithOffsets_wp(void* h, const char* s, int n, int32_t* ids, int* starts, int* ends, const int max_ids, const int unk) {
  Model* m = (Model*)h;
  if (!m || !m->has_wbd || m->has_seg) return 0;
  return TextToIdsWithOffsets(h, s, n, ids, starts, ends, max_ids, unk);
}

int TextToIdsWithOffsets_sp(void* h, const char* s, int n, int32_t* ids, int* starts, int* ends, const int max_ids, const int unk) {
  Model* m = (Model*)h;
  if (!m || !m->has_seg) return 0;
  return TextToIdsWithOffsets(h, s, n, ids, starts, ends, max_ids, unk);
}

int TextToIds_wp(void* h, const char* s, int n, int32_t* ids, const int max_ids, const int unk) {
  Model* m = (Model*)h;
  if (!m || !m->has_wbd || m->has_seg) return 0;
  return TextToIds(h, s, n, ids, max_ids, unk);
}

int TextToIds_sp(void* h, const char* s, int n, int32_t* ids, const int max_ids, const int unk) {
  Model* m = (Model*)h;
  if (!m || !m->has_seg) return 0;
  return TextToIds(h, s, n, ids, max_ids, unk);
}

// The reference embeds wbd.bin and sbd.bin as byte arrays and initialises them once under a mutex
// (blingfiretokdll.cpp:114-133, :426-434).  Here the same models are loaded from $BLINGFIRE_B200_WBD /
// $BLINGFIRE_B200_SBD, or from wbd.bin / sbd.bin next to this shared library, on first use.
static Model* default_model(int which) {        // 0 = word breaker, 1 = sentence breaker
  static std::mutex mu;
  static Model* model[2] = {nullptr, nullptr};
  static bool tried[2] = {false, false};
  std::lock_guard<std::mutex> lock(mu);
  if (tried[which]) return model[which];
  tried[which] = true;
  std::string path;
  if (const char* env = std::getenv(which ? "BLINGFIRE_B200_SBD" : "BLINGFIRE_B200_WBD")) path = env;
  if (path.empty()) {
    Dl_info info;
    if (dladdr((const void*)&default_model, &info) && info.dli_fname) {
      path = info.dli_fname;
      const size_t slash = path.find_last_of('/');
      path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + (which ? "/sbd.bin" : "/wbd.bin");
    }
  }
  model[which] = (Model*)LoadModel(path.c_str());
  if (!model[which]) set_error(std::string("default ") + (which ? "sentence" : "word") + "-breaking model not found (" + path + "): " + g_last_error);
  return model[which];
}

namespace {

// What the host needs to put a lexer's output back into text: the (Tag, From, To) triples and the byte
// offset of every code point.  The lexer itself (decode, classes, Process_int with every nested call)
// runs on the GPU, on the caller's leased context.  Returns false on any failure (invalid UTF-8 included).
struct LexedText { int ncps = 0, rn = 0; const int32_t* tri = nullptr; std::vector<int> cp_off; };

bool lex_on_gpu(Model* m, Ctx* ctx, const char* s, int n, LexedText* R) {
  if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return false;
  Slot& sl = ctx->slots[0];
  if (!ensure_stream(sl)) return false;
  const size_t nb = (size_t)n, tb = (nb + 3) & ~(size_t)3;
  // pinned staging: [offsets 16 B][text] in, [ncps, tri_count, triples...] out
  const size_t in_words = 4 + (tb + 64) / 4;
  if (!sl.text.reserve(16 + tb + 64) || !sl.lex_cls.reserve(nb + 8) || !sl.lex_ncps.reserve(1) ||
      !sl.lex_tri_count.reserve(1) || !sl.lex_tri.reserve(3 * nb + 8) || !ctx->h_words.reserve(in_words + 3 * nb + 16))
    return false;
  int32_t* hin = ctx->h_words.p;
  int64_t* ho = reinterpret_cast<int64_t*>(hin);
  ho[0] = 0; ho[1] = n;
  std::memcpy(hin + 4, s, nb);
  if (!cuda_ok(cudaMemcpyAsync(sl.text.p, hin, 16 + nb, cudaMemcpyHostToDevice, sl.stream), "H2D")) return false;   // offsets + text, one copy
  LexLaunch X = make_lex_launch(sl, sl.text.p + 16, reinterpret_cast<const int64_t*>(sl.text.p), 0, n, 1, m->d_cls_words, 1);   // MaxOut = 3 * MaxBuffSize (:492-499, :249-251)
  int nl = 0;
  if (!cuda_ok(lex_launch(X, make_lex_model(m), sl.stream, &nl), "lexer launch")) return false;
  g_launches += nl;
  int32_t* hw = hin + in_words;
  if (!cuda_ok(cudaMemcpyAsync(hw, sl.lex_ncps.p, 4, cudaMemcpyDeviceToHost, sl.stream), "D2H")) return false;
  if (!cuda_ok(cudaMemcpyAsync(hw + 1, sl.lex_tri_count.p, 4, cudaMemcpyDeviceToHost, sl.stream), "D2H")) return false;
  // short inputs: the triples travel with the counts (at most 3 per code point), one synchronisation
  const bool eager = nb <= (8u << 10);
  if (eager && !cuda_ok(cudaMemcpyAsync(hw + 2, sl.lex_tri.p, 3 * nb * 4, cudaMemcpyDeviceToHost, sl.stream), "D2H triples")) return false;
  if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return false;
  const int ncps = hw[0], rn = hw[1];
  if (ncps <= 0) return false;                                      // invalid UTF-8 or nothing decoded (:475-478, :231-234)
  if (rn < 0 || rn > 3 * ncps || rn % 3 != 0) return false;         // :500-502, :252-254
  if (rn > 0 && !eager) {
    if (!cuda_ok(cudaMemcpyAsync(hw + 2, sl.lex_tri.p, (size_t)rn * 4, cudaMemcpyDeviceToHost, sl.stream), "D2H triples")) return false;
    if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return false;
  }
  // code point -> byte offset of the (already validated) input, BOM skipped like the decoder
  R->cp_off.resize((size_t)ncps + 1);
  int p = (n >= 3 && (uint8_t)s[0] == 0xEF && (uint8_t)s[1] == 0xBB && (uint8_t)s[2] == 0xBF) ? 3 : 0;
  for (int i = 0; i < ncps; ++i) {
    R->cp_off[(size_t)i] = p;
    const uint8_t c = (uint8_t)s[p];
    p += c < 0x80 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : 4;
  }
  R->cp_off[(size_t)ncps] = p;
  R->ncps = ncps; R->rn = rn; R->tri = hw + 2;
  return true;
}

inline bool white_cp_at(const char* s, int b) {                      // __FAIsWhiteSpace__ (blingfiretokdll.h:17-21) of the code point at byte b
  const uint8_t c0 = (uint8_t)s[b];
  uint32_t cp;
  if (c0 < 0x80) cp = c0;
  else if ((c0 & 0xE0) == 0xC0) cp = ((c0 & 0x1Fu) << 6) | ((uint8_t)s[b + 1] & 0x3Fu);
  else if ((c0 & 0xF0) == 0xE0) cp = ((c0 & 0x0Fu) << 12) | (((uint8_t)s[b + 1] & 0x3Fu) << 6) | ((uint8_t)s[b + 2] & 0x3Fu);
  else cp = ((c0 & 0x07u) << 18) | (((uint8_t)s[b + 1] & 0x3Fu) << 12) | (((uint8_t)s[b + 2] & 0x3Fu) << 6) | ((uint8_t)s[b + 3] & 0x3Fu);
  if (cp == 0) cp = 0x20;                                            // U+0000 was replaced before the lexer ran (:482, :237)
  return cp <= 0x20 || cp == 0xa0 || (cp >= 0x2000 && cp <= 0x200f) || cp == 0x202f || cp == 0x205f || cp == 0x2060 ||
         cp == 0x2420 || cp == 0x2424 || cp == 0x3000 || cp == 0xfeff;
}

inline int end_offset_of(const char* s, int off) {                   // ToOffset + FAUtf8Size(last char) - 1 (:526-530)
  const uint8_t c = (uint8_t)s[off];
  const int cs = c < 0x80 ? 1 : (c & 0xE0) == 0xC0 ? 2 : (c & 0xF0) == 0xE0 ? 3 : (c & 0xF8) == 0xF0 ? 4 : 0;
  return off + (cs > 0 ? cs - 1 : 0);
}

int finish_text(const std::string& os, char* out, int max_out) {
  const int len = (int)os.size();                                     // includes the trailing NUL (:555, :343)
  if (len <= max_out && out) std::memcpy(out, os.data(), os.size());
  return len;
}

}  // namespace

// blingfiretokdll.cpp:415-566.  The host only turns the (Tag, From, To) triples back into the ' '-joined
// UTF-8 string, like the reference's ostringstream loop (:507-552), and fills the offset arrays.
int TextToWordsWithOffsetsWithModel(const char* s, int n, char* out, int* starts, int* ends, const int max_out, void* hModel) {
  try {
    g_last_error.clear();
    Model* m = hModel ? (Model*)hModel : default_model(0);
    if (!m) return -1;
    if (n == 0) return 0;                                           // :446-448
    if (n < 0 || n > 1000000000 || !s) return -1;                   // :449-454
    if (!m->has_wbd || !m->lex_ok) { set_error("model has no lexer engine"); return -1; }
    if (starts && max_out > 0) std::memset(starts, 0, (size_t)max_out * sizeof(int));   // :467-472
    if (ends && max_out > 0) std::memset(ends, 0, (size_t)max_out * sizeof(int));
    CtxLease lease(m);
    LexedText R;
    if (!lex_on_gpu(m, lease.c, s, n, &R)) return -1;
    std::string os;
    os.reserve((size_t)n + (size_t)R.rn / 3 + 2);
    bool added = false;
    int words = 0;
    for (int i = 0; i < R.rn; i += 3) {
      if (R.tri[i] == 4) continue;                                  // WBD_IGNORE_TAG (:511-514)
      const int from = R.tri[i + 1], to = R.tri[i + 2];
      if (from < 0 || from > R.ncps || to >= R.ncps || to < -1) return -1;
      if (to >= 0 && from < R.ncps) {
        if (starts && words < max_out) starts[words] = R.cp_off[(size_t)from];
        if (ends && words < max_out) ends[words] = end_offset_of(s, R.cp_off[(size_t)to]);
      }
      ++words;
      if (added) os.push_back(' ');
      for (int b = R.cp_off[(size_t)from]; b < R.cp_off[(size_t)to + 1]; ++b) {
        char c = s[b];
        if (c == 0) c = 0x20;                                       // U+0000 -> U+0020 (:482)
        if (c == ' ') c = '_';                                      // ' ' is the delimiter (:546)
        os.push_back(c);
      }
      added = true;
    }
    os.push_back('\0');                                             // :555
    return finish_text(os, out, max_out);
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}
int TextToWordsWithOffsets(const char* s, int n, char* out, int* starts, int* ends, const int max_out) {
  return TextToWordsWithOffsetsWithModel(s, n, out, starts, ends, max_out, nullptr);
}
int TextToWordsWithModel(const char* s, int n, char* out, const int max_out, void* hModel) {
  return TextToWordsWithOffsetsWithModel(s, n, out, nullptr, nullptr, max_out, hModel);
}
int TextToWords(const char* s, int n, char* out, const int max_out) { return TextToWordsWithOffsetsWithModel(s, n, out, nullptr, nullptr, max_out, nullptr); }

// Additive: TextToWords[WithModel] for a batch (CSR documents in, CSR strings out).  The lexer AND the string building
// (blingfiretokdll.cpp:507-555) run on the GPU; documents go through in chunks of a few MB.
//   results[i]      what TextToWordsWithModel returns for document i: -1 error, 0 empty input, else the byte length of
//                   its string including the NUL
//   out_offsets[i]  where that string starts in `out` (documents without one take no bytes); out_offsets[ndocs] = total
// Returns the total bytes, or -total when `capacity` is too small (out_offsets and results are complete then), -1 on error.
static int64_t text_batch(bool sentences, void* hModel, const char* utf8, const int64_t* offsets, int64_t ndocs, char* out, int64_t capacity,
                          int64_t* out_offsets, int32_t* results) {
  try {
    g_last_error.clear();
    Model* m = hModel ? (Model*)hModel : default_model(sentences ? 1 : 0);
    if (!m) return -1;
    if (ndocs < 0 || !out_offsets || !results || (ndocs > 0 && (!utf8 || !offsets)) || (capacity > 0 && !out)) { set_error("bad batch arguments"); return -1; }
    if (!m->has_wbd || !m->lex_ok) { set_error("model has no lexer engine"); return -1; }
    if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return -1;
    CtxLease lease(m);
    Slot& sl = lease.c->slots[0];
    if (!ensure_stream(sl)) return -1;
    int64_t total = 0;
    bool overflow = false;
    out_offsets[0] = 0;
    const int64_t kChunkBytes = 8ll << 20;
    for (int64_t d0 = 0; d0 < ndocs;) {
      int64_t d1 = d0;
      while (d1 < ndocs && (d1 == d0 || offsets[d1 + 1] - offsets[d0] <= kChunkBytes) && d1 - d0 < (1 << 20)) ++d1;
      const int64_t nd = d1 - d0;
      const int64_t b0 = offsets[d0] & ~(int64_t)3, b1 = offsets[d1];
      const size_t nb = (size_t)(b1 - b0), span = (size_t)(b1 - offsets[d0]);
      if (!sl.text.reserve(nb + 64) || !sl.offsets.reserve((size_t)nd + 1) || !sl.lex_cls.reserve(span + 8) || !sl.lex_ncps.reserve((size_t)nd) ||
          !sl.lex_tri_count.reserve((size_t)nd) || !sl.lex_tri.reserve(3 * span + 8) || !sl.lex_boff.reserve(span + 8) ||
          !sl.counts.reserve((size_t)nd + 1) || !sl.ids.reserve((size_t)nd + 1) || !sl.row_off.reserve((size_t)nd + 1) ||
          !sl.h_row_off.reserve((size_t)nd + 2) || !sl.h_csr.reserve((size_t)nd + 1))
        return -1;
      if (nb && !cuda_ok(cudaMemcpyAsync(sl.text.p, utf8 + b0, nb, cudaMemcpyHostToDevice, sl.stream), "H2D text")) return -1;
      if (!cuda_ok(cudaMemcpyAsync(sl.offsets.p, offsets + d0, ((size_t)nd + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, sl.stream), "H2D offsets")) return -1;
      LexLaunch X = make_lex_launch(sl, sl.text.p - b0, sl.offsets.p, offsets[d0], b1, nd, m->d_cls_words, 1);   // MaxOut = 3 * MaxBuffSize (:492-499)
      X.boff_buf = sl.lex_boff.p;
      int nl = 0;
      if (!cuda_ok(lex_launch(X, make_lex_model(m), sl.stream, &nl), "lexer launch")) return -1;
      int32_t* d_lens = sl.counts.p;
      int32_t* d_results = sl.ids.p;
      if (!cuda_ok(lex_words_len_launch(X, sentences, d_lens, d_results, sl.stream), "words length launch")) return -1;
      if (!cuda_ok(wp_scan_counts(d_lens, sl.row_off.p, nd, sl.stream), "scan")) return -1;
      g_launches += nl + 2;
      if (!cuda_ok(cudaMemcpyAsync(sl.h_row_off.p, sl.row_off.p, ((size_t)nd + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, sl.stream), "D2H offsets")) return -1;
      if (!cuda_ok(cudaMemcpyAsync(sl.h_csr.p, d_results, (size_t)nd * sizeof(int32_t), cudaMemcpyDeviceToHost, sl.stream), "D2H results")) return -1;
      if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return -1;
      const int64_t nout = sl.h_row_off.p[nd];
      for (int64_t i = 0; i < nd; ++i) { out_offsets[d0 + i + 1] = total + sl.h_row_off.p[i + 1]; results[d0 + i] = sl.h_csr.p[i]; }
      if (total + nout > capacity) overflow = true;
      else if (nout > 0) {
        if (!sl.csr.reserve((size_t)(nout + 3) / 4 + 1)) return -1;
        if (!cuda_ok(lex_words_write_launch(X, sentences, sl.row_off.p, d_results, reinterpret_cast<char*>(sl.csr.p), sl.stream), "words write launch")) return -1;
        g_launches += 1;
        if (!cuda_ok(cudaMemcpyAsync(out + total, sl.csr.p, (size_t)nout, cudaMemcpyDeviceToHost, sl.stream), "D2H text")) return -1;
        if (!cuda_ok(cudaStreamSynchronize(sl.stream), "sync")) return -1;
      }
      total += nout;
      d0 = d1;
    }
    return overflow ? -total : total;
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}
int64_t TextToWordsBatch(void* hModel, const char* utf8, const int64_t* offsets, int64_t ndocs, char* out, int64_t capacity,
                         int64_t* out_offsets, int32_t* results) {
  return text_batch(false, hModel, utf8, offsets, ndocs, out, capacity, out_offsets, results);
}
// Additive: TextToSentences[WithModel] for a batch, same conventions (blingfiretokdll.cpp:163-355; sentences joined by '\n').
int64_t TextToSentencesBatch(void* hModel, const char* utf8, const int64_t* offsets, int64_t ndocs, char* out, int64_t capacity,
                             int64_t* out_offsets, int32_t* results) {
  return text_batch(true, hModel, utf8, offsets, ndocs, out, capacity, out_offsets, results);
}

// blingfiretokdll.cpp:163-355.  One sentence per triple of the sentence-breaking lexer: it starts right
// after the previous one (tags and Froms are ignored, :262-266), leading white space is dropped, '\n'
// inside becomes ' ', sentences are joined by '\n'; what follows the last boundary is the last sentence.
int TextToSentencesWithOffsetsWithModel(const char* s, int n, char* out, int* starts, int* ends, const int max_out, void* hModel) {
  try {
    g_last_error.clear();
    Model* m = hModel ? (Model*)hModel : default_model(1);
    if (!m) return -1;
    if (n == 0) return 0;                                           // :206-208
    if (n < 0 || n > 1000000000 || !s) return -1;                   // :209-214
    if (!m->has_wbd || !m->lex_ok) { set_error("model has no lexer engine"); return -1; }
    if (starts && max_out > 0) std::memset(starts, 0, (size_t)max_out * sizeof(int));   // :227-232
    if (ends && max_out > 0) std::memset(ends, 0, (size_t)max_out * sizeof(int));
    CtxLease lease(m);
    LexedText R;
    if (!lex_on_gpu(m, lease.c, s, n, &R)) return -1;
    std::string os;
    os.reserve((size_t)n + 2);
    bool added = false;
    int count = 0, prev_end = -1;
    auto sentence = [&](int from, int to) {
      int first = from;
      while (first <= to && white_cp_at(s, R.cp_off[(size_t)first])) ++first;   // FAGetFirstNonWhiteSpace (:138-150)
      if (first > to) return;
      if (starts && count < max_out) starts[count] = R.cp_off[(size_t)first];
      if (ends && count < max_out) ends[count] = end_offset_of(s, R.cp_off[(size_t)to]);
      ++count;
      if (added) os.push_back('\n');
      for (int b = R.cp_off[(size_t)first]; b < R.cp_off[(size_t)to + 1]; ++b) {
        char c = s[b];
        if (c == 0) c = 0x20;                                       // :237
        if (c == '\n') c = ' ';                                     // '\n' is the delimiter (:292)
        os.push_back(c);
      }
      added = true;
    };
    for (int i = 0; i < R.rn; i += 3) {
      const int to = R.tri[i + 2];
      if (to < -1 || to >= R.ncps) return -1;                      // (never: To is a position of the input)
      sentence(prev_end + 1, to);
      prev_end = to;
    }
    if (prev_end + 1 < R.ncps) sentence(prev_end + 1, R.ncps - 1);   // the end of the paragraph ends a sentence (:303-338)
    os.push_back('\0');                                             // :343
    return finish_text(os, out, max_out);
  } catch (const std::exception& e) { set_error(e.what()); return -1; }
}
int TextToSentencesWithOffsets(const char* s, int n, char* out, int* starts, int* ends, const int max_out) {
  return TextToSentencesWithOffsetsWithModel(s, n, out, starts, ends, max_out, nullptr);
}
int TextToSentencesWithModel(const char* s, int n, char* out, const int max_out, void* hModel) {
  return TextToSentencesWithOffsetsWithModel(s, n, out, nullptr, nullptr, max_out, hModel);
}
int TextToSentences(const char* s, int n, char* out, const int max_out) { return TextToSentencesWithOffsetsWithModel(s, n, out, nullptr, nullptr, max_out, nullptr); }

}  // extern "C"
