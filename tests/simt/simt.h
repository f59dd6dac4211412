// simt.h -- TEST INFRASTRUCTURE ONLY.
//
// Runs the SOURCE of the product's CUDA kernels on the CPU: every lane of a warp is an OS thread and
// every full-mask warp intrinsic (__shfl*_sync, __ballot_sync, __any_sync, __reduce_max_sync,
// __syncwarp) is a rendezvous of the 32 threads of the warp.  A kernel whose lanes do not all reach the
// same intrinsics in the same order deadlocks here (the tests run under a timeout), which is exactly the
// property the real hardware needs from full-mask intrinsics.  Slow (a rendezvous costs microseconds), so
// the tests use a few hundred short documents; the GPU suite remains the parity test proper.
//
// The kernels are compiled with -DBF_SIMT_HOST, which only swaps the declaration of the dynamic shared
// memory for a pointer, removes the launch wrappers, and replaces two inline-PTX helpers.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
inline int2 make_int2(int a, int b) { return int2{a, b}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
struct dim3s { unsigned x = 0, y = 0, z = 0; };

typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;

namespace simt {

struct Warp {
  std::barrier<> bar{32};
  uint64_t buf[2][32];
  unsigned phase = 0;   // only read/written between rendezvous by all lanes consistently (each lane keeps its own copy)
};

struct LaneState {
  int lane = 0;
  Warp* warp = nullptr;
  unsigned phase = 0;
  uint8_t* smem = nullptr;
  std::barrier<>* cta = nullptr;       // all emulated threads of the CTA (__syncthreads)
};
inline thread_local LaneState tl;
inline uint8_t* shared_base() { return tl.smem; }

// one rendezvous: publish v, wait, return the 32 published values (valid until this lane's next-but-one rendezvous)
inline const uint64_t* exchange(uint64_t v) {
  Warp* w = tl.warp;
  uint64_t* b = w->buf[tl.phase & 1];
  tl.phase++;
  b[tl.lane] = v;
  w->bar.arrive_and_wait();
  return b;
}
template <class T> inline uint64_t pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unpack(uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }

// Runs fn(lane) on 32 lane threads per warp, `warps` warps, sharing one block of dynamic shared memory.
inline void run_cta(int warps, size_t smem_bytes, const std::function<void()>& kernel_body);

}  // namespace simt

inline thread_local dim3s threadIdx, blockIdx;
inline dim3s blockDim, gridDim;

template <class T> inline T __shfl_sync(unsigned, T v, int src) { const uint64_t* b = simt::exchange(simt::pack(v)); return simt::unpack<T>(b[src & 31]); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) { const uint64_t* b = simt::exchange(simt::pack(v)); const int s = simt::tl.lane - (int)d; return simt::unpack<T>(b[s < 0 ? simt::tl.lane : s]); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { const uint64_t* b = simt::exchange(simt::pack(v)); const int s = simt::tl.lane + (int)d; return simt::unpack<T>(b[s > 31 ? simt::tl.lane : s]); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { const uint64_t* b = simt::exchange(simt::pack(v)); return simt::unpack<T>(b[(simt::tl.lane ^ m) & 31]); }
inline unsigned __ballot_sync(unsigned, bool p) { const uint64_t* b = simt::exchange(p ? 1u : 0u); unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (unsigned)(b[i] & 1u) << i; return r; }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline int __reduce_max_sync(unsigned, int v) { const uint64_t* b = simt::exchange(simt::pack(v)); int r = simt::unpack<int>(b[0]); for (int i = 1; i < 32; ++i) r = std::max(r, simt::unpack<int>(b[i])); return r; }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { const uint64_t* b = simt::exchange(simt::pack(v)); unsigned r = 0; for (int i = 0; i < 32; ++i) r += simt::unpack<unsigned>(b[i]); return r; }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::exchange(0); }
inline void __syncthreads() { simt::tl.cta->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {     // PRMT, default mode (selectors 0..7, no sign replication)
  const uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline float __int_as_float(int v) { return simt::unpack<float>((uint64_t)(uint32_t)v); }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
template <class T> inline T __ldg(const T* p) { return *p; }

template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int cmp, int val) { __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

using std::max;
using std::min;

namespace simt {
inline void run_cta(int warps, size_t smem_bytes, const std::function<void()>& kernel_body) {
  std::vector<uint8_t> smem(smem_bytes + 64, 0xCD);   // shared memory is NOT zero at kernel start on the device either
  uint8_t* base = smem.data() + ((64 - (reinterpret_cast<uintptr_t>(smem.data()) & 63)) & 63);
  std::vector<Warp> ws((size_t)warps);
  std::barrier<> cta_bar(warps * 32);
  std::vector<std::thread> th;
  for (int w = 0; w < warps; ++w)
    for (int l = 0; l < 32; ++l)
      th.emplace_back([&, w, l] {
        tl.lane = l; tl.warp = &ws[(size_t)w]; tl.phase = 0; tl.smem = base; tl.cta = &cta_bar;
        threadIdx.x = (unsigned)(w * 32 + l); blockIdx.x = 0;
        kernel_body();
      });
  for (auto& t : th) t.join();
}
}  // namespace simt
