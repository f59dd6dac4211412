// copy_pool.h -- library-owned host threads that move bytes between the caller's PAGEABLE buffers and the
// pinned staging buffers of the host batch pipeline (capi.cu).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through one small bounce buffer (~2.4 GB/s
// measured on the bench box against ~55 GB/s for pinned memory), so a caller that owns plain malloc / numpy
// memory -- which is every caller of the reference's bindings (dist-pypi/blingfire/__init__.py:243-253) --
// would otherwise see a twentieth of the PCIe rate.  The pool splits one memcpy into slices, one per
// thread; the threads are bound to the CPUs of the GPU's NUMA node so that the pinned staging buffers
// (first touched by them) and the DMA engine stay on the same socket.
#pragma once

#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace bfb200 {

// CPUs of the NUMA node a PCI device hangs off ("0000:1b:00.0"), intersected with the CPUs this process
// may run on.  Empty set: unknown topology (no binding is done then).
inline bool numa_cpus_of_pci(const char* bus_id, cpu_set_t* out) {
  CPU_ZERO(out);
  std::string id(bus_id ? bus_id : "");
  for (auto& c : id) c = (char)tolower((unsigned char)c);
  if (id.size() > 12) id = id.substr(id.size() - 12);   // cuda prints an 8-digit domain, sysfs has 4
  std::ifstream f("/sys/bus/pci/devices/" + id + "/numa_node");
  int node = -1;
  if (!(f >> node) || node < 0) return false;
  std::ifstream g("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  std::string list;
  if (!std::getline(g, list) || list.empty()) return false;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return false;
  size_t i = 0;
  int n = 0;
  while (i < list.size()) {
    char* end = nullptr;
    long a = std::strtol(list.c_str() + i, &end, 10), b = a;
    i = (size_t)(end - list.c_str());
    if (i < list.size() && list[i] == '-') { b = std::strtol(list.c_str() + i + 1, &end, 10); i = (size_t)(end - list.c_str()); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET((int)c, &allowed)) { CPU_SET((int)c, out); ++n; }
    if (i < list.size() && list[i] == ',') ++i; else if (i < list.size() && !isdigit((unsigned char)list[i])) break;
  }
  return n > 0;
}

// Completion of one (or several) submitted copies.
struct CopyJob {
  std::atomic<size_t> left{0};
  bool done() const { return left.load(std::memory_order_acquire) == 0; }
};

class CopyPool {
 public:
  // `threads` workers in addition to the calling thread; `cpus` (may be null) = where they may run
  CopyPool(int threads, const cpu_set_t* cpus) {
    for (int t = 0; t < threads; ++t) {
      workers_.emplace_back([this] { run(); });
      if (cpus) pthread_setaffinity_np(workers_.back().native_handle(), sizeof(cpu_set_t), cpus);
    }
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int threads() const { return (int)workers_.size(); }

  // Queues the copy in slices and returns; `job` counts them down.  The buffers must stay valid until wait(job).
  void submit(void* dst, const void* src, size_t n, CopyJob* job) {
    if (n == 0) return;
    const size_t parts = (n + kSlice - 1) / kSlice;
    job->left.fetch_add(parts, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> l(mu_);
      for (size_t k = 0; k < parts; ++k) {
        const size_t a = k * kSlice, b = a + kSlice < n ? a + kSlice : n;
        q_.push_back(Slice{(uint8_t*)dst + a, (const uint8_t*)src + a, b - a, job, nullptr});
      }
    }
    if (parts > 1) cv_.notify_all(); else cv_.notify_one();
  }
  // Blocks until every slice of `job` is through; the caller takes slices (of any job) meanwhile.
  void wait(CopyJob* job) {
    while (!job->done()) {
      Slice s;
      if (pop(&s)) run_slice(s); else std::this_thread::yield();
    }
  }
  // blocking parallel loop: fn(k) for k in [0, parts), the calling thread takes its share
  void parallel_for(size_t parts, const std::function<void(size_t)>& fn) {
    if (parts == 0) return;
    if (parts == 1 || workers_.empty()) { for (size_t k = 0; k < parts; ++k) fn(k); return; }
    CopyJob job;
    job.left.fetch_add(parts, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> l(mu_);
      for (size_t k = 0; k < parts; ++k) q_.push_back(Slice{nullptr, nullptr, k, &job, &fn});
    }
    cv_.notify_all();
    wait(&job);
  }
  // blocking parallel memcpy
  void copy(void* dst, const void* src, size_t n) {
    if (n <= 2 * kSlice || workers_.empty()) { std::memcpy(dst, src, n); return; }
    CopyJob job;
    submit(dst, src, n, &job);
    wait(&job);
  }

 private:
  static constexpr size_t kSlice = 1u << 20;
  struct Slice { uint8_t* dst; const uint8_t* src; size_t n; CopyJob* job; const std::function<void(size_t)>* fn = nullptr; };
  static void run_slice(const Slice& s) {
    if (s.fn) (*s.fn)(s.n); else std::memcpy(s.dst, s.src, s.n);
    s.job->left.fetch_sub(1, std::memory_order_release);
  }
  bool pop(Slice* out) {
    std::lock_guard<std::mutex> l(mu_);
    if (q_.empty()) return false;
    *out = q_.front(); q_.pop_front();
    return true;
  }
  void run() {
    for (;;) {
      Slice s;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || !q_.empty(); });
        if (stop_) return;
        s = q_.front(); q_.pop_front();
      }
      run_slice(s);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Slice> q_;
  bool stop_ = false;
};

}  // namespace bfb200
