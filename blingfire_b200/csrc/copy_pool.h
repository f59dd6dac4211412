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
    { std::lock_guard<std::mutex> l(mu_); stop_ = true; ++epoch_; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int threads() const { return (int)workers_.size(); }

  // blocking parallel loop: fn(k) for k in [0, parts), the calling thread takes its share
  void parallel_for(size_t parts, const std::function<void(size_t)>& fn) {
    if (parts == 0) return;
    if (parts == 1 || workers_.empty()) { for (size_t k = 0; k < parts; ++k) fn(k); return; }
    auto job = std::make_shared<Job>();
    job->fn = &fn; job->parts = parts;
    { std::lock_guard<std::mutex> l(mu_); job_ = job; ++epoch_; }
    cv_.notify_all();
    work(*job);
    // the parts are short: spin until the last worker is through
    while (job->done.load(std::memory_order_acquire) < job->parts) std::this_thread::yield();
  }

  // blocking parallel memcpy
  void copy(void* dst, const void* src, size_t n) {
    if (n <= 2 * kSlice || workers_.empty()) { std::memcpy(dst, src, n); return; }
    uint8_t* d = (uint8_t*)dst;
    const uint8_t* s = (const uint8_t*)src;
    parallel_for((n + kSlice - 1) / kSlice, [&](size_t k) {
      const size_t a = k * kSlice, b = a + kSlice < n ? a + kSlice : n;
      std::memcpy(d + a, s + a, b - a);
    });
  }

 private:
  static constexpr size_t kSlice = 1u << 20;
  struct Job {
    const std::function<void(size_t)>* fn = nullptr;   // outlives the job: parallel_for blocks until done == parts
    size_t parts = 0;
    std::atomic<size_t> next{0}, done{0};
  };
  static void work(Job& j) {
    for (;;) {
      const size_t k = j.next.fetch_add(1, std::memory_order_relaxed);
      if (k >= j.parts) break;
      (*j.fn)(k);
      j.done.fetch_add(1, std::memory_order_release);
    }
  }
  void run() {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        job = job_;      // a worker that wakes late works on (the leftovers of) the job it saw
      }
      if (job) work(*job);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t epoch_ = 0;
  bool stop_ = false;
  std::shared_ptr<Job> job_;
};

}  // namespace bfb200
