// tsan_main.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the kernel source over the SIMT shim under ThreadSanitizer
// (`make -C tests/simt tsan`): with one OS thread per lane and the warp intrinsics as the only synchronisation, a
// shared-memory access that is not ordered by an intrinsic shows up as a data race -- a CPU-side racecheck.
// Usage: tsan_run sp|spo|wp <model.bin> <text file>      (spo: the [pos-dict] kernels with offsets)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
extern "C" {
void* spsim_load(const char*); const char* spsim_error(void*);
int spsim_batch(void*, const char*, const int64_t*, int64_t, int32_t*, int32_t*, int32_t*, int32_t*, int, int, int);
void* wpsim_load(const char*); const char* wpsim_error(void*);
int wpsim_batch(void*, const char*, const int64_t*, int64_t, int32_t*, int32_t*, int, int, int);
}
int main(int argc, char** argv) {
  std::ifstream f(argv[3]); std::string line; std::vector<std::string> docs; std::string acc;
  int k = 0;
  while (std::getline(f, line) && docs.size() < 40) { acc += line + " "; if (++k % 6 == 0) { docs.push_back(acc); acc.clear(); } }
  docs.push_back(std::string(900, 'a') + " tail"); docs.push_back("abc \xff def");
  std::string text; std::vector<int64_t> offs{0};
  for (auto& d : docs) { text += d; offs.push_back((int64_t)text.size()); }
  const int max_ids = 2048; std::vector<int32_t> ids(docs.size() * max_ids), counts(docs.size());
  if (std::string(argv[1]) == "sp" || std::string(argv[1]) == "spo") {
    const bool off = std::string(argv[1]) == "spo";
    std::vector<int32_t> st(off ? ids.size() : 0), en(off ? ids.size() : 0);
    void* h = spsim_load(argv[2]); if (*spsim_error(h)) { puts(spsim_error(h)); return 1; }
    printf("flag %d\n", spsim_batch(h, text.c_str(), offs.data(), (int64_t)docs.size(), ids.data(), counts.data(), off ? st.data() : nullptr,
                                    off ? en.data() : nullptr, max_ids, 0, 2));
  } else {
    void* h = wpsim_load(argv[2]); if (*wpsim_error(h)) { puts(wpsim_error(h)); return 1; }
    printf("rc %d\n", wpsim_batch(h, text.c_str(), offs.data(), (int64_t)docs.size(), ids.data(), counts.data(), max_ids, 100, 2));
  }
  long tot = 0; for (auto c : counts) tot += c; printf("docs %zu tokens %ld\n", docs.size(), tot);
}
