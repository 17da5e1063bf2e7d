// TEST INFRASTRUCTURE ONLY -- host executor for the UL-UNAS launch sequence (csrc/ulunas_ops.cuh, the sequence libadn runs on the
// GPU): plain loops per functor, Linear through translate() + gemm_ref as on the GPU.  Never linked into libadn.so.
#include "ulunas_ops.cuh"

#include <cmath>
#include <map>
#include <string>
#include <vector>

typedef void (*dump_fn)(const char* name, const float* data, long long count);

struct HostExec {
  dump_fn dump = nullptr;
  int launches = 0;
  template <class F>
  void run(long long n, const F& f) {
    ++launches;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) f(i);
  }
  void run(long long n, const gan::Linear& f) {
    ++launches;
    gan::GemmOp op;
    gan::translate(f, n, &op);
    gan::gemm_ref(op);
  }
  void mark(const char* tag, const char* name, const float* p, long long count) {
    if (!dump) return;
    std::string key = tag[0] ? std::string(tag) + "." + name : std::string(name);
    dump(key.c_str(), p, count);
  }
};

// spec: (B, 514, T) packed STFT; out: (B, 514, T) masked spectrum (the ISTFT operand)
extern "C" int ulunas_host_forward(const char* const* names, const unsigned long long* offsets, const unsigned long long* counts,
                                   int n_tensors, const float* blob, int B, int T, const float* spec, float* out, dump_fn dump,
                                   char* errbuf, int errlen) {
  std::map<std::string, std::pair<unsigned long long, unsigned long long>> index;
  for (int i = 0; i < n_tensors; ++i) index[names[i]] = {offsets[i], counts[i]};
  std::string err;
  auto lk = [&](const char* name, size_t expect) -> const float* {
    auto it = index.find(name);
    if (it == index.end() || (expect && it->second.second != expect)) {
      if (err.empty()) err = std::string("tensor '") + name + "' missing or wrong size (" +
                             (it == index.end() ? "absent" : std::to_string(it->second.second) + " != " + std::to_string(expect)) + ")";
      return nullptr;
    }
    return blob + it->second.first;
  };
  uln::Weights W;
  if (!uln::bind(W, lk)) {
    snprintf(errbuf, errlen, "%s", err.c_str());
    return -1;
  }
  std::vector<std::vector<float>> bufs;
  auto alloc = [&](size_t n) { bufs.emplace_back(n ? n : 1, std::nanf("")); return bufs.back().data(); };   // NaN poison: cudaMalloc does not zero either
  uln::Workspace ws;
  if (!uln::alloc_ws(ws, B, T, alloc)) return -2;
  HostExec ex;
  ex.dump = dump;
  uln::forward(ex, ws, W, spec, out, B, T);
  return ex.launches;
}
