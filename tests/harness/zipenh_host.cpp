// TEST INFRASTRUCTURE ONLY -- host executor for the ZipEnhancer launch sequence.
// Instantiates zip::forward (csrc/zipenh_ops.cuh, the sequence libadn runs on the GPU) with a plain loop per functor and
// zip::lin_ref per LinOp, so every functor's index arithmetic, the implicit-GEMM addressing of the convs, the weight
// layouts of adn/zipenh_params.py and the buffer plumbing are checked against the oracle on a machine without a GPU
// (tests/test_zipenh_host.py).  Never linked into libadn.so.
#include "zipenh_ops.cuh"

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
  void gemm(const zip::LinOp& g, const char*) {
    ++launches;
    zip::lin_ref(g);
  }
  void mark(const char* name, const float* p, long long count) {
    if (dump) dump(name, p, count);
  }
  void mark_strided(const char* name, const float* src, long long pixels, int ld, int coff, int width) {
    if (!dump) return;
    std::vector<float> v((size_t)pixels * width);
    for (long long p = 0; p < pixels; ++p)
      for (int c = 0; c < width; ++c) v[p * width + c] = src[p * ld + coff + c];
    dump(name, v.data(), (long long)v.size());
  }
};

extern "C" int zipenh_host_forward(const char* const* names, const unsigned long long* offsets, const unsigned long long* counts,
                                   int n_tensors, const float* blob, const int* ds, int B, int T, const float* feat, float* mx,
                                   float* ri, dump_fn dump, char* errbuf, int errlen) {
  std::map<std::string, std::pair<unsigned long long, unsigned long long>> index;
  for (int i = 0; i < n_tensors; ++i) index[names[i]] = {offsets[i], counts[i]};
  std::string err;
  auto lk = [&](const char* name, size_t expect) -> const float* {
    auto it = index.find(name);
    if (it == index.end() || (expect && it->second.second != expect)) {
      if (err.empty()) err = std::string("tensor '") + name + "' missing or wrong size";
      return nullptr;
    }
    return blob + it->second.first;
  };
  zip::Weights W;
  if (!zip::bind(W, T, ds, lk)) {
    snprintf(errbuf, errlen, "%s", err.c_str());
    return -1;
  }
  std::vector<std::vector<float>> bufs;
  auto alloc = [&](size_t n) { bufs.emplace_back(n ? n : 1, std::nanf("")); return bufs.back().data(); };   // NaN poison: cudaMalloc does not zero either
  zip::Workspace ws;
  if (!zip::alloc_ws(ws, B, T, alloc)) return -2;
  HostExec ex;
  ex.dump = dump;
  zip::forward(ex, ws, W, feat, mx, ri, B, T);
  return ex.launches;
}
