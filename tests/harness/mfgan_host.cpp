// TEST INFRASTRUCTURE ONLY -- host executor for the MossFormerGAN-SE-16K launch sequence.
// Instantiates gan::forward (csrc/mfgan_ops.cuh, the sequence libadn runs on the GPU) with a plain loop per
// operator, so every functor's index arithmetic and the buffer plumbing are checked against the oracle on a
// machine without a GPU (tests/test_mfgan_host.py).  Never linked into libadn.so.
#include "mfgan_gemm.cuh"

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
  // use_gemm: the contractions go through translate() + gemm_ref, i.e. the GemmOp the CUDA executor launches
  bool use_gemm = false;
  template <class F>
  void run_gemm(long long n, const F& f) {
    if (!use_gemm) { run<F>(n, f); return; }
    ++launches;
    gan::GemmOp ops[3];
    const int k = gan::translate(f, n, ops);
    for (int i = 0; i < k; ++i) gan::gemm_ref(ops[i]);
  }
  void run(long long n, const gan::Linear& f) { run_gemm(n, f); }
  void run(long long n, const gan::SimLocal& f) { run_gemm(n, f); }
  void run(long long n, const gan::SimCross& f) { run_gemm(n, f); }
  void run(long long n, const gan::LinKV& f) { run_gemm(n, f); }
  void run(long long n, const gan::Att& f) { run_gemm(n, f); }
  void run(long long n, const gan::GateConvT& f) { run_gemm(n, f); }
  void run(long long n, const gan::TaScores& f) { run_gemm(n, f); }
  void run(long long n, const gan::TaAV& f) { run_gemm(n, f); }
  void run(long long n, const gan::Conv2d& f) {
    if (f.Cout >= 16 && f.Cin % 16 == 0) run_gemm(n, f);
    else run<gan::Conv2d>(n, f);
  }
  void mark(const char* tag, const char* name, const float* p, long long count) {
    if (!dump) return;
    std::string key = tag[0] ? std::string(tag) + "." + name : std::string(name);
    dump(key.c_str(), p, count);
  }
};

extern "C" int mfgan_host_forward(const char* const* names, const unsigned long long* offsets, const unsigned long long* counts,
                                  int n_tensors, const float* blob, int layers, int B, int T, const float* feat, float* mask,
                                  float* cplx, dump_fn dump, char* errbuf, int errlen, int use_gemm) {
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
  gan::Weights W;
  if (!gan::bind(W, layers, T, lk)) {
    snprintf(errbuf, errlen, "%s", err.c_str());
    return -1;
  }
  std::vector<std::vector<float>> bufs;
  auto alloc = [&](size_t n) { bufs.emplace_back(n ? n : 1, std::nanf("")); return bufs.back().data(); };   // NaN poison: cudaMalloc does not zero either
  gan::Workspace ws;
  if (!gan::alloc_ws(ws, B, T, alloc)) return -2;
  HostExec ex;
  ex.dump = dump;
  ex.use_gemm = use_gemm != 0;
  gan::forward(ex, ws, W, feat, mask, cplx, B, T);
  return ex.launches;
}
