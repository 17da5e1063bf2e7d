// TEST INFRASTRUCTURE ONLY -- host executor for the DFSMN launch sequence (csrc/dfsmn_ops.cuh, the sequence libadn runs on
// the GPU): plain loops per functor, the contractions through translate() + gemm_ref.  Never linked into libadn.so.
#include "dfsmn_ops.cuh"

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
  void run(long long n, const gan::Linear& f) {     // as on the GPU: the GemmOp of translate()
    ++launches;
    gan::GemmOp op;
    gan::translate(f, n, &op);
    gan::gemm_ref(op);
  }
  void gemm(const gan::GemmOp& g) { ++launches; gan::gemm_ref(g); }
  void mark(const char* tag, const char* name, const float* p, long long count) {
    if (!dump) return;
    std::string key = tag[0] ? std::string(tag) + "." + name : std::string(name);
    dump(key.c_str(), p, count);
  }
};

// audio: (B, L) fp32 or int16 (is_i16); out: (B, 2 * 961, T) masked packed spectrum (the ISTFT operand)
extern "C" int dfsmn_host_forward(const char* const* names, const unsigned long long* offsets, const unsigned long long* counts,
                                  int n_tensors, const float* blob, int layers, int lorder, int B, int L, const void* audio,
                                  int is_i16, float* spec, dump_fn dump, char* errbuf, int errlen) {
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
  dfs::Weights W;
  if (!dfs::bind(W, layers, lorder, lk)) {
    snprintf(errbuf, errlen, "%s", err.c_str());
    return -1;
  }
  const int T = (L - dfs::FRAME) / dfs::HOP + 1;
  std::vector<std::vector<float>> bufs;
  auto alloc = [&](size_t n) { bufs.emplace_back(n ? n : 1, std::nanf("")); return bufs.back().data(); };   // NaN poison: cudaMalloc does not zero either
  dfs::Workspace ws;
  if (!dfs::alloc_ws(ws, B, T, alloc)) return -2;
  std::vector<float> x((size_t)B * L);
  HostExec ex;
  ex.dump = dump;
  if (is_i16) ex.run((long long)B * L, dfs::Prep<int16_t>{(const int16_t*)audio, 1.0f / 32768.0f, x.data()});
  else ex.run((long long)B * L, dfs::Prep<float>{(const float*)audio, 1.0f, x.data()});
  dfs::forward(ex, ws, W, x.data(), spec, B, L, T);
  return ex.launches;
}

extern "C" void dfsmn_host_out_i16(const float* wave, short* out, long long n) {
  dfs::OutI16 f{wave, out};
  for (long long i = 0; i < n; ++i) f(i);
}
