// DFSMN 48 kHz causal denoiser behind the C ABI: model family "dfsmn".
// Reference: DFSMN/Export_DFSMN.py `DFSMN.forward` (:191-250).  The launch sequence is csrc/dfsmn_ops.cuh (functors shared with
// the CPU host harness); here: input cast, the sequence on the CUDA executor, ISTFT (periodic-hamming synthesis window,
// divide by the overlap-added w^2: `inverse_packed`, DFSMN/STFT_Process.py:304-310) and the output rule.
#include "dfsmn_ops.cuh"

#include "common.cuh"
#include "gan_exec.cuh"
#include "model_impl.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace gan {
GAN_OP_NAME(dfs::Power, "dfsmn_power");
GAN_OP_NAME(dfs::DwCausal, "dfsmn_memory");
GAN_OP_NAME(dfs::MaskApply, "dfsmn_mask_apply");
GAN_OP_NAME(dfs::OutI16, "dfsmn_out_i16");
GAN_OP_NAME(dfs::Prep<float>, "dfsmn_prep");
GAN_OP_NAME(dfs::Prep<int16_t>, "dfsmn_prep");
GAN_OP_NAME(dfs::Prep<__half>, "dfsmn_prep");
}  // namespace gan

namespace dfs {

struct OutF16 {
  const float* w; __half* out;
  __device__ void operator()(long long i) const { out[i] = __float2half_rn(w[i]); }
};

class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, T = 0, layers = 9, lorder = 20;
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;
  Weights W;
  Workspace ws;
  adn_stft* stft = nullptr;
  std::vector<void*> allocs;
  int cap = 0;
  float *x = nullptr, *spec = nullptr, *wave = nullptr;
  int stop_after = 0, last_launches = 0, last_batch = 0;
  std::map<std::string, std::vector<float>> dumps;

  ~Model() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    if (stft) adn_stft_destroy(stft);
  }
  gan::TcCache tcc;               // tcgen05 plans + weight operand planes of the analysis transform and the Linear layers (gan_exec.cuh)
  void free_ws() {
    adn_note_free();
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    tcc.plans.clear();
    cap = 0;
  }
  float* dalloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) { err = "dfsmn: out of device memory for the workspace"; return nullptr; }
    allocs.push_back(p);
    return (float*)p;
  }
  bool init(const std::map<std::string, std::string>& meta, const float* h_blob) {
    auto geti = [&](const char* k, int& v) {
      auto it = meta.find(k);
      if (it == meta.end() || it->second.empty()) { err = std::string("Required metadata key ") + k + " is missing."; return false; }
      v = atoi(it->second.c_str());
      return true;
    };
    auto gets = [&](const char* k, std::string& v) {
      auto it = meta.find(k);
      if (it == meta.end()) { err = std::string("Required metadata key ") + k + " is missing."; return false; }
      v = it->second;
      return true;
    };
    int nfft = 0, hop = 0, nmels = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", L) || !geti("nfft", nfft) || !geti("hop_length", hop) || !geti("n_mels", nmels) ||
        !geti("dfsmn_layers", layers) || !geti("dfsmn_lorder", lorder) || !gets("input_audio_dtype", sin) ||
        !gets("output_audio_dtype", sout))
      return false;
    if (nfft != FRAME || hop != HOP || nmels != NM || L < FRAME || (L - FRAME) % HOP) {
      err = "dfsmn needs nfft=1920, hop_length=960, n_mels=120 and a window of 1920 + k*960 samples (snip-edges framing)";
      return false;
    }
    {
      int in_sr = 48000, out_sr = 48000;
      auto opt = [&](const char* k, int& v) { auto it = meta.find(k); if (it != meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
      opt("in_sample_rate", in_sr); opt("out_sample_rate", out_sr);
      if (in_sr != 48000 || out_sr != 48000) { err = "dfsmn runs at 48 kHz I/O only"; return false; }
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = (L - FRAME) / HOP + 1;
    err.clear();
    auto lk = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || (expect && it->second.count != expect)) {
        if (err.empty()) err = std::string("weight blob: tensor '") + name + "' missing or wrong size";
        return nullptr;
      }
      return d_blob + it->second.offset;
    };
    if (!bind(W, layers, lorder, lk)) { if (err.empty()) err = "dfsmn: 1..16 layers, lorder 1..64"; return false; }
    auto host = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || it->second.count != expect) { err = std::string("weight blob: tensor '") + name + "' missing or wrong size"; return nullptr; }
      return h_blob + it->second.offset;
    };
    const float* an = host("analysis_w", (size_t)AN * FRAME);
    const float* inv = host("stft.inv", (size_t)2 * SB * FRAME);
    const float* nrm = host("stft.norm", (size_t)L);
    if (!an || !inv || !nrm) return false;
    adn_stft_geom g;
    memset(&g, 0, sizeof(g));
    g.nfft = FRAME; g.hop = HOP; g.center = 0; g.pad_reflect = 0; g.norm_multiply = 0;
    // forward basis = the mask-STFT rows of the fused analysis kernel (unused here: the analysis runs in the fused GEMM)
    if (adn_stft_create(&stft, &g, an + (size_t)2 * KB * FRAME, inv, nrm, T, device) != ADN_OK) {
      err = std::string("dfsmn: ") + adn_last_error(nullptr);
      return false;
    }
    { const char* e = getenv("ADN_STFT_TC"); if (!(e && e[0] == '0')) adn_stft_enable_tc(stft, sms); }   // ISTFT (40 % of the step on the exact GEMM) on tcgen05
    return true;
  }
  bool ensure(int B) {
    if (B <= cap) return true;
    cudaDeviceSynchronize();
    free_ws();
    auto a = [&](size_t n) { return dalloc(n); };
    if (!alloc_ws(ws, B, T, a)) return false;
    const size_t b = (size_t)B;
    if (!(x = dalloc(b * L)) || !(spec = dalloc(b * 2 * SB * T)) || !(wave = dalloc(b * L))) return false;
    cap = B;
    return true;
  }
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);        // Export_DFSMN.py:293
    in->dtype = in_dtype; in->channels = 1; in->length = L;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);   // :294
    out->dtype = out_dtype; out->channels = 1; out->length = L;
  }
  size_t workspace_bytes(int batch) override {
    const size_t rows = (size_t)batch * T;
    return (rows * (AN + KB + NM + 4 * H + SB) + (size_t)batch * (2 * L + 2 * SB * T)) * sizeof(float);
  }
  int launches(int) override { return 1 + 6 + 3 * layers + 2 + (out_dtype == ADN_F32 ? 0 : 1); }
  void set_stop_after(int n) override { stop_after = n; }

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    gan::CudaExec ex;
    ex.st = st; ex.tick = tick; ex.tick_ctx = tick_ctx;
    ex.capture = stop_after != 0; ex.dumps = &dumps;
    if (ex.capture) dumps.clear();
    tcc.sms = sms; tcc.min_rows = 1;
    { const char* e = getenv("ADN_GAN_TC"); tcc.enabled = !(e && e[0] == '0'); }
    ex.tc = &tcc; ex.tc_pass = B; ex.tc_idx = 0;
    const long long n = (long long)B * L;
    if (in_dtype == ADN_I16) ex.run(n, Prep<int16_t>{(const int16_t*)d_in, 1.0f / 32768.0f, x});
    else if (in_dtype == ADN_F16) ex.run(n, Prep<__half>{(const __half*)d_in, 1.0f, x});
    else ex.run(n, Prep<float>{(const float*)d_in, 1.0f, x});
    forward(ex, ws, W, x, spec, B, L, T);
    float* y = out_dtype == ADN_F32 ? (float*)d_out : wave;
    if (adn_stft_inverse(stft, spec, y, B, T, st) != ADN_OK) { err = std::string("dfsmn istft: ") + adn_last_error(nullptr); return ADN_ERR_CUDA; }
    ex.launches += 2;
    if (tick) tick(tick_ctx, "istft");
    if (out_dtype == ADN_I16) ex.run(n, OutI16{wave, (int16_t*)d_out});
    else if (out_dtype == ADN_F16) ex.run(n, OutF16{wave, (__half*)d_out});
    last_launches = ex.launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("dfsmn run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    if (!last_batch) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    if (!strcmp(name, "launches")) {
      if (actual) *actual = 1;
      if (h_dst && count) h_dst[0] = (float)last_launches;
      return ADN_OK;
    }
    auto it = dumps.find(name);
    if (it == dumps.end()) { err = std::string("adn_debug_read: unknown tensor '") + name + "' (stage dumps need adn_debug_stop_after(m, -1) before the run)"; return ADN_ERR_INVALID; }
    if (actual) *actual = it->second.size();
    if (!h_dst) return ADN_OK;
    const size_t nc = count < it->second.size() ? count : it->second.size();
    memcpy(h_dst, it->second.data(), nc * sizeof(float));
    return ADN_OK;
  }
};

}  // namespace dfs

ModelImpl* dfsmn_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  dfs::Model* m = new dfs::Model();
  m->device = device;
  m->sms = sms;
  m->d_blob = d_blob;
  m->index = index;
  if (!m->init(meta, h_blob)) {
    err = m->err;
    delete m;
    return nullptr;
  }
  return m;
}
