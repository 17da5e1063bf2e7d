// UL-UNAS 16 kHz denoiser behind the C ABI: model family "ulunas".
// Reference: UL-UNAS/Export_UL_UNAS.py `ULUNAS_CUSTOM.forward` (:849-912).  The launch sequence is csrc/ulunas_ops.cuh (functors
// shared with the CPU host harness); here: input cast, STFT 512 / 256 hann with the int16 scale folded into the basis, the sequence
// on the CUDA executor, ISTFT x (PCM scale / window sum) and the output rule.
#include "ulunas_ops.cuh"

#include "common.cuh"
#include "gan_exec.cuh"

#include <stdio.h>
#include <string.h>

namespace gan {
GAN_OP_NAME(uln::PowerLogErb, "ulunas_log_erb");
GAN_OP_NAME(uln::ConvG, "ulunas_conv");
GAN_OP_NAME(uln::DeconvG, "ulunas_deconv");
GAN_OP_NAME(uln::AffAct, "ulunas_act");
GAN_OP_NAME(uln::Shuffle, "ulunas_shuffle");
GAN_OP_NAME(uln::Add2, "ulunas_add");
GAN_OP_NAME(uln::MeanF2, "ulunas_mean_f");
GAN_OP_NAME(uln::MeanC2, "ulunas_mean_c");
GAN_OP_NAME(uln::CtfaApply, "ulunas_ctfa_apply");
GAN_OP_NAME(uln::GruSeq, "ulunas_gru");
GAN_OP_NAME(uln::ToBTFC, "ulunas_to_btfc");
GAN_OP_NAME(uln::ToNCHW, "ulunas_to_nchw");
GAN_OP_NAME(uln::LnStats, "ulunas_ln_stats");
GAN_OP_NAME(uln::LnApplyRes, "ulunas_ln_apply");
GAN_OP_NAME(uln::ErbMask, "ulunas_erb_mask");
GAN_OP_NAME(uln::OutRule, "ulunas_out");
}  // namespace gan

namespace uln {

struct CastF16 {
  const float* w; __half* out;
  __device__ void operator()(long long i) const { out[i] = __float2half_rn(w[i]); }
};
struct CountExec {                      // dry run of the sequence: how many launches it makes
  int launches = 0;
  template <class F> void run(long long, const F&) { ++launches; }
  void mark(const char*, const char*, const float*, long long) {}
};

// nn.GRU with a hidden state wide enough to share (the cTFA time-attention GRUs: hidden 2C = 24 ... 64 over the 63 frames, ONE
// sequence per window): one CTA = one (sequence, group, direction), thread j = hidden unit j, both weight matrices transposed in
// shared memory ([input][3H]: thread j reads consecutive words), h double-buffered in shared memory, two barriers per step.
// Summation order = the GruSeq functor's (bias first, inputs in order), so the two paths give identical bits.
__global__ void gru_coop_kernel(GruSeq f) {
  extern __shared__ float sh[];
  const int H = f.H, I = f.I, H3 = 3 * f.H;
  float* wiT = sh;                  // [I][3H]
  float* whT = wiT + I * H3;        // [H][3H]
  float* hs = whT + H * H3;         // [2][H]
  float* xv = hs + 2 * H;           // [I]
  const long long idx = blockIdx.x;
  const int d = (int)(idx % f.dirs);
  long long r = idx / f.dirs;
  const int g = (int)(r % f.ngroups);
  const long long n = r / f.ngroups;
  const long long gd = (long long)g * f.dirs + d;
  const float* wi = f.w_ih + gd * H3 * I;
  const float* wh = f.w_hh + gd * H3 * H;
  for (int e = threadIdx.x; e < H3 * I; e += blockDim.x) wiT[(e % I) * H3 + e / I] = __ldg(wi + e);
  for (int e = threadIdx.x; e < H3 * H; e += blockDim.x) whT[(e % H) * H3 + e / H] = __ldg(wh + e);
  const int j = threadIdx.x;
  float bir = 0.f, biz = 0.f, bin = 0.f, bhr = 0.f, bhz = 0.f, bhn = 0.f;
  if (j < H) {
    const float* bi = f.b_ih + gd * H3;
    const float* bh = f.b_hh + gd * H3;
    bir = bi[j]; biz = bi[H + j]; bin = bi[2 * H + j];
    bhr = bh[j]; bhz = bh[H + j]; bhn = bh[2 * H + j];
    hs[j] = 0.f;
  }
  const float* xs = f.x + (n / f.n2) * f.x_sA + (n % f.n2) * f.x_sB + (long long)g * f.x_grp;
  float* os = f.out + (n / f.n2) * f.o_sA + (n % f.n2) * f.o_sB + (long long)g * f.o_grp + (long long)d * H;
  int cur = 0;
  for (int k = 0; k < f.steps; ++k) {
    const int s = d ? f.steps - 1 - k : k;
    if (j < I) xv[j] = (f.x_limit > 0 && s * I + j >= f.x_limit) ? 0.f : xs[(long long)s * f.x_step + j];
    __syncthreads();
    if (j < H) {
      float ir = bir, iz = biz, in_ = bin;
      for (int q = 0; q < I; ++q) {
        const float v = xv[q];
        ir += wiT[q * H3 + j] * v; iz += wiT[q * H3 + H + j] * v; in_ += wiT[q * H3 + 2 * H + j] * v;
      }
      float hr = bhr, hz = bhz, hh = bhn;
      const float* hc = hs + cur * H;
      for (int q = 0; q < H; ++q) {
        const float v = hc[q];
        hr += whT[q * H3 + j] * v; hz += whT[q * H3 + H + j] * v; hh += whT[q * H3 + 2 * H + j] * v;
      }
      const float rg = gan::sigmoidf_(ir + hr), zg = gan::sigmoidf_(iz + hz);
      const float ng = tanhf(in_ + rg * hh);
      const float hn = (1.f - zg) * ng + zg * hc[j];
      hs[(cur ^ 1) * H + j] = hn;
      os[(long long)s * f.o_step + j] = hn;
    }
    __syncthreads();
    cur ^= 1;
  }
}

// the functor executor + the cooperative GRU for wide hidden states
struct UlExec : gan::CudaExec {
  using gan::CudaExec::run;
  void run(long long n, const GruSeq& f) {
    if (f.H < 16 || n <= 0) { gan::CudaExec::run<GruSeq>(n, f); return; }
    const int threads = ((f.H > f.I ? f.H : f.I) + 31) / 32 * 32;
    const size_t smem = ((size_t)(f.I + f.H) * 3 * f.H + 2 * f.H + f.I) * sizeof(float);
    static unsigned long long configured = 0;
    if (adn_first_use_on_device(configured))
      cudaFuncSetAttribute(gru_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)(GRU_HMAX + GRU_HMAX) * 3 * GRU_HMAX + 3 * GRU_HMAX) * sizeof(float)));
    gru_coop_kernel<<<(unsigned)n, threads, smem, st>>>(f);
    ++launches;
    if (tick) tick(tick_ctx, "ulunas_gru");
  }
};

class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, T = 0, Lout = 0, seq_launches = 0;
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;
  Weights W;
  Workspace ws;
  adn_stft* stft = nullptr;
  std::vector<void*> allocs;
  int cap = 0;
  float *x = nullptr, *spec = nullptr, *spec2 = nullptr, *wave = nullptr, *wave2 = nullptr;
  int stop_after = 0, last_launches = 0, last_batch = 0;
  std::map<std::string, std::vector<float>> dumps;

  ~Model() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    if (stft) adn_stft_destroy(stft);
  }
  void free_ws() {
    adn_note_free();
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    cap = 0;
  }
  float* dalloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) { err = "ulunas: out of device memory for the workspace"; return nullptr; }
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
    int nfft = 0, hop = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", L) || !geti("nfft", nfft) || !geti("hop_length", hop) || !gets("input_audio_dtype", sin) ||
        !gets("output_audio_dtype", sout))
      return false;
    if (nfft != NFFT || hop != HOP || L < NFFT) { err = "ulunas needs nfft=512, hop_length=256, input_audio_length >= 512"; return false; }
    {
      int in_sr = 16000, out_sr = 16000;
      auto opt = [&](const char* k, int& v) { auto it = meta.find(k); if (it != meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
      opt("in_sample_rate", in_sr); opt("out_sample_rate", out_sr);
      if (in_sr != 16000 || out_sr != 16000) { err = "ulunas runs at 16 kHz I/O only"; return false; }
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = L / HOP + 1;
    Lout = HOP * (T - 1);
    err.clear();
    auto lk = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || (expect && it->second.count != expect)) {
        if (err.empty()) err = std::string("weight blob: tensor '") + name + "' missing or wrong size";
        return nullptr;
      }
      return d_blob + it->second.offset;
    };
    if (!bind(W, lk)) { if (err.empty()) err = "ulunas: weight binding failed"; return false; }
    auto host = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || it->second.count != expect) { err = std::string("weight blob: tensor '") + name + "' missing or wrong size"; return nullptr; }
      return h_blob + it->second.offset;
    };
    const float* fwd = host("stft.fwd", (size_t)2 * NB * NFFT);
    const float* inv = host("stft.inv", (size_t)2 * NB * NFFT);
    const float* nrm = host("stft.norm", (size_t)Lout);
    if (!fwd || !inv || !nrm) return false;
    adn_stft_geom g;
    memset(&g, 0, sizeof(g));
    g.nfft = NFFT; g.hop = HOP; g.center = 1; g.pad_reflect = 1; g.norm_multiply = 1;    // x (scale / window sum), STFT_Process.py:264
    if (adn_stft_create(&stft, &g, fwd, inv, nrm, T, device) != ADN_OK) { err = std::string("ulunas: ") + adn_last_error(nullptr); return false; }
    CountExec cnt;                                  // dry run over distinct, never dereferenced addresses
    Workspace fake;
    uintptr_t cur = 4096;
    auto fa = [&](size_t n) { float* p = reinterpret_cast<float*>(cur); cur += n * sizeof(float) + 256; return p; };
    alloc_ws(fake, 1, T, fa);
    forward(cnt, fake, W, nullptr, nullptr, 1, T);
    seq_launches = cnt.launches;
    return true;
  }
  bool ensure(int B) {
    if (B <= cap) return true;
    cudaDeviceSynchronize();
    free_ws();
    auto a = [&](size_t n) { return dalloc(n); };
    if (!alloc_ws(ws, B, T, a)) return false;
    const size_t b = (size_t)B;
    if (!(x = dalloc(b * L)) || !(spec = dalloc(b * 2 * NB * T)) || !(spec2 = dalloc(b * 2 * NB * T)) || !(wave = dalloc(b * Lout)) ||
        !(wave2 = dalloc(b * Lout)))
      return false;
    cap = B;
    return true;
  }
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);
    in->dtype = in_dtype; in->channels = 1; in->length = L;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);
    out->dtype = out_dtype; out->channels = 1; out->length = Lout;
  }
  size_t workspace_bytes(int batch) override {
    const size_t b = (size_t)batch, map = b * CMAX * T * WMAX;
    return (9 * map + b * T * (4 * CMAX + 4 * (WMAX + 3) + 2) + b * ((size_t)L + 4 * NB * T + 2 * Lout)) * sizeof(float);
  }
  int launches(int) override { return 1 + 2 + seq_launches + 2 + (out_dtype == ADN_F16 ? 2 : 1); }
  void set_stop_after(int n) override { stop_after = n; }

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    UlExec ex;
    ex.st = st; ex.tick = tick; ex.tick_ctx = tick_ctx;
    ex.capture = stop_after != 0; ex.dumps = &dumps;
    if (ex.capture) dumps.clear();
    const long long n = (long long)B * L, no = (long long)B * Lout;
    if (in_dtype == ADN_I16) ex.run(n, Cast<int16_t>{(const int16_t*)d_in, x});      // the 1/32768 sits in the STFT basis
    else if (in_dtype == ADN_F16) ex.run(n, Cast<__half>{(const __half*)d_in, x});
    else ex.run(n, Cast<float>{(const float*)d_in, x});
    if (adn_stft_forward(stft, x, spec, B, L, st) != ADN_OK) { err = std::string("ulunas stft: ") + adn_last_error(nullptr); return ADN_ERR_CUDA; }
    ex.launches += 2;
    if (tick) tick(tick_ctx, "stft");
    forward(ex, ws, W, spec, spec2, B, T);
    if (adn_stft_inverse(stft, spec2, wave, B, T, st) != ADN_OK) { err = std::string("ulunas istft: ") + adn_last_error(nullptr); return ADN_ERR_CUDA; }
    ex.launches += 2;
    if (tick) tick(tick_ctx, "istft");
    const int fix = in_dtype == ADN_I16 ? 0 : 1;                                         // nan_to_num for float inputs (:906-907)
    if (out_dtype == ADN_F16) {
      ex.run(no, OutRule{wave, wave2, fix, 0});
      ex.run(no, CastF16{wave2, (__half*)d_out});
    } else {
      ex.run(no, OutRule{wave, d_out, fix, out_dtype == ADN_I16 ? 1 : 0});
    }
    last_launches = ex.launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("ulunas run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
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

}  // namespace uln

ModelImpl* ulunas_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                         const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  uln::Model* m = new uln::Model();
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
