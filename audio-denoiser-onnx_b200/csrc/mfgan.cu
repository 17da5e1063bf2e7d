// MossFormerGAN-SE-16K (the enhancement half of BASELINE configs[4]) behind the C ABI.
// Reference: MossFormerGAN_SE_16K/Export_MossFormer_SE.py `MOSSFORMER_SE.forward` (:532-897).
//   RMS norm + wrap pad (:564-568) -> STFT 400/100 hamming (:570) -> power-law features (:578-586)      [ends.cu operators]
//   -> dense encoder, 6 x (intra path, inter path, triple attention), mask / complex decoders           [mfgan_ops.cuh]
//   -> mask * compressed + complex, decompress (:863-868) -> ISTFT -> x norm factor, output rule (:880-897) [ends.cu]
// First-correct implementation: one grid per operator (fp32 FFMA, one output per thread), all intermediates in HBM.
#include "gan_exec.cuh"

#include "common.cuh"
#include "dfsmn_ops.cuh"
#include "model_impl.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace gan {

struct CastF16 {
  const float* w; __half* out;
  __device__ void operator()(long long i) const { out[i] = __float2half_rn(w[i]); }
};

class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, Lpad = 0, T = 0, layers = 6, Lsrc = 0;
  int L_in = 0, L_final = 0;     // I/O window lengths at the in / out sample rates (== L at 16 kHz)
  bool rs_in = false, rs_out = false;
  float *xr = nullptr, *ymod = nullptr, *yout = nullptr;
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;
  Weights W;
  Workspace ws;
  adn_stft* stft = nullptr;
  std::vector<void*> allocs;
  int cap = 0;                  // windows the workspace holds
  float *xn = nullptr, *nf = nullptr, *spec = nullptr, *feat = nullptr, *keep = nullptr, *mask = nullptr, *cplx = nullptr,
        *spec2 = nullptr, *wave = nullptr;
  int stop_after = 0, last_launches = 0, last_batch = 0;
  std::map<std::string, std::vector<float>> dumps;
  TcCache tcc;                     // tcgen05 plans + weight operand planes of the Linear / Conv2d operators (gan_exec.cuh)
  static constexpr int SUB = 16;   // windows per pass of the backbone (bounds the workspace)

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
    tcc.plans.clear();             // the plans hold workspace addresses
    cap = 0;
  }
  float* dalloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) { err = "mossformergan_se: out of device memory for the workspace"; return nullptr; }
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
    if (!geti("input_audio_length", L) || !geti("nfft", nfft) || !geti("hop_length", hop) || !geti("gan_layers", layers) ||
        !gets("input_audio_dtype", sin) || !gets("output_audio_dtype", sout))
      return false;
    if (nfft != 400 || hop != 100 || L < 1 || layers < 1 || layers > 8) {
      err = "mossformergan_se needs nfft=400, hop_length=100, input_audio_length >= 400, 1..8 layers";
      return false;
    }
    {   // optional linear resampling either side of the model (:542-549, :884-891): input_audio_length is at in_sample_rate
      int in_sr = 16000, out_sr = 16000, model_sr = 16000;
      auto opt = [&](const char* k, int& v) { auto it = meta.find(k); if (it != meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
      opt("in_sample_rate", in_sr); opt("out_sample_rate", out_sr); opt("model_sample_rate", model_sr);
      if (model_sr != 16000 || in_sr <= 0 || out_sr <= 0) { err = "mossformergan_se runs at model_sample_rate 16000"; return false; }
      L_in = L;
      rs_in = in_sr != 16000;
      rs_out = out_sr != 16000;
      if (rs_in) L = (int)llround((double)L_in * 16000.0 / in_sr);                 // MODEL_AUDIO_LENGTH (:37)
      L_final = rs_out ? (int)llround((double)L_in * out_sr / 16000.0) : L;        // OUTPUT_AUDIO_LENGTH (:38): input length x out / model rate
      if (L < 400 || L_final < 1) { err = "mossformergan_se: the model-rate window must cover one STFT frame (400 samples)"; return false; }
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    Lpad = L + (hop - L % hop) % hop;             // wrap-around pad to a hop multiple (:566-568)
    T = Lpad / hop + 1;
    Lsrc = hop * (T - 1);                         // centred ISTFT length
    auto lk = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || (expect && it->second.count != expect)) {
        if (err.empty()) err = std::string("weight blob: tensor '") + name + "' missing or wrong size";
        return nullptr;
      }
      return d_blob + it->second.offset;
    };
    err.clear();
    if (!bind(W, layers, T, lk)) { if (err.empty()) err = "mossformergan_se: weight binding failed"; return false; }
    auto host = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || it->second.count != expect) { err = std::string("weight blob: tensor '") + name + "' missing or wrong size"; return nullptr; }
      return h_blob + it->second.offset;
    };
    const float* fwd = host("stft.fwd", (size_t)2 * FB * 400);
    const float* inv = host("stft.inv", (size_t)2 * FB * 400);
    const float* nrm = host("stft.norm", (size_t)Lsrc);
    if (!fwd || !inv || !nrm) return false;
    adn_stft_geom g;
    memset(&g, 0, sizeof(g));
    g.nfft = 400; g.hop = 100; g.center = 1; g.pad_reflect = 1; g.norm_multiply = 0;
    if (adn_stft_create(&stft, &g, fwd, inv, nrm, T, device) != ADN_OK) { err = std::string("mossformergan_se: ") + adn_last_error(nullptr); return false; }
    return true;
  }
  bool ensure(int B) {
    const int sub = B < SUB ? B : SUB;
    if (B <= cap) return true;
    cudaDeviceSynchronize();
    free_ws();
    auto a = [&](size_t n) { return dalloc(n); };
    if (!alloc_ws(ws, sub, T, a)) return false;
    const size_t b = (size_t)B;
    if (!(xn = dalloc(b * Lpad)) || !(nf = dalloc(b)) || !(spec = dalloc(b * 2 * FB * T)) || !(feat = dalloc(b * 3 * T * FB)) ||
        !(keep = dalloc(b * 2 * FB * T)) || !(mask = dalloc(b * FB * T)) || !(cplx = dalloc(b * 2 * FB * T)) ||
        !(spec2 = dalloc(b * 2 * FB * T)) || !(wave = dalloc(b * Lsrc)))
      return false;
    if ((rs_in && !(xr = dalloc(b * L))) || (rs_out && (!(ymod = dalloc(b * L)) || !(yout = dalloc(b * L_final))))) return false;
    cap = B;
    return true;
  }
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);
    in->dtype = in_dtype; in->channels = 1; in->length = L_in;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);
    out->dtype = out_dtype; out->channels = 1; out->length = L_final;
  }
  size_t workspace_bytes(int batch) override {
    const size_t sub = batch < SUB ? batch : SUB, S = T > FQ ? T : FQ, rows = sub * T * FQ, px2 = sub * T * (FB + 1);
    size_t f = px2 * (SKIPC + 3 * C + 1) + rows * (3 * C + 2 + PI + 2 * UV + 3 * UV + 2 * C + 2 + HUV + 4 * QK + 2 * S + HID + HID / 2 + C + QKV + 2 * C) +
               sub * S * QK * HID + sub * HEADS * T * T;
    f += (size_t)batch * (Lpad + 1 + 9 * FB * T + Lsrc);
    if (rs_in) f += (size_t)batch * L;
    if (rs_out) f += (size_t)batch * ((size_t)L + L_final);
    return f * sizeof(float);
  }
  int launches(int batch) override {
    const int per_dense = DEPTH * 7, per_path = 28, per_ta = 11;   // Att is three GEMM launches
    const int bb = 8 + per_dense + layers * (2 * per_path + per_ta) + 2 * (6 + per_dense);
    return 6 + (rs_in ? 1 : 0) + (rs_out ? (out_dtype == ADN_F32 ? 1 : 2) : 0) + ((batch + SUB - 1) / SUB) * bb;
  }
  void set_stop_after(int n) override { stop_after = n; }

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    auto chk = [&](adn_status s, const char* what) {
      if (s != ADN_OK) err = std::string("mossformergan_se ") + what + ": " + adn_last_error(nullptr);
      return s == ADN_OK;
    };
    auto tk = [&](const char* name) { if (tick) tick(tick_ctx, name); };
    // F32 / F16 inputs are in [-1, 1]: x32768 (:540-541).  With a resampled input the reference lifts, then interpolates
    // (:542-549); the lift is a power of two, so interpolating the raw samples and lifting inside the RMS pass is bit-identical.
    const void* src = d_in;
    int src_dtype = in_dtype;
    int extra = 0;
    if (rs_in) {
      if (!chk(adn_resample_linear(d_in, in_dtype, xr, B, L_in, L, 0.0, st), "input resampler")) return ADN_ERR_CUDA;
      tk("resample_in");
      src = xr; src_dtype = ADN_F32; ++extra;
    }
    if (!chk(adn_rms_normalize(src, src_dtype, in_dtype == ADN_I16 ? 1.0f : 32768.0f, 1e-6f, xn, nf, B, L, Lpad, st), "rms_normalize")) return ADN_ERR_CUDA;
    tk("rms_normalize");
    if (!chk(adn_stft_forward(stft, xn, spec, B, Lpad, st), "stft")) return ADN_ERR_CUDA;
    tk("stft");
    if (!chk(adn_spec_features(ADN_FAMILY_MOSSFORMERGAN, spec, feat, keep, B, FB, T, st), "spec_features")) return ADN_ERR_CUDA;
    tk("spec_features");
    CudaExec ex;
    ex.st = st; ex.tick = tick; ex.tick_ctx = tick_ctx;
    ex.capture = stop_after != 0; ex.dumps = &dumps;
    if (ex.capture) dumps.clear();
    tcc.sms = sms;
    { const char* e = getenv("ADN_GAN_TC"); tcc.enabled = !(e && e[0] == '0'); }     // ADN_GAN_TC=0: the round-1 FFMA tiles
    ex.tc = &tcc;
    for (int b0 = 0; b0 < B; b0 += SUB) {
      const int nb = B - b0 < SUB ? B - b0 : SUB;
      ex.tc_pass = nb; ex.tc_idx = 0;
      forward(ex, ws, W, feat + (size_t)b0 * 3 * T * FB, mask + (size_t)b0 * FB * T, cplx + (size_t)b0 * 2 * FB * T, nb, T);
      ex.capture = false;                          // stage dumps cover the first pass only
    }
    last_launches = ex.launches + 6 + extra;
    if (!chk(adn_spec_recombine(ADN_FAMILY_MOSSFORMERGAN, mask, cplx, keep, spec2, B, FB, T, st), "spec_recombine")) return ADN_ERR_CUDA;
    tk("spec_recombine");
    if (!chk(adn_stft_inverse(stft, spec2, wave, B, T, st), "istft")) return ADN_ERR_CUDA;
    tk("istft");
    if (!rs_out) {
      if (!chk(adn_condition_output(ADN_FAMILY_MOSSFORMERGAN, wave, Lsrc, nf, 1, d_out, out_dtype, B, L, st), "condition_output")) return ADN_ERR_CUDA;
      tk("condition_output");
    } else {
      // x norm_factor -> interpolate -> output rule (:877-897).  The F32 rule of adn_condition_output (x nf x 2^-15) commutes
      // exactly with the interpolation (power-of-two scale), so: condition to fp32, resample, then restore x 2^15 / clamp /
      // truncate for int16.
      if (!chk(adn_condition_output(ADN_FAMILY_MOSSFORMERGAN, wave, Lsrc, nf, 1, ymod, ADN_F32, B, L, st), "condition_output")) return ADN_ERR_CUDA;
      tk("condition_output");
      float* dst = out_dtype == ADN_F32 ? (float*)d_out : yout;
      if (!chk(adn_resample_linear(ymod, ADN_F32, dst, B, L, L_final, 0.0, st), "output resampler")) return ADN_ERR_CUDA;
      tk("resample_out");
      const long long no = (long long)B * L_final;
      if (out_dtype == ADN_I16) ex.run(no, dfs::OutI16{yout, (int16_t*)d_out});
      else if (out_dtype == ADN_F16) ex.run(no, CastF16{yout, (__half*)d_out});
      last_launches += out_dtype == ADN_F32 ? 1 : 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mossformergan_se run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
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

}  // namespace gan

ModelImpl* mfgan_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  gan::Model* m = new gan::Model();
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
