// DFSMN (48 kHz, causal) -- SURVEY 8f rank 3: the operators of `DFSMN.forward` (reference DFSMN/Export_DFSMN.py:191-250)
// as functors + the launch sequence over them, templated on the executor exactly like csrc/mfgan_ops.cuh: libadn runs it
// with the CUDA executor (csrc/dfsmn.cu), tests/harness/dfsmn_host.cpp with a host loop (CPU check against the oracle).
//   fused analysis conv [Kaldi fbank (2 x 1025 rows) | mask STFT (2 x 961 rows)], frame 1920, hop 960 (:132-140, :212)
//   -> power x 32768^2 -> mel (120) -> log floor (:220-221) -> linear1 + ReLU -> layers x UniDeepFsmn (affine + ReLU,
//   projection, causal depthwise memory with the inner residual folded into the last tap, outer residual) -> linear2 +
//   sigmoid (:228-234) -> mask x packed spectrum (:239-240); the ISTFT and the output rule live in csrc/dfsmn.cu.
#pragma once
#include "mfgan_gemm.cuh"

namespace dfs {

using gan::GemmOp;
using gan::Linear;

constexpr int FRAME = 1920, HOP = 960, KB = 1025, SB = 961, NM = 120, H = 256;
constexpr int AN = 2 * KB + 2 * SB;   // 3972 analysis rows

template <class TIn>
struct Prep {                          // cast (+ 1/32768 for int16 input, :193-197)
  const TIn* in; float scale; float* out;
  GAN_HD void operator()(long long i) const { out[i] = (float)in[i] * scale; }
};
struct Power {                         // (re^2 + im^2) x 32768^2 of the Kaldi rows (:220)
  const float* an; float* pw; float scale;
  GAN_HD void operator()(long long i) const {
    const int f = (int)(i % KB); const long long r = i / KB;
    const float re = an[r * AN + f], im = an[r * AN + KB + f];
    pw[i] = (re * re + im * im) * scale;
  }
};
struct DwCausal {                      // h + causal depthwise memory over frames (left zero pad lorder - 1, :231-232)
  const float* src; const float* taps; const float* res; float* dst; int T, k;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % H); const long long r = i / H; const int t = (int)(r % T); const long long n = r / T;
    float m = 0.f;
    for (int j = 0; j < k; ++j) {
      const int tj = t + j - (k - 1);
      if (tj >= 0) m += taps[j * H + c] * src[(n * T + tj) * H + c];
    }
    dst[i] = res[i] + m;
  }
};
struct MaskApply {                     // packed [re ; im] spectrum x [mask ; mask], frame-minor for the ISTFT (:239-240)
  const float* an; const float* mask; float* spec; int T;
  GAN_HD void operator()(long long i) const {
    const int t = (int)(i % T); const long long q = i / T; const int r = (int)(q % (2 * SB)); const long long b = q / (2 * SB);
    const long long row = b * T + t;
    spec[i] = an[row * AN + 2 * KB + r] * mask[row * SB + (r % SB)];
  }
};
struct OutI16 {                        // x 32768, clamp, truncate (:246-248)
  const float* w; int16_t* out;
  GAN_HD void operator()(long long i) const {
    float v = w[i] * 32768.0f;
    v = v < -32768.0f ? -32768.0f : (v > 32767.0f ? 32767.0f : v);
    out[i] = (int16_t)(int)v;
  }
};

struct LayerW { const float *lin_w, *lin_b, *proj_w, *conv_w; };
struct Weights {
  const float *analysis_w, *mel_t, *lin1_w, *lin1_b, *lin2_w, *lin2_b;
  int layers, lorder;
  LayerW uf[16];
};
template <class Lookup>
bool bind(Weights& W, int layers, int lorder, Lookup& lk) {
  if (layers < 1 || layers > 16 || lorder < 1 || lorder > 64) return false;
  bool ok = true;
  char nm[64];
  auto g = [&](const char* name, size_t n) { const float* p = lk(name, n); ok = ok && p; return p; };
  W.layers = layers; W.lorder = lorder;
  W.analysis_w = g("analysis_w", (size_t)AN * FRAME); W.mel_t = g("mel_t", (size_t)KB * NM);
  W.lin1_w = g("lin1_w", NM * H); W.lin1_b = g("lin1_b", H); W.lin2_w = g("lin2_w", H * SB); W.lin2_b = g("lin2_b", SB);
  for (int i = 0; i < layers; ++i) {
    auto gl = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "uf%d.%s", i, k); return g(nm, n); };
    W.uf[i].lin_w = gl("lin_w", H * H); W.uf[i].lin_b = gl("lin_b", H); W.uf[i].proj_w = gl("proj_w", H * H);
    W.uf[i].conv_w = gl("conv_w", (size_t)lorder * H);
  }
  return ok;
}

struct Workspace { float *an, *pw, *feat, *h0, *h1, *f1, *p1, *mask; };
template <class Alloc>
bool alloc_ws(Workspace& w, int B, int T, Alloc& alloc) {
  const long long rows = (long long)B * T;
  bool ok = true;
  auto a = [&](float*& p, long long n) { p = alloc((size_t)n); ok = ok && p; };
  a(w.an, rows * AN); a(w.pw, rows * KB); a(w.feat, rows * NM); a(w.h0, rows * H); a(w.h1, rows * H); a(w.f1, rows * H);
  a(w.p1, rows * H); a(w.mask, rows * SB);
  return ok;
}

// x (B, L) fp32 (int16 input already x 1/32768) -> spec (B, 2 * 961, T) masked packed spectrum.
// `ex.run(count, functor)`, `ex.gemm(GemmOp)`, `ex.mark(tag, name, ptr, count)` as in mfgan_ops.cuh.
template <class Exec>
void forward(Exec& ex, const Workspace& w, const Weights& W, const float* x, float* spec, int B, int L, int T) {
  const long long rows = (long long)B * T;
  GemmOp g = gan::gemm_blank();                       // framing as a row-overlapping A operand: frame t starts at t * hop
  g.batch = B;
  g.A = x; g.a_b1 = L; g.a_m = HOP; g.a_k = 1;
  g.B = W.analysis_w; g.b_k = 1; g.b_n = FRAME;
  g.C = w.an; g.c_b1 = (long long)T * AN; g.c_m = AN; g.c_n = 1;
  g.M = T; g.N = AN; g.K = FRAME;
  ex.gemm(g);
  ex.run(rows * KB, Power{w.an, w.pw, 32768.0f * 32768.0f});
  ex.run(rows * NM, Linear{w.pw, KB, nullptr, W.mel_t, nullptr, w.feat, NM, KB, NM, gan::ACT_LOGCLAMP, nullptr});
  ex.mark("", "feat", w.feat, rows * NM);
  ex.run(rows * H, Linear{w.feat, NM, nullptr, W.lin1_w, W.lin1_b, w.h0, H, NM, H, gan::ACT_RELU, nullptr});
  ex.mark("", "lin1", w.h0, rows * H);
  float *h = w.h0, *hn = w.h1;
  char tag[16];
  for (int i = 0; i < W.layers; ++i) {
    ex.run(rows * H, Linear{h, H, nullptr, W.uf[i].lin_w, W.uf[i].lin_b, w.f1, H, H, H, gan::ACT_RELU, nullptr});
    ex.run(rows * H, Linear{w.f1, H, nullptr, W.uf[i].proj_w, nullptr, w.p1, H, H, H, gan::ACT_NONE, nullptr});
    ex.run(rows * H, DwCausal{w.p1, W.uf[i].conv_w, h, hn, T, W.lorder});
    float* t = h; h = hn; hn = t;
    snprintf(tag, sizeof(tag), "uf%d", i);
    ex.mark("", tag, h, rows * H);
  }
  ex.run(rows * SB, Linear{h, H, nullptr, W.lin2_w, W.lin2_b, w.mask, SB, H, SB, gan::ACT_SIGMOID, nullptr});
  ex.mark("", "mask", w.mask, rows * SB);
  ex.run((long long)B * 2 * SB * T, MaskApply{w.an, w.mask, spec, T});
}

}  // namespace dfs
