// MossFormerGAN-SE-16K backbone: the operators of `MOSSFORMER_SE.forward` /
// `_mossformer_block` (reference MossFormerGAN_SE_16K/Export_MossFormer_SE.py:137-244, :588-868) as
// one-output-per-thread functors, and the launch sequence over them (`gan::forward`), templated on the
// executor.  libadn instantiates it with the CUDA executor only (csrc/mfgan.cu: one grid per functor);
// tests/harness/mfgan_host.cu instantiates the SAME sequence with a host loop so the index arithmetic of
// every functor is checked against the oracle without a GPU.  First-correct design: fp32 FFMA, every
// intermediate materialised in HBM, feature map channel-last (window, frame, sub-band, 64).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDACC__)
#define GAN_HD __host__ __device__ __forceinline__
#else
#define GAN_HD inline
#endif

namespace gan {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2, ACT_PRELU = 3, ACT_LOGCLAMP = 4, ACT_SIGMOID = 5 };

constexpr int C = 64;        // emb
constexpr int KS = 2;        // emb_ks
constexpr int PI = 128;      // emb * emb_ks
constexpr int UV = 128;      // to_u / to_v width
constexpr int HID = 256;     // MossFormer to_hidden width ([v | u])
constexpr int QK = 128;
constexpr int HUV = HID + QK;
constexpr int ROT = 32;
constexpr int DW = 31;       // FFConvM depthwise kernel
constexpr int LORDER = 20;   // path UniDeepFsmn
constexpr int DLORDER = 5;   // dense-block FSMN
constexpr int HEADS = 4, AE = 6, VC = 16, QKV = 2 * HEADS * AE + C;   // 112
constexpr int FQ = 101, FB = 201;
constexpr int DEPTH = 4;
constexpr int SKIPC = C * (DEPTH + 1);
constexpr float EPS = 1e-5f;

GAN_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// a / d and a % d for a >= 0, d > 0: 32-bit arithmetic whenever a fits (a 64-bit division by a run-time divisor is ~100 GPU
// instructions; the element-wise functors do two to four per output, which made them instruction-bound, not memory-bound)
GAN_HD long long idiv(long long a, int d) {
  return (unsigned long long)a < 0x100000000ull ? (long long)((unsigned)a / (unsigned)d) : a / d;
}
GAN_HD long long idiv(long long a, long long d) {
  return (unsigned long long)(a | d) < 0x100000000ull ? (long long)((unsigned)a / (unsigned)d) : a / d;
}
GAN_HD int imod(long long a, int d) {
  return (unsigned long long)a < 0x100000000ull ? (int)((unsigned)a % (unsigned)d) : (int)(a % d);
}
GAN_HD float act(float v, int a, const float* slope, int n) {
  if (a == ACT_RELU) return v > 0.f ? v : 0.f;
  if (a == ACT_SILU) return v * sigmoidf_(v);
  if (a == ACT_PRELU) return v >= 0.f ? v : slope[n] * v;
  if (a == ACT_LOGCLAMP) return logf(v > 1.1920928955078125e-07f ? v : 1.1920928955078125e-07f);   // log(max(v, fp32 eps)), Kaldi log floor
  if (a == ACT_SIGMOID) return sigmoidf_(v);
  return v;
}

// sequence (n, s) -> pixel index of the channel-last feature map
struct PixMap {
  int n2;
  long long sA, sB, sS;
  GAN_HD long long pix(long long n, int s) const { return idiv(n, n2) * sA + imod(n, n2) * sB + (long long)s * sS; }
};

// out[r, n] = act(bias[n] + sum_k xin(r, k) * Wt[k, n]);  xin normalised by the row's (mean, rstd) when stat != null
struct Linear {
  const float* in; int ldi; const float* stat; const float* Wt; const float* bias; float* out; int ldo; int K, N, a;
  const float* slope;
  GAN_HD void operator()(long long i) const {
    const long long r = i / N; const int n = (int)(i % N);
    const float* x = in + r * ldi;
    float acc = 0.f;
    if (stat) {
      const float mu = stat[2 * r], rs = stat[2 * r + 1];
      for (int k = 0; k < K; ++k) acc += ((x[k] - mu) * rs) * Wt[(long long)k * N + n];
    } else {
      for (int k = 0; k < K; ++k) acc += x[k] * Wt[(long long)k * N + n];
    }
    if (bias) acc += bias[n];
    out[r * ldo + n] = act(acc, a, slope, n);
  }
};

// (mean, rstd) of each row over K values: LayerNorm without affine / the per-pixel channel norm (K % 4 == 0, 16-byte rows)
struct alignas(16) F4 { float x, y, z, w; };
struct RowStats {
  const float* in; int ldi; int K; float* stat;
  GAN_HD void operator()(long long r) const {
    const F4* x = reinterpret_cast<const F4*>(in + r * ldi);
    float s = 0.f;
    for (int k = 0; k < K / 4; ++k) { const F4 v = x[k]; s += v.x; s += v.y; s += v.z; s += v.w; }
    const float mu = s / (float)K;
    float v = 0.f;
    for (int k = 0; k < K / 4; ++k) {
      const F4 q = x[k];
      float d = q.x - mu; v += d * d;
      d = q.y - mu; v += d * d;
      d = q.z - mu; v += d * d;
      d = q.w - mu; v += d * d;
    }
    stat[2 * r] = mu;
    stat[2 * r + 1] = 1.0f / sqrtf(v / (float)K + EPS);
  }
};

// channel norm + grouped (1, ks) conv (intra, :639-641) / one-hot unfold (inter, :666-668): seq[n, s, ch*ks + j]
struct Gather {
  const float* x; const float* pst; PixMap pm; const float* w; const float* b; float* out; int S;
  GAN_HD void operator()(long long i) const {
    const int o = (int)(i % PI); const long long r = i / PI; const int s = imod(r, S); const long long n = idiv(r, S);
    const int ch = o / KS;
    float acc = b[o];
    for (int k = 0; k < KS; ++k) {
      const long long p = pm.pix(n, s + k);
      acc += w[o * KS + k] * ((x[p * C + ch] - pst[2 * p]) * pst[2 * p + 1]);
    }
    out[i] = acc;
  }
};

// dst[n, s, c] = res[n, s, c] + src[n, s, c] + sum_k taps[k, c] * src[n, s + k - padl, c]   (zero outside 0..S-1)
// One thread = one channel x DWS consecutive positions: the KT taps and the DWS + KT - 1 inputs sit in registers
// (KT * DWS FMAs for DWS + 2 KT - 1 loads); adjacent threads = adjacent channels (count = dw_count(N, S, Cn)).
constexpr int DWS = 8;
GAN_HD long long dw_count(long long N, int S, int Cn) { return N * ((S + DWS - 1) / DWS) * Cn; }
template <int KT>
struct DwConv {
  const float* src; int lds; const float* res; int ldr; const float* taps; float* dst; int ldd; int use_pm;
  PixMap pm; int S, Cn;
  GAN_HD void operator()(long long i) const {
    constexpr int padl = (KT - 1) / 2;             // 'same' padding of every depthwise conv of the model
    const int c = imod(i, Cn); const long long r = idiv(i, Cn);
    const int sg = (S + DWS - 1) / DWS; const int s0 = (int)(r % sg) * DWS; const long long n = r / sg;
    float t[KT], x[DWS + KT - 1];
#pragma unroll
    for (int j = 0; j < KT; ++j) t[j] = taps[j * Cn + c];
#pragma unroll
    for (int j = 0; j < DWS + KT - 1; ++j) {
      const int sj = s0 + j - padl;
      x[j] = (sj >= 0 && sj < S) ? src[(n * S + sj) * lds + c] : 0.f;
    }
#pragma unroll
    for (int o = 0; o < DWS; ++o) {
      const int s = s0 + o;
      if (s < S) {
        const long long row = n * S + s;
        float acc = x[o + padl];
        if (res) acc += res[row * ldr + c];
        float m = 0.f;
#pragma unroll
        for (int j = 0; j < KT; ++j) m += t[j] * x[o + j];
        const long long d = use_pm ? pm.pix(n, s) * ldd : row * ldd;
        dst[d + c] = acc + m;
      }
    }
  }
};

// gate iv * iu, then ConvTranspose1d(k = ks, stride 1) back to Q = S + ks - 1 positions (:660-661)
struct GateConvT {
  const float* iu; int ldu; const float* iv; int ldv; const float* w; const float* b; float* out; int S, Q;
  GAN_HD void operator()(long long i) const {
    const int co = (int)(i % C); const long long r = i / C; const int q = (int)(r % Q); const long long n = r / Q;
    float acc = b[co];
    for (int k = 0; k < KS; ++k) {
      const int s = q - k;
      if (s < 0 || s >= S) continue;
      const float* u = iu + (n * S + s) * ldu;
      const float* v = iv + (n * S + s) * ldv;
      const float* wk = w + (long long)k * UV * C;
      for (int ci = 0; ci < UV; ++ci) acc += (v[ci] * u[ci]) * wk[ci * C + co];
    }
    out[i] = acc;
  }
};

// token shift of the first half of the channels (:139-141)
struct Shift {
  const float* x; float* out; int Q;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long r = i / C; const int q = (int)(r % Q);
    out[i] = c < C / 2 ? (q > 0 ? x[i - C] : 0.f) : x[i];
  }
};

// OffsetScale (4 heads) + rotary on the first 32 dims (:150-159): heads[n, q, h, j]
struct OffsetRot {
  const float* huv; const float* gamma; const float* beta; const float* cs; const float* sn; float* heads; int Q;
  GAN_HD void operator()(long long i) const {
    const int j = (int)(i % QK); const int h = (int)((i / QK) % 4); const long long r = i / (4 * QK); const int q = imod(r, Q);
    const float* z = huv + r * HUV + HID;
    float v = z[j] * gamma[h * QK + j] + beta[h * QK + j];
    if (j < ROT) {
      const int p = j ^ 1;
      const float vp = z[p] * gamma[h * QK + p] + beta[h * QK + p];
      v = v * cs[q * ROT + j] + vp * sn[q * ROT + j];
    }
    heads[i] = v;
  }
};

// local quadratic attention weights relu(quad_q . quad_k)^2 within a sequence (:170, :176)
struct SimLocal {
  const float* heads; float* A; int Q;
  GAN_HD void operator()(long long i) const {
    const int q2 = (int)(i % Q); const long long r = i / Q; const int q = (int)(r % Q); const long long n = r / Q;
    const float* a = heads + ((n * Q + q) * 4 + 0) * QK;
    const float* b = heads + ((n * Q + q2) * 4 + 2) * QK;
    float acc = 0.f;
    for (int d = 0; d < QK; ++d) acc += a[d] * b[d];
    acc = acc > 0.f ? acc : 0.f;
    A[i] = acc * acc;
  }
};

// cross-token weights: same position q, across the BT sequences of a window, diagonal removed (:171-178)
struct SimCross {
  const float* heads; float* Ac; int Q, BT; float scale;
  GAN_HD void operator()(long long i) const {
    const int t2 = (int)(i % BT); long long r = i / BT; const int t1 = (int)(r % BT); r /= BT; const int q = (int)(r % Q);
    const long long b = r / Q;
    if (t1 == t2) { Ac[i] = 0.f; return; }
    const float* a = heads + (((b * BT + t1) * Q + q) * 4 + 0) * QK;
    const float* k = heads + (((b * BT + t2) * Q + q) * 4 + 2) * QK;
    float acc = 0.f;
    for (int d = 0; d < QK; ++d) acc += a[d] * k[d];
    acc *= scale;
    acc = acc > 0.f ? acc : 0.f;
    Ac[i] = acc * acc;
  }
};

// linear attention state lin_k^T hs per sequence (:182): kv[n, d, e]
struct LinKV {
  const float* heads; const float* huv; float* kv; int Q;
  GAN_HD void operator()(long long i) const {
    const int e = (int)(i % HID); const long long r = i / HID; const int d = (int)(r % QK); const long long n = r / QK;
    float acc = 0.f;
    for (int q = 0; q < Q; ++q) acc += heads[((n * Q + q) * 4 + 3) * QK + d] * huv[(n * Q + q) * HUV + e];
    kv[i] = acc;
  }
};

// att[n, q, e] = local + cross-token + linear (:179-182)
struct Att {
  const float* A; const float* Ac; const float* heads; const float* kv; const float* huv; float* att; int Q, BT;
  GAN_HD void operator()(long long i) const {
    const int e = (int)(i % HID); const long long r = i / HID; const int q = (int)(r % Q); const long long n = r / Q;
    const long long b = n / BT; const int t1 = (int)(n % BT);
    float acc = 0.f;
    const float* a = A + (n * Q + q) * Q;
    for (int q2 = 0; q2 < Q; ++q2) acc += a[q2] * huv[(n * Q + q2) * HUV + e];
    float cr = 0.f;
    const float* ac = Ac + ((b * Q + q) * BT + t1) * BT;
    for (int t2 = 0; t2 < BT; ++t2) cr += ac[t2] * huv[((b * BT + t2) * Q + q) * HUV + e];
    float ln = 0.f;
    const float* lq = heads + ((n * Q + q) * 4 + 1) * QK;
    for (int d = 0; d < QK; ++d) ln += lq[d] * kv[(n * QK + d) * HID + e];
    att[i] = (acc + cr) + ln;
  }
};

// (att_u * v) * sigmoid(att_v * u) (:183)
struct GateOut {
  const float* att; const float* huv; float* out;
  GAN_HD void operator()(long long i) const {
    const int j = (int)(i % (HID / 2)); const long long r = i / (HID / 2);
    const float* a = att + r * HID; const float* h = huv + r * HUV;
    out[i] = (a[HID / 2 + j] * h[j]) * sigmoidf_(a[j] * h[HID / 2 + j]);
  }
};

// SELayer (:689-696): per (window, frame, channel) sum / max over sub-bands, then per (window, channel), then the two MLPs
struct SePool1 {
  const float* x; float* part; int Fw;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long bt = i / C;
    const float* p = x + bt * Fw * C + c;
    float s = 0.f, m = -INFINITY;
    for (int f = 0; f < Fw; ++f) { const float v = p[f * C]; s += v; m = v > m ? v : m; }
    part[2 * i] = s; part[2 * i + 1] = m;
  }
};
struct SePool2 {
  const float* part; float* pooled; int T, Fw;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long b = i / C;
    double s = 0.0; float m = -INFINITY;
    for (int t = 0; t < T; ++t) { const float* p = part + 2 * ((b * T + t) * C + c); s += (double)p[0]; m = p[1] > m ? p[1] : m; }
    pooled[(b * 2 + 0) * C + c] = (float)(s / ((double)T * Fw));
    pooled[(b * 2 + 1) * C + c] = m;
  }
};
struct SeMlp {
  const float* pooled; const float* w0[2]; const float* b0[2]; const float* w2[2]; const float* b2[2]; float* scale;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long b = i / C;
    float tot = 0.f;
    for (int kd = 0; kd < 2; ++kd) {
      const float* p = pooled + (b * 2 + kd) * C;
      float o = b2[kd][c];
      for (int j = 0; j < C; ++j) {
        float h = b0[kd][j];
        for (int k = 0; k < C; ++k) h += w0[kd][j * C + k] * p[k];
        h = h > 0.f ? h : 0.f;
        o += w2[kd][c * C + j] * h;
      }
      tot += sigmoidf_(o);
    }
    scale[i] = tot;
  }
};
struct ScaleRes {
  const float* t; const float* scale; const float* x; float* out; long long per_window;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long b = idiv(i, per_window);
    out[i] = scale[b * C + c] * t[i] + x[i];
  }
};

// LayerNorm over (channel group, sub-band) per (window, frame) (:751-784): stats, then affine (+ residual)
struct GroupBounds { int n; int lo[13]; };
// stats in two steps: per (window, frame, sub-band, group) double sums over the group's channels, then over sub-bands
struct GroupPart {
  const float* x; int ld; GroupBounds g; double* part; 
  GAN_HD void operator()(long long i) const {
    const int gi = (int)(i % g.n); const long long p = i / g.n;
    const float* v = x + p * ld;
    double s = 0.0, s2 = 0.0;
    for (int c = g.lo[gi]; c < g.lo[gi + 1]; ++c) { const double d = (double)v[c]; s += d; s2 += d * d; }
    part[2 * i] = s; part[2 * i + 1] = s2;
  }
};
struct GroupFin {
  const double* part; GroupBounds g; float* stat; int Fw;
  GAN_HD void operator()(long long i) const {
    const int gi = (int)(i % g.n); const long long bt = i / g.n;
    double s = 0.0, s2 = 0.0;
    for (int f = 0; f < Fw; ++f) { const double* p = part + 2 * ((bt * Fw + f) * g.n + gi); s += p[0]; s2 += p[1]; }
    const double cnt = (double)Fw * (g.lo[gi + 1] - g.lo[gi]), mu = s / cnt;
    double var = s2 / cnt - mu * mu;
    var = var > 0.0 ? var : 0.0;
    stat[2 * i] = (float)mu;
    stat[2 * i + 1] = (float)(1.0 / sqrt(var + (double)EPS));
  }
};
struct GroupNorm {
  float* x; int ld; GroupBounds g; const float* stat; const float* gam; const float* bet; const float* res; float* out; int Fw;
  GAN_HD void operator()(long long i) const {
    const int c = imod(i, ld); const long long p = idiv(i, ld); const int f = imod(p, Fw); const long long bt = idiv(p, Fw);
    int gi = 0;
    while (gi + 1 < g.n && c >= g.lo[gi + 1]) ++gi;
    const float* st = stat + 2 * (bt * g.n + gi);
    float v = (x[i] - st[0]) * st[1] * gam[c * Fw + f] + bet[c * Fw + f];
    if (res) v += res[i];
    out[i] = v;
  }
};
// scores over time with (channel, sub-band) flattened features; softmax; weighted values
struct TaScores {
  const float* qkv; float* a; int T, Fw;
  GAN_HD void operator()(long long i) const {
    const int t2 = (int)(i % T); long long r = i / T; const int t1 = (int)(r % T); r /= T; const int h = (int)(r % HEADS);
    const long long b = r / HEADS;
    const float* q = qkv + (b * T + t1) * Fw * QKV + h * AE;
    const float* k = qkv + (b * T + t2) * Fw * QKV + HEADS * AE + h * AE;
    float acc = 0.f;
    for (int f = 0; f < Fw; ++f)
      for (int e = 0; e < AE; ++e) acc += q[f * QKV + e] * k[f * QKV + e];
    a[i] = acc;
  }
};
struct Softmax {
  float* a; int n;
  GAN_HD void operator()(long long r) const {
    float* p = a + r * n;
    float m = -INFINITY;
    for (int j = 0; j < n; ++j) m = p[j] > m ? p[j] : m;
    float s = 0.f;
    for (int j = 0; j < n; ++j) { const float e = expf(p[j] - m); p[j] = e; s += e; }
    const float inv = 1.0f / s;
    for (int j = 0; j < n; ++j) p[j] *= inv;
  }
};
struct TaAV {
  const float* a; const float* qkv; float* out; int T, Fw;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); long long r = i / C; const int f = (int)(r % Fw); r /= Fw; const int t1 = (int)(r % T);
    const long long b = r / T;
    const int h = c / VC;
    const float* w = a + ((b * HEADS + h) * T + t1) * T;
    const float* v = qkv + (b * T * (long long)Fw + f) * QKV + 2 * HEADS * AE + c;
    float acc = 0.f;
    for (int t2 = 0; t2 < T; ++t2) acc += w[t2] * v[(long long)t2 * Fw * QKV];
    out[i] = acc;
  }
};

// InstanceNorm2d (biased variance over frames x sub-bands per window and channel) + PReLU
struct InPart {
  const float* x; int ld; int Cn; double* part; int Fw;
  GAN_HD void operator()(long long i) const {
    const int c = imod(i, Cn); const long long bt = idiv(i, Cn);
    const float* p = x + bt * Fw * ld + c;
    double s = 0.0, s2 = 0.0;
    for (int f = 0; f < Fw; ++f) { const double v = (double)p[(long long)f * ld]; s += v; s2 += v * v; }
    part[2 * i] = s; part[2 * i + 1] = s2;
  }
};
struct InFin {
  const double* part; float* stat; int T, Fw, Cn;
  GAN_HD void operator()(long long i) const {
    const int c = imod(i, Cn); const long long b = idiv(i, Cn);
    double s = 0.0, s2 = 0.0;
    for (int t = 0; t < T; ++t) { const double* p = part + 2 * ((b * T + t) * Cn + c); s += p[0]; s2 += p[1]; }
    const double cnt = (double)T * Fw, mu = s / cnt;
    double var = s2 / cnt - mu * mu;
    var = var > 0.0 ? var : 0.0;
    stat[2 * i] = (float)mu;
    stat[2 * i + 1] = (float)(1.0 / sqrt(var + (double)EPS));
  }
};
struct InApply {
  const float* x; int ld; int Cn; const float* stat; const float* w; const float* b; const float* slope; float* out; int ldo;
  long long per_window;      // frames * sub-bands
  GAN_HD void operator()(long long i) const {
    const int c = imod(i, Cn); const long long p = idiv(i, Cn); const long long bb = idiv(p, per_window);
    const float* st = stat + 2 * (bb * Cn + c);
    const float v = (x[p * ld + c] - st[0]) * st[1] * w[c] + b[c];
    out[p * ldo + c] = v >= 0.f ? v : slope[c] * v;
  }
};

// channel-last conv: kernel (KT, KF), causal dilation over frames, stride / zero pad over sub-bands
struct Conv2d {
  const float* in; int ldi; int Fin; const float* W; const float* bias; float* out; int ldo; int Fout; int T; int Cin, Cout;
  int KT, KF, dil, sf, pf;
  GAN_HD void operator()(long long i) const {
    const int co = (int)(i % Cout); long long r = i / Cout; const int fo = (int)(r % Fout); r /= Fout; const int t = (int)(r % T);
    const long long b = r / T;
    float acc = bias ? bias[co] : 0.f;
    for (int kt = 0; kt < KT; ++kt) {
      const int ti = t - (KT - 1 - kt) * dil;
      if (ti < 0) continue;
      for (int kf = 0; kf < KF; ++kf) {
        const int fi = fo * sf + kf - pf;
        if (fi < 0 || fi >= Fin) continue;
        const float* x = in + ((b * T + ti) * Fin + fi) * ldi;
        const float* w = W + (long long)(kt * KF + kf) * Cin * Cout + co;
        for (int ci = 0; ci < Cin; ++ci) acc += x[ci] * w[(long long)ci * Cout];
      }
    }
    out[((b * T + t) * Fout + fo) * ldo + co] = acc;
  }
};

// encoder conv_1: 1x1 over the 3 channel-first feature planes (window, 3, frame, bin) -> channel-last (:590)
struct FeatConv {
  const float* feat; const float* w; const float* b; float* out; int T, Fw;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long p = i / C; const long long tf = p % ((long long)T * Fw); const long long bb = p / ((long long)T * Fw);
    float acc = b[c];
    for (int k = 0; k < 3; ++k) acc += w[c * 3 + k] * feat[(bb * 3 + k) * (long long)T * Fw + tf];
    out[i] = acc;
  }
};
struct CopyCh {
  const float* in; int ldi; float* out; int ldo; int Cn;
  GAN_HD void operator()(long long i) const {
    const int c = imod(i, Cn); const long long p = idiv(i, Cn);
    out[p * ldo + c] = in[p * ldi + c];
  }
};
// mask tail: InstanceNorm2d(1) + PReLU(1) + 1x1 final conv + per-bin PReLU, transposed to (window, bin, frame) (:818-828)
struct MaskTail {
  const float* xm; const float* stat; const float* nw; const float* nb; const float* pa; const float* fw; const float* fb;
  const float* pout; float* mask; int T, Fb;
  GAN_HD void operator()(long long i) const {
    const int t = (int)(i % T); long long r = i / T; const int f = (int)(r % Fb); const long long b = r / Fb;
    float v = (xm[(b * T + t) * Fb + f] - stat[2 * b]) * stat[2 * b + 1] * nw[0] + nb[0];
    v = v >= 0.f ? v : pa[0] * v;
    v = v * fw[0] + fb[0];
    mask[i] = v >= 0.f ? v : pout[f] * v;
  }
};
// complex tail: (1, 2) conv 64 -> 2, transposed to (window, 2, bin, frame) (:846-848)
struct CplxTail {
  const float* xc; const float* w; const float* b; float* out; int T, Fb;
  GAN_HD void operator()(long long i) const {
    const int t = (int)(i % T); long long r = i / T; const int f = (int)(r % Fb); r /= Fb; const int k = (int)(r % 2);
    const long long bb = r / 2;
    float acc = b[k];
    for (int kf = 0; kf < 2; ++kf) {
      const float* x = xc + ((bb * T + t) * (Fb + 1) + f + kf) * C;
      for (int ci = 0; ci < C; ++ci) acc += x[ci] * w[(kf * C + ci) * 2 + k];
    }
    out[i] = acc;
  }
};

// ------------------------------------------------------------------------------------------------ weights
struct DenseW { const float *conv_w[DEPTH], *conv_b[DEPTH], *nw[DEPTH], *nb[DEPTH], *pa[DEPTH], *fl_w[DEPTH], *fl_b[DEPTH], *fp_w[DEPTH], *fm_w[DEPTH]; };
struct MfW { const float *in_w, *in_b, *in_c, *out_w, *out_b, *out_c, *gamma, *beta; float cross_scale; };
struct PathW {
  const float *gw, *gb, *uv_w, *uv_b, *uv_c, *rl_w, *rl_b, *rp_w, *rm_w, *lin_w, *lin_b;
  MfW mf;
  const float *se_w0[2], *se_b0[2], *se_w2[2], *se_b2[2];
};
struct AttW { const float *w, *b, *a, *g, *beta, *p_w, *p_b, *p_a, *p_g, *p_beta; };
struct BlockW { PathW intra, inter; AttW att; };
struct DecW { DenseW dd; const float *sp_w, *sp_b, *nw, *nb, *pa; };
struct Weights {
  const float *c1_w, *c1_b, *n1_w, *n1_b, *p1, *c2_w, *c2_b, *n2_w, *n2_b, *p2, *rot_cos, *rot_sin;
  DenseW enc_dd;
  int layers;
  BlockW blocks[8];
  DecW md, cd;
  const float *md_c1_w, *md_c1_b, *md_fin_w, *md_fin_b, *md_pout, *cd_c_w, *cd_c_b;
};

// Look weights up by name (`lk(name, expected_count)` returns a pointer or null and records the error itself).
template <class Lookup>
bool bind_dense(DenseW& d, const char* pre, Lookup& lk) {
  char nm[96];
  bool ok = true;
  for (int i = 0; i < DEPTH; ++i) {
    const int cin = C * (i + 1);
    auto get = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "%s%d.%s", pre, i, k); const float* p = lk(nm, n); ok = ok && p; return p; };
    d.conv_w[i] = get("conv_w", (size_t)6 * cin * C); d.conv_b[i] = get("conv_b", C);
    d.nw[i] = get("nw", C); d.nb[i] = get("nb", C); d.pa[i] = get("pa", C);
    d.fl_w[i] = get("fl_w", C * C); d.fl_b[i] = get("fl_b", C); d.fp_w[i] = get("fp_w", C * C); d.fm_w[i] = get("fm_w", (2 * DLORDER - 1) * C);
  }
  return ok;
}
template <class Lookup>
bool bind(Weights& W, int layers, int T, Lookup& lk) {
  bool ok = true;
  char nm[96];
  auto g = [&](const char* name, size_t n) { const float* p = lk(name, n); ok = ok && p; return p; };
  if (layers < 1 || layers > 8) return false;
  W.layers = layers;
  W.c1_w = g("enc.c1_w", C * 3); W.c1_b = g("enc.c1_b", C); W.n1_w = g("enc.n1_w", C); W.n1_b = g("enc.n1_b", C); W.p1 = g("enc.p1", C);
  W.c2_w = g("enc.c2_w", 3 * C * C); W.c2_b = g("enc.c2_b", C); W.n2_w = g("enc.n2_w", C); W.n2_b = g("enc.n2_b", C); W.p2 = g("enc.p2", C);
  const int maxseq = (T > FQ ? T : FQ) + 2;
  W.rot_cos = g("rot_cos", (size_t)maxseq * ROT); W.rot_sin = g("rot_sin", (size_t)maxseq * ROT);
  ok = bind_dense(W.enc_dd, "enc.dd", lk) && ok;
  for (int i = 0; i < layers; ++i) {
    BlockW& b = W.blocks[i];
    for (int pi = 0; pi < 2; ++pi) {
      PathW& p = pi ? b.inter : b.intra;
      auto gp = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "B%d.%s.%s", i, pi ? "inter" : "intra", k); return g(nm, n); };
      p.gw = gp("gw", PI * KS); p.gb = gp("gb", PI);
      p.uv_w = gp("uv_w", PI * 2 * UV); p.uv_b = gp("uv_b", 2 * UV); p.uv_c = gp("uv_c", DW * 2 * UV);
      p.rl_w = gp("rl_w", UV * UV); p.rl_b = gp("rl_b", UV); p.rp_w = gp("rp_w", UV * UV); p.rm_w = gp("rm_w", (2 * LORDER - 1) * UV);
      p.lin_w = gp("lin_w", KS * UV * C); p.lin_b = gp("lin_b", C);
      p.mf.in_w = gp("mf.in_w", C * HUV); p.mf.in_b = gp("mf.in_b", HUV); p.mf.in_c = gp("mf.in_c", DW * HUV);
      p.mf.out_w = gp("mf.out_w", (HID / 2) * C); p.mf.out_b = gp("mf.out_b", C); p.mf.out_c = gp("mf.out_c", DW * C);
      p.mf.gamma = gp("mf.gamma", 4 * QK); p.mf.beta = gp("mf.beta", 4 * QK);
      p.mf.cross_scale = (float)((double)(pi ? T : FQ) / (double)(pi ? FQ : T));       // Q / BT (:171-175)
      const char* kinds[2] = {"avg", "max"};
      for (int kd = 0; kd < 2; ++kd) {
        char k[32];
        snprintf(k, sizeof(k), "se_%s0_w", kinds[kd]); p.se_w0[kd] = gp(k, C * C);
        snprintf(k, sizeof(k), "se_%s0_b", kinds[kd]); p.se_b0[kd] = gp(k, C);
        snprintf(k, sizeof(k), "se_%s2_w", kinds[kd]); p.se_w2[kd] = gp(k, C * C);
        snprintf(k, sizeof(k), "se_%s2_b", kinds[kd]); p.se_b2[kd] = gp(k, C);
      }
    }
    auto ga = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "B%d.att.%s", i, k); return g(nm, n); };
    b.att.w = ga("w", C * QKV); b.att.b = ga("b", QKV); b.att.a = ga("a", QKV);
    b.att.g = ga("g", QKV * FQ); b.att.beta = ga("beta", QKV * FQ);
    b.att.p_w = ga("p_w", C * C); b.att.p_b = ga("p_b", C); b.att.p_a = ga("p_a", C);
    b.att.p_g = ga("p_g", C * FQ); b.att.p_beta = ga("p_beta", C * FQ);
  }
  for (int di = 0; di < 2; ++di) {
    DecW& d = di ? W.cd : W.md;
    const char* o = di ? "cd" : "md";
    snprintf(nm, sizeof(nm), "%s.dd", o);
    char pre[16];
    snprintf(pre, sizeof(pre), "%s.dd", o);
    ok = bind_dense(d.dd, pre, lk) && ok;
    auto gd = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "%s.%s", o, k); return g(nm, n); };
    d.sp_w = gd("sp_w", 3 * C * 2 * C); d.sp_b = gd("sp_b", 2 * C);
    d.nw = gd("nw", di ? C : 1); d.nb = gd("nb", di ? C : 1); d.pa = gd("pa", di ? C : 1);
  }
  W.md_c1_w = g("md.c1_w", 2 * C); W.md_c1_b = g("md.c1_b", 1); W.md_fin_w = g("md.fin_w", 1); W.md_fin_b = g("md.fin_b", 1);
  W.md_pout = g("md.pout", FB); W.cd_c_w = g("cd.c_w", 2 * C * 2); W.cd_c_b = g("cd.c_b", 2);
  return ok;
}

// ------------------------------------------------------------------------------------------------ workspace
// Per-window float counts; `alloc(count)` returns a buffer of that many floats (or null).
struct Workspace {
  float *skip, *o201, *h201, *p201, *x, *xa, *xb, *pst, *seq, *huv, *fh, *fp, *iu, *t0, *sh, *rst, *mhuv, *heads, *A, *Ac, *kv,
      *att, *go, *ho, *separt, *sepool, *sescale, *qkv, *gst, *sc, *av, *pr, *ist, *xm;
  double *dpart, *gpart;
};
template <class Alloc>
bool alloc_ws(Workspace& w, int B, int T, Alloc& alloc) {
  const long long px = (long long)B * T * FQ, px2 = (long long)B * T * (FB + 1);
  const int S = T > FQ ? T : FQ;                   // longest sequence
  const long long rows = (long long)B * T * FQ;   // rows of either path, upper bound (Q positions x BT sequences)
  bool ok = true;
  auto a = [&](float*& p, long long n) { p = alloc((size_t)n); ok = ok && p; };
  a(w.skip, px2 * SKIPC); a(w.o201, px2 * C); a(w.h201, px2 * C); a(w.p201, px2 * C);
  a(w.x, px * C); a(w.xa, px * C); a(w.xb, px * C); a(w.pst, px * 2);
  a(w.seq, rows * PI); a(w.huv, rows * 2 * UV); a(w.fh, rows * UV); a(w.fp, rows * UV); a(w.iu, rows * UV);
  a(w.t0, rows * C); a(w.sh, rows * C); a(w.rst, rows * 2); a(w.mhuv, rows * HUV); a(w.heads, rows * 4 * QK);
  a(w.A, rows * S); a(w.Ac, rows * S); a(w.kv, (long long)B * S * QK * HID); a(w.att, rows * HID); a(w.go, rows * (HID / 2));
  a(w.ho, rows * C); a(w.separt, (long long)B * T * C * 2); a(w.sepool, (long long)B * 2 * C); a(w.sescale, (long long)B * C);
  a(w.qkv, px * QKV); a(w.gst, (long long)B * T * 12 * 2); a(w.sc, (long long)B * HEADS * T * T); a(w.av, px * C); a(w.pr, px * C);
  a(w.ist, (long long)B * C * 2); a(w.xm, px2);
  float* d = alloc((size_t)B * T * C * 2 * 2);     // doubles
  ok = ok && d;
  w.dpart = reinterpret_cast<double*>(d);
  float* d2 = alloc((size_t)px * 12 * 2 * 2);      // doubles
  ok = ok && d2;
  w.gpart = reinterpret_cast<double*>(d2);
  return ok;
}

// ------------------------------------------------------------------------------------------------ launch sequence
// `ex.run(count, functor)` evaluates functor(i) for i in [0, count); `ex.mark(name, ptr, count)` is a stage hook.
template <class Exec>
void inst_norm(Exec& ex, const Workspace& w, const float* x, int ld, int Cn, int B, int T, int Fw, const float* nw, const float* nb,
               const float* pa, float* out, int ldo) {
  ex.run((long long)B * T * Cn, InPart{x, ld, Cn, w.dpart, Fw});
  ex.run((long long)B * Cn, InFin{w.dpart, w.ist, T, Fw, Cn});
  ex.run((long long)B * T * Fw * Cn, InApply{x, ld, Cn, w.ist, nw, nb, pa, out, ldo, (long long)T * Fw});
}

// DilatedDenseNet + FSMN along sub-bands (:601-623).  The input must already sit in skip slot DEPTH; result -> out (ld C).
template <class Exec>
void dense_block(Exec& ex, const Workspace& w, const DenseW& d, int B, int T, int Fw, float* out) {
  const long long px = (long long)B * T * Fw;
  for (int i = 0; i < DEPTH; ++i) {
    const int cin = C * (i + 1);
    ex.run(px * C, Conv2d{w.skip + (DEPTH - i) * C, SKIPC, Fw, d.conv_w[i], d.conv_b[i], w.o201, C, Fw, T, cin, C, 2, 3, 1 << i, 1, 1});
    inst_norm(ex, w, w.o201, C, C, B, T, Fw, d.nw[i], d.nb[i], d.pa[i], w.o201, C);
    ex.run(px * C, Linear{w.o201, C, nullptr, d.fl_w[i], d.fl_b[i], w.h201, C, C, C, ACT_RELU, nullptr});
    ex.run(px * C, Linear{w.h201, C, nullptr, d.fp_w[i], nullptr, w.p201, C, C, C, ACT_NONE, nullptr});
    float* dst = i + 1 < DEPTH ? w.skip + (DEPTH - 1 - i) * C : out;
    ex.run(dw_count((long long)B * T, Fw, C), DwConv<2 * DLORDER - 1>{w.p201, C, w.o201, C, d.fm_w[i], dst, i + 1 < DEPTH ? SKIPC : C, 0, PixMap{1, 0, 0, 0},
                          Fw, C});
  }
}

// One intra / inter path (:639-679) + MossFormer block (:137-244) + SE + residual.  xin -> xout (channel-last maps).
template <class Exec>
void path(Exec& ex, const Workspace& w, const Weights& W, const PathW& p, const float* xin, float* xout, int B, int T, bool inter,
          const char* tag) {
  const int Q = inter ? T : FQ, BT = inter ? FQ : T, S = Q - KS + 1;
  const long long N = (long long)B * BT, px = (long long)B * T * FQ;
  const PixMap pm = inter ? PixMap{FQ, (long long)T * FQ, 1, FQ} : PixMap{T, (long long)T * FQ, FQ, 1};
  ex.run(px, RowStats{xin, C, C, w.pst});
  ex.run(N * S * PI, Gather{xin, w.pst, pm, p.gw, p.gb, w.seq, S});
  ex.run(N * S, RowStats{w.seq, PI, PI, w.rst});
  ex.run(N * S * 2 * UV, Linear{w.seq, PI, w.rst, p.uv_w, p.uv_b, w.att, 2 * UV, PI, 2 * UV, ACT_SILU, nullptr});
  ex.run(dw_count(N, S, 2 * UV), DwConv<DW>{w.att, 2 * UV, nullptr, 0, p.uv_c, w.huv, 2 * UV, 0, pm, S, 2 * UV});
  ex.mark(tag, "huv", w.huv, N * S * 2 * UV);
  ex.run(N * S * UV, Linear{w.huv, 2 * UV, nullptr, p.rl_w, p.rl_b, w.fh, UV, UV, UV, ACT_RELU, nullptr});
  ex.run(N * S * UV, Linear{w.fh, UV, nullptr, p.rp_w, nullptr, w.fp, UV, UV, UV, ACT_NONE, nullptr});
  ex.run(dw_count(N, S, UV), DwConv<2 * LORDER - 1>{w.fp, UV, w.huv, 2 * UV, p.rm_w, w.iu, UV, 0, pm, S, UV});
  ex.run(N * Q * C, GateConvT{w.iu, UV, w.huv + UV, 2 * UV, p.lin_w, p.lin_b, w.t0, S, Q});
  ex.mark(tag, "lin", w.t0, N * Q * C);
  // MossFormer block on t0 (N, Q, 64)
  ex.run(N * Q * C, Shift{w.t0, w.sh, Q});
  ex.run(N * Q, RowStats{w.sh, C, C, w.rst});
  ex.run(N * Q * HUV, Linear{w.sh, C, w.rst, p.mf.in_w, p.mf.in_b, w.heads, HUV, C, HUV, ACT_SILU, nullptr});
  ex.run(dw_count(N, Q, HUV), DwConv<DW>{w.heads, HUV, nullptr, 0, p.mf.in_c, w.mhuv, HUV, 0, pm, Q, HUV});
  ex.mark(tag, "mf.huv", w.mhuv, N * Q * HUV);
  ex.run(N * Q * 4 * QK, OffsetRot{w.mhuv, p.mf.gamma, p.mf.beta, W.rot_cos, W.rot_sin, w.heads, Q});
  ex.run(N * Q * Q, SimLocal{w.heads, w.A, Q});
  ex.run((long long)B * Q * BT * BT, SimCross{w.heads, w.Ac, Q, BT, p.mf.cross_scale});
  ex.run(N * QK * HID, LinKV{w.heads, w.mhuv, w.kv, Q});
  ex.run(N * Q * HID, Att{w.A, w.Ac, w.heads, w.kv, w.mhuv, w.att, Q, BT});
  ex.mark(tag, "mf.att", w.att, N * Q * HID);
  ex.run(N * Q * (HID / 2), GateOut{w.att, w.mhuv, w.go});
  ex.run(N * Q, RowStats{w.go, HID / 2, HID / 2, w.rst});
  ex.run(N * Q * C, Linear{w.go, HID / 2, w.rst, p.mf.out_w, p.mf.out_b, w.ho, C, HID / 2, C, ACT_SILU, nullptr});
  ex.run(dw_count(N, Q, C), DwConv<DW>{w.ho, C, w.t0, C, p.mf.out_c, w.pr, C, 1, pm, Q, C});     // back to the channel-last map
  // SE (:689-696) + residual
  ex.run((long long)B * T * C, SePool1{w.pr, w.separt, FQ});
  ex.run((long long)B * C, SePool2{w.separt, w.sepool, T, FQ});
  ex.run((long long)B * C, SeMlp{w.sepool, {p.se_w0[0], p.se_w0[1]}, {p.se_b0[0], p.se_b0[1]}, {p.se_w2[0], p.se_w2[1]},
                                 {p.se_b2[0], p.se_b2[1]}, w.sescale});
  ex.run(px * C, ScaleRes{w.pr, w.sescale, xin, xout, (long long)T * FQ * C});
}

template <class Exec>
void triple_attention(Exec& ex, const Workspace& w, const AttW& a, const float* xin, float* xout, int B, int T) {
  const long long px = (long long)B * T * FQ;
  GroupBounds g12{12, {0, 6, 12, 18, 24, 30, 36, 42, 48, 64, 80, 96, 112}}, g1{1, {0, 64}};
  ex.run(px * QKV, Linear{xin, C, nullptr, a.w, a.b, w.qkv, QKV, C, QKV, ACT_PRELU, a.a});
  ex.run(px * 12, GroupPart{w.qkv, QKV, g12, w.gpart});
  ex.run((long long)B * T * 12, GroupFin{w.gpart, g12, w.gst, FQ});
  ex.run(px * QKV, GroupNorm{w.qkv, QKV, g12, w.gst, a.g, a.beta, nullptr, w.qkv, FQ});
  ex.run((long long)B * HEADS * T * T, TaScores{w.qkv, w.sc, T, FQ});
  ex.run((long long)B * HEADS * T, Softmax{w.sc, T});
  ex.run(px * C, TaAV{w.sc, w.qkv, w.av, T, FQ});
  ex.run(px * C, Linear{w.av, C, nullptr, a.p_w, a.p_b, w.pr, C, C, C, ACT_PRELU, a.p_a});
  ex.run(px, GroupPart{w.pr, C, g1, w.gpart});
  ex.run((long long)B * T, GroupFin{w.gpart, g1, w.gst, FQ});
  ex.run(px * C, GroupNorm{w.pr, C, g1, w.gst, a.p_g, a.p_beta, xin, xout, FQ});
}

// feat (B, 3, T, 201) -> mask (B, 201, T), cplx (B, 2, 201, T)
template <class Exec>
void forward(Exec& ex, const Workspace& w, const Weights& W, const float* feat, float* mask, float* cplx, int B, int T) {
  const long long px = (long long)B * T * FQ, pxb = (long long)B * T * FB, px2 = (long long)B * T * (FB + 1);
  char tag[32];
  // dense encoder (:590-627)
  ex.run(pxb * C, FeatConv{feat, W.c1_w, W.c1_b, w.o201, T, FB});
  inst_norm(ex, w, w.o201, C, C, B, T, FB, W.n1_w, W.n1_b, W.p1, w.skip + DEPTH * C, SKIPC);
  dense_block(ex, w, W.enc_dd, B, T, FB, w.h201);
  ex.run(px * C, Conv2d{w.h201, C, FB, W.c2_w, W.c2_b, w.xa, C, FQ, T, C, C, 1, 3, 1, 2, 1});
  inst_norm(ex, w, w.xa, C, C, B, T, FQ, W.n2_w, W.n2_b, W.p2, w.x, C);
  ex.mark("", "enc", w.x, px * C);
  for (int i = 0; i < W.layers; ++i) {
    const BlockW& b = W.blocks[i];
    snprintf(tag, sizeof(tag), "B%d.intra", i);
    path(ex, w, W, b.intra, w.x, w.xa, B, T, false, tag);
    ex.mark("", tag, w.xa, px * C);
    snprintf(tag, sizeof(tag), "B%d.inter", i);
    path(ex, w, W, b.inter, w.xa, w.xb, B, T, true, tag);
    ex.mark("", tag, w.xb, px * C);
    triple_attention(ex, w, b.att, w.xb, w.x, B, T);
    snprintf(tag, sizeof(tag), "B%d.x", i);
    ex.mark("", tag, w.x, px * C);
  }
  // mask decoder (:792-828)
  ex.run(px * C, CopyCh{w.x, C, w.skip + DEPTH * C, SKIPC, C});
  dense_block(ex, w, W.md.dd, B, T, FQ, w.xa);
  ex.run(px * 2 * C, Conv2d{w.xa, C, FQ, W.md.sp_w, W.md.sp_b, w.o201, 2 * C, FQ, T, C, 2 * C, 1, 3, 1, 1, 1});     // == (B, T, 202, 64)
  ex.run(pxb, Conv2d{w.o201, C, FB + 1, W.md_c1_w, W.md_c1_b, w.xm, 1, FB, T, C, 1, 1, 2, 1, 1, 0});
  ex.run((long long)B * T, InPart{w.xm, 1, 1, w.dpart, FB});
  ex.run((long long)B, InFin{w.dpart, w.ist, T, FB, 1});
  ex.run(pxb, MaskTail{w.xm, w.ist, W.md.nw, W.md.nb, W.md.pa, W.md_fin_w, W.md_fin_b, W.md_pout, mask, T, FB});
  ex.mark("", "mask", mask, pxb);
  // complex decoder (:830-848)
  ex.run(px * C, CopyCh{w.x, C, w.skip + DEPTH * C, SKIPC, C});
  dense_block(ex, w, W.cd.dd, B, T, FQ, w.xa);
  ex.run(px * 2 * C, Conv2d{w.xa, C, FQ, W.cd.sp_w, W.cd.sp_b, w.o201, 2 * C, FQ, T, C, 2 * C, 1, 3, 1, 1, 1});
  inst_norm(ex, w, w.o201, C, C, B, T, FB + 1, W.cd.nw, W.cd.nb, W.cd.pa, w.o201, C);
  ex.run(pxb * 2, CplxTail{w.o201, W.cd_c_w, W.cd_c_b, cplx, T, FB});
  ex.mark("", "complex", cplx, pxb * 2);
  (void)px2;
}

}  // namespace gan
