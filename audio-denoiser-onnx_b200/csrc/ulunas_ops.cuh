// UL-UNAS (16 kHz) -- SURVEY 8f rank 3: the operators of `ULUNAS.forward` / `ULUNAS_CUSTOM.forward` (reference
// UL-UNAS/Export_UL_UNAS.py:51-912) as functors + the launch sequence over them, templated on the executor like
// csrc/mfgan_ops.cuh / csrc/dfsmn_ops.cuh: libadn runs it on the CUDA executor (csrc/ulunas.cu), tests/harness/ulunas_host.cpp
// with a host loop (CPU check against oracle/ulunas_oracle.py, stage by stage).  First-correct design: feature maps NCHW
// (window, channel, frame, band) as the reference lays them out, one output per thread, GRUs as one thread per sequence.
#pragma once
#include "mfgan_gemm.cuh"

namespace uln {

using gan::Linear;

constexpr int NB = 257, ERB_LO = 65, ERB_HI = 64, NE = ERB_LO + ERB_HI, HOP = 256, NFFT = 512;
constexpr int GRU_HMAX = 64;

// log power -> ERB bands (:719-721): feat[b, 0, t, j]
struct PowerLogErb {
  const float* spec; const float* erb; float* feat; int T;      // spec (B, 514, T); erb (64, 192) = erb_fc.weight
  GAN_HD float lp(long long b, int f, int t) const {
    const float re = spec[(b * 2 * NB + f) * T + t], im = spec[(b * 2 * NB + NB + f) * T + t];
    const float p = re * re + im * im;
    return logf(p > 1e-24f ? p : 1e-24f);
  }
  GAN_HD void operator()(long long i) const {
    const int j = (int)(i % NE); const long long bt = i / NE; const int t = (int)(bt % T); const long long b = bt / T;
    if (j < ERB_LO) { feat[i] = lp(b, j, t); return; }
    const float* w = erb + (long long)(j - ERB_LO) * (NB - ERB_LO);
    float acc = 0.f;
    for (int k = 0; k < NB - ERB_LO; ++k) acc += lp(b, ERB_LO + k, t) * w[k];
    feat[i] = acc;
  }
};

// grouped causal conv, NCHW, stride / zero pad over bands; weights (Cout, Cin / groups, KT, KF)
struct ConvG {
  const float* x; const float* w; const float* bias; float* y; int Cin, Cout, T, Fin, Fout, KT, KF, sf, pf, groups;
  GAN_HD void operator()(long long i) const {
    const int fo = (int)(i % Fout); long long r = i / Fout; const int t = (int)(r % T); r /= T; const int co = (int)(r % Cout);
    const long long b = r / Cout;
    const int ipg = Cin / groups, g = co / (Cout / groups);
    float acc = bias[co];
    for (int ci = 0; ci < ipg; ++ci) {
      const float* xc = x + ((b * Cin + g * ipg + ci) * T) * Fin;
      const float* wc = w + ((long long)co * ipg + ci) * KT * KF;
      for (int kt = 0; kt < KT; ++kt) {
        const int ti = t - (KT - 1 - kt);
        if (ti < 0) continue;
        for (int kf = 0; kf < KF; ++kf) {
          const int fi = fo * sf + kf - pf;
          if (fi >= 0 && fi < Fin) acc += xc[(long long)ti * Fin + fi] * wc[kt * KF + kf];
        }
      }
    }
    y[i] = acc;
  }
};

// grouped causal transposed conv (stride over bands, padding (0, pf), last KT - 1 frames trimmed); weights (Cin, Cout / groups, KT, KF)
struct DeconvG {
  const float* x; const float* w; const float* bias; float* y; int Cin, Cout, T, Fin, Fout, KT, KF, sf, pf, groups;
  GAN_HD void operator()(long long i) const {
    const int fo = (int)(i % Fout); long long r = i / Fout; const int t = (int)(r % T); r /= T; const int co = (int)(r % Cout);
    const long long b = r / Cout;
    const int opg = Cout / groups, ipg = Cin / groups, g = co / opg, col = co - g * opg;
    float acc = bias[co];
    for (int ci = 0; ci < ipg; ++ci) {
      const int cin = g * ipg + ci;
      const float* xc = x + ((b * Cin + cin) * T) * Fin;
      const float* wc = w + ((long long)cin * opg + col) * KT * KF;
      for (int kt = 0; kt < KT; ++kt) {
        const int ti = t - kt;
        if (ti < 0) continue;
        for (int kf = 0; kf < KF; ++kf) {
          const int num = fo + pf - kf;
          if (num < 0 || num % sf) continue;
          const int fi = num / sf;
          if (fi < Fin) acc += xc[(long long)ti * Fin + fi] * wc[kt * KF + kf];
        }
      }
    }
    y[i] = acc;
  }
};

// AffinePReLU after fuse_for_export_ (:122-129): per (channel, band) slopes and bias, in place
struct AffAct {
  float* x; const float* pos; const float* neg; const float* bias; int Cn, T, Fw;
  GAN_HD void operator()(long long i) const {
    const int f = (int)(i % Fw); const int c = (int)((i / ((long long)Fw * T)) % Cn);
    const float v = x[i];
    x[i] = (v > 0.f ? pos[c * Fw + f] : neg[c * Fw + f]) * v + bias[c * Fw + f];
  }
};
// channel shuffle (:197-208): out[:, j] = in[:, (j % 2) * half + j / 2]
struct Shuffle {
  const float* x; float* y; int Cn; long long plane;    // plane = T * F
  GAN_HD void operator()(long long i) const {
    const long long p = i % plane; long long r = i / plane; const int j = (int)(r % Cn); const long long b = r / Cn;
    const int src = (j & 1) * (Cn / 2) + (j >> 1);
    y[i] = x[(b * Cn + src) * plane + p];
  }
};
struct Add2 {
  const float* a; const float* b; float* y;
  GAN_HD void operator()(long long i) const { y[i] = a[i] + (b ? b[i] : 0.f); }
};

// cTFA statistics (:185-186, :156): mean over bands / over channels of x^2
struct MeanF2 {                         // zt[b, t, c]
  const float* x; float* zt; int Cn, T, Fw;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % Cn); const long long bt = i / Cn; const int t = (int)(bt % T); const long long b = bt / T;
    const float* p = x + ((b * Cn + c) * T + t) * Fw;
    float s = 0.f;
    for (int f = 0; f < Fw; ++f) s += p[f] * p[f];
    zt[i] = s / (float)Fw;
  }
};
struct MeanC2 {                         // zf[b, t, f]
  const float* x; float* zf; int Cn, T, Fw;
  GAN_HD void operator()(long long i) const {
    const int f = (int)(i % Fw); const long long bt = i / Fw; const int t = (int)(bt % T); const long long b = bt / T;
    float s = 0.f;
    for (int c = 0; c < Cn; ++c) { const float v = x[((b * Cn + c) * T + t) * Fw + f]; s += v * v; }
    zf[i] = s / (float)Cn;
  }
};
struct CtfaApply {                      // at[b, t, c] * x * af[(b, t), f]   (af rows are Fp = padded band count wide)
  float* x; const float* at; const float* af; int Cn, T, Fw, Fp;
  GAN_HD void operator()(long long i) const {
    const int f = (int)(i % Fw); long long r = i / Fw; const int t = (int)(r % T); r /= T; const int c = (int)(r % Cn);
    const long long b = r / Cn;
    x[i] = at[(b * T + t) * Cn + c] * x[i] * af[(b * T + t) * Fp + f];
  }
};

// nn.GRU, zero initial state, gate order (r, z, n); one thread = one (sequence, group, direction), hidden state in local memory.
//   sequence n -> base offsets n / n2 * sA + n % n2 * sB (input and output); step stride; group channel offset; elements contiguous.
struct GruSeq {
  const float* x; long long x_sA, x_sB, x_step; int x_grp; int x_limit;      // x_limit > 0: element s * I + i beyond it reads 0 (FA pad)
  float* out; long long o_sA, o_sB, o_step; int o_grp;                         // direction d writes at + d * H
  const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;   // per (group, direction): (3H, I), (3H, H), (3H), (3H)
  int n2, ngroups, dirs, steps, I, H;
  GAN_HD void operator()(long long idx) const {
    const int d = (int)(idx % dirs); long long r = idx / dirs; const int g = (int)(r % ngroups); const long long n = r / ngroups;
    const long long gd = (long long)g * dirs + d;
    const float* wi = w_ih + gd * 3 * H * I; const float* wh = w_hh + gd * 3 * H * H;
    const float* bi = b_ih + gd * 3 * H; const float* bh = b_hh + gd * 3 * H;
    const float* xs = x + (n / n2) * x_sA + (n % n2) * x_sB + (long long)g * x_grp;
    float* os = out + (n / n2) * o_sA + (n % n2) * o_sB + (long long)g * o_grp + (long long)d * H;
    float h[GRU_HMAX], hn[GRU_HMAX];
    for (int j = 0; j < H; ++j) h[j] = 0.f;
    for (int k = 0; k < steps; ++k) {
      const int s = d ? steps - 1 - k : k;
      const float* xv = xs + (long long)s * x_step;
      for (int j = 0; j < H; ++j) {
        float ir = bi[j], iz = bi[H + j], in_ = bi[2 * H + j];
        for (int q = 0; q < I; ++q) {
          const float v = (x_limit > 0 && s * I + q >= x_limit) ? 0.f : xv[q];
          ir += wi[(long long)j * I + q] * v; iz += wi[(long long)(H + j) * I + q] * v; in_ += wi[(long long)(2 * H + j) * I + q] * v;
        }
        float hr = bh[j], hz = bh[H + j], hh = bh[2 * H + j];
        for (int q = 0; q < H; ++q) {
          hr += wh[(long long)j * H + q] * h[q]; hz += wh[(long long)(H + j) * H + q] * h[q]; hh += wh[(long long)(2 * H + j) * H + q] * h[q];
        }
        const float rg = gan::sigmoidf_(ir + hr), zg = gan::sigmoidf_(iz + hz);
        const float ng = tanhf(in_ + rg * hh);
        hn[j] = (1.f - zg) * ng + zg * h[j];
      }
      for (int j = 0; j < H; ++j) { h[j] = hn[j]; os[(long long)s * o_step + j] = hn[j]; }
    }
  }
};

// NCHW <-> (window, frame, band, channel) for the dual-path blocks (:726-731)
struct ToBTFC {
  const float* x; float* y; int Cn, T, Fw;
  GAN_HD void operator()(long long i) const {
    const int c = (int)(i % Cn); long long r = i / Cn; const int f = (int)(r % Fw); r /= Fw; const int t = (int)(r % T); const long long b = r / T;
    y[i] = x[((b * Cn + c) * T + t) * Fw + f];
  }
};
struct ToNCHW {
  const float* x; float* y; int Cn, T, Fw;
  GAN_HD void operator()(long long i) const {
    const int f = (int)(i % Fw); long long r = i / Fw; const int t = (int)(r % T); r /= T; const int c = (int)(r % Cn); const long long b = r / Cn;
    y[i] = x[((b * T + t) * Fw + f) * Cn + c];
  }
};
// LayerNorm over (band, channel) per (window, frame), eps 1e-8, affine (band, channel), + residual (:564-566)
struct LnStats {
  const float* x; float* stat; int n;
  GAN_HD void operator()(long long r) const {
    const float* p = x + r * n;
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += (double)p[k];
    const float mu = (float)(s / n);
    double v = 0.0;
    for (int k = 0; k < n; ++k) { const float d = p[k] - mu; v += (double)(d * d); }
    stat[2 * r] = mu;
    stat[2 * r + 1] = 1.0f / sqrtf((float)(v / n) + 1e-8f);
  }
};
struct LnApplyRes {
  const float* x; const float* stat; const float* g; const float* b; const float* res; float* y; int n;
  GAN_HD void operator()(long long i) const {
    const long long r = i / n; const int k = (int)(i % n);
    y[i] = res[i] + ((x[i] - stat[2 * r]) * stat[2 * r + 1] * g[k] + b[k]);
  }
};

// sigmoid -> ERB merge (:733-737) -> mask on the packed spectrum (:878-879): out (B, 514, T)
struct ErbMask {
  const float* m; const float* ierb; const float* spec; float* out; int T;     // m (B, 1, T, 129) pre-sigmoid; ierb (192, 64) = ierb_fc.weight
  GAN_HD void operator()(long long i) const {
    const int t = (int)(i % T); long long r = i / T; const int f = (int)(r % NB); const long long b = r / NB;
    const float* mv = m + (b * T + t) * NE;
    float g;
    if (f < ERB_LO) g = gan::sigmoidf_(mv[f]);
    else {
      const float* w = ierb + (long long)(f - ERB_LO) * ERB_HI;
      g = 0.f;
      for (int k = 0; k < ERB_HI; ++k) g += gan::sigmoidf_(mv[ERB_LO + k]) * w[k];
    }
    const long long re = (b * 2 * NB + f) * T + t, im = (b * 2 * NB + NB + f) * T + t;
    out[re] = spec[re] * g;
    out[im] = spec[im] * g;
  }
};
template <class TIn>
struct Cast {
  const TIn* in; float* out;
  GAN_HD void operator()(long long i) const { out[i] = (float)in[i]; }
};
// output rule (:906-912): nan_to_num for float inputs; int16: clamp + truncate (the PCM scale sits in the ISTFT reciprocal)
struct OutRule {
  const float* w; void* out; int fix_nan, i16;
  GAN_HD void operator()(long long i) const {
    float v = w[i];
    if (fix_nan) {
      if (v != v) v = 0.f;
      else if (v > 3.4028234e38f) v = 32767.f;
      else if (v < -3.4028234e38f) v = -32768.f;
    }
    if (i16) {
      v = v < -32768.0f ? -32768.0f : (v > 32767.0f ? 32767.0f : v);
      reinterpret_cast<int16_t*>(out)[i] = (int16_t)(int)v;
    } else {
      reinterpret_cast<float*>(out)[i] = v;
    }
  }
};

// ------------------------------------------------------------------------------------------------ architecture (:656-668)
struct BlockCfg { int type, cin, cout, kt, kf, stride, groups, deconv, last, win, wout; };
constexpr int N_ENC = 5;
inline void arch(BlockCfg* enc, BlockCfg* dec) {
  const int types[5] = {0, 2, 1, 2, 1}, strides[5] = {2, 2, 1, 1, 1}, groups[5] = {1, 2, 2, 2, 2}, ch[5] = {12, 24, 24, 32, 16};
  const int kt[5] = {3, 2, 2, 1, 1}, kf[5] = {3, 3, 3, 5, 5}, widths[5] = {65, 33, 33, 33, 33};
  int cin = 1, win = NE;
  for (int i = 0; i < 5; ++i) {
    enc[i] = BlockCfg{types[i], cin, ch[i], kt[i], kf[i], strides[i], groups[i], 0, 0, win, widths[i]};
    cin = ch[i]; win = widths[i];
  }
  for (int i = 0; i < 5; ++i) {
    const int k = 4 - i, cout = k > 0 ? ch[k - 1] : 1, wout = k > 0 ? widths[k - 1] : NE;
    dec[i] = BlockCfg{types[k], cin, cout, kt[k], kf[k], strides[k], groups[k], 1, k == 0, win, wout};
    cin = cout; win = wout;
  }
}

struct CtfaW { const float *ta_wih, *ta_whh, *ta_bih, *ta_bhh, *ta_fc_w, *ta_fc_b, *fa_wih, *fa_whh, *fa_bih, *fa_bhh, *fa_fc_w, *fa_fc_b; };
struct ConvW { const float *w, *b; };
struct ActW { const float *pos, *neg, *bias; };
struct BlockW { ConvW c0, c1, c2; ActW a0, a1; CtfaW ctfa; };      // c0/a0: conv | pconv(1); c1/a1: dconv; c2: pconv2
struct DpW { const float *i_wih, *i_whh, *i_bih, *i_bhh, *i_fc_w, *i_fc_b, *i_ln_g, *i_ln_b, *e_wih, *e_whh, *e_bih, *e_bhh, *e_fc_w, *e_fc_b, *e_ln_g, *e_ln_b; };
struct Weights { const float *erb, *ierb; BlockW enc[5], dec[5]; DpW dp[2]; };

template <class Lookup>
bool bind(Weights& W, Lookup& lk) {
  bool ok = true;
  char nm[64];
  BlockCfg enc[5], dec[5];
  arch(enc, dec);
  auto g = [&](const char* name, size_t n) { const float* p = lk(name, n); ok = ok && p; return p; };
  W.erb = g("erb", (size_t)ERB_HI * (NB - ERB_LO)); W.ierb = g("ierb", (size_t)(NB - ERB_LO) * ERB_HI);
  for (int side = 0; side < 2; ++side)
    for (int i = 0; i < 5; ++i) {
      const BlockCfg& c = side ? dec[i] : enc[i];
      BlockW& b = side ? W.dec[i] : W.enc[i];
      auto gb = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "%s%d.%s", side ? "dec" : "enc", i, k); return g(nm, n); };
      const int C = c.cout;
      const size_t dw = (size_t)C * c.kt * c.kf;                                   // depthwise kernel
      if (c.type == 0) {
        const size_t full = c.deconv ? (size_t)c.cin * (C / c.groups) * c.kt * c.kf : (size_t)C * (c.cin / c.groups) * c.kt * c.kf;
        b.c0.w = gb("c0_w", full); b.c0.b = gb("c0_b", C);
        if (!c.last) { b.a0.pos = gb("a0_pos", (size_t)C * c.wout); b.a0.neg = gb("a0_neg", (size_t)C * c.wout); b.a0.bias = gb("a0_bias", (size_t)C * c.wout); }
      } else {
        b.c0.w = gb("c0_w", (size_t)C * (c.cin / c.groups)); b.c0.b = gb("c0_b", C);
        b.a0.pos = gb("a0_pos", (size_t)C * c.win); b.a0.neg = gb("a0_neg", (size_t)C * c.win); b.a0.bias = gb("a0_bias", (size_t)C * c.win);
        b.c1.w = gb("c1_w", dw); b.c1.b = gb("c1_b", C);
        if (!(c.type == 1 && c.last)) { b.a1.pos = gb("a1_pos", (size_t)C * c.wout); b.a1.neg = gb("a1_neg", (size_t)C * c.wout); b.a1.bias = gb("a1_bias", (size_t)C * c.wout); }
        if (c.type == 2) { b.c2.w = gb("c2_w", (size_t)C * (C / c.groups)); b.c2.b = gb("c2_b", C); }
      }
      const int Hh = 2 * C;
      b.ctfa.ta_wih = gb("ta_wih", (size_t)3 * Hh * C); b.ctfa.ta_whh = gb("ta_whh", (size_t)3 * Hh * Hh);
      b.ctfa.ta_bih = gb("ta_bih", 3 * Hh); b.ctfa.ta_bhh = gb("ta_bhh", 3 * Hh);
      b.ctfa.ta_fc_w = gb("ta_fc_w", (size_t)Hh * C); b.ctfa.ta_fc_b = gb("ta_fc_b", C);
      b.ctfa.fa_wih = gb("fa_wih", 2 * 12 * 4); b.ctfa.fa_whh = gb("fa_whh", 2 * 12 * 4); b.ctfa.fa_bih = gb("fa_bih", 2 * 12); b.ctfa.fa_bhh = gb("fa_bhh", 2 * 12);
      b.ctfa.fa_fc_w = gb("fa_fc_w", 8 * 4); b.ctfa.fa_fc_b = gb("fa_fc_b", 4);
    }
  for (int j = 0; j < 2; ++j) {
    DpW& d = W.dp[j];
    auto gd = [&](const char* k, size_t n) { snprintf(nm, sizeof(nm), "dp%d.%s", j, k); return g(nm, n); };
    d.i_wih = gd("i_wih", 2 * 2 * 12 * 8); d.i_whh = gd("i_whh", 2 * 2 * 12 * 4); d.i_bih = gd("i_bih", 2 * 2 * 12); d.i_bhh = gd("i_bhh", 2 * 2 * 12);
    d.i_fc_w = gd("i_fc_w", 16 * 16); d.i_fc_b = gd("i_fc_b", 16); d.i_ln_g = gd("i_ln_g", 33 * 16); d.i_ln_b = gd("i_ln_b", 33 * 16);
    d.e_wih = gd("e_wih", 2 * 24 * 8); d.e_whh = gd("e_whh", 2 * 24 * 8); d.e_bih = gd("e_bih", 2 * 24); d.e_bhh = gd("e_bhh", 2 * 24);
    d.e_fc_w = gd("e_fc_w", 16 * 16); d.e_fc_b = gd("e_fc_b", 16); d.e_ln_g = gd("e_ln_g", 33 * 16); d.e_ln_b = gd("e_ln_b", 33 * 16);
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------ workspace / sequence
constexpr int CMAX = 32, WMAX = NE;
struct Workspace { float *a, *b, *c, *d, *skip[5], *zt, *zf, *g1, *at, *g2, *af, *stat; };
template <class Alloc>
bool alloc_ws(Workspace& w, int B, int T, Alloc& alloc) {
  const long long map = (long long)B * CMAX * T * WMAX;
  bool ok = true;
  auto a = [&](float*& p, long long n) { p = alloc((size_t)n); ok = ok && p; };
  a(w.a, map); a(w.b, map); a(w.c, map); a(w.d, map);
  for (int i = 0; i < 5; ++i) a(w.skip[i], map);
  a(w.zt, (long long)B * T * CMAX); a(w.zf, (long long)B * T * (WMAX + 3)); a(w.g1, (long long)B * T * 2 * CMAX); a(w.at, (long long)B * T * CMAX);
  a(w.g2, (long long)B * T * (WMAX + 3) * 2); a(w.af, (long long)B * T * (WMAX + 3)); a(w.stat, (long long)B * T * 2);
  return ok;
}

// x (B, C, T, Fw) in place: cTFA (:183-195)
template <class Exec>
void ctfa(Exec& ex, const Workspace& w, const CtfaW& c, float* x, int B, int C, int T, int Fw) {
  const int Fp = (Fw + 3) / 4 * 4, Hf = Fp / 4, Hh = 2 * C;
  const long long bt = (long long)B * T;
  ex.run(bt * C, MeanF2{x, w.zt, C, T, Fw});
  ex.run(bt * Fw, MeanC2{x, w.zf, C, T, Fw});
  // time attention: one sequence per window over T frames, input C, hidden 2C
  ex.run((long long)B, GruSeq{w.zt, (long long)T * C, 0, C, 0, 0, w.g1, (long long)T * Hh, 0, Hh, 0, c.ta_wih, c.ta_whh, c.ta_bih, c.ta_bhh,
                              1, 1, 1, T, C, Hh});
  ex.run(bt * C, Linear{w.g1, Hh, nullptr, c.ta_fc_w, c.ta_fc_b, w.at, C, Hh, C, gan::ACT_SIGMOID, nullptr});
  // frequency attention: one bidirectional sequence per (window, frame) over the groups of 4 bands (zero padded), hidden 4
  ex.run(bt * 2, GruSeq{w.zf, (long long)Fw, 0, 4, 0, Fw, w.g2, (long long)Hf * 8, 0, 8, 0, c.fa_wih, c.fa_whh, c.fa_bih, c.fa_bhh,
                        1, 1, 2, Hf, 4, 4});
  ex.run(bt * Hf * 4, Linear{w.g2, 8, nullptr, c.fa_fc_w, c.fa_fc_b, w.af, 4, 8, 4, gan::ACT_SIGMOID, nullptr});
  ex.run((long long)B * C * T * Fw, CtfaApply{x, w.at, w.af, C, T, Fw, Fp});
}

template <class Exec>
void conv(Exec& ex, const ConvW& cw, const float* x, float* y, int B, int T, int cin, int cout, int fin, int fout, int kt, int kf, int stride,
          int groups, int deconv) {
  const long long n = (long long)B * cout * T * fout;
  if (deconv) ex.run(n, DeconvG{x, cw.w, cw.b, y, cin, cout, T, fin, fout, kt, kf, stride, kf / 2, groups});
  else ex.run(n, ConvG{x, cw.w, cw.b, y, cin, cout, T, fin, fout, kt, kf, stride, kf / 2, groups});
}

// one X-block (:266-274, :344-357, :432-453): x -> out (distinct buffers; t1 / t2 scratch)
template <class Exec>
void block(Exec& ex, const Workspace& w, const BlockCfg& c, const BlockW& bw, const float* x, float* out, float* t1, float* t2, int B, int T) {
  const int C = c.cout;
  const long long no = (long long)B * C * T * c.wout, ni = (long long)B * C * T * c.win;
  if (c.type == 0) {
    float* y = (!c.last && c.groups == 2) ? t1 : out;
    conv(ex, bw.c0, x, y, B, T, c.cin, C, c.win, c.wout, c.kt, c.kf, c.stride, c.groups, c.deconv);
    if (!c.last) ex.run(no, AffAct{y, bw.a0.pos, bw.a0.neg, bw.a0.bias, C, T, c.wout});
    ctfa(ex, w, bw.ctfa, y, B, C, T, c.wout);
    if (y != out) ex.run(no, Shuffle{y, out, C, (long long)T * c.wout});
    return;
  }
  // pointwise (grouped 1x1) + act (+ shuffle) at the INPUT width
  conv(ex, bw.c0, x, t1, B, T, c.cin, C, c.win, c.win, 1, 1, 1, c.groups, 0);
  ex.run(ni, AffAct{t1, bw.a0.pos, bw.a0.neg, bw.a0.bias, C, T, c.win});
  float* h = t1;
  if (c.groups == 2) { ex.run(ni, Shuffle{t1, t2, C, (long long)T * c.win}); h = t2; }
  float* o = h == t1 ? t2 : t1;
  if (c.type == 1) {                                   // XDWSBlock: depthwise (de)conv + act + cTFA
    conv(ex, bw.c1, h, out, B, T, C, C, c.win, c.wout, c.kt, c.kf, c.stride, C, c.deconv);
    if (!c.last) ex.run(no, AffAct{out, bw.a1.pos, bw.a1.neg, bw.a1.bias, C, T, c.wout});
    ctfa(ex, w, bw.ctfa, out, B, C, T, c.wout);
    return;
  }
  // XMBBlocks: depthwise (de)conv + act, pointwise, cTFA, residual, shuffle
  conv(ex, bw.c1, h, o, B, T, C, C, c.win, c.wout, c.kt, c.kf, c.stride, C, c.deconv);
  ex.run(no, AffAct{o, bw.a1.pos, bw.a1.neg, bw.a1.bias, C, T, c.wout});
  float* p = h;                                        // h is free again
  conv(ex, bw.c2, o, p, B, T, C, C, c.wout, c.wout, 1, 1, 1, c.groups, 0);
  ctfa(ex, w, bw.ctfa, p, B, C, T, c.wout);
  if (c.cin == C && c.stride == 1) ex.run(no, Add2{p, x, p});
  if (!c.last && c.groups == 2) ex.run(no, Shuffle{p, out, C, (long long)T * c.wout});
  else ex.run(no, Add2{p, nullptr, out});              // (never taken: every XMB block of this architecture shuffles)
}

// dual-path block on (B, T, 33, 16) (:557-574): x -> y (distinct), scratch s1 / s2
template <class Exec>
void dpgrnn(Exec& ex, const Workspace& w, const DpW& d, const float* x, float* y, float* s1, float* s2, int B, int T) {
  const int Wd = 33, C = 16;
  const long long px = (long long)B * T * Wd, bt = (long long)B * T;
  // intra: per (window, frame), two groups of 8 channels, bidirectional hidden 4 -> [g0: fwd4 bwd4 | g1: fwd4 bwd4]
  ex.run(bt * 2 * 2, GruSeq{x, (long long)Wd * C, 0, C, 8, 0, s1, (long long)Wd * C, 0, C, 8, d.i_wih, d.i_whh, d.i_bih, d.i_bhh, 1, 2, 2, Wd, 8, 4});
  ex.run(px * C, Linear{s1, C, nullptr, d.i_fc_w, d.i_fc_b, s2, C, C, C, gan::ACT_NONE, nullptr});
  ex.run(bt, LnStats{s2, w.stat, Wd * C});
  ex.run(px * C, LnApplyRes{s2, w.stat, d.i_ln_g, d.i_ln_b, x, s1, Wd * C});                 // s1 = intra_out
  // inter: per (window, band), two groups of 8 channels, unidirectional hidden 8, over frames
  ex.run((long long)B * Wd * 2, GruSeq{s1, (long long)T * Wd * C, C, (long long)Wd * C, 8, 0, s2, (long long)T * Wd * C, C, (long long)Wd * C, 8,
                                       d.e_wih, d.e_whh, d.e_bih, d.e_bhh, Wd, 2, 1, T, 8, 8});
  ex.run(px * C, Linear{s2, C, nullptr, d.e_fc_w, d.e_fc_b, y, C, C, C, gan::ACT_NONE, nullptr});
  ex.run(bt, LnStats{y, w.stat, Wd * C});
  ex.run(px * C, LnApplyRes{y, w.stat, d.e_ln_g, d.e_ln_b, s1, y, Wd * C});
}

// spec (B, 514, T) packed STFT -> out (B, 514, T) masked spectrum
template <class Exec>
void forward(Exec& ex, const Workspace& w, const Weights& W, const float* spec, float* out, int B, int T) {
  BlockCfg enc[5], dec[5];
  arch(enc, dec);
  char tag[16];
  ex.run((long long)B * T * NE, PowerLogErb{spec, W.erb, w.a, T});
  const float* x = w.a;
  for (int i = 0; i < 5; ++i) {
    block(ex, w, enc[i], W.enc[i], x, w.skip[i], w.b, w.c, B, T);
    x = w.skip[i];
    snprintf(tag, sizeof(tag), "enc%d", i);
    ex.mark("", tag, x, (long long)B * enc[i].cout * T * enc[i].wout);
  }
  const long long px = (long long)B * T * 33 * 16;
  ex.run(px, ToBTFC{x, w.a, 16, T, 33});
  dpgrnn(ex, w, W.dp[0], w.a, w.b, w.c, w.d, B, T);
  ex.mark("", "dp0", w.b, px);
  dpgrnn(ex, w, W.dp[1], w.b, w.a, w.c, w.d, B, T);
  ex.mark("", "dp1", w.a, px);
  ex.run(px, ToNCHW{w.a, w.b, 16, T, 33});
  float* cur = w.b;
  float* other = w.a;
  for (int i = 0; i < 5; ++i) {
    const int k = 4 - i;
    const long long n = (long long)B * dec[i].cin * T * dec[i].win;
    ex.run(n, Add2{cur, w.skip[k], cur});
    float* t1 = w.c;
    float* t2 = w.skip[k];                               // the skip tensor is consumed: its buffer is scratch from here on
    block(ex, w, dec[i], W.dec[i], cur, other, t1, t2, B, T);
    float* t = cur; cur = other; other = t;
    snprintf(tag, sizeof(tag), "dec%d", i);
    ex.mark("", tag, cur, (long long)B * dec[i].cout * T * dec[i].wout);
  }
  ex.run((long long)B * NB * T, ErbMask{cur, W.ierb, spec, out, T});
}

}  // namespace uln
