// ZipEnhancer backbone (BASELINE configs[1]): the launch sequence of `ZipEnhancer.forward`
// (reference ZipEnhancer/Export_ZipEnhancer.py:818-927) between the spectral features and the recombine step --
// DenseEncoder (:851-853, `_dense_block` :701-723), four dual-path Zipformer2 encoders (`_dualpath_encoder` :770-781,
// `_downsampled_encoder` :783-816, layer forward :143-187 with the overrides :118-339), mask / phase decoders
// (:725-768, :866-880) -- templated on the executor.
//
// * Dense contractions are `LinOp`s: C = A W^T with A in fp32, either plain rows or a K-concatenation of shifted channel
//   windows of a padded channel-last map (the (2,3) dilated causal convs as implicit GEMMs: no im2col buffer).  libadn runs
//   them on the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) in its fp32-A mode: TMA brings the fp32 tile into shared memory and
//   converter warps split it into the tf32 hi / lo operand tiles in place, so activations exist in HBM once, as fp32.  The
//   host harness (tests/harness/zipenh_host.cpp) runs the same LinOps with plain loops, so the addressing is checked
//   without a GPU.
// * Everything else is a one-output-per-thread functor (`ex.run(count, functor)`); the CUDA executor replaces the
//   attention functors (AttnW, SaApply, NlApply), the gated depthwise conv and the final norm by cooperative kernels (csrc/zipenh.cu), which the GPU
//   tests compare against the same stage dumps.
// * Layout: tokens are channel-last rows of 64 floats in (window, frame, sub-band) order for BOTH path directions; a layer
//   over sub-bands and a layer over frames differ only in the `SeqMap` that turns (sequence, position) into a token row --
//   the reference's permute / contiguous round trips (:775-781) do not exist here.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDACC__)
#define ZIP_HD __host__ __device__ __forceinline__
#else
#define ZIP_HD inline
#endif

namespace zip {

constexpr int C = 64;                 // dense_channel / encoder_dim
constexpr int HEADS = 4, QD = 12, PD = 4, VD = 12;
constexpr int HB = 2 * QD + PD;       // per-head block [q | k | p] of the attention projection
constexpr int AP = HEADS * HB;        // 112
constexpr int SV = HEADS * VD;        // 48
constexpr int NH = 3 * C / 4;         // NonlinAttention hidden (48)
constexpr int FF1 = 192, FF2 = 256, FF3 = 320;
constexpr int DWK = 15;               // ConvolutionModule depthwise kernel
constexpr int FB = 201, FQ = 101;     // bins, sub-bands
constexpr int FPE = 204;              // padded width of the full-resolution maps: [0 | f 0..200 | 0 0]
constexpr int FPD = 104;              // padded width of the decoder maps:        [0 | f 0..100 | 0 0]
constexpr int FH = 102;               // row grid of the stride-2 conv (FPE / 2)
constexpr int DEPTH = 4;
constexpr int SLOTC = C * DEPTH;      // dense-block buffer: channels [d2 | d1 | d0 | x]
constexpr int UPF = 2;                // sub-pixel factor
constexpr int FU = FQ * UPF;          // 202
constexpr int NENC = 4;
constexpr float IN_EPS = 1e-5f;

enum { ACT_NONE = 0, ACT_SWOOSH_L = 7, ACT_SWOOSH_R = 8 };   // values shared with tc::ACT_*

ZIP_HD float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// softplus(x - o) - 0.08 x; the activation's constant lives in the next bias (:131-140, :446-455)
ZIP_HD float swoosh(float x, float o) {
  const float y = x - o;
  const float sp = y > 20.f ? y : log1pf(expf(y));
  return sp - 0.08f * x;
}
ZIP_HD float actf(float v, int a) { return a == ACT_SWOOSH_L ? swoosh(v, 4.0f) : a == ACT_SWOOSH_R ? swoosh(v, 1.0f) : v; }
ZIP_HD void split_tf32(float v, float& hi, float& lo) {
  union { float f; uint32_t u; } x;
  x.f = v;
  x.u &= 0xFFFFE000u;
  hi = x.f;
  lo = v - hi;
}

// (sequence n, position s) -> token row
struct SeqMap {
  int n2; long long sA, sB, sS; int S;
  ZIP_HD long long tok(long long n, int s) const { return (n / n2) * sA + (n % n2) * sB + (long long)s * sS; }
};
inline SeqMap seq_over_f(int T, int F) { (void)T; return SeqMap{1 << 30, 0, (long long)F, 1, F}; }          // sequences = (window, frame)
inline SeqMap seq_over_t(int T, int F) { return SeqMap{F, (long long)T * F, 1, (long long)F, T}; }           // sequences = (window, sub-band)

// ------------------------------------------------------------------------------------------------ dense contraction
struct LinW {
  const float* w;          // (n_pad, k_pad) fp32, zero padded
  const float* b;          // (n_pad) or null
  int n_pad, k_pad;
};
struct LinOp {
  const float* a;
  long long a_sB, a_sR;    // element (chunk b, row r, k) at a[b*a_sB + (r + a_r0)*a_sR + k]
  int a_r0;
  int a_ke;                // readable k extent of one A row (beyond it: zeros)
  int a_rows;              // rows of one chunk that exist in memory (outside: zeros), counted in units of a_sR
  int chunks, rows;        // output rows m = b*rows + r
  int K, N;
  int taps, tap_c, a_k0;   // taps > 0: k = tap*tap_c + c reads channel a_k0 + c of row r + a_r0 + tap_shift[tap]
  int tap_shift[6];
  LinW W;
  int act;
  const float* resid;      // v += resid[m*ldc + n]
  const float* resid2;     // v = resid2[m*ldc + n] + (v - resid2[m*ldc + n]) * colscale[n]
  const float* colscale;
  float* Cf;               // fp32 output
  long long ldc;
};
inline LinOp lin_rows(const float* a, long long lda, int K, long long M, const LinW& W, int N) {
  LinOp g{};
  g.a = a; g.a_sB = 0; g.a_sR = lda; g.a_r0 = 0; g.a_ke = K; g.a_rows = (int)M;
  g.chunks = 1; g.rows = (int)M; g.K = K; g.N = N; g.taps = 0; g.W = W; g.act = ACT_NONE; g.ldc = N;
  return g;
}

// reference evaluation of a LinOp (host harness; also documents the semantics the tcgen05 path implements)
inline void lin_ref(const LinOp& g) {
#pragma omp parallel for schedule(static)
  for (long long m = 0; m < (long long)g.chunks * g.rows; ++m) {
    const long long b = m / g.rows; const int r = (int)(m % g.rows);
    for (int n = 0; n < g.N; ++n) {
      double acc = 0.0;
      const float* w = g.W.w + (long long)n * g.W.k_pad;
      if (g.taps > 0) {
        for (int tp = 0; tp < g.taps; ++tp) {
          const long long rr = (long long)r + g.a_r0 + g.tap_shift[tp];
          if (rr < 0 || rr >= g.a_rows) continue;
          const long long o = b * g.a_sB + rr * g.a_sR + g.a_k0;
          for (int c = 0; c < g.tap_c; ++c) acc += (double)g.a[o + c] * (double)w[tp * g.tap_c + c];
        }
      } else {
        const long long rr = (long long)r + g.a_r0;
        if (rr >= 0 && rr < g.a_rows) {
          const long long o = b * g.a_sB + rr * g.a_sR;
          for (int k = 0; k < g.K && k < g.a_ke; ++k) acc += (double)g.a[o + k] * (double)w[k];
        }
      }
      float v = (float)acc;
      if (g.W.b) v += g.W.b[n];
      v = actf(v, g.act);
      const long long oc = m * g.ldc + n;
      if (g.resid) v += g.resid[oc];
      if (g.resid2) v = g.resid2[oc] + (v - g.resid2[oc]) * g.colscale[n];
      g.Cf[oc] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ functors
// p -> p / d, returns p % d; 32-bit arithmetic whenever p fits (a 64-bit division by a run-time divisor is ~100 GPU instructions,
// and the one-output-per-thread functors below do two or three of them per element)
ZIP_HD int split_mod(long long& p, int d) {
  if ((unsigned long long)p < 0x100000000ull) {
    const unsigned q = (unsigned)p / (unsigned)d;
    const int r = (int)((unsigned)p - q * (unsigned)d);
    p = (long long)q;
    return r;
  }
  const int r = (int)(p % d);
  p /= d;
  return r;
}

// dense_conv_1 (:851): 1x1 conv over the planar features (window, 2, frame, bin) -> raw channel-last padded map
struct FeatConv {
  const float* feat; const float* w; const float* b; float* raw; int T;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % C); long long p = i / C; const int f = split_mod(p, FB); const int t = split_mod(p, T); const long long bb = p;
    const float* x = feat + (bb * 2 * T + t) * FB + f;
    raw[((bb * T + t) * FPE + f + 1) * C + c] = b[c] + w[c * 2] * x[0] + w[c * 2 + 1] * x[(long long)T * FB];
  }
};

// InstanceNorm2d statistics of a raw conv output on a padded grid: per (window, frame, raw channel) double sums over the valid
// columns, then per (window, channel) over frames and over the `pool` raw channels that the sub-pixel shuffle merges
struct InPart {
  const float* raw; int ld; int W; int lo, hi; double* part;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % ld); const long long bt = i / ld;
    const float* p = raw + (bt * W) * ld + c;
    double s = 0.0, s2 = 0.0;
    for (int f = lo; f < hi; ++f) { const double v = (double)p[(long long)f * ld]; s += v; s2 += v * v; }
    part[2 * i] = s; part[2 * i + 1] = s2;
  }
};
struct InFin {
  const double* part; int ld; int T; int nvalid; int pool; float* stat;
  ZIP_HD void operator()(long long i) const {
    const int cn = ld / pool; const int c = (int)(i % cn); const long long b = i / cn;
    double s = 0.0, s2 = 0.0;
    for (int t = 0; t < T; ++t)
      for (int u = 0; u < pool; ++u) { const double* p = part + 2 * ((b * T + t) * ld + c * pool + u); s += p[0]; s2 += p[1]; }
    const double cnt = (double)T * nvalid * pool, mu = s / cnt;
    double var = s2 / cnt - mu * mu;
    var = var > 0.0 ? var : 0.0;
    stat[2 * i] = (float)mu;
    stat[2 * i + 1] = (float)(1.0 / sqrt(var + (double)IN_EPS));
  }
};
// normalise + affine + PReLU; source grid (T, Ws) with valid columns from src_lo, `pool` raw channels per output channel
// (output column = source column * pool + u); destination grid (T, Wd) with the outputs at columns dst_lo.., zeros elsewhere;
// written at channel offset coff of ldd-wide pixels
struct InApply {
  const float* raw; int ld; int Ws; int src_lo; int pool; const float* stat; const float* w; const float* b; const float* slope;
  float* of; int ldd; int coff; int Wd; int dst_lo; int nout; int T;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % C); long long p = i / C; const int fd = split_mod(p, Wd); const int t = split_mod(p, T); const long long bb = p;
    float v = 0.f;
    const int k = fd - dst_lo;
    if (k >= 0 && k < nout) {
      const int fs = src_lo + k / pool, u = k % pool;
      const float* st = stat + 2 * (bb * C + c);
      v = (raw[((bb * T + t) * Ws + fs) * ld + c * pool + u] - st[0]) * st[1] * w[c] + b[c];
      v = v >= 0.f ? v : slope[c] * v;
    }
    of[((bb * T + t) * Wd + fd) * ldd + coff + c] = v;
  }
};

// compact tokens (window, frame, FQ, C) -> the x slot of the two decoder dense-block buffers (padded grid, zero pad columns)
struct PadCopy {
  const float* x; float *da, *db; int T;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % C); long long p = i / C; const int fd = split_mod(p, FPD); const long long bt = p;
    float v = 0.f;
    if (fd >= 1 && fd <= FQ) v = x[(bt * FQ + fd - 1) * C + c];
    const long long o = (bt * FPD + fd) * SLOTC + (DEPTH - 1) * C + c;
    da[o] = v; db[o] = v;
  }
};

// rows of the attention-weight tensor are padded to a multiple of 4 floats (16-byte aligned rows for vector loads); pad = 0
ZIP_HD int aw_ld(int S) { return (S + 3) & ~3; }

// RelPositionMultiheadAttentionWeights (:232-296): one thread = one (sequence, head, query) row of softmax(q.k + p.R[j - i])
struct AttnW {
  const float* ap; SeqMap sm; const float* pos; float* aw;      // pos (HEADS, PD, 2S-1); aw (sequence, head, S, aw_ld(S))
  ZIP_HD void operator()(long long idx) const {
    const int S = sm.S;
    const int i = (int)(idx % S); long long r = idx / S; const int h = (int)(r % HEADS); const long long n = r / HEADS;
    const float* qi = ap + sm.tok(n, i) * AP + h * HB;
    float q[QD], p[PD];
    for (int d = 0; d < QD; ++d) q[d] = qi[d];
    for (int d = 0; d < PD; ++d) p[d] = qi[2 * QD + d];
    float* row = aw + ((n * HEADS + h) * S + i) * (long long)aw_ld(S);
    for (int j = S; j < aw_ld(S); ++j) row[j] = 0.f;
    const float* ph = pos + (long long)h * PD * (2 * S - 1) + (S - 1 - i);
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) {
      const float* kj = ap + sm.tok(n, j) * AP + h * HB + QD;
      float a = 0.f;
      for (int d = 0; d < QD; ++d) a += q[d] * kj[d];
      float e = 0.f;
      for (int d = 0; d < PD; ++d) e += p[d] * ph[(long long)d * (2 * S - 1) + j];
      a += e;
      row[j] = a;
      mx = a > mx ? a : mx;
    }
    float sum = 0.f;
    for (int j = 0; j < S; ++j) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
    const float inv = 1.0f / sum;
    for (int j = 0; j < S; ++j) row[j] *= inv;
  }
};

// SelfAttention value product (:298-308): out[tok(n,i), c] = sum_j aw[n, c / VD, i, j] * v[tok(n,j), c]   (width SV)
struct SaApply {
  const float* aw; SeqMap sm; const float* v; float* out;
  ZIP_HD void operator()(long long idx) const {
    const int S = sm.S;
    const int c = (int)(idx % SV); long long r = idx / SV; const int i = (int)(r % S); const long long n = r / S;
    const float* a = aw + ((n * HEADS + c / VD) * S + i) * (long long)aw_ld(S);
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc += a[j] * v[sm.tok(n, j) * SV + c];
    out[sm.tok(n, i) * SV + c] = acc;
  }
};
// NonlinAttention core (:310-326) on the fused projection np = [s | x_mid | y] (width 3 NH): head 0 of the attention weights
// mixes x_mid * tanh(s) over the sequence, then the y gate   (width NH)
struct NlApply {
  const float* aw; SeqMap sm; const float* np; float* out;
  ZIP_HD void operator()(long long idx) const {
    const int S = sm.S;
    const int c = (int)(idx % NH); long long r = idx / NH; const int i = (int)(r % S); const long long n = r / S;
    const float* a = aw + ((n * HEADS) * S + i) * (long long)aw_ld(S);
    float acc = 0.f;
    for (int j = 0; j < S; ++j) {
      const float* pj = np + sm.tok(n, j) * (3 * NH);
      acc += a[j] * (pj[NH + c] * tanhf(pj[c]));
    }
    const long long o = sm.tok(n, i);
    out[o * NH + c] = acc * np[o * (3 * NH) + 2 * NH + c];
  }
};

// ConvolutionModule core (:328-339) on the fused projection cp = [x_mid | gate] (width 2 C): u = x_mid * sigmoid(gate),
// depthwise Conv1d(k 15, 'same') along the sequence, then the SwooshR of the out projection (:131-140)
struct GluDwConv {
  const float* cp; SeqMap sm; const float* w; const float* b; float* out;     // w (C, DWK)
  ZIP_HD void operator()(long long idx) const {
    const int S = sm.S;
    const int c = (int)(idx % C); long long r = idx / C; const int s = (int)(r % S); const long long n = r / S;
    float acc = b[c];
    for (int k = 0; k < DWK; ++k) {
      const int sj = s + k - DWK / 2;
      if (sj < 0 || sj >= S) continue;
      const float* pj = cp + sm.tok(n, sj) * (2 * C);
      acc += w[c * DWK + k] * (pj[c] * sigmoidf_(pj[C + c]));
    }
    out[sm.tok(n, s) * C + c] = swoosh(acc, 1.0f);
  }
};

// final BiasNorm + layer bypass + dual-path bypass, fused as the reference fuses them (:176-184, :659-676):
// x0 <- x / ||x - bias||_2 * nscale + x0 * rscale   (in place on the layer input)
struct NormBypass {
  const float* x; float* x0; const float* nbias; const float* nscale; const float* rscale;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % C); const long long r = i / C;
    const float* xr = x + r * C;
    float ss = 0.f;
    for (int k = 0; k < C; ++k) { const float d = xr[k] - nbias[k]; ss += d * d; }
    x0[i] = (xr[c] / sqrtf(ss)) * nscale[c] + x0[i] * rscale[c];
  }
};

// SimpleDownsample over frames then sub-bands (:194-220, :788-791): last position repeated up to a multiple of ds
struct Down {
  const float* x; const float* wt; const float* wf; int ds; int T, F, Td, Fd; float* of;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % C); long long p = i / C; const int fj = split_mod(p, Fd); const int ti = split_mod(p, Td); const long long bb = p;
    float acc = 0.f;
    for (int e = 0; e < ds; ++e) {
      const int f = fj * ds + e < F ? fj * ds + e : F - 1;
      float a = 0.f;
      for (int k = 0; k < ds; ++k) {
        const int t = ti * ds + k < T ? ti * ds + k : T - 1;
        a += x[((bb * T + t) * F + f) * C + c] * wt[k];
      }
      acc += a * wf[e];
    }
    of[i] = acc;
  }
};
// out-combiner (:806-816): x0 <- x0 * (1 - scale) + up(y) * scale, nearest-neighbour upsampling of both axes
struct UpCombine {
  const float* y; const float* scale; const float* rscale; int ds; int T, F, Td, Fd; float* x0;
  ZIP_HD void operator()(long long i) const {
    const int c = (int)(i % C); long long p = i / C; const int f = split_mod(p, F); const int t = split_mod(p, T); const long long bb = p;
    x0[i] = x0[i] * rscale[c] + (y[((bb * Td + t / ds) * Fd + f / ds) * C + c] * scale[c]);
  }
};

// decoder heads (:866-880): (1,2) convs over the up-sampled map (window, frame, FU, C): NOUT = 1 mask / 2 phase (r, i);
// out planar (window, NOUT, frame, bin)
struct Head {
  const float* up; const float* w; const float* b; int nout; float* out; int T;     // w (nout, 2, C)
  ZIP_HD void operator()(long long i) const {
    const int f = (int)(i % FB); long long p = i / FB; const int t = split_mod(p, T); const int o = split_mod(p, nout); const long long bb = p;
    const float* x = up + ((bb * T + t) * FU + f) * C;
    const float* wo = w + o * 2 * C;
    float acc = b[o];
    for (int k = 0; k < 2 * C; ++k) acc += wo[k] * x[k];
    out[i] = acc;
  }
};

// ------------------------------------------------------------------------------------------------ weights
struct NormAct { const float *w, *b, *slope; };
struct DenseW { LinW conv[DEPTH]; NormAct na[DEPTH]; };
struct LayerW {
  LinW attn_in, ff1_in, ff1_out, nl_in, nl_out, sa1_in, sa1_out, cv1_in, cv1_out, ff2_in, ff2_out, sa2_in, sa2_out, cv2_in, cv2_out,
      ff3_in, ff3_out;
  const float *dw1_w, *dw1_b, *dw2_w, *dw2_b, *mid_scale, *norm_bias, *norm_scale, *res_scale, *pos;
};
struct EncW {
  int ds;
  LayerW f, t;
  const float *down_t, *down_f, *comb_scale, *comb_rscale;
};
struct Weights {
  const float *c1_w, *c1_b; NormAct c1_na;
  DenseW enc_dense;
  LinW c2; NormAct c2_na;
  EncW enc[NENC];
  DenseW mask_dense, phase_dense;
  LinW mask_up, phase_up; NormAct mask_up_na, phase_up_na;
  const float *mask_out_w, *mask_out_b, *phase_out_w, *phase_out_b;
};

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int pad_to(int a, int b) { return ceil_div(a, b) * b; }

// Lookup: const float* lk(const char* name, size_t expected_count)  (0 = any; returns null and records the error)
template <class Lookup>
bool bind(Weights& W, int T, const int* ds, Lookup& lk) {
  char nm[96];
  bool ok = true;
  auto get = [&](const char* pre, const char* leaf, size_t n) -> const float* {
    snprintf(nm, sizeof(nm), "%s.%s", pre, leaf);
    const float* p = lk(nm, n);
    if (!p) ok = false;
    return p;
  };
  auto lin = [&](const char* pre, const char* leaf, int N, int K, bool bias = true) {
    LinW w{};
    w.n_pad = pad_to(N, 64); w.k_pad = pad_to(K, 32);
    char p2[64];
    snprintf(p2, sizeof(p2), "%s.%s", pre, leaf);
    w.w = get(p2, "w", (size_t)w.n_pad * w.k_pad);
    w.b = bias ? get(p2, "b", (size_t)w.n_pad) : nullptr;
    return w;
  };
  auto na = [&](const char* pre, const char* leaf) {
    char p2[64];
    snprintf(p2, sizeof(p2), "%s.%s", pre, leaf);
    NormAct n{};
    n.w = get(p2, "in_w", C); n.b = get(p2, "in_b", C); n.slope = get(p2, "prelu", C);
    return n;
  };
  auto dense = [&](const char* pre, DenseW& d) {
    for (int i = 0; i < DEPTH; ++i) {
      char leaf[16];
      snprintf(leaf, sizeof(leaf), "d%d", i);
      d.conv[i] = lin(pre, leaf, C, 6 * C * (i + 1));
      d.na[i] = na(pre, leaf);
    }
  };
  W.c1_w = get("enc.c1", "w", 2 * C); W.c1_b = get("enc.c1", "b", C); W.c1_na = na("enc", "c1");
  dense("enc", W.enc_dense);
  W.c2 = lin("enc", "c2", C, 3 * C); W.c2_na = na("enc", "c2");
  int Tc = T, Fc = FQ;
  for (int k = 0; k < NENC; ++k) {
    EncW& e = W.enc[k];
    e.ds = ds[k];
    const int Tk = ceil_div(Tc, e.ds), Fk = ceil_div(Fc, e.ds);
    char pre[32];
    snprintf(pre, sizeof(pre), "ts%d", k);
    if (e.ds > 1) {
      e.down_t = get(pre, "down_t", e.ds); e.down_f = get(pre, "down_f", e.ds);
      e.comb_scale = get(pre, "comb_scale", C); e.comb_rscale = get(pre, "comb_rscale", C);
    }
    for (int dir = 0; dir < 2; ++dir) {
      LayerW& l = dir ? e.t : e.f;
      const int S = dir ? Tk : Fk;
      char lp[40];
      snprintf(lp, sizeof(lp), "ts%d.%c", k, dir ? 't' : 'f');
      l.attn_in = lin(lp, "attn_in", AP, C); l.ff1_in = lin(lp, "ff1_in", FF1, C); l.ff1_out = lin(lp, "ff1_out", C, FF1);
      l.nl_in = lin(lp, "nl_in", 3 * NH, C); l.nl_out = lin(lp, "nl_out", C, NH);
      l.sa1_in = lin(lp, "sa1_in", SV, C); l.sa1_out = lin(lp, "sa1_out", C, SV);
      l.cv1_in = lin(lp, "cv1_in", 2 * C, C); l.cv1_out = lin(lp, "cv1_out", C, C);
      l.ff2_in = lin(lp, "ff2_in", FF2, C); l.ff2_out = lin(lp, "ff2_out", C, FF2);
      l.sa2_in = lin(lp, "sa2_in", SV, C); l.sa2_out = lin(lp, "sa2_out", C, SV);
      l.cv2_in = lin(lp, "cv2_in", 2 * C, C); l.cv2_out = lin(lp, "cv2_out", C, C);
      l.ff3_in = lin(lp, "ff3_in", FF3, C); l.ff3_out = lin(lp, "ff3_out", C, FF3);
      l.dw1_w = get(lp, "dw1.w", C * DWK); l.dw1_b = get(lp, "dw1.b", C);
      l.dw2_w = get(lp, "dw2.w", C * DWK); l.dw2_b = get(lp, "dw2.b", C);
      l.mid_scale = get(lp, "mid_scale", C); l.norm_bias = get(lp, "norm_bias", C);
      l.norm_scale = get(lp, "norm_scale", C); l.res_scale = get(lp, "res_scale", C);
      l.pos = get(lp, "pos", (size_t)HEADS * PD * (2 * S - 1));
    }
  }
  dense("mask", W.mask_dense); dense("phase", W.phase_dense);
  W.mask_up = lin("mask", "up", UPF * C, 3 * C); W.mask_up_na = na("mask", "up");
  W.phase_up = lin("phase", "up", UPF * C, 3 * C); W.phase_up_na = na("phase", "up");
  W.mask_out_w = get("mask.out", "w", 2 * C); W.mask_out_b = get("mask.out", "b", 1);
  W.phase_out_w = get("phase.out", "w", 2 * 2 * C); W.phase_out_b = get("phase.out", "b", 2);
  return ok;
}

// ------------------------------------------------------------------------------------------------ workspace
struct Workspace {
  float *encbuf, *d3, *hp, *p64, *decm, *decp;
  float *raw, *stat, *x0, *x, *t1, *t2, *aw, *x0d, *xd, *up;
  double* part;
};
inline size_t aw_floats(int B, int T) {
  const size_t a = (size_t)B * FQ * HEADS * T * aw_ld(T), b = (size_t)B * T * HEADS * FQ * aw_ld(FQ);
  return a > b ? a : b;
}
// Alloc: float* alloc(size_t n_floats)  (null on failure).  Every buffer of the dense-block / stride-conv A operands gets
// slack behind it: the row-gather of the last rows reads (zero-weighted or masked) elements past the logical end.
template <class Alloc>
bool alloc_ws(Workspace& w, int B, int T, Alloc& alloc) {
  const size_t b = (size_t)B, pxE = b * T * FPE, pxD = b * T * FPD, M = b * T * FQ, Md = b * ceil_div(T, 2) * ceil_div(FQ, 2);
  const size_t slack = 4096;
  auto pl = [&](float*& p, size_t n) { p = alloc(n + slack); return p != nullptr; };
  bool ok = pl(w.encbuf, pxE * SLOTC) && pl(w.d3, pxE * C) && pl(w.hp, M * FF3) && pl(w.p64, M * C) && pl(w.decm, pxD * SLOTC) &&
            pl(w.decp, pxD * SLOTC);
  if (!ok) return false;
  const size_t raw_n = pxE * C > pxD * UPF * C ? pxE * C : pxD * UPF * C;
  w.raw = alloc(raw_n + slack);
  w.stat = alloc(b * C * 2);
  w.part = reinterpret_cast<double*>(alloc(b * T * UPF * C * 2 * 2));
  w.x0 = alloc(M * C); w.x = alloc(M * C); w.t1 = alloc(M * 3 * NH); w.t2 = alloc(M * C);
  w.aw = alloc(aw_floats(B, T));
  w.x0d = alloc(Md * C); w.xd = alloc(Md * C);
  w.up = alloc(b * T * FU * C);
  return w.raw && w.stat && w.part && w.x0 && w.x && w.t1 && w.t2 && w.aw && w.x0d && w.xd && w.up;
}
inline size_t ws_floats(int B, int T) {
  const size_t b = (size_t)B, pxE = b * T * FPE, pxD = b * T * FPD, M = b * T * FQ, Md = b * ceil_div(T, 2) * ceil_div(FQ, 2);
  const size_t raw_n = pxE * C > pxD * UPF * C ? pxE * C : pxD * UPF * C;
  return pxE * SLOTC + pxE * C + M * FF3 + M * C + 2 * pxD * SLOTC + raw_n + b * C * 2 + b * T * UPF * C * 4 +
         M * (2 * C + 3 * NH + C) + aw_floats(B, T) + 2 * Md * C + b * T * FU * C;
}

// ------------------------------------------------------------------------------------------------ launch sequence
// InstanceNorm2d + PReLU of a raw conv output (see InPart / InFin / InApply)
template <class Exec>
void inorm(Exec& ex, Workspace& w, int B, int T, int ld, int Ws, int src_lo, int nsrc, int pool, const NormAct& na, float* of,
           int ldd, int coff, int Wd, int dst_lo) {
  ex.run((long long)B * T * ld, InPart{w.raw, ld, Ws, src_lo, src_lo + nsrc, w.part});
  ex.run((long long)B * (ld / pool), InFin{w.part, ld, T, nsrc, pool, w.stat});
  ex.run((long long)B * T * Wd * C, InApply{w.raw, ld, Ws, src_lo, pool, w.stat, na.w, na.b, na.slope, of, ldd, coff, Wd, dst_lo, nsrc * pool, T});
}

// DenseBlockV2 on a padded grid of width Wp (valid columns 1..nvalid): layer i reads the last C*(i+1) channels of `buf`
// and writes slot DEPTH-2-i, the last layer writes `last` (C-wide pixels)
template <class Exec>
void dense_block(Exec& ex, Workspace& w, const DenseW& d, float* buf, float* last, int B, int T, int Wp, int nvalid, const char* tag) {
  for (int i = 0; i < DEPTH; ++i) {
    const int dil = 1 << i, cin = C * (i + 1);
    LinOp g{};
    g.a = buf; g.a_sB = (long long)T * Wp * SLOTC; g.a_sR = SLOTC; g.a_r0 = 0; g.a_ke = SLOTC; g.a_rows = T * Wp;
    g.chunks = B; g.rows = T * Wp; g.K = 6 * cin; g.N = C;
    g.taps = 6; g.tap_c = cin; g.a_k0 = SLOTC - cin;
    for (int kt = 0; kt < 2; ++kt)
      for (int kf = 0; kf < 3; ++kf) g.tap_shift[kt * 3 + kf] = (kt - 1) * dil * Wp + (kf - 1);
    g.W = d.conv[i]; g.act = ACT_NONE; g.Cf = w.raw; g.ldc = C;
    ex.gemm(g, "zip_dense_conv");
    if (i + 1 < DEPTH) inorm(ex, w, B, T, C, Wp, 1, nvalid, 1, d.na[i], buf, SLOTC, (DEPTH - 2 - i) * C, Wp, 1);
    else inorm(ex, w, B, T, C, Wp, 1, nvalid, 1, d.na[i], last, C, 0, Wp, 1);
    char nm[32];
    snprintf(nm, sizeof(nm), "%s.d%d", tag, i);
    ex.mark_strided(nm, i + 1 < DEPTH ? buf : last, (long long)B * T * Wp, i + 1 < DEPTH ? SLOTC : C, i + 1 < DEPTH ? (DEPTH - 2 - i) * C : 0, C);
  }
}

// one Zipformer2EncoderLayer over M tokens with the sequence structure `sm` (:143-187); x0 holds the layer input on entry and
// the layer output on exit, x is the running residual stream
template <class Exec>
void zip_layer(Exec& ex, Workspace& w, const LayerW& L, float* x0, float* x, long long M, const SeqMap& sm, long long nseq, const char* tag) {
  const int S = sm.S;
  char nm[48];
  auto mark = [&](const char* leaf, const float* p, int width) {
    snprintf(nm, sizeof(nm), "%s.%s", tag, leaf);
    ex.mark(nm, p, M * width);
  };
  // attention projection and weights
  LinOp g = lin_rows(x0, C, C, M, L.attn_in, AP);
  g.Cf = w.t1; g.ldc = AP;
  ex.gemm(g, "zip_attn_in");
  ex.run(nseq * HEADS * S, AttnW{w.t1, sm, L.pos, w.aw});
  snprintf(nm, sizeof(nm), "%s.aw", tag);
  ex.mark_strided(nm, w.aw, nseq * HEADS * S, aw_ld(S), 0, S);
  // feed-forward modules: in projection + SwooshL, out projection + residual (+ bypass_mid)
  auto ff = [&](const float* in, const LinW& win, const LinW& wout, int hidden, const float* resid, const float* resid2, const float* cs) {
    LinOp a = lin_rows(in, C, C, M, win, hidden);
    a.act = ACT_SWOOSH_L; a.Cf = w.hp; a.ldc = hidden;
    ex.gemm(a, "zip_ff_in");
    LinOp o = lin_rows(w.hp, hidden, hidden, M, wout, C);
    o.resid = resid; o.resid2 = resid2; o.colscale = cs; o.Cf = x; o.ldc = C;
    ex.gemm(o, "zip_ff_out");
  };
  ff(x0, L.ff1_in, L.ff1_out, FF1, x0, nullptr, nullptr);
  mark("ff1", x, C);
  // NonlinAttention
  g = lin_rows(x, C, C, M, L.nl_in, 3 * NH);
  g.Cf = w.t1; g.ldc = 3 * NH;
  ex.gemm(g, "zip_nl_in");
  ex.run(M * NH, NlApply{w.aw, sm, w.t1, w.p64});
  g = lin_rows(w.p64, NH, NH, M, L.nl_out, C);
  g.resid = x; g.Cf = x; g.ldc = C;
  ex.gemm(g, "zip_nl_out");
  mark("nla", x, C);
  auto self_attn = [&](const LinW& win, const LinW& wout) {
    LinOp a = lin_rows(x, C, C, M, win, SV);
    a.Cf = w.t2; a.ldc = SV;
    ex.gemm(a, "zip_sa_in");
    ex.run(M * SV, SaApply{w.aw, sm, w.t2, w.p64});
    LinOp o = lin_rows(w.p64, SV, SV, M, wout, C);
    o.resid = x; o.Cf = x; o.ldc = C;
    ex.gemm(o, "zip_sa_out");
  };
  auto conv_module = [&](const LinW& win, const LinW& wout, const float* dw_w, const float* dw_b) {
    LinOp a = lin_rows(x, C, C, M, win, 2 * C);
    a.Cf = w.t1; a.ldc = 2 * C;
    ex.gemm(a, "zip_cv_in");
    ex.run(nseq * S * C, GluDwConv{w.t1, sm, dw_w, dw_b, w.p64});
    LinOp o = lin_rows(w.p64, C, C, M, wout, C);
    o.resid = x; o.Cf = x; o.ldc = C;
    ex.gemm(o, "zip_cv_out");
  };
  self_attn(L.sa1_in, L.sa1_out);
  mark("sa1", x, C);
  conv_module(L.cv1_in, L.cv1_out, L.dw1_w, L.dw1_b);
  mark("cv1", x, C);
  ff(x, L.ff2_in, L.ff2_out, FF2, x, x0, L.mid_scale);      // + bypass_mid against the layer input
  mark("mid", x, C);
  self_attn(L.sa2_in, L.sa2_out);
  conv_module(L.cv2_in, L.cv2_out, L.dw2_w, L.dw2_b);
  ff(x, L.ff3_in, L.ff3_out, FF3, x, nullptr, nullptr);
  mark("ff3", x, C);
  ex.run(M * C, NormBypass{x, x0, L.norm_bias, L.norm_scale, L.res_scale});
}

template <class Exec>
void dual_path(Exec& ex, Workspace& w, const EncW& e, float* x0, float* x, int B, int T, int F, const char* tag) {
  const long long M = (long long)B * T * F;
  char nm[32];
  snprintf(nm, sizeof(nm), "%s.f", tag);
  zip_layer(ex, w, e.f, x0, x, M, seq_over_f(T, F), (long long)B * T, nm);
  snprintf(nm, sizeof(nm), "%s.f.out", tag);
  ex.mark(nm, x0, M * C);
  snprintf(nm, sizeof(nm), "%s.t", tag);
  zip_layer(ex, w, e.t, x0, x, M, seq_over_t(T, F), (long long)B * F, nm);
}

// feat (B, 2, T, FB) planar -> mx (B, 1, T, FB) (mask-decoder output before the ReLU), ri (B, 2, T, FB)
template <class Exec>
void forward(Exec& ex, Workspace& w, const Weights& W, const float* feat, float* mx, float* ri, int B, int T) {
  // ---- DenseEncoder (:851-853)
  ex.run((long long)B * T * FB * C, FeatConv{feat, W.c1_w, W.c1_b, w.raw, T});
  inorm(ex, w, B, T, C, FPE, 1, FB, 1, W.c1_na, w.encbuf, SLOTC, (DEPTH - 1) * C, FPE, 1);
  ex.mark_strided("enc0", w.encbuf, (long long)B * T * FPE, SLOTC, (DEPTH - 1) * C, C);
  dense_block(ex, w, W.enc_dense, w.encbuf, w.d3, B, T, FPE, FB, "enc");
  {   // dense_conv_2: (1,3) stride (1,2) pad (0,1): row (t, f') starts at padded pixel t*FPE + 2f' and spans 3 pixels of C channels
    LinOp g{};
    g.a = w.d3; g.a_sB = (long long)T * FPE * C; g.a_sR = 2 * C; g.a_r0 = 0; g.a_ke = 3 * C; g.a_rows = T * FH;
    g.chunks = B; g.rows = T * FH; g.K = 3 * C; g.N = C; g.taps = 0; g.W = W.c2; g.act = ACT_NONE; g.Cf = w.raw; g.ldc = C;
    ex.gemm(g, "zip_stride_conv");
    inorm(ex, w, B, T, C, FH, 0, FQ, 1, W.c2_na, w.x0, C, 0, FQ, 0);
  }
  ex.mark("enc", w.x0, (long long)B * T * FQ * C);
  // ---- four dual-path encoders (:863-867)
  for (int k = 0; k < NENC; ++k) {
    const EncW& e = W.enc[k];
    char tag[16];
    snprintf(tag, sizeof(tag), "ts%d", k);
    if (e.ds == 1) {
      dual_path(ex, w, e, w.x0, w.x, B, T, FQ, tag);
    } else {
      const int Td = ceil_div(T, e.ds), Fd = ceil_div(FQ, e.ds);
      ex.run((long long)B * Td * Fd * C, Down{w.x0, e.down_t, e.down_f, e.ds, T, FQ, Td, Fd, w.x0d});
      char nm[32];
      snprintf(nm, sizeof(nm), "%s.down", tag);
      ex.mark(nm, w.x0d, (long long)B * Td * Fd * C);
      dual_path(ex, w, e, w.x0d, w.xd, B, Td, Fd, tag);
      ex.run((long long)B * T * FQ * C, UpCombine{w.x0d, e.comb_scale, e.comb_rscale, e.ds, T, FQ, Td, Fd, w.x0});
    }
    ex.mark(tag, w.x0, (long long)B * T * FQ * C);
  }
  // ---- decoders (:868-880): the two dense blocks share their input
  ex.run((long long)B * T * FPD * C, PadCopy{w.x0, w.decm, w.decp, T});
  for (int dec = 0; dec < 2; ++dec) {
    const DenseW& dw = dec ? W.phase_dense : W.mask_dense;
    dense_block(ex, w, dw, dec ? w.decp : w.decm, w.d3, B, T, FPD, FQ, dec ? "phase" : "mask");
    // sub-pixel conv (1,3) pad (0,1) to UPF*C channels: row of pixel p starts at pixel p-1 and spans 3 pixels
    LinOp g{};
    g.a = w.d3; g.a_sB = (long long)T * FPD * C; g.a_sR = C; g.a_r0 = -1; g.a_ke = 3 * C; g.a_rows = T * FPD;
    g.chunks = B; g.rows = T * FPD; g.K = 3 * C; g.N = UPF * C; g.taps = 0; g.W = dec ? W.phase_up : W.mask_up; g.act = ACT_NONE;
    g.Cf = w.raw; g.ldc = UPF * C;
    ex.gemm(g, "zip_up_conv");
    inorm(ex, w, B, T, UPF * C, FPD, 1, FQ, UPF, dec ? W.phase_up_na : W.mask_up_na, w.up, C, 0, FU, 0);
    ex.mark(dec ? "phase_up" : "mask_up", w.up, (long long)B * T * FU * C);
    if (dec) ex.run((long long)B * 2 * T * FB, Head{w.up, W.phase_out_w, W.phase_out_b, 2, ri, T});
    else ex.run((long long)B * T * FB, Head{w.up, W.mask_out_w, W.mask_out_b, 1, mx, T});
  }
}

}  // namespace zip
