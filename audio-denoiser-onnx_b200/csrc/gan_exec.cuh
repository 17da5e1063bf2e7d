// CUDA executor of the functor launch sequences (csrc/mfgan_ops.cuh, csrc/dfsmn_ops.cuh): one grid per functor, the
// contractions on the tiled GEMM of csrc/mfgan_gemm.cuh; per-launch tick (CUDA-event timing) and stage dumps.
#pragma once
#include "mfgan_gemm.cuh"

#include "common.cuh"
#include "gemm_tc.cuh"
#include "model_impl.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace gan {

// ---- tcgen05 path for the operators that are plain contractions (Linear; Conv2d with unit stride): the fp32-A mode of
// csrc/gemm_tc.cu -- fp32 activations through TMA, operand split + LayerNorm-on-load + conv edge masking in the converter warps,
// TMA-store epilogue.  Weights arrive as (K, N) matrices (adjacent threads of the functor read adjacent outputs); their
// transposed, zero-padded tf32 hi / lo planes are built on first use and kept in the model's TcCache together with the plans.
static __global__ void transpose_split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int K, int N,
                                              int k_pad, long long total, int kmajor) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i / k_pad), k = (int)(i % k_pad);
  const float v = (n < N && k < K) ? (kmajor ? w[(long long)n * K + k] : w[(long long)k * N + n]) : 0.f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = v - h;
}

struct TcEntry {
  bool valid = false;
  const float *in = nullptr, *w = nullptr; float* out = nullptr; const float* stat = nullptr;
  long long rows = 0; int chunks = 0, K = 0, N = 0, lda = 0, ldc = 0, kind = 0, dil = 0;
  tc::TcPlan plan;
  tc::TcArgs args;
};
struct TcWeight { float* planes = nullptr; int n_pad = 0, k_pad = 0; };
struct TcCache {
  int min_rows = 128;             // Linear layers with fewer rows stay on the FFMA tiles (DFSMN sets 1: a window's result must not depend on the batch it rides in)
  int sms = 148;
  bool enabled = true;
  std::map<int, std::vector<TcEntry>> plans;           // per windows-in-pass
  std::map<const float*, TcWeight> weights;
  std::map<const float*, float*> tables_t;             // (rows, cols) parameter tables kept transposed for coalesced reads
  std::string err;
  ~TcCache() { clear(); }
  void clear() {
    for (auto& kv : weights) cudaFree(kv.second.planes);
    for (auto& kv : tables_t) cudaFree(kv.second);
    weights.clear();
    tables_t.clear();
    plans.clear();
  }
  // t[c * cols + f] -> (cols, rows) copy, built on first use (an eager run: adn_run never captures the first run of a batch)
  const float* transposed(const float* t, int rows, int cols, cudaStream_t st);
  // kmajor: w is (N, K) row-major already; otherwise (K, N)
  const TcWeight* weight(const float* w, int K, int N, cudaStream_t st, bool kmajor = false) {
    auto it = weights.find(w);
    if (it != weights.end()) return &it->second;
    TcWeight t;
    t.n_pad = (N + 63) / 64 * 64; t.k_pad = (K + 31) / 32 * 32;
    const long long total = (long long)t.n_pad * t.k_pad;
    if (cudaMalloc((void**)&t.planes, 2 * total * sizeof(float)) != cudaSuccess) { err = "out of device memory (operand planes)"; return nullptr; }
    transpose_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, t.planes, t.planes + total, K, N, t.k_pad, total, kmajor ? 1 : 0);
    return &(weights[w] = t);
  }
};

template <class F>
__global__ void __launch_bounds__(256) op_kernel(long long n, F f) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) f(i);
}

template <class F> struct OpName { static const char* get() { return "op"; } };
#define GAN_OP_NAME(T, s) template <> struct OpName<T> { static const char* get() { return s; } }
template <int KT> struct OpName<DwConv<KT>> { static const char* get() { return "gan_dw_conv"; } };
GAN_OP_NAME(Linear, "gan_linear");
GAN_OP_NAME(Conv2d, "gan_conv2d");
GAN_OP_NAME(Att, "gan_att");
GAN_OP_NAME(SimLocal, "gan_sim_local");
GAN_OP_NAME(SimCross, "gan_sim_cross");
GAN_OP_NAME(LinKV, "gan_lin_k_v");
GAN_OP_NAME(TaScores, "gan_ta_scores");
GAN_OP_NAME(TaAV, "gan_ta_a_v");
GAN_OP_NAME(GateConvT, "gan_gate_conv_t");
GAN_OP_NAME(RowStats, "gan_row_stats");
GAN_OP_NAME(Gather, "gan_gather");
GAN_OP_NAME(Shift, "gan_shift");
GAN_OP_NAME(OffsetRot, "gan_offset_rot");
GAN_OP_NAME(GateOut, "gan_gate_out");
GAN_OP_NAME(SePool1, "gan_se_pool1");
GAN_OP_NAME(SePool2, "gan_se_pool2");
GAN_OP_NAME(SeMlp, "gan_se_mlp");
GAN_OP_NAME(ScaleRes, "gan_scale_res");
GAN_OP_NAME(GroupPart, "gan_group_part");
GAN_OP_NAME(GroupFin, "gan_group_fin");
GAN_OP_NAME(GroupNorm, "gan_group_norm");
GAN_OP_NAME(Softmax, "gan_softmax");
GAN_OP_NAME(InPart, "gan_in_part");
GAN_OP_NAME(InFin, "gan_in_fin");
GAN_OP_NAME(InApply, "gan_in_apply");
GAN_OP_NAME(FeatConv, "gan_feat_conv");
GAN_OP_NAME(CopyCh, "gan_copy_ch");
GAN_OP_NAME(MaskTail, "gan_mask_tail");
GAN_OP_NAME(CplxTail, "gan_cplx_tail");

static __global__ void transpose_table_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < rows * cols) dst[(i % cols) * rows + i / cols] = src[i];
}
inline const float* TcCache::transposed(const float* t, int rows, int cols, cudaStream_t st) {
  auto it = tables_t.find(t);
  if (it != tables_t.end()) return it->second;
  float* d = nullptr;
  if (cudaMalloc((void**)&d, (size_t)rows * cols * sizeof(float)) != cudaSuccess) { err = "out of device memory (transposed table)"; return nullptr; }
  transpose_table_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, st>>>(t, d, rows, cols);
  tables_t[t] = d;
  return d;
}

// GroupNorm apply (the GroupNorm functor): one thread = four channels of a pixel.  The functor reads its (channel, sub-band)
// affine tables as gam[c * Fw + f] -- 32 cache lines per warp load, twice per element: the kernel was bound by L1 wavefronts.  Here
// the tables are read from (sub-band, channel) copies (TcCache::transposed) as float4.  Same arithmetic per element.
static __global__ void __launch_bounds__(256) group_norm_kernel(GroupNorm f, const float* __restrict__ gamT, const float* __restrict__ betT,
                                                               long long n4) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  const int l4 = f.ld >> 2;
  const int c = (int)(idx % l4) << 2;
  const long long p = idx / l4;
  const int fq = (int)(p % f.Fw);
  const long long bt = p / f.Fw;
  const long long i = p * f.ld + c;
  const float4 x = *reinterpret_cast<const float4*>(f.x + i);
  const float4 g4 = *reinterpret_cast<const float4*>(gamT + (long long)fq * f.ld + c);
  const float4 b4 = *reinterpret_cast<const float4*>(betT + (long long)fq * f.ld + c);
  const float xv[4] = {x.x, x.y, x.z, x.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
  float v[4];
  int gi = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    while (gi + 1 < f.g.n && c + e >= f.g.lo[gi + 1]) ++gi;
    const float* st = f.stat + 2 * (bt * f.g.n + gi);
    v[e] = (xv[e] - st[0]) * st[1] * gv[e] + bv[e];
  }
  if (f.res) {
    const float4 r = *reinterpret_cast<const float4*>(f.res + i);
    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
  }
  *reinterpret_cast<float4*>(f.out + i) = make_float4(v[0], v[1], v[2], v[3]);
}

// OffsetScale (4 heads) + rotary (the OffsetRot functor): one thread = four consecutive dims of one (token, head) -- a rotary pair
// (j, j ^ 1) lies inside the float4.  Same arithmetic per element.
static __global__ void __launch_bounds__(256) offset_rot_kernel(OffsetRot f, long long n4) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  const int j = (int)(idx % (QK / 4)) << 2;
  const int h = (int)((idx / (QK / 4)) % 4);
  const long long r = idx / QK;                      // 4 heads x QK / 4 threads per token
  const int q = (int)(r % f.Q);
  const float4 z4 = *reinterpret_cast<const float4*>(f.huv + r * HUV + HID + j);
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(f.gamma + h * QK + j));
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(f.beta + h * QK + j));
  float v[4] = {z4.x * g4.x + b4.x, z4.y * g4.y + b4.y, z4.z * g4.z + b4.z, z4.w * g4.w + b4.w};
  if (j < ROT) {
    const float4 cs = __ldg(reinterpret_cast<const float4*>(f.cs + q * ROT + j));
    const float4 sn = __ldg(reinterpret_cast<const float4*>(f.sn + q * ROT + j));
    const float o[4] = {v[0] * cs.x + v[1] * sn.x, v[1] * cs.y + v[0] * sn.y, v[2] * cs.z + v[3] * sn.z, v[3] * cs.w + v[2] * sn.w};
    v[0] = o[0]; v[1] = o[1]; v[2] = o[2]; v[3] = o[3];
  }
  *reinterpret_cast<float4*>(f.heads + idx * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

// LayerNorm statistics of a row (the RowStats functor): eight lanes per row, float4 loads (a quarter-warp reads a contiguous 128-byte
// piece of the row), butterfly sums inside the lane group.  One thread per row made every load instruction touch 32 cache lines.
// Two passes like the functor (mean, then centred squares); the summation ORDER differs from the functor's sequential one.
static __global__ void __launch_bounds__(256) row_stats_kernel(RowStats f, long long rows) {
  const long long r = ((long long)blockIdx.x * 256 + threadIdx.x) >> 3;
  const int l = threadIdx.x & 7;
  const bool ok = r < rows;
  const F4* x = reinterpret_cast<const F4*>(f.in + (ok ? r : 0) * f.ldi);
  const int k4 = f.K >> 2;
  float s = 0.f;
  if (ok) for (int k = l; k < k4; k += 8) { const F4 v = x[k]; s += (v.x + v.y) + (v.z + v.w); }
  s += __shfl_xor_sync(0xffffffffu, s, 4); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1);
  const float mu = s / (float)f.K;
  float v2 = 0.f;
  if (ok) for (int k = l; k < k4; k += 8) {
    const F4 q = x[k];
    const float a = q.x - mu, b = q.y - mu, c = q.z - mu, d = q.w - mu;
    v2 += (a * a + b * b) + (c * c + d * d);
  }
  v2 += __shfl_xor_sync(0xffffffffu, v2, 4); v2 += __shfl_xor_sync(0xffffffffu, v2, 2); v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
  if (ok && l == 0) {
    f.stat[2 * r] = mu;
    f.stat[2 * r + 1] = 1.0f / sqrtf(v2 / (float)f.K + EPS);
  }
}

// SELayer MLPs (the SeMlp functor): one CTA per window.  The functor recomputes the 64-wide hidden layer for each of its 64 output
// channels with one thread per output (4 096 threads in all, 8 k dependent loads each: 0.7 ms per launch); here thread (kd, j)
// computes hidden unit j of MLP kd once, then thread c combines both MLPs' outputs.  Same sums in the same order.
static __global__ void __launch_bounds__(128) se_mlp_kernel(SeMlp f) {
  __shared__ float pool[2][C], hid[2][C];
  const long long b = blockIdx.x;
  const int kd = threadIdx.x >> 6, j = threadIdx.x & 63;
  pool[kd][j] = f.pooled[(b * 2 + kd) * C + j];
  __syncthreads();
  float h = f.b0[kd][j];
  for (int k = 0; k < C; ++k) h += f.w0[kd][j * C + k] * pool[kd][k];
  hid[kd][j] = h > 0.f ? h : 0.f;
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float tot = 0.f;
    for (int m = 0; m < 2; ++m) {
      float o = f.b2[m][c];
      for (int jj = 0; jj < C; ++jj) o += f.w2[m][c * C + jj] * hid[m][jj];
      tot += sigmoidf_(o);
    }
    f.scale[b * C + c] = tot;
  }
}

// Gated attention output (the GateOut functor), four columns per thread; the sigmoid on ex2.approx / rcp.approx (~2 ulp).
static __global__ void __launch_bounds__(256) gate_out_kernel(GateOut f, long long n4) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  const int j = (int)(idx % (HID / 8)) << 2;
  const long long r = idx / (HID / 8);
  const float* a = f.att + r * HID; const float* h = f.huv + r * HUV;
  const float4 a0 = *reinterpret_cast<const float4*>(a + j), a1 = *reinterpret_cast<const float4*>(a + HID / 2 + j);
  const float4 h0 = *reinterpret_cast<const float4*>(h + j), h1 = *reinterpret_cast<const float4*>(h + HID / 2 + j);
  auto sg = [](float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); };
  *reinterpret_cast<float4*>(f.out + idx * 4) = make_float4((a1.x * h0.x) * sg(a0.x * h1.x), (a1.y * h0.y) * sg(a0.y * h1.y),
                                                            (a1.z * h0.z) * sg(a0.z * h1.z), (a1.w * h0.w) * sg(a0.w * h1.w));
}

// Depthwise conv along a sequence, one CTA per (sequence, 32-channel group).  The S x 32 strip goes ONCE from global into shared
// memory with 16-byte cp.async copies (zero halos; no register staging, so nothing waits on a load until the one wait before the
// barrier), the taps sit in shared memory too (lanes = channels: every access conflict-free), and each warp computes DWS
// consecutive outputs per pass from a DWS + KT - 1 register window -- the blocking of the DwConv functor without its KT tap
// registers (126 registers, 25 % occupancy, inputs re-fetched through L1: ncu showed FFMA at a quarter of the instructions and IPC
// 1.9).  Same arithmetic in the same order as the functor: bit-equal.
template <int KT>
__global__ void __launch_bounds__(256, 3) dw_conv_seq_kernel(DwConv<KT> f) {
  extern __shared__ __align__(16) float sx[];        // [S + KT - 1 + DWS][32] inputs (zero halos), then [KT][32] taps
  constexpr int padl = (KT - 1) / 2;
  const int S = f.S, groups = f.Cn >> 5, rows = S + KT - 1 + DWS;
  float* stp = sx + rows * 32;
  const long long n = blockIdx.x / groups;
  const int c0 = (int)(blockIdx.x % groups) << 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = c0 + lane;
  for (int idx = threadIdx.x; idx < S * 8; idx += 256) {
    const int r = idx >> 3, q = idx & 7;
    const float* g = f.src + (n * S + r) * f.lds + c0 + 4 * q;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(sx + (r + padl) * 32 + 4 * q);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int idx = threadIdx.x; idx < (KT - 1 + DWS) * 32; idx += 256) {          // halo rows in front of and behind the strip
    const int r = idx >> 5;
    sx[(r < padl ? r : S + r) * 32 + (idx & 31)] = 0.f;
  }
  for (int j = warp; j < KT; j += 8) stp[j * 32 + lane] = f.taps[j * f.Cn + c];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const long long pix0 = f.use_pm ? f.pm.pix(n, 0) : n * S;
  const long long pixs = f.use_pm ? f.pm.sS : 1;
  for (int s0 = warp * DWS; s0 < S; s0 += 8 * DWS) {
    float x[DWS + KT - 1], m[DWS];
#pragma unroll
    for (int j = 0; j < DWS + KT - 1; ++j) x[j] = sx[(s0 + j) * 32 + lane];
#pragma unroll
    for (int o = 0; o < DWS; ++o) m[o] = 0.f;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float tj = stp[j * 32 + lane];
#pragma unroll
      for (int o = 0; o < DWS; ++o) m[o] += tj * x[o + j];
    }
#pragma unroll
    for (int o = 0; o < DWS; ++o) {
      const int s = s0 + o;
      if (s < S) {
        float acc = x[o + padl];
        if (f.res) acc += f.res[(n * S + s) * f.ldr + c];
        f.dst[(pix0 + s * pixs) * f.ldd + c] = acc + m[o];
      }
    }
  }
}

struct CudaExec {
  cudaStream_t st = nullptr;
  int launches = 0;
  ImplTickFn tick = nullptr;
  void* tick_ctx = nullptr;
  bool capture = false;
  std::map<std::string, std::vector<float>>* dumps = nullptr;
  template <class F>
  void run(long long n, const F& f) {
    if (n <= 0) return;
    op_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f);
    ++launches;
    if (tick) tick(tick_ctx, OpName<F>::get());
  }
  void run(long long n, const GroupNorm& f) {
    static const bool functor = getenv("ADN_GAN_EW") && !strcmp(getenv("ADN_GAN_EW"), "functor");
    const float *gt = nullptr, *bt = nullptr;
    if (!functor && tc && !(f.ld & 3) && n > 0) {
      gt = tc->transposed(f.gam, f.ld, f.Fw, st);
      bt = tc->transposed(f.bet, f.ld, f.Fw, st);
    }
    if (!gt || !bt) { run<GroupNorm>(n, f); return; }
    group_norm_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, gt, bt, n / 4);
    ++launches;
    if (tick) tick(tick_ctx, "gan_group_norm");
  }
  void run(long long n, const RowStats& f) {
    static const bool functor = getenv("ADN_GAN_EW") && !strcmp(getenv("ADN_GAN_EW"), "functor");
    if (functor || n <= 0 || (f.K & 3) || (f.ldi & 3)) { run<RowStats>(n, f); return; }
    row_stats_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, st>>>(f, n);
    ++launches;
    if (tick) tick(tick_ctx, "gan_row_stats");
  }
  void run(long long n, const SeMlp& f) {
    static const bool functor = getenv("ADN_GAN_EW") && !strcmp(getenv("ADN_GAN_EW"), "functor");
    if (functor || n <= 0) { run<SeMlp>(n, f); return; }
    se_mlp_kernel<<<(unsigned)(n / C), 128, 0, st>>>(f);
    ++launches;
    if (tick) tick(tick_ctx, "gan_se_mlp");
  }
  void run(long long n, const GateOut& f) {
    static const bool functor = getenv("ADN_GAN_EW") && !strcmp(getenv("ADN_GAN_EW"), "functor");
    if (functor || n <= 0) { run<GateOut>(n, f); return; }
    gate_out_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, n / 4);
    ++launches;
    if (tick) tick(tick_ctx, "gan_gate_out");
  }
  void run(long long n, const OffsetRot& f) {
    static const bool functor = getenv("ADN_GAN_EW") && !strcmp(getenv("ADN_GAN_EW"), "functor");
    if (functor || n <= 0) { run<OffsetRot>(n, f); return; }
    offset_rot_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, n / 4);
    ++launches;
    if (tick) tick(tick_ctx, "gan_offset_rot");
  }
  template <int KT>
  void run(long long n, const DwConv<KT>& f) {
    if (n <= 0) return;
    static const bool functor = getenv("ADN_GAN_DW") && !strcmp(getenv("ADN_GAN_DW"), "functor");
    const long long nseq = n / ((long long)((f.S + DWS - 1) / DWS) * f.Cn);
    const size_t smem = (size_t)(f.S + KT - 1 + DWS + KT) * 32 * sizeof(float);
    if (functor || (f.Cn & 31) || (f.lds & 3) || ((uintptr_t)f.src & 15) || smem > 48 * 1024 || nseq * (f.Cn >> 5) > 0x7fffffffLL) {
      op_kernel<DwConv<KT>><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f);
    } else {
      dw_conv_seq_kernel<KT><<<(unsigned)(nseq * (f.Cn >> 5)), 256, smem, st>>>(f);
    }
    ++launches;
    if (tick) tick(tick_ctx, "gan_dw_conv");
  }
  // contractions run on the shared-memory tiled GEMM (mfgan_gemm.cuh) instead of the one-output-per-thread functor
  template <class F>
  void run_gemm(long long n, const F& f) {
    if (n <= 0) return;
    GemmOp ops[3];
    const int k = translate(f, n, ops);
    for (int i = 0; i < k; ++i) launch_gemm(ops[i], st);
    launches += k;
    if (tick) tick(tick_ctx, OpName<F>::get());
  }
  void gemm(const GemmOp& g, const char* name = "gemm") {
    if (try_tc(g, name)) return;
    launch_gemm(g, st);
    ++launches;
    if (tick) tick(tick_ctx, name);
  }
  // ---- tcgen05 dispatch
  TcCache* tc = nullptr;
  int tc_pass = 0, tc_idx = 0;       // plans are cached per (windows in this pass, position in the launch sequence)
  static int tc_act(int a) {
    return a == ACT_NONE ? tc::ACT_NONE : a == ACT_RELU ? tc::ACT_RELU : a == ACT_SILU ? tc::ACT_SILU : a == ACT_PRELU ? tc::ACT_PRELU_VEC
           : a == ACT_SIGMOID ? tc::ACT_SIGMOID : -1;
  }
  bool tc_launch(TcEntry& e, const char* name) {
    if (tc::launch(e.plan, e.args, EPI_LIN, tc->sms, st) != cudaSuccess) { cudaGetLastError(); return false; }
    ++launches;
    if (tick) tick(tick_ctx, name);
    return true;
  }
  TcEntry& tc_slot() {
    std::vector<TcEntry>& v = tc->plans[tc_pass];
    if ((int)v.size() <= tc_idx) v.resize(tc_idx + 1);
    return v[tc_idx++];
  }
  bool tc_plan(TcEntry& e, const float* a, int a_ke, long long rows, long long a_sr, int chunks, long long a_sb, const TcWeight* w, int N, int K,
               float* out, int ldc) {
    const int bn = N <= 64 ? 64 : N <= 128 ? 128 : 256;
    const int bt = rows >= 128 ? 128 : (int)rows;
    e.plan = tc::TcPlan{};
    e.plan.bn = bn;
    e.plan.a_f32 = true;
    const long long plane = (long long)w->n_pad * w->k_pad;
    std::string& err = tc->err;
    if (!tc::make_row_map(&e.plan.map_a_hi, a, a_ke, (int)rows, a_sr, chunks, a_sb, bt, 1, err) ||
        !tc::make_weight_map(&e.plan.map_w_hi, w->planes, w->k_pad, w->n_pad, bn, err) ||
        !tc::make_weight_map(&e.plan.map_w_lo, w->planes + plane, w->k_pad, w->n_pad, bn, err) ||
        !tc::make_store_map(&e.plan.map_c, out, N, (int)rows, ldc, chunks, rows * ldc, err))
      return false;
    e.plan.map_a_lo = e.plan.map_a_hi;
    e.plan.map_w2_hi = e.plan.map_w_hi;
    e.plan.map_w2_lo = e.plan.map_w_lo;
    tc::TcArgs& g = e.args;
    g = tc::TcArgs{};
    g.bb = 1; g.bt = bt; g.tiles_per_chunk = (int)((rows + 127) / 128); g.t0 = 0;
    g.B = chunks; g.TM = (int)rows; g.N = N; g.K = K;
    g.m_tiles = chunks * g.tiles_per_chunk;
    g.C = out; g.ldc = ldc;
    return true;
  }
  bool try_tc(long long n, const Linear& f) {
    if (!tc || !tc->enabled || tc_act(f.a) < 0 || f.N % 4 || f.ldo % 4 || f.ldi % 4 || f.K % 4 || f.N < 16 || ((uintptr_t)f.in & 15) || ((uintptr_t)f.out & 15))
      return false;
    const long long M = n / f.N;
    if (M < tc->min_rows) return false;
    const TcWeight* w = tc->weight(f.Wt, f.K, f.N, st);
    if (!w) return false;
    TcEntry& e = tc_slot();
    if (!(e.valid && e.kind == 1 && e.in == f.in && e.out == f.out && e.w == f.Wt && e.stat == f.stat && e.rows == M && e.K == f.K && e.N == f.N &&
          e.lda == f.ldi && e.ldc == f.ldo)) {
      e.valid = false;
      if (!tc_plan(e, f.in, f.K, M, f.ldi, 1, M * f.ldi, w, f.N, f.K, f.out, f.ldo)) return false;
      e.args.bias = f.bias; e.args.act = tc_act(f.a); e.args.act_vec = f.slope; e.args.a_rowstat = f.stat;
      e.kind = 1; e.in = f.in; e.out = f.out; e.w = f.Wt; e.stat = f.stat; e.rows = M; e.chunks = 1; e.K = f.K; e.N = f.N; e.lda = f.ldi; e.ldc = f.ldo;
      e.valid = true;
    }
    return tc_launch(e, "gan_linear_tc");
  }
  // unit-stride channel-last conv as an implicit GEMM over the UN-padded map: K = taps x Cin, tap (kt, kf) reads the rows shifted by
  // (kt - (KT-1)) * dil * F + (kf - pf); rows before the window start read as zeros (TMA), rows whose sub-band leaves [0, F) are zeroed
  // by the converter warps
  bool try_tc(long long n, const Conv2d& f) {
    if (!tc || !tc->enabled || f.sf != 1 || f.Fin != f.Fout || f.Cin % 32 || f.Cout % 4 || f.Cout < 16 || f.KT * f.KF > 6 || f.ldi % 4 || f.ldo % 4 ||
        ((uintptr_t)f.in & 15) || ((uintptr_t)f.out & 15))
      return false;
    const long long px = n / f.Cout;                       // B * T * F
    const long long rows = (long long)f.T * f.Fin;
    const int B = (int)(px / rows);
    if (rows < 128) return false;
    const int K = f.KT * f.KF * f.Cin;
    const TcWeight* w = tc->weight(f.W, K, f.Cout, st);
    if (!w) return false;
    TcEntry& e = tc_slot();
    if (!(e.valid && e.kind == 2 && e.in == f.in && e.out == f.out && e.w == f.W && e.rows == rows && e.chunks == B && e.K == K && e.N == f.Cout &&
          e.lda == f.ldi && e.ldc == f.ldo && e.dil == f.dil)) {
      e.valid = false;
      if (!tc_plan(e, f.in, f.Cin, rows, f.ldi, B, rows * f.ldi, w, f.Cout, K, f.out, f.ldo)) return false;
      tc::TcArgs& g = e.args;
      g.bias = f.bias;
      g.taps = f.KT * f.KF; g.tap_kb = f.Cin / 32; g.tap_k0 = 0; g.tap_w = f.Fin;
      for (int kt = 0; kt < f.KT; ++kt)
        for (int kf = 0; kf < f.KF; ++kf) {
          g.tap_shift[kt * f.KF + kf] = (kt - (f.KT - 1)) * f.dil * f.Fin + (kf - f.pf);
          g.tap_df[kt * f.KF + kf] = kf - f.pf;
        }
      e.kind = 2; e.in = f.in; e.out = f.out; e.w = f.W; e.stat = nullptr; e.rows = rows; e.chunks = B; e.K = K; e.N = f.Cout; e.lda = f.ldi;
      e.ldc = f.ldo; e.dil = f.dil;
      e.valid = true;
    }
    return tc_launch(e, "gan_conv2d_tc");
  }
  // a plain GemmOp whose operands are K-contiguous (the DFSMN analysis transform: frames as overlapping rows of the waveform)
  bool try_tc(const GemmOp& g, const char* name) {
    if (!tc || !tc->enabled || g.a_k != 1 || g.b_k != 1 || g.c_n != 1 || g.stat || g.cv.on || g.gt.on || g.epi != GEPI_ACT || g.act != ACT_NONE ||
        g.bias || g.nb2 != 1 || g.a_kin || g.b_kin || g.b_nin || g.c_nin || g.b_b1 || g.a_m % 4 || g.a_b1 % 4 || g.c_m % 4 || g.K % 4 || g.N % 4 ||
        g.b_n != g.K || g.c_b1 != (long long)g.M * g.c_m || ((uintptr_t)g.A & 15) || ((uintptr_t)g.C & 15))
      return false;
    const TcWeight* w = tc->weight(g.B, g.K, g.N, st, true);
    if (!w) return false;
    TcEntry& e = tc_slot();
    if (!(e.valid && e.kind == 3 && e.in == g.A && e.out == g.C && e.w == g.B && e.rows == g.M && e.chunks == g.batch && e.K == g.K && e.N == g.N &&
          e.lda == (int)g.a_m && e.ldc == (int)g.c_m)) {
      e.valid = false;
      if (!tc_plan(e, g.A, g.K, g.M, g.a_m, g.batch, g.a_b1, w, g.N, g.K, g.C, (int)g.c_m)) return false;
      e.kind = 3; e.in = g.A; e.out = g.C; e.w = g.B; e.stat = nullptr; e.rows = g.M; e.chunks = g.batch; e.K = g.K; e.N = g.N;
      e.lda = (int)g.a_m; e.ldc = (int)g.c_m;
      e.valid = true;
    }
    return tc_launch(e, name);
  }
  void run(long long n, const Linear& f) {
    if (try_tc(n, f)) return;
    run_gemm(n, f);
  }
  void run(long long n, const SimLocal& f) { run_gemm(n, f); }
  void run(long long n, const SimCross& f) { run_gemm(n, f); }
  void run(long long n, const LinKV& f) { run_gemm(n, f); }
  void run(long long n, const Att& f) { run_gemm(n, f); }
  void run(long long n, const GateConvT& f) { run_gemm(n, f); }
  void run(long long n, const TaScores& f) { run_gemm(n, f); }
  void run(long long n, const TaAV& f) { run_gemm(n, f); }
  void run(long long n, const Conv2d& f) {
    if (try_tc(n, f)) return;
    if (f.Cout >= 16 && f.Cin % GK == 0) run_gemm(n, f);
    else run<Conv2d>(n, f);
  }
  void mark(const char* tag, const char* name, const float* p, long long count) {
    if (!capture || !dumps) return;
    std::string key = tag[0] ? std::string(tag) + "." + name : std::string(name);
    std::vector<float>& v = (*dumps)[key];
    v.resize((size_t)count);
    cudaStreamSynchronize(st);
    cudaMemcpy(v.data(), p, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost);
  }
};

}  // namespace gan
