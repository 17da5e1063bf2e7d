// MossFormer2-SE-48K: Kaldi-fbank + STFT frontend, 24 x (FLASH attention block + gated FSMN block),
// mask tail and ISTFT (reference MossFormer2_SE_48K/Export_MossFormer_SE.py:312-507).
//
// Layout: activations are token-major fp32, row m = window*T + frame.  Every dense contraction
// (frontend DFT, 7 linear layers per block, the four attention products, mask tail, ISTFT
// overlap-add) runs on the tcgen05 3xTF32 GEMM (gemm_tc.cu); producers write the tf32 hi/lo
// operand planes directly.  The attention products are batched per window with the window's own
// keys / values as the "weight" operand:
//     S1  = Ql Kl^T                    A = lin_q  (T x 128)      W = lin_k    (T x 128)
//     S   = relu(Qq Kq^T)^2 + S1       A = quad_q (T x 128)      W = quad_k   (T x 128)
//     O   = S [v|u]                    A = S      (T x Tp)       W = [v|u]^T  (2048 x Tp)
// A window is one FLASH group (T <= 256 frames), so the reference's linear branch
// Ql (Kl^T [v|u]) (:422-427) is re-associated to (Ql Kl^T) [v|u] and shares the value product with
// the quadratic branch: T <= 256 makes the T x T form the cheaper one, and it removes two GEMMs and
// the 2048 x 128 per-window KV round trip.  The depthwise-conv kernel that follows the input
// projection emits [v|u]^T through a shared-memory transpose.  Zero-padded keys contribute exactly
// nothing (:417-420) and are never stored.
//
// Kernel <-> reference map:
//   feat_kernel       power spectrum, mel filterbank, log (:337-341)
//   featnorm_kernel   deltas (:304-310, :345-347), GroupNorm `norm` (:348), position table (:350)
//   shiftnorm_kernel  token shift + ScaleNorm denominator (:393-397)
//   dwconv_in_kernel  ConvModule residual (:399-400), OffsetScale + rotary (:404-409)
//   gate_kernel       (att_u*v)*sigmoid(att_v*u) (:431-432) + ScaleNorm denominator (:435)
//   dwconv_kernel     ConvModule residual of to_out (:437-438) / to_u||to_v (:451-452)
//   ln2_kernel        norm1 + affine-free LayerNorm (:446-449)
//   fsmn_mem_kernel   UniDeepFsmn memory conv (:459-462), gate (:464), norm2 (:467)
//   tail_norm_kernel  final LayerNorm, GroupNorm, skip, PReLU (:474-482)
//   tail_gate_kernel  tanh * sigmoid (:484-485)
//   mask_apply_kernel mask x STFT rows into the zero-framed ISTFT operand (:487)
#include "adn.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "gtcrn.cuh"
#include "model_impl.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace mf2 {

using gtcrn::split_tf32_store;

constexpr int D = 512, VU = 1024, VU2 = 2048, QK = 128, PROJ = 2176, FI = 256;
constexpr int DW = 17, DWH = 8, MEMK = 39, MEMH = 19;
constexpr int NM = 60, FEAT = 180, FEATP = 192;
constexpr int NFFT = 1920, HOP = 384, KB = 1025, KROWS = 2050, BINS = 961, SROWS = 1922, FRONT = 3972;
constexpr int BINSP = 964, ROT = 32;
constexpr int SPEC_LD = 1928, R_OLA = 5, PADF = 4;
constexpr float EPS_IN = 1e-5f * 22.62741699796952f;     // eps / dim^-0.5   (:151)
constexpr float EPS_OUT = 1e-5f * 32.0f;                 // eps / 1024^-0.5  (:152)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void split4(float4 v, float* hi, float* lo, long long i) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
  st4(hi + i, h);
  st4(lo + i, l);
}

// ---------------------------------------------------------------------------------
// One CTA per frame: Kaldi power spectrum -> 60 mel bands -> log (+ int16-domain offset).
__global__ void __launch_bounds__(256)
feat_kernel(const float* __restrict__ fr, const float* __restrict__ banks, const int* __restrict__ mel_lo,
            const int* __restrict__ mel_hi, float* __restrict__ mel, float floor_v, float log_off) {
  __shared__ float pw[KB + 3];
  const long long m = blockIdx.x;
  const float* row = fr + m * FRONT;
  for (int i = threadIdx.x; i < KB; i += 256) {
    const float re = __ldg(row + i), im = __ldg(row + KB + i);
    pw[i] = re * re + im * im;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < NM; j += 8) {
    const int lo = mel_lo[j], hi = mel_hi[j];
    float acc = 0.f;
    for (int i = lo + lane; i < hi; i += 32) acc += __ldg(banks + j * KB + i) * pw[i];
    acc = warp_sum(acc);
    if (lane == 0) mel[m * NM + j] = logf(fmaxf(acc, floor_v)) + log_off;
  }
}

// One CTA per window: deltas, delta-deltas, GroupNorm(1, 180) over (180, T), operand planes for the
// 1x1 encoder GEMM; also seeds z with the position table (the GEMM adds onto it).
__global__ void __launch_bounds__(256)
featnorm_kernel(const float* __restrict__ mel, const float* __restrict__ gw, const float* __restrict__ gb,
                const float* __restrict__ emb, float* __restrict__ fhi, float* __restrict__ flo,
                float* __restrict__ z, int T) {
  extern __shared__ float sm[];
  float* s0 = sm;                 // [T][60] log-mel
  float* s1 = sm + T * NM;        // [T][60] delta
  __shared__ double red[2][8];
  __shared__ float stat[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = mel + (long long)b * T * NM;
  for (int i = tid; i < T * NM; i += 256) s0[i] = __ldg(src + i);
  __syncthreads();
  auto delta = [&](const float* s, int t, int j) {
    float acc = 0.f;
#pragma unroll
    for (int k = -2; k <= 2; ++k) {
      int tt = t + k;
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      acc += ((float)k * 0.1f) * s[tt * NM + j];
    }
    return acc;
  };
  for (int i = tid; i < T * NM; i += 256) s1[i] = delta(s0, i / NM, i % NM);
  __syncthreads();
  double su = 0.0, sq = 0.0;
  for (int i = tid; i < T * NM; i += 256) {
    const float a = s0[i], d1 = s1[i], d2 = delta(s1, i / NM, i % NM);
    su += (double)a + (double)d1 + (double)d2;
    sq += (double)a * a + (double)d1 * d1 + (double)d2 * d2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { su += __shfl_xor_sync(0xffffffffu, su, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
  if ((tid & 31) == 0) { red[0][tid >> 5] = su; red[1][tid >> 5] = sq; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; q += red[1][w]; }
    const double n = (double)T * FEAT, mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + 1e-8));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (int i = tid; i < T * FEAT; i += 256) {
    const int t = i / FEAT, c = i - t * FEAT;
    const int part = c / NM, j = c - part * NM;
    const float v = part == 0 ? s0[t * NM + j] : (part == 1 ? s1[t * NM + j] : delta(s1, t, j));
    const float y = (v - mean) * rstd * __ldg(gw + c) + __ldg(gb + c);
    split_tf32_store(y, fhi, flo, ((long long)b * T + t) * FEATP + c);
  }
  float* zb = z + (long long)b * T * D;
  for (int i = tid; i < T * D / 4; i += 256) st4(zb + 4 * i, __ldg(reinterpret_cast<const float4*>(emb) + i));
}

// One warp per token: first half of the channels comes from the previous frame (zero at t = 0);
// rs = 1 / (||x|| + eps) is applied as a row scale by the consuming GEMM.
__global__ void __launch_bounds__(256)
shiftnorm_kernel(const float* __restrict__ h, float* __restrict__ xhi, float* __restrict__ xlo,
                 float* __restrict__ rs, long long M, int T) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const int t = (int)(m % T);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c >= D / 2) v = ld4(h + m * D + c);
    else if (t > 0) v = ld4(h + (m - 1) * D + c);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    split4(v, xhi, xlo, m * D + c);
  }
  ss = warp_sum(ss);
  if (lane == 0) rs[m] = 1.0f / (sqrtf(ss) + EPS_IN);
}

// Depthwise k=17 'same' conv over time + residual, streamed: one warp owns a 32-channel strip of one
// window (lane = channel) and walks the frames once.  The 17-frame window plus the prefetched frames
// live in a RING-register ring (static indices after unrolling), so every frame costs one coalesced
// 128-byte global load issued RING-16 frames ahead of its first use, 17 FMAs and no shared memory or
// CTA barrier.  emit(t, value) consumes frame t; flush(t0) runs after every 32 frames.
template <int RING, typename F, typename G>
__device__ __forceinline__ void dwconv_stream(const float* __restrict__ src, long long ld, const float (&w)[DW], int T,
                                              F&& emit, G&& flush) {
  static_assert(RING % 32 == 0 && RING > DW, "ring = whole 32-frame groups");
  float ring[RING];
#pragma unroll
  for (int i = 0; i < RING; ++i) {
    const int r = i - DWH;
    ring[i] = (r >= 0 && r < T) ? __ldg(src + (long long)r * ld) : 0.f;
  }
  for (int t0 = 0; t0 < T; t0 += RING) {
#pragma unroll
    for (int j = 0; j < RING; ++j) {
      const int t = t0 + j;
      if (t < T) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < DW; ++k) acc += w[k] * ring[(j + k) % RING];
        acc += ring[(j + DWH) % RING];
        emit(t, acc);
      }
      const int r = t + RING - DWH;
      ring[j] = r < T ? __ldg(src + (long long)r * ld) : 0.f;
      if ((j & 31) == 31 && t0 + j - 31 < T) flush(t0 + j - 31);
    }
  }
}

// ConvModule residual on the fused to_hidden||to_qk projection; CTA = 4 warps = 4 adjacent 32-channel
// strips of one window.  Strips of the 2048 value channels write [v|u] (token-major fp32, for the
// gate) and [v|u]^T (tf32 planes, the attention operand; 32x32 per-warp transposes); the last CTA
// column holds the 128 qk channels: four OffsetScale heads + rotary embedding -> quad_q / lin_q /
// quad_k / lin_k (token-major planes).
constexpr int DWI_WARPS = 4;
__global__ void __launch_bounds__(DWI_WARPS * 32)
dwconv_in_kernel(const float* __restrict__ proj, const float* __restrict__ taps, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ rcos, const float* __restrict__ rsin,
                 float* __restrict__ vu, float* __restrict__ vuT_hi, float* __restrict__ vuT_lo,
                 float* __restrict__ qq_hi, float* __restrict__ qq_lo, float* __restrict__ lq_hi,
                 float* __restrict__ lq_lo, float* __restrict__ qk_hi, float* __restrict__ qk_lo,
                 float* __restrict__ lk_hi, float* __restrict__ lk_lo, int T, int Tp, int Tn) {
  __shared__ float stage[DWI_WARPS][32 * 33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * DWI_WARPS + warp) * 32, b = blockIdx.y;
  const float* src = proj + (long long)b * T * PROJ + c0 + lane;
  float w[DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) w[k] = __ldg(taps + k * PROJ + c0 + lane);
  if (c0 < VU2) {
    float* st = stage[warp];
    dwconv_stream<64>(src, PROJ, w, T,
                  [&](int t, float acc) {
                    vu[((long long)b * T + t) * VU2 + c0 + lane] = acc;
                    st[(t & 31) * 33 + lane] = acc;
                  },
                  [&](int t0) {
                    __syncwarp();
                    const int t = t0 + lane;
                    if (t < T) {
#pragma unroll 8
                      for (int c = 0; c < 32; ++c)
                        split_tf32_store(st[lane * 33 + c], vuT_hi, vuT_lo, ((long long)b * VU2 + c0 + c) * Tp + t);
                    }
                    __syncwarp();
                  });
  } else {
    const int q = c0 - VU2 + lane;             // qk channel
    float g4[4], b4[4];
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) { g4[hd] = __ldg(gamma + hd * QK + q); b4[hd] = __ldg(beta + hd * QK + q); }
    dwconv_stream<64>(src, PROJ, w, T,
                  [&](int t, float acc) {
                    float s[4];
#pragma unroll
                    for (int hd = 0; hd < 4; ++hd) s[hd] = acc * g4[hd] + b4[hd];
                    if (c0 == VU2) {           // rotary on the first 32 qk channels, interleaved pairs
                      const float cs = __ldg(rcos + t * ROT + lane), sn = __ldg(rsin + t * ROT + lane);
#pragma unroll
                      for (int hd = 0; hd < 4; ++hd) {
                        const float other = __shfl_xor_sync(0xffffffffu, s[hd], 1);
                        const float rot = (lane & 1) ? other : -other;
                        s[hd] = s[hd] * cs + rot * sn;
                      }
                    }
                    const long long m = (long long)b * T + t;
                    split_tf32_store(s[0], qq_hi, qq_lo, m * QK + q);
                    split_tf32_store(s[1], lq_hi, lq_lo, m * QK + q);
                    split_tf32_store(s[2], qk_hi, qk_lo, ((long long)b * Tn + t) * QK + q);
                    split_tf32_store(s[3], lk_hi, lk_lo, ((long long)b * Tn + t) * QK + q);
                  },
                  [](int) {});
  }
}

// One warp per token: gate and ScaleNorm denominator of to_out.
__global__ void __launch_bounds__(256)
gate_kernel(const float* __restrict__ att, const float* __restrict__ vu, float* __restrict__ ghi,
            float* __restrict__ glo, float* __restrict__ rs, long long M) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const float* a = att + m * VU2;
  const float* x = vu + m * VU2;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 av = ld4(a + c), au = ld4(a + VU + c), v = ld4(x + c), u = ld4(x + VU + c);
    float4 o;
    o.x = (au.x * v.x) * adn_sigmoid(av.x * u.x);
    o.y = (au.y * v.y) * adn_sigmoid(av.y * u.y);
    o.z = (au.z * v.z) * adn_sigmoid(av.z * u.z);
    o.w = (au.w * v.w) * adn_sigmoid(av.w * u.w);
    ss += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    split4(o, ghi, glo, m * VU + c);
  }
  ss = warp_sum(ss);
  if (lane == 0) rs[m] = 1.0f / (sqrtf(ss) + EPS_OUT);
}

// out = x + dwconv17(x) (+ resid); optional tf32 planes of the first `plane_cols` channels.
// CTA = 4 warps = 4 adjacent 32-channel strips of one window (see dwconv_stream).
__global__ void __launch_bounds__(DWI_WARPS * 32)
dwconv_kernel(const float* __restrict__ x, const float* __restrict__ taps, const float* __restrict__ resid,
              float* __restrict__ out, float* __restrict__ phi, float* __restrict__ plo, int plane_cols, int T, int C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * DWI_WARPS + warp) * 32 + lane, b = blockIdx.y;
  float w[DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) w[k] = __ldg(taps + k * C + c);
  const bool planes = phi && c < plane_cols;
  dwconv_stream<32>(x + (long long)b * T * C + c, C, w, T,
                [&](int t, float acc) {
                  const long long m = (long long)b * T + t;
                  if (resid) acc += __ldg(resid + m * C + c);
                  out[m * C + c] = acc;
                  if (planes) split_tf32_store(acc, phi, plo, m * plane_cols + c);
                },
                [](int) {});
}

__device__ __forceinline__ void ln256(const float (&v)[8], float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  mean = warp_sum(s) * (1.0f / FI);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
  rstd = rsqrtf(warp_sum(q) * (1.0f / FI) + 1e-5f);
}

// One warp per token: g_in = LayerNorm(c1y) (affine) -> fp32; xn = LayerNorm(g_in) (no affine) -> planes.
__global__ void __launch_bounds__(256)
ln2_kernel(const float* __restrict__ c1y, const float* __restrict__ w, const float* __restrict__ bvec,
           float* __restrict__ gin, float* __restrict__ xhi, float* __restrict__ xlo, long long M) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  float v[8];
  const float4 a0 = ld4(c1y + m * FI + lane * 4), a1 = ld4(c1y + m * FI + 128 + lane * 4);
  v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
  float mean, rstd;
  ln256(v, mean, rstd);
  const float4 w0 = ld4(w + lane * 4), w1 = ld4(w + 128 + lane * 4), b0 = ld4(bvec + lane * 4), b1 = ld4(bvec + 128 + lane * 4);
  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
  st4(gin + m * FI + lane * 4, make_float4(v[0], v[1], v[2], v[3]));
  st4(gin + m * FI + 128 + lane * 4, make_float4(v[4], v[5], v[6], v[7]));
  ln256(v, mean, rstd);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd;
  split4(make_float4(v[0], v[1], v[2], v[3]), xhi, xlo, m * FI + lane * 4);
  split4(make_float4(v[4], v[5], v[6], v[7]), xhi, xlo, m * FI + 128 + lane * 4);
}

// CTA = (8 frames, window), thread = channel: memory conv (k = 39, zero padded) on the projected
// branch, gate with xv, residual g_in, norm2 -> operand planes of conv2.
constexpr int FM_TOK = 8;
__global__ void __launch_bounds__(256)
fsmn_mem_kernel(const float* __restrict__ xp, const float* __restrict__ uv, const float* __restrict__ gin,
                const float* __restrict__ taps, const float* __restrict__ w, const float* __restrict__ bvec,
                float* __restrict__ yhi, float* __restrict__ ylo, int T) {
  extern __shared__ float fsm[];
  float (*tile)[FI] = reinterpret_cast<float (*)[FI]>(fsm);                              // [FM_TOK + 2*MEMH][FI]
  float (*ys)[FI] = reinterpret_cast<float (*)[FI]>(fsm + (FM_TOK + 2 * MEMH) * FI);       // [FM_TOK][FI]
  const int t0 = blockIdx.x * FM_TOK, b = blockIdx.y, c = threadIdx.x;
  const long long base = (long long)b * T;
  for (int r = 0; r < FM_TOK + 2 * MEMH; ++r) {
    const int t = t0 + r - MEMH;
    tile[r][c] = (t >= 0 && t < T) ? __ldg(xp + (base + t) * FI + c) : 0.f;
  }
  float k[MEMK];
#pragma unroll
  for (int i = 0; i < MEMK; ++i) k[i] = __ldg(taps + i * FI + c);
  // own column only: no barrier needed between the tile fill and the taps
#pragma unroll
  for (int tt = 0; tt < FM_TOK; ++tt) {
    const int t = t0 + tt;
    float y = 0.f;
    if (t < T) {
      float conv = 0.f;
#pragma unroll
      for (int i = 0; i < MEMK; ++i) conv += k[i] * tile[tt + i][c];
      const long long m = base + t;
      const float xu = __ldg(uv + m * (2 * FI) + c) + (tile[tt + MEMH][c] + conv);
      y = __ldg(uv + m * (2 * FI) + FI + c) * xu + __ldg(gin + m * FI + c);
    }
    ys[tt][c] = y;
  }
  __syncthreads();
  const int warp = c >> 5, lane = c & 31;
  const int t = t0 + warp;
  if (t >= T) return;
  float v[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = ys[warp][lane * 4 + i]; v[4 + i] = ys[warp][128 + lane * 4 + i]; }
  float mean, rstd;
  ln256(v, mean, rstd);
  const float4 w0 = ld4(w + lane * 4), w1 = ld4(w + 128 + lane * 4), b0 = ld4(bvec + lane * 4), b1 = ld4(bvec + 128 + lane * 4);
  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
  const long long m = base + t;
  split4(make_float4(v[0], v[1], v[2], v[3]), yhi, ylo, m * FI + lane * 4);
  split4(make_float4(v[4], v[5], v[6], v[7]), yhi, ylo, m * FI + 128 + lane * 4);
}

// One CTA per window: LayerNorm(512) per frame, GroupNorm(1, 512) over the window, + encoder output,
// PReLU -> operand planes of the tail gate GEMM.  `hn` is fp32 scratch.
__global__ void __launch_bounds__(512)
tail_norm_kernel(const float* __restrict__ h, const float* __restrict__ z, const float* __restrict__ lw,
                 const float* __restrict__ lb, const float* __restrict__ gw, const float* __restrict__ gb,
                 const float* __restrict__ slope, float* __restrict__ hn, float* __restrict__ thi,
                 float* __restrict__ tlo, int T) {
  __shared__ double red[2][16];
  __shared__ float stat[2];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long base = (long long)b * T;
  double su = 0.0, sq = 0.0;
  for (int t = warp; t < T; t += 16) {
    const float* row = h + (base + t) * D;
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a = ld4(row + (i * 32 + lane) * 4);
      v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 w4 = ld4(lw + c), b4 = ld4(lb + c);
      float4 o;
      o.x = (v[4 * i] - mean) * rstd * w4.x + b4.x;
      o.y = (v[4 * i + 1] - mean) * rstd * w4.y + b4.y;
      o.z = (v[4 * i + 2] - mean) * rstd * w4.z + b4.z;
      o.w = (v[4 * i + 3] - mean) * rstd * w4.w + b4.w;
      st4(hn + (base + t) * D + c, o);
      su += (double)o.x + (double)o.y + (double)o.z + (double)o.w;
      sq += (double)o.x * o.x + (double)o.y * o.y + (double)o.z * o.z + (double)o.w * o.w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { su += __shfl_xor_sync(0xffffffffu, su, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
  if (lane == 0) { red[0][warp] = su; red[1][warp] = sq; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 16; ++w) { a += red[0][w]; q += red[1][w]; }
    const double n = (double)T * D, mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + 1e-8));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1], a = __ldg(slope);
  for (int i = tid; i < T * D / 4; i += 512) {
    const int c = (i * 4) % D;
    const long long o = base * D + (long long)i * 4;
    const float4 x = ld4(hn + o), zz = ld4(z + o), w4 = ld4(gw + c), b4 = ld4(gb + c);
    float4 y;
    y.x = adn_prelu((x.x - mean) * rstd * w4.x + b4.x + zz.x, a);
    y.y = adn_prelu((x.y - mean) * rstd * w4.y + b4.y + zz.y, a);
    y.z = adn_prelu((x.z - mean) * rstd * w4.z + b4.z + zz.z, a);
    y.w = adn_prelu((x.w - mean) * rstd * w4.w + b4.w + zz.w, a);
    split4(y, thi, tlo, o);
  }
}

__global__ void __launch_bounds__(256)
tail_gate_kernel(const float* __restrict__ g, float* __restrict__ thi, float* __restrict__ tlo, long long M) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;       // one float4 of the 512 outputs
  if (i >= M * (D / 4)) return;
  const long long m = i / (D / 4);
  const int c = (int)(i - m * (D / 4)) * 4;
  const float4 a = ld4(g + m * (2 * D) + c), s = ld4(g + m * (2 * D) + D + c);
  float4 y;
  y.x = tanhf(a.x) * adn_sigmoid(s.x);
  y.y = tanhf(a.y) * adn_sigmoid(s.y);
  y.z = tanhf(a.z) * adn_sigmoid(s.z);
  y.w = tanhf(a.w) * adn_sigmoid(s.w);
  split4(y, thi, tlo, m * D + c);
}

// One CTA per frame: real mask on both STFT row blocks, written into the zero-framed ISTFT operand.
__global__ void __launch_bounds__(256)
mask_apply_kernel(const float* __restrict__ fr, const float* __restrict__ mask, float* __restrict__ ehi,
                  float* __restrict__ elo, int T) {
  const long long m = blockIdx.x;
  const int b = (int)(m / T), t = (int)(m - (long long)b * T);
  const float* st = fr + m * FRONT + KROWS;
  const float* mk = mask + m * BINSP;
  const long long o = ((long long)b * (T + 2 * PADF) + PADF + t) * SPEC_LD;
  for (int r = threadIdx.x; r < SROWS; r += 256) {
    const int f = r >= BINS ? r - BINS : r;
    split_tf32_store(__ldg(st + r) * __ldg(mk + f), ehi, elo, o + r);
  }
}

// (rows, cols) fp32 -> zero-padded (rows_pad, cols_pad) tf32 hi/lo planes
__global__ void pad_split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                 int rows, int cols, int cols_pad, long long total) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int r = (int)(i / cols_pad), c = (int)(i - (long long)r * cols_pad);
  float v = 0.f;
  if (r < rows && c < cols) v = src[(long long)r * cols + c];
  split_tf32_store(v, hi, lo, i);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int choose_bn(int N) {
  const int cands[3] = {256, 176, 128};
  int best = 128, best_pad = 1 << 30;
  for (int c : cands) {
    int pad = round_up(N, c);
    if (pad < best_pad) { best_pad = pad; best = c; }
  }
  return best;
}

// A weight operand: (n_pad, k_pad) tf32 planes per batch + tensor maps.
struct Lin {
  int N = 0, K = 0, n_pad = 0, k_pad = 0, bn = 0, batches = 1;
  float* planes = nullptr;       // owned (weights) or null (activation operand)
  CUtensorMap w_hi, w_lo;
};
struct Gemm {
  tc::TcPlan plan;
  tc::TcArgs args;
};

class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, Lp = 0, T = 0, Tp = 0, T4 = 0, Tn = 0, layers = 24;
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;

  int *d_mel_lo = nullptr, *d_mel_hi = nullptr;
  const float *banks = nullptr, *norm_w = nullptr, *norm_b = nullptr, *emb = nullptr, *rcos = nullptr, *rsin = nullptr;
  const float *mm_w = nullptr, *mm_b = nullptr, *in_w = nullptr, *in_b = nullptr, *prelu_a = nullptr, *gate_b = nullptr;
  const float* d_norm = nullptr;
  struct Layer {
    Lin in, out, c1, uv, ul, up, c2;
    const float *in_b, *in_c, *gamma, *beta, *out_b, *out_c, *c1_b, *c1_a, *n1_w, *n1_b, *uv_b, *uv_c, *ul_b, *mem_c,
        *n2_w, *n2_b, *c2_b;
  };
  std::vector<Layer> lw;
  Lin front, enc, gate, dec, ola;

  int planned = 0;
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  float *xp = nullptr, *xpl = nullptr, *fr = nullptr, *mel = nullptr, *featpl = nullptr, *z = nullptr, *h = nullptr;
  float *xs = nullptr, *rs = nullptr, *proj = nullptr, *vu = nullptr, *vuT = nullptr, *qq = nullptr, *lq = nullptr;
  float *qk = nullptr, *lk = nullptr, *s1 = nullptr, *ppl = nullptr, *att = nullptr, *gated = nullptr, *rs2 = nullptr;
  float *y = nullptr, *hpl = nullptr, *c1y = nullptr, *gin = nullptr, *xn = nullptr, *uvp = nullptr, *uv = nullptr;
  float *xupl = nullptr, *f1 = nullptr, *xp2 = nullptr, *yn = nullptr, *hn = nullptr, *tpl = nullptr, *gbuf = nullptr;
  float *tg = nullptr, *mask = nullptr, *enh = nullptr;
  size_t enh_plane = 0;
  Lin a_qk, a_lk, a_vuT;                   // per-window activation operands (W side)
  Gemm g_front, g_enc, g_lk, g_qk, g_pv, g_gate, g_dec, g_istft;
  struct LayerG { Gemm in, out, c1, uv, ul, up, c2; };
  std::vector<LayerG> lg;
  int stop_after = 0, last_batch = 0;

  ~Model() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    auto fl = [](Lin& l) { if (l.planes) cudaFree(l.planes); };
    for (auto& w : lw) { fl(w.in); fl(w.out); fl(w.c1); fl(w.uv); fl(w.ul); fl(w.up); fl(w.c2); }
    fl(front); fl(enc); fl(gate); fl(dec); fl(ola);
    cudaFree(d_mel_lo); cudaFree(d_mel_hi);
  }
  void free_ws() {
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    ws_bytes = 0;
    planned = 0;
  }

  const float* dptr(const std::string& name, size_t expect, bool& ok) {
    auto it = index.find(name);
    if (it == index.end() || (expect && it->second.count != expect)) {
      if (ok) err = "weight blob: tensor '" + name + "' missing or wrong size";
      ok = false;
      return nullptr;
    }
    return d_blob + it->second.offset;
  }

  // weights (N, K) fp32 on the device -> zero-padded tf32 planes; n_valid_pad: N rounded to 4 for the epilogue
  bool make_lin(Lin& l, const float* src, int N, int K) {
    l.N = N; l.K = K; l.batches = 1;
    l.bn = choose_bn(N);
    l.n_pad = round_up(N, l.bn);
    l.k_pad = round_up(K, 32);
    const long long plane = (long long)l.n_pad * l.k_pad;
    if (cudaMalloc((void**)&l.planes, 2 * plane * sizeof(float)) != cudaSuccess) { err = "out of memory (weights)"; return false; }
    pad_split_kernel<<<(unsigned)((plane + 255) / 256), 256>>>(src, l.planes, l.planes + plane, N, K, l.k_pad, plane);
    return tc::make_weight_map(&l.w_hi, l.planes, l.k_pad, l.n_pad, l.bn, err, 1) &&
           tc::make_weight_map(&l.w_lo, l.planes + plane, l.k_pad, l.n_pad, l.bn, err, 1);
  }
  // activation planes used as the per-window W operand
  bool make_act_lin(Lin& l, float* hi, float* lo, int N, int n_pad, int K, int k_pad, int bn, int batches) {
    l.N = N; l.K = K; l.n_pad = n_pad; l.k_pad = k_pad; l.bn = bn; l.batches = batches; l.planes = nullptr;
    return tc::make_weight_map(&l.w_hi, hi, k_pad, n_pad, bn, err, batches) &&
           tc::make_weight_map(&l.w_lo, lo, k_pad, n_pad, bn, err, batches);
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
        !geti("mf2_layers", layers) || !gets("input_audio_dtype", sin) || !gets("output_audio_dtype", sout))
      return false;
    if (nfft != NFFT || hop != HOP || nmels != NM || L < NFFT || (L - NFFT) % HOP) {
      err = "mossformer2_se needs nfft=1920, hop=384, n_mels=60 and input_audio_length = 1920 + k*384";
      return false;
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = (L - NFFT) / HOP + 1;
    if (T > 256) { err = "mossformer2_se: windows longer than one FLASH group (256 frames) are not supported; fold the audio"; return false; }
    Lp = round_up(L, 4);
    Tp = round_up(T, 32);
    T4 = round_up(T, 4);
    Tn = T4 <= 128 ? 128 : 256;

    bool ok = true;
    banks = dptr("mel_banks", (size_t)NM * KB, ok);
    norm_w = dptr("norm.w", FEAT, ok); norm_b = dptr("norm.b", FEAT, ok);
    emb = dptr("emb_pos", (size_t)T * D, ok);
    rcos = dptr("rot_cos", (size_t)T * ROT, ok); rsin = dptr("rot_sin", (size_t)T * ROT, ok);
    mm_w = dptr("mm_norm.w", D, ok); mm_b = dptr("mm_norm.b", D, ok);
    in_w = dptr("intra_norm.w", D, ok); in_b = dptr("intra_norm.b", D, ok);
    prelu_a = dptr("prelu_a", 1, ok);
    gate_b = dptr("gate_b", 2 * D, ok);
    d_norm = dptr("istft.norm", (size_t)L, ok);
    if (!ok) return false;
    {   // non-zero span of every mel filter
      auto it = index.find("mel_banks");
      const float* bk = h_blob + it->second.offset;
      std::vector<int> lo(NM), hi(NM);
      for (int j = 0; j < NM; ++j) {
        int a = KB, b = 0;
        for (int i = 0; i < KB; ++i)
          if (bk[j * KB + i] != 0.f) { if (i < a) a = i; b = i + 1; }
        if (a > b) a = b = 0;
        lo[j] = a; hi[j] = b;
      }
      if (cudaMalloc((void**)&d_mel_lo, NM * 4) != cudaSuccess || cudaMalloc((void**)&d_mel_hi, NM * 4) != cudaSuccess ||
          cudaMemcpy(d_mel_lo, lo.data(), NM * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(d_mel_hi, hi.data(), NM * 4, cudaMemcpyHostToDevice) != cudaSuccess) { err = "mel table upload failed"; return false; }
    }
    const float* w;
    w = dptr("frontend", (size_t)FRONT * NFFT, ok); if (ok && !make_lin(front, w, FRONT, NFFT)) return false;
    w = dptr("enc.w", (size_t)D * FEAT, ok);        if (ok && !make_lin(enc, w, D, FEAT)) return false;
    w = dptr("gate_w", (size_t)2 * D * D, ok);      if (ok && !make_lin(gate, w, 2 * D, D)) return false;
    w = dptr("dec_w", (size_t)BINS * D, ok);        if (ok && !make_lin(dec, w, BINS, D)) return false;
    lw.resize(layers);
    for (int i = 0; i < layers && ok; ++i) {
      const std::string p = "L" + std::to_string(i) + ".";
      Layer& Y = lw[i];
      w = dptr(p + "in_w", (size_t)PROJ * D, ok);  if (ok && !make_lin(Y.in, w, PROJ, D)) return false;
      w = dptr(p + "out_w", (size_t)D * VU, ok);   if (ok && !make_lin(Y.out, w, D, VU)) return false;
      w = dptr(p + "c1_w", (size_t)FI * D, ok);    if (ok && !make_lin(Y.c1, w, FI, D)) return false;
      w = dptr(p + "uv_w", (size_t)2 * FI * FI, ok); if (ok && !make_lin(Y.uv, w, 2 * FI, FI)) return false;
      w = dptr(p + "ul_w", (size_t)FI * FI, ok);   if (ok && !make_lin(Y.ul, w, FI, FI)) return false;
      w = dptr(p + "up_w", (size_t)FI * FI, ok);   if (ok && !make_lin(Y.up, w, FI, FI)) return false;
      w = dptr(p + "c2_w", (size_t)D * FI, ok);    if (ok && !make_lin(Y.c2, w, D, FI)) return false;
      Y.in_b = dptr(p + "in_b", PROJ, ok); Y.in_c = dptr(p + "in_c", (size_t)DW * PROJ, ok);
      Y.gamma = dptr(p + "qk_gamma", 4 * QK, ok); Y.beta = dptr(p + "qk_beta", 4 * QK, ok);
      Y.out_b = dptr(p + "out_b", D, ok); Y.out_c = dptr(p + "out_c", (size_t)DW * D, ok);
      Y.c1_b = dptr(p + "c1_b", FI, ok); Y.c1_a = dptr(p + "c1_a", 1, ok);
      Y.n1_w = dptr(p + "n1_w", FI, ok); Y.n1_b = dptr(p + "n1_b", FI, ok);
      Y.uv_b = dptr(p + "uv_b", 2 * FI, ok); Y.uv_c = dptr(p + "uv_c", (size_t)DW * 2 * FI, ok);
      Y.ul_b = dptr(p + "ul_b", FI, ok); Y.mem_c = dptr(p + "mem_c", (size_t)MEMK * FI, ok);
      Y.n2_w = dptr(p + "n2_w", FI, ok); Y.n2_b = dptr(p + "n2_b", FI, ok);
      Y.c2_b = dptr(p + "c2_b", D, ok);
    }
    if (!ok) return false;
    {   // overlap-add weight: raw hop-block j = sum over the R frames that cover it (see api.cu build_ola_weight)
      auto inv = index.find("istft.inv");
      if (inv == index.end() || inv->second.count != (size_t)SROWS * NFFT) { err = "missing istft.inv"; return false; }
      const float* ib = h_blob + inv->second.offset;
      const int K = R_OLA * SPEC_LD;
      std::vector<float> wv((size_t)HOP * K, 0.f);
      for (int n = 0; n < HOP; ++n)
        for (int q = 0; q < R_OLA; ++q) {
          const int src = n + (R_OLA - 1 - q) * HOP;
          if (src >= NFFT) continue;
          for (int r = 0; r < SROWS; ++r) wv[(size_t)n * K + (size_t)q * SPEC_LD + r] = ib[(size_t)r * NFFT + src];
        }
      float* tmp = nullptr;
      if (cudaMalloc((void**)&tmp, wv.size() * 4) != cudaSuccess ||
          cudaMemcpy(tmp, wv.data(), wv.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { err = "ola upload failed"; return false; }
      const bool r = make_lin(ola, tmp, HOP, K);
      cudaDeviceSynchronize();
      cudaFree(tmp);
      if (!r) return false;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "weight split failed"; return false; }
    return true;
  }

  bool alloc(float*& p, size_t nfloats, bool zero) {
    if (cudaMalloc((void**)&p, nfloats * sizeof(float)) != cudaSuccess) { err = "out of device memory (workspace)"; return false; }
    allocs.push_back(p);
    ws_bytes += nfloats * sizeof(float);
    if (zero) cudaMemset(p, 0, nfloats * sizeof(float));
    return true;
  }

  bool plan_gemm(Gemm& g, const float* a_planes, long long a_plane_stride, int K, int rows, long long row_stride,
                 int batches, long long batch_stride, const Lin& l) {
    const int bt = rows >= 128 ? 128 : rows;
    g.plan.bn = l.bn;
    g.plan.map_w_hi = l.w_hi;
    g.plan.map_w_lo = l.w_lo;
    if (!tc::make_row_map(&g.plan.map_a_hi, a_planes, K, rows, row_stride, batches, batch_stride, bt, 1, err) ||
        !tc::make_row_map(&g.plan.map_a_lo, a_planes + a_plane_stride, K, rows, row_stride, batches, batch_stride, bt, 1, err))
      return false;
    tc::TcArgs& a = g.args;
    a = tc::TcArgs{};
    a.bb = 1; a.bt = bt; a.tiles_per_chunk = (rows + 127) / 128; a.t0 = 0;
    a.B = batches; a.TM = rows; a.N = l.N; a.K = l.K;
    a.m_tiles = batches * a.tiles_per_chunk;
    a.w_batched = l.batches > 1;
    return true;
  }

  size_t floats_needed(size_t B) const {
    const size_t M = B * T;
    return B * Lp * 3 + M * FRONT + M * NM + 2 * M * FEATP + 2 * M * D + 2 * M * D + 2 * M + M * PROJ + M * VU2 +
           2 * B * VU2 * Tp + 4 * M * QK + 4 * B * Tn * QK + 3 * M * Tp + M * VU2 +
           2 * M * VU + M * D + 2 * M * D + 2 * M * FI + 2 * M * FI + 2 * M * D + 2 * M * FI + 2 * M * FI + M * FI +
           2 * M * FI + M * D + 2 * M * D + M * 2 * D + 2 * M * D + M * BINSP + 2 * (B * (T + 2 * PADF) * SPEC_LD + 9664);
  }

  bool ensure(int B) {
    if (B == planned) return true;
    cudaDeviceSynchronize();
    free_ws();
    const long long M = (long long)B * T;
    const size_t xplane = (size_t)B * Lp;
    enh_plane = (size_t)B * (T + 2 * PADF) * SPEC_LD + ola.k_pad;       // + slack for the K overrun of the last block
    if (!alloc(xp, xplane, false) || !alloc(xpl, 2 * xplane, false) || !alloc(fr, (size_t)M * FRONT, false) ||
        !alloc(mel, (size_t)M * NM, false) || !alloc(featpl, 2 * (size_t)M * FEATP, true) || !alloc(z, (size_t)M * D, false) ||
        !alloc(h, (size_t)M * D, false) || !alloc(xs, 2 * (size_t)M * D, false) || !alloc(rs, (size_t)M, false) ||
        !alloc(proj, (size_t)M * PROJ, false) || !alloc(vu, (size_t)M * VU2, false) ||
        !alloc(vuT, 2 * (size_t)B * VU2 * Tp, true) || !alloc(qq, 2 * (size_t)M * QK, false) ||
        !alloc(lq, 2 * (size_t)M * QK, false) || !alloc(qk, 2 * (size_t)B * Tn * QK, true) ||
        !alloc(lk, 2 * (size_t)B * Tn * QK, true) || !alloc(s1, (size_t)M * Tp, true) ||
        !alloc(ppl, 2 * (size_t)M * Tp, true) || !alloc(att, (size_t)M * VU2, false) ||
        !alloc(gated, 2 * (size_t)M * VU, false) || !alloc(rs2, (size_t)M, false) || !alloc(y, (size_t)M * D, false) ||
        !alloc(hpl, 2 * (size_t)M * D, false) || !alloc(c1y, (size_t)M * FI, false) || !alloc(gin, (size_t)M * FI, false) ||
        !alloc(xn, 2 * (size_t)M * FI, false) || !alloc(uvp, (size_t)M * 2 * FI, false) || !alloc(uv, (size_t)M * 2 * FI, false) ||
        !alloc(xupl, 2 * (size_t)M * FI, false) || !alloc(f1, 2 * (size_t)M * FI, false) || !alloc(xp2, (size_t)M * FI, false) ||
        !alloc(yn, 2 * (size_t)M * FI, false) || !alloc(hn, (size_t)M * D, false) || !alloc(tpl, 2 * (size_t)M * D, false) ||
        !alloc(gbuf, (size_t)M * 2 * D, false) || !alloc(tg, 2 * (size_t)M * D, false) || !alloc(mask, (size_t)M * BINSP, false) ||
        !alloc(enh, 2 * enh_plane, true))
      return false;

    // frontend: rows = frames (stride hop) of the raw window
    if (!plan_gemm(g_front, xpl, (long long)xplane, NFFT, T, HOP, B, Lp, front)) return false;
    g_front.args.C = fr; g_front.args.ldc = FRONT;
    if (!plan_gemm(g_enc, featpl, M * FEATP, FEATP, (int)M, FEATP, 1, M * FEATP, enc)) return false;
    g_enc.args.K = FEATP; g_enc.args.resid = z; g_enc.args.C = z; g_enc.args.ldc = D;

    // attention operands that live in activations
    const int bn_qk = Tn;                    // 128 or 256 keys per tile
    if (!make_act_lin(a_qk, qk, qk + (size_t)B * Tn * QK, T4, Tn, QK, QK, bn_qk, B) ||
        !make_act_lin(a_lk, lk, lk + (size_t)B * Tn * QK, T4, Tn, QK, QK, bn_qk, B) ||
        !make_act_lin(a_vuT, vuT, vuT + (size_t)B * VU2 * Tp, VU2, VU2, Tp, Tp, 256, B))
      return false;
    if (!plan_gemm(g_lk, lq, M * QK, QK, T, QK, B, (long long)T * QK, a_lk)) return false;
    g_lk.args.C = s1; g_lk.args.ldc = Tp;
    if (!plan_gemm(g_qk, qq, M * QK, QK, T, QK, B, (long long)T * QK, a_qk)) return false;
    g_qk.args.act = tc::ACT_RELU2; g_qk.args.resid = s1; g_qk.args.Chi = ppl; g_qk.args.Clo = ppl + M * Tp; g_qk.args.ldc = Tp;
    if (!plan_gemm(g_pv, ppl, M * Tp, Tp, T, Tp, B, (long long)T * Tp, a_vuT)) return false;
    g_pv.args.C = att; g_pv.args.ldc = VU2;

    lg.assign(layers, LayerG{});
    for (int i = 0; i < layers; ++i) {
      LayerG& G = lg[i];
      const Layer& Y = lw[i];
      if (!plan_gemm(G.in, xs, M * D, D, (int)M, D, 1, M * D, Y.in)) return false;
      G.in.args.rowscale = rs; G.in.args.bias = Y.in_b; G.in.args.act = tc::ACT_SILU; G.in.args.C = proj; G.in.args.ldc = PROJ;
      if (!plan_gemm(G.out, gated, M * VU, VU, (int)M, VU, 1, M * VU, Y.out)) return false;
      G.out.args.rowscale = rs2; G.out.args.bias = Y.out_b; G.out.args.act = tc::ACT_SILU; G.out.args.C = y; G.out.args.ldc = D;
      if (!plan_gemm(G.c1, hpl, M * D, D, (int)M, D, 1, M * D, Y.c1)) return false;
      G.c1.args.bias = Y.c1_b; G.c1.args.act = tc::ACT_PRELU; G.c1.args.act_param = Y.c1_a; G.c1.args.C = c1y; G.c1.args.ldc = FI;
      if (!plan_gemm(G.uv, xn, M * FI, FI, (int)M, FI, 1, M * FI, Y.uv)) return false;
      G.uv.args.bias = Y.uv_b; G.uv.args.act = tc::ACT_SILU; G.uv.args.C = uvp; G.uv.args.ldc = 2 * FI;
      if (!plan_gemm(G.ul, xupl, M * FI, FI, (int)M, FI, 1, M * FI, Y.ul)) return false;
      G.ul.args.bias = Y.ul_b; G.ul.args.act = tc::ACT_RELU; G.ul.args.Chi = f1; G.ul.args.Clo = f1 + M * FI; G.ul.args.ldc = FI;
      if (!plan_gemm(G.up, f1, M * FI, FI, (int)M, FI, 1, M * FI, Y.up)) return false;
      G.up.args.C = xp2; G.up.args.ldc = FI;
      if (!plan_gemm(G.c2, yn, M * FI, FI, (int)M, FI, 1, M * FI, Y.c2)) return false;
      G.c2.args.bias = Y.c2_b; G.c2.args.resid = h; G.c2.args.C = h; G.c2.args.ldc = D;
    }
    if (!plan_gemm(g_gate, tpl, M * D, D, (int)M, D, 1, M * D, gate)) return false;
    g_gate.args.bias = gate_b; g_gate.args.C = gbuf; g_gate.args.ldc = 2 * D;
    if (!plan_gemm(g_dec, tg, M * D, D, (int)M, D, 1, M * D, dec)) return false;
    g_dec.args.N = BINSP; g_dec.args.act = tc::ACT_RELU; g_dec.args.C = mask; g_dec.args.ldc = BINSP;

    {   // inverse: rows = raw hop blocks, each a run of R consecutive zero-framed spectrum frames
      const int rows = T + 2 * PADF;
      const int TM = T + R_OLA - 1;          // blocks 0 .. (raw-1)/hop
      if (!plan_gemm(g_istft, enh, (long long)enh_plane, ola.k_pad, rows, SPEC_LD, B, (long long)rows * SPEC_LD, ola)) return false;
      tc::TcArgs& c = g_istft.args;
      const int bt = TM >= 128 ? 128 : TM;
      c.bt = bt; c.tiles_per_chunk = (TM + 127) / 128; c.m_tiles = B * c.tiles_per_chunk; c.TM = TM; c.t0 = 0;
      c.N = HOP; c.K = R_OLA * SPEC_LD;
      c.norm = d_norm; c.norm_mul = 0; c.hop = HOP; c.shift = 0; c.out_len = L; c.out_dtype = out_dtype; c.i16_mode = 1;
      // the A box must match bt rows
      if (!tc::make_row_map(&g_istft.plan.map_a_hi, enh, ola.k_pad, rows, SPEC_LD, B, (long long)rows * SPEC_LD, bt, 1, err) ||
          !tc::make_row_map(&g_istft.plan.map_a_lo, enh + enh_plane, ola.k_pad, rows, SPEC_LD, B, (long long)rows * SPEC_LD, bt, 1, err))
        return false;
    }
    planned = B;
    return true;
  }

  // ---- ModelImpl
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);        // Export_MossFormer_SE.py:537
    in->dtype = in_dtype; in->channels = 1; in->length = L;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);   // :538
    out->dtype = out_dtype; out->channels = 1; out->length = L;
  }
  size_t workspace_bytes(int batch) override { return floats_needed((size_t)batch) * sizeof(float); }
  int launches(int) override { return 5 + layers * 17 + 6; }
  void set_stop_after(int n) override { stop_after = n; }

#define MF_TICK(name) do { ++n; if (tick) tick(tick_ctx, name); if (stop_after > 0 && n >= stop_after) return ADN_OK; } while (0)
#define MF_GEMM(G, epi, name) do { cudaError_t e_ = tc::launch((G).plan, (G).args, epi, sms, st); \
    if (e_ != cudaSuccess) { err = std::string("gemm launch (") + name + "): " + cudaGetErrorString(e_); return ADN_ERR_CUDA; } MF_TICK(name); } while (0)

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    int n = 0;
    const long long M = (long long)B * T;
    const unsigned wtok = (unsigned)((M + 7) / 8);
    static bool cfg = false;
    if (!cfg) {
      cudaFuncSetAttribute(featnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 2 * NM * 4);
      cudaFuncSetAttribute(fsmn_mem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * FM_TOK + 2 * MEMH) * FI * 4);
      cfg = true;
    }

    // 1-3: cast (+1/32768 for int16, :315-317), fused Kaldi||STFT frontend (:335), log-mel (:337-341)
    gtcrn::launch_prep(d_in, in_dtype, xp, xpl, xpl + (size_t)B * Lp, B, L, Lp, 0, 0, 0, st);
    MF_TICK("prep");
    MF_GEMM(g_front, EPI_LIN, "frontend_gemm");
    feat_kernel<<<(unsigned)M, 256, 0, st>>>(fr, banks, d_mel_lo, d_mel_hi, mel, 1.1920929e-07f * (1.0f / 32768.0f) * (1.0f / 32768.0f),
                                            20.794415416798357f);
    MF_TICK("feat");
    featnorm_kernel<<<B, 256, (size_t)T * 2 * NM * sizeof(float), st>>>(mel, norm_w, norm_b, emb, featpl, featpl + M * FEATP, z, T);
    MF_TICK("featnorm");
    MF_GEMM(g_enc, EPI_LIN, "enc_gemm");

    for (int i = 0; i < layers; ++i) {
      LayerG& G = lg[i];
      const Layer& Y = lw[i];
      const float* hin = i == 0 ? z : h;
      shiftnorm_kernel<<<wtok, 256, 0, st>>>(hin, xs, xs + M * D, rs, M, T);
      MF_TICK("shiftnorm");
      MF_GEMM(G.in, EPI_LIN, "fl_in");
      dwconv_in_kernel<<<dim3(PROJ / 32 / DWI_WARPS, B), DWI_WARPS * 32, 0, st>>>(
          proj, Y.in_c, Y.gamma, Y.beta, rcos, rsin, vu, vuT, vuT + (size_t)B * VU2 * Tp, qq, qq + M * QK, lq, lq + M * QK,
          qk, qk + (size_t)B * Tn * QK, lk, lk + (size_t)B * Tn * QK, T, Tp, Tn);
      MF_TICK("dwconv_in");
      MF_GEMM(g_lk, EPI_LIN, "att_lk");
      MF_GEMM(g_qk, EPI_LIN, "att_qk");
      MF_GEMM(g_pv, EPI_LIN, "att_pv");
      gate_kernel<<<wtok, 256, 0, st>>>(att, vu, gated, gated + M * VU, rs2, M);
      MF_TICK("gate");
      MF_GEMM(G.out, EPI_LIN, "fl_out");
      dwconv_kernel<<<dim3(D / 32 / DWI_WARPS, B), DWI_WARPS * 32, 0, st>>>(y, Y.out_c, hin, h, hpl, hpl + M * D, D, T, D);
      MF_TICK("dwconv_out");
      MF_GEMM(G.c1, EPI_LIN, "fsmn_conv1");
      ln2_kernel<<<wtok, 256, 0, st>>>(c1y, Y.n1_w, Y.n1_b, gin, xn, xn + M * FI, M);
      MF_TICK("ln2");
      MF_GEMM(G.uv, EPI_LIN, "fsmn_uv");
      dwconv_kernel<<<dim3(2 * FI / 32 / DWI_WARPS, B), DWI_WARPS * 32, 0, st>>>(uvp, Y.uv_c, nullptr, uv, xupl, xupl + M * FI, FI, T, 2 * FI);
      MF_TICK("dwconv_uv");
      MF_GEMM(G.ul, EPI_LIN, "fsmn_linear");
      MF_GEMM(G.up, EPI_LIN, "fsmn_project");
      fsmn_mem_kernel<<<dim3((T + FM_TOK - 1) / FM_TOK, B), 256, (2 * FM_TOK + 2 * MEMH) * FI * sizeof(float), st>>>(xp2, uv, gin, Y.mem_c, Y.n2_w, Y.n2_b, yn, yn + M * FI, T);
      MF_TICK("fsmn_mem");
      MF_GEMM(G.c2, EPI_LIN, "fsmn_conv2");
    }

    tail_norm_kernel<<<B, 512, 0, st>>>(layers ? h : z, z, mm_w, mm_b, in_w, in_b, prelu_a, hn, tpl, tpl + M * D, T);
    MF_TICK("tail_norm");
    MF_GEMM(g_gate, EPI_LIN, "tail_gate_gemm");
    tail_gate_kernel<<<(unsigned)((M * (D / 4) + 255) / 256), 256, 0, st>>>(gbuf, tg, tg + M * D, M);
    MF_TICK("tail_gate");
    MF_GEMM(g_dec, EPI_LIN, "mask_gemm");
    mask_apply_kernel<<<(unsigned)M, 256, 0, st>>>(fr, mask, enh, enh + enh_plane, T);
    MF_TICK("mask_apply");
    g_istft.args.out = d_out;
    MF_GEMM(g_istft, EPI_ISTFT, "istft_gemm");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mf2se run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    const size_t B = last_batch;
    if (!B) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    const size_t M = B * T;
    std::map<std::string, std::pair<const float*, size_t>> tbl = {
        {"fr", {fr, M * FRONT}}, {"mel", {mel, M * NM}}, {"z", {z, M * D}}, {"h", {h, M * D}}, {"proj", {proj, M * PROJ}},
        {"vu", {vu, M * VU2}}, {"att", {att, M * VU2}}, {"y", {y, M * D}}, {"c1y", {c1y, M * FI}}, {"gin", {gin, M * FI}},
        {"uv", {uv, M * 2 * FI}}, {"xp2", {xp2, M * FI}}, {"gate", {gbuf, M * 2 * D}}, {"mask", {mask, M * BINSP}},
        {"rs", {rs, M}}, {"rs2", {rs2, M}},
    };
    auto it = tbl.find(name);
    if (it == tbl.end()) { err = std::string("adn_debug_read: unknown tensor '") + name + "'"; return ADN_ERR_INVALID; }
    if (actual) *actual = it->second.second;
    if (!h_dst) return ADN_OK;
    const size_t nc = count < it->second.second ? count : it->second.second;
    cudaDeviceSynchronize();
    if (cudaMemcpy(h_dst, it->second.first, nc * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
      err = "debug copy failed";
      return ADN_ERR_CUDA;
    }
    return ADN_OK;
  }
};

}  // namespace mf2

ModelImpl* mf2se_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  mf2::Model* m = new mf2::Model();
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
