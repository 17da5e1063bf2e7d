// Shared device kernels and host helpers of the MossFormer2 family (SE-48K: mf2se.cu, SS-16K: mf2ss.cu):
// FLASH block pieces (token shift + ScaleNorm, streamed depthwise convs, OffsetScale + rotary, gate), gated-FSMN
// pieces (LayerNorm pair), mask-tail pieces, and the tcgen05 GEMM planning helpers.  Included by exactly those two
// translation units; every kernel has internal linkage.
#pragma once
#include "adn.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "gtcrn.cuh"
#include "model_impl.h"
#include "tma_utils.cuh"

#include <cuda_bf16.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace mf2 {

// GEMM operand stores.  3xTF32 W-side operands: value -> tf32 hi / lo planes.  bf16 matmul mode (lo == nullptr; only where
// the workload spec licenses bf16 matmuls): ONE bf16 plane living in the hi buffer at the same element offsets.  fp32-A mode
// (lo == hi): the consumer is a GEMM that splits its A tile itself (gemm_tc.cu, TcPlan::a_f32), so the value is stored once, as fp32.
__device__ __forceinline__ void split_tf32_store(float v, float* hi, float* lo, long long i) {
  if (lo == hi) { hi[i] = v; return; }
  if (lo == nullptr) {
    reinterpret_cast<__nv_bfloat16*>(hi)[i] = __float2bfloat16_rn(v);
    return;
  }
  const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = v - h;
}

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
  if (lo == hi) { st4(hi + i, v); return; }   // fp32-A mode
  if (lo == nullptr) {                        // bf16 plane: 4 values = 8 bytes
    __nv_bfloat162 p01 = __floats2bfloat162_rn(v.x, v.y), p23 = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p01);
    pk.y = *reinterpret_cast<uint32_t*>(&p23);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(hi) + i) = pk;
    return;
  }
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
  st4(hi + i, h);
  st4(lo + i, l);
}

// One warp per token: first half of the channels comes from the previous frame (zero at t = 0);
// rs = 1 / (||x|| + eps) (or 1 / max(||x||, eps)) is applied as a row scale by the consuming GEMM.
static __global__ void __launch_bounds__(256)
shiftnorm_kernel(const float* __restrict__ h, float* __restrict__ xhi, float* __restrict__ xlo,
                 float* __restrict__ rs, long long M, int T, int clamp) {
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
  // SE: x / (|x| + eps) (Export_MossFormer_SE.py:397); SS: x / clamp(|x|, min=eps) (Export_MossFormer2_SS_16K.py:467)
  if (lane == 0) rs[m] = clamp ? 1.0f / fmaxf(sqrtf(ss), EPS_IN) : 1.0f / (sqrtf(ss) + EPS_IN);
}

// Depthwise k=17 'same' conv over time + residual, streamed: one warp owns a 32-channel strip of one
// window (lane = channel) and walks the frames [t_begin, t_end) once (long windows are cut into segments so
// that enough warps are in flight; the halo frames come from global memory, zeros outside [0, T)).  The
// 17-frame window plus the prefetched frames live in a RING-register ring (static indices after unrolling),
// so every frame costs one coalesced 128-byte global load issued RING-16 frames ahead of its first use,
// 17 FMAs and no shared memory or CTA barrier.  emit(t, value) consumes frame t; flush(t0) runs after every
// 32 frames (t_begin must be a multiple of 32).
template <int RING, typename F, typename G>
__device__ __forceinline__ void dwconv_stream(const float* __restrict__ src, long long ld, const float (&w)[DW], int T,
                                              int t_begin, int t_end, F&& emit, G&& flush) {
  static_assert(RING % 32 == 0 && RING > DW, "ring = whole 32-frame groups");
  float ring[RING];
#pragma unroll
  for (int i = 0; i < RING; ++i) {
    const int r = t_begin + i - DWH;
    ring[i] = (r >= 0 && r < T) ? __ldg(src + (long long)r * ld) : 0.f;
  }
  for (int t0 = t_begin; t0 < t_end; t0 += RING) {
#pragma unroll
    for (int j = 0; j < RING; ++j) {
      const int t = t0 + j;
      if (t < t_end) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < DW; ++k) acc += w[k] * ring[(j + k) % RING];
        acc += ring[(j + DWH) % RING];
        emit(t, acc);
      }
      const int r = t + RING - DWH;
      ring[j] = r < T ? __ldg(src + (long long)r * ld) : 0.f;
      if ((j & 31) == 31 && t0 + j - 31 < t_end) flush(t0 + j - 31);
    }
  }
}

// frames per time segment of the streamed depthwise convs (blockIdx.z)
constexpr int DW_SEG = 256;

// ConvModule residual on the fused to_hidden||to_qk projection; CTA = 4 warps = 4 adjacent 32-channel
// strips of one window.  Strips of the 2048 value channels write [v|u] (token-major fp32, for the
// gate) and [v|u]^T (tf32 planes, the attention operand; 32x32 per-warp transposes); the last CTA
// column holds the 128 qk channels: four OffsetScale heads + rotary embedding -> quad_q / lin_q /
// quad_k / lin_k (token-major planes; queries at row window*Tq + frame, keys at window*Tn + frame).  With lkT_hi set
// (windows of several FLASH groups) lin_k is emitted transposed, (window, 128, Tp), for the global K^T[v|u] product.
constexpr int DWI_WARPS = 4;
static __global__ void __launch_bounds__(DWI_WARPS * 32)
dwconv_in_kernel(const float* __restrict__ proj, const float* __restrict__ taps, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ rcos, const float* __restrict__ rsin,
                 float* __restrict__ vu, float* __restrict__ vuT_hi, float* __restrict__ vuT_lo,
                 float* __restrict__ qq_hi, float* __restrict__ qq_lo, float* __restrict__ lq_hi,
                 float* __restrict__ lq_lo, float* __restrict__ qk_hi, float* __restrict__ qk_lo,
                 float* __restrict__ lk_hi, float* __restrict__ lk_lo, float* __restrict__ lkT_hi,
                 float* __restrict__ lkT_lo, int T, int Tp, int Tn, int Tq, int lq_ld) {
  __shared__ float stage[DWI_WARPS][32 * 33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * DWI_WARPS + warp) * 32, b = blockIdx.y;
  const int t_begin = blockIdx.z * DW_SEG, t_end = min(T, t_begin + DW_SEG);
  const float* src = proj + (long long)b * T * PROJ + c0 + lane;
  float w[DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) w[k] = __ldg(taps + k * PROJ + c0 + lane);
  if (c0 < VU2) {
    float* st = stage[warp];
    dwconv_stream<64>(src, PROJ, w, T, t_begin, t_end,
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
    dwconv_stream<64>(src, PROJ, w, T, t_begin, t_end,
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
                    const long long m = (long long)b * Tq + t;
                    split_tf32_store(s[0], qq_hi, qq_lo, m * QK + q);
                    split_tf32_store(s[1], lq_hi, lq_lo, m * lq_ld + q);     // lq_ld > QK: columns of a wider operand
                    split_tf32_store(s[2], qk_hi, qk_lo, ((long long)b * Tn + t) * QK + q);
                    if (lkT_hi) stage[warp][(t & 31) * 33 + lane] = s[3];      // multi-group windows: lin_k^T (SS)
                    else split_tf32_store(s[3], lk_hi, lk_lo, ((long long)b * Tn + t) * QK + q);
                  },
                  [&](int t0) {
                    if (!lkT_hi) return;
                    __syncwarp();
                    const int t = t0 + lane;
                    if (t < T) {
#pragma unroll 8
                      for (int c = 0; c < 32; ++c)
                        split_tf32_store(stage[warp][lane * 33 + c], lkT_hi, lkT_lo, ((long long)b * QK + (c0 - VU2) + c) * Tp + t);
                    }
                    __syncwarp();
                  });
  }
}

// TMA-fed, double-buffered form of dwconv_in_kernel.  CTA = 128 channels x up to DP_TILES consecutive 64-frame tiles
// of one window.  Each (64 + 16) x 128 input tile arrives as ONE bulk tensor copy (512-byte row segments; rows before
// the window start or past its end are zero-filled by the TMA unit, which is the conv's zero padding) into one of two
// shared-memory buffers, so tile i+1 is in flight while tile i is processed.  Every thread owns one channel and 32
// frames (48-value register window, 17 FMAs per output); the transposed operands ([v|u]^T, lin_k^T) leave through a
// 128 x 65 tile that aliases the buffer just consumed, as 128-byte row segments.  Same arguments and arithmetic order
// as the streamed kernel; `map` is a tile map over proj: dims (2176, T, windows), box (128, 80, 1).
constexpr int DP_F = 64, DP_C = 128, DP_ROWS = DP_F + 2 * DWH, DP_TILES = 4;
constexpr size_t DP_SMEM = 2 * DP_ROWS * DP_C * sizeof(float) + 64;
static __global__ void __launch_bounds__(256, 2)
dwconv_in_tma_kernel(const __grid_constant__ CUtensorMap map, const float* __restrict__ taps, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ rcos, const float* __restrict__ rsin,
                     float* __restrict__ vu, float* __restrict__ vuT_hi, float* __restrict__ vuT_lo,
                     float* __restrict__ qq_hi, float* __restrict__ qq_lo, float* __restrict__ lq_hi,
                     float* __restrict__ lq_lo, float* __restrict__ qk_hi, float* __restrict__ qk_lo,
                     float* __restrict__ lk_hi, float* __restrict__ lk_lo, float* __restrict__ lkT_hi,
                     float* __restrict__ lkT_lo, int T, int Tp, int Tn, int Tq, int lq_ld) {
  extern __shared__ __align__(128) float dsm[];
  float* buf[2] = {dsm, dsm + DP_ROWS * DP_C};
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + 2 * DP_ROWS * DP_C);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.x * DP_C, b = blockIdx.y, seg = blockIdx.z * (DP_TILES * DP_F);
  const int ntiles = min(DP_TILES, (T - seg + DP_F - 1) / DP_F);
  if (tid == 0) {
    tc::mbar_init(&full[0], 1);
    tc::mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < 2 && i < ntiles; ++i) {
      tc::mbar_expect_tx(&full[i], DP_ROWS * DP_C * 4);
      tc::tma_load_3d(buf[i], &map, &full[i], c0, seg + i * DP_F - DWH, b);
    }
  }
  const int ch = tid & (DP_C - 1), f0 = (tid >> 7) * (DP_F / 2), c = c0 + ch;
  float w[DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) w[k] = __ldg(taps + k * PROJ + c);
  float g4[4], b4[4];
  if (c0 >= VU2) {
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) { g4[hd] = __ldg(gamma + hd * QK + ch); b4[hd] = __ldg(beta + hd * QK + ch); }
  }
  for (int i = 0; i < ntiles; ++i) {
    const int s = i & 1, t0 = seg + i * DP_F;
    tc::mbar_wait(&full[s], (i >> 1) & 1);
    float x[DP_F / 2 + 2 * DWH];
#pragma unroll
    for (int k = 0; k < DP_F / 2 + 2 * DWH; ++k) x[k] = buf[s][(f0 + k) * DP_C + ch];
    __syncthreads();                             // every column of buffer s is in registers: it can be reused
    float y[DP_F / 2];
#pragma unroll
    for (int j = 0; j < DP_F / 2; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < DW; ++k) acc += w[k] * x[j + k];
      y[j] = acc + x[j + DWH];
    }
    float* tt = buf[s];                          // [DP_C][DP_F + 1]
    if (c0 < VU2) {
#pragma unroll
      for (int j = 0; j < DP_F / 2; ++j) {
        const int t = t0 + f0 + j;
        if (t < T) vu[((long long)b * T + t) * VU2 + c] = y[j];
        tt[ch * (DP_F + 1) + f0 + j] = y[j];
      }
      __syncthreads();
      for (int cc = warp * (DP_C / 8); cc < (warp + 1) * (DP_C / 8); ++cc) {
#pragma unroll
        for (int h = 0; h < DP_F / 32; ++h) {
          const int f = lane + 32 * h, t = t0 + f;
          if (t < T) split_tf32_store(tt[cc * (DP_F + 1) + f], vuT_hi, vuT_lo, ((long long)b * VU2 + c0 + cc) * Tp + t);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < DP_F / 2; ++j) {
        const int t = t0 + f0 + j;
        const int tc_ = t < T ? t : T - 1;       // (whole warps stay converged for the shuffles)
        float sv[4];
#pragma unroll
        for (int hd = 0; hd < 4; ++hd) sv[hd] = y[j] * g4[hd] + b4[hd];
        if (ch < ROT) {                          // rotary on the first 32 qk channels, interleaved pairs (warp-uniform)
          const float cs = __ldg(rcos + tc_ * ROT + lane), sn = __ldg(rsin + tc_ * ROT + lane);
#pragma unroll
          for (int hd = 0; hd < 4; ++hd) {
            const float other = __shfl_xor_sync(0xffffffffu, sv[hd], 1);
            const float rot = (lane & 1) ? other : -other;
            sv[hd] = sv[hd] * cs + rot * sn;
          }
        }
        if (lkT_hi) tt[ch * (DP_F + 1) + f0 + j] = sv[3];
        if (t < T) {
          const long long m = (long long)b * Tq + t;
          split_tf32_store(sv[0], qq_hi, qq_lo, m * QK + ch);
          split_tf32_store(sv[1], lq_hi, lq_lo, m * lq_ld + ch);
          split_tf32_store(sv[2], qk_hi, qk_lo, ((long long)b * Tn + t) * QK + ch);
          if (!lkT_hi) split_tf32_store(sv[3], lk_hi, lk_lo, ((long long)b * Tn + t) * QK + ch);
        }
      }
      __syncthreads();
      if (lkT_hi) {                              // multi-group windows: lin_k^T (SS)
        for (int cc = warp * (DP_C / 8); cc < (warp + 1) * (DP_C / 8); ++cc) {
#pragma unroll
          for (int h = 0; h < DP_F / 32; ++h) {
            const int f = lane + 32 * h, t = t0 + f;
            if (t < T) split_tf32_store(tt[cc * (DP_F + 1) + f], lkT_hi, lkT_lo, ((long long)b * QK + cc) * Tp + t);
          }
        }
      }
    }
    __syncthreads();                             // transpose tile drained: buffer s may be overwritten
    if (tid == 0 && i + 2 < ntiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc::mbar_expect_tx(&full[s], DP_ROWS * DP_C * 4);
      tc::tma_load_3d(buf[s], &map, &full[s], c0, seg + (i + 2) * DP_F - DWH, b);
    }
  }
}

// One warp per token: gate and ScaleNorm denominator of to_out.
static __global__ void __launch_bounds__(256)
gate_kernel(const float* __restrict__ att, const float* __restrict__ vu, float* __restrict__ ghi,
            float* __restrict__ glo, float* __restrict__ rs, long long M, int T, int Tq, int clamp) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  // attention rows live in the group-padded layout (window*Tq + frame); Tq == T for one-group windows
  const float* a = att + ((m / T) * Tq + (m % T)) * VU2;
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
  if (lane == 0) rs[m] = clamp ? 1.0f / fmaxf(sqrtf(ss), EPS_OUT) : 1.0f / (sqrtf(ss) + EPS_OUT);
}

// out = x + dwconv17(x) (+ resid); optional tf32 planes of the first `plane_cols` channels.
// CTA = 4 warps = 4 adjacent 32-channel strips of one window (see dwconv_stream).
static __global__ void __launch_bounds__(DWI_WARPS * 32)
dwconv_kernel(const float* __restrict__ x, const float* __restrict__ taps, const float* __restrict__ resid,
              float* __restrict__ out, float* __restrict__ phi, float* __restrict__ plo, int plane_cols, int T, int C) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * DWI_WARPS + warp) * 32 + lane, b = blockIdx.y;
  float w[DW];
#pragma unroll
  for (int k = 0; k < DW; ++k) w[k] = __ldg(taps + k * C + c);
  const bool planes = phi && c < plane_cols;
  const int t_begin = blockIdx.z * DW_SEG, t_end = min(T, t_begin + DW_SEG);
  dwconv_stream<32>(x + (long long)b * T * C + c, C, w, T, t_begin, t_end,
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
static __global__ void __launch_bounds__(256)
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

// One CTA per window: LayerNorm(512) per frame, GroupNorm(1, 512) over the window, + encoder output,
// PReLU -> operand planes of the tail gate GEMM.  `hn` is fp32 scratch.
static __global__ void __launch_bounds__(512)
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

static __global__ void __launch_bounds__(256)
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

// (rows, cols) fp32 -> zero-padded (rows_pad, cols_pad) tf32 hi/lo planes
static __global__ void pad_split_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
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
  // widest tile first: a 256-column tile amortises the A operand twice as well as a 128-column one, so it wins
  // unless its zero padding wastes more than ~8 % of the columns; otherwise least padding
  if (round_up(N, 256) * 100 <= N * 108) return 256;
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
  bool bf16 = false;             // one bf16 plane instead of tf32 hi / lo (w_lo unused)
  CUtensorMap w_hi, w_lo;
};
struct Gemm {
  tc::TcPlan plan;
  tc::TcArgs args;
};

// Host-side state shared by both families: weight blob lookup, tf32 operand planes + tensor maps, workspace
// allocation and GEMM planning.
struct Base : public ModelImpl {
  int device = 0, sms = 148;
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  int planned = 0;
  bool af = false;               // fp32-A GEMMs: activations are stored once as fp32, the GEMM splits its A tile in shared memory

  void free_ws() {
    adn_note_free();
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
  bool make_lin(Lin& l, const float* src, int N, int K, bool bf16 = false) {
    l.N = N; l.K = K; l.batches = 1; l.bf16 = bf16;
    l.bn = choose_bn(N);
    if (bf16 && l.bn == 176) l.bn = 128;
    if (af && !bf16 && l.bn == 176) l.bn = 256;     // the TMA-store epilogue writes 32-column boxes: tiles are multiples of 32
    l.n_pad = round_up(N, l.bn);
    l.k_pad = round_up(K, bf16 ? 64 : 32);
    const long long plane = (long long)l.n_pad * l.k_pad;
    if (cudaMalloc((void**)&l.planes, 2 * plane * sizeof(float)) != cudaSuccess) { err = "out of memory (weights)"; return false; }
    pad_split_kernel<<<(unsigned)((plane + 255) / 256), 256>>>(src, l.planes, bf16 ? nullptr : l.planes + plane, N, K, l.k_pad, plane);
    if (bf16) return tc::make_weight_map(&l.w_hi, l.planes, l.k_pad, l.n_pad, l.bn, err, 1, true);
    return tc::make_weight_map(&l.w_hi, l.planes, l.k_pad, l.n_pad, l.bn, err, 1) &&
           tc::make_weight_map(&l.w_lo, l.planes + plane, l.k_pad, l.n_pad, l.bn, err, 1);
  }
  // activation planes used as the per-window W operand (lo == nullptr: one bf16 plane in `hi`)
  bool make_act_lin(Lin& l, float* hi, float* lo, int N, int n_pad, int K, int k_pad, int bn, int batches) {
    l.N = N; l.K = K; l.n_pad = n_pad; l.k_pad = k_pad; l.bn = bn; l.batches = batches; l.planes = nullptr;
    l.bf16 = lo == nullptr;
    if (l.bf16) return tc::make_weight_map(&l.w_hi, hi, k_pad, n_pad, bn, err, batches, true);
    return tc::make_weight_map(&l.w_hi, hi, k_pad, n_pad, bn, err, batches) &&
           tc::make_weight_map(&l.w_lo, lo, k_pad, n_pad, bn, err, batches);
  }

  bool alloc(float*& p, size_t nfloats, bool zero) {
    if (cudaMalloc((void**)&p, nfloats * sizeof(float)) != cudaSuccess) { err = "out of device memory (workspace)"; return false; }
    allocs.push_back(p);
    ws_bytes += nfloats * sizeof(float);
    if (zero) cudaMemset(p, 0, nfloats * sizeof(float));
    return true;
  }

  // A operand: tf32 hi plane at a_planes, lo plane a_plane_stride floats later; bf16 weights (l.bf16): ONE bf16 plane at
  // a_planes (2-byte elements, same element strides)
  bool plan_gemm(Gemm& g, const float* a_planes, long long a_plane_stride, int K, int rows, long long row_stride,
                 int batches, long long batch_stride, const Lin& l, bool allow_af = true) {
    const int bt = rows >= 128 ? 128 : rows;
    g.plan = tc::TcPlan{};
    g.plan.bn = l.bn;
    g.plan.bf16 = l.bf16;
    g.plan.map_w_hi = l.w_hi;
    g.plan.map_w_lo = l.w_lo;
    g.plan.a_f32 = af && allow_af && !l.bf16;
    if (g.plan.a_f32) {                          // a_planes is the fp32 tensor itself
      if (!tc::make_row_map(&g.plan.map_a_hi, a_planes, K, rows, row_stride, batches, batch_stride, bt, 1, err)) return false;
      g.plan.map_a_lo = g.plan.map_a_hi;
    } else if (l.bf16) {
      if (!tc::make_row_map(&g.plan.map_a_hi, a_planes, K, rows, row_stride, batches, batch_stride, bt, 1, err, true)) return false;
    } else if (!tc::make_row_map(&g.plan.map_a_hi, a_planes, K, rows, row_stride, batches, batch_stride, bt, 1, err) ||
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
  // fp32-A plans write their output through a TMA store: call once the caller has set args.C / Chi / ldc / N.  An operand-plane
  // output (Chi / Clo) becomes the fp32 output.
  bool finish_af(Gemm& g) {
    if (!g.plan.a_f32) return true;
    tc::TcArgs& a = g.args;
    if (!a.C && a.Chi) { a.C = a.Chi; }
    a.Chi = a.Clo = nullptr;
    if (!a.C) { err = "fp32-A GEMM without an fp32 output"; return false; }
    return tc::make_store_map(&g.plan.map_c, a.C, a.N, a.TM, a.ldc, a.B, (long long)a.TM * a.ldc, err);
  }
};

}  // namespace mf2
