// Mel-Band-Roformer (stereo) on sm_100a.
// Reference: Mel_Band_Roformer/Stereo/Export_MelBandRoformer.py  forward :629-680, _core :585-627,
// _attention :545-563, _transformer :568-571, _band_split :573-576, _mask_estimator :578-583.
//
// Token layout: x is (n_bands, B*T, dim) row-major ("band-major"): row = band*Mf + (b*T + t).  The
// reference permutes between (bands*B, T) and (T*B, bands) views for the time / frequency
// transformers; here nothing is permuted -- the Linear layers are per-token and the attention
// kernel walks the sequence axis with a (base, stride) pair.
//
// Every dense contraction (band-split, qkv+gates, attention out, feed-forward, mask-estimator
// MLPs: > 95 % of the flops) runs on the tcgen05 3xTF32 GEMM (gemm_tc.cu); producers emit tf32
// hi/lo planes directly from their epilogues.  RMS normalisation is applied as a per-row scale in
// the consuming GEMM's epilogue: (x/|x|) W == (x W)/|x|.
#include "adn.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "gtcrn.cuh"     // launch_prep, split_tf32_store
#include "model_impl.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mbr {

constexpr int NFFT = 2048, HOP = 441, FB = 1025, CH = 2, FC = FB * CH;
constexpr int LD = 2056;          // frame stride of the packed spectrum (2050 used)
constexpr int DHEAD = 64;
constexpr float EPS = 1e-12f;

// ---------------------------------------------------------------------------------
// gather: spectrum (B*2, T, LD) -> band-split input planes (B*T, SD) + per-(band, token) 1/|x_band|
// (:592-596 index_select on the (freq,chan)-interleaved axis; _normalize of every band :533-538)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_kernel(const float* __restrict__ spec, const int* __restrict__ freq_idx, const int* __restrict__ band_off,
              float* __restrict__ xg_hi, float* __restrict__ xg_lo, float* __restrict__ rs, int T, int Mf, int SD,
              int nb) {
  extern __shared__ float vals[];
  const int tk = blockIdx.x, b = tk / T, t = tk - b * T;
  for (int j = threadIdx.x; j < SD; j += 256) {
    const int s = j >> 1, ri = j & 1;
    const int fi = __ldg(freq_idx + s);
    const int f = fi >> 1, c = fi & 1;
    const float v = __ldg(spec + ((long long)(b * CH + c) * T + t) * LD + ri * FB + f);
    vals[j] = v;
    gtcrn::split_tf32_store(v, xg_hi, xg_lo, (long long)tk * SD + j);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < nb; i += 8) {
    const int lo = __ldg(band_off + i), hi = __ldg(band_off + i + 1);
    float ss = 0.f;
    for (int j = lo + lane; j < hi; j += 32) ss = fmaf(vals[j], vals[j], ss);
    ss = warp_sum(ss);
    if (lane == 0) rs[(long long)i * Mf + tk] = 1.0f / fmaxf(sqrtf(ss), EPS);
  }
}

// rn[m] = 1 / max(|x_m|, eps); warp per row
__global__ void __launch_bounds__(256)
rownorm_kernel(const float* __restrict__ x, float* __restrict__ rn, long long M, int D) {
  const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + row * D;
  float ss = 0.f;
  for (int j = lane; j < D; j += 32) { const float v = __ldg(xr + j); ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) rn[row] = 1.0f / fmaxf(sqrtf(ss), EPS);
}

// x <- x / max(|x|, eps) * g  (transformer output norm :571), + tf32 planes + 1/|x_new| for the next layer
__global__ void __launch_bounds__(256)
renorm_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ hi, float* __restrict__ lo,
              float* __restrict__ rn, long long M, int D) {
  const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float* xr = x + row * D;
  float ss = 0.f;
  for (int j = lane; j < D; j += 32) { const float v = xr[j]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), EPS);
  float s2 = 0.f;
  for (int j = lane; j < D; j += 32) {
    const float y = xr[j] * inv * __ldg(g + j);     // == (x / norm) * g up to one rounding
    xr[j] = y;
    gtcrn::split_tf32_store(y, hi, lo, row * D + j);
    s2 = fmaf(y, y, s2);
  }
  s2 = warp_sum(s2);
  if (lane == 0) rn[row] = 1.0f / fmaxf(sqrtf(s2), EPS);
}

// ---------------------------------------------------------------------------------
// attention (:545-563): rotary on q,k (pair swap, sign folded in the sin table), softmax(q k^T) v,
// sigmoid head gates.  One CTA per (sequence, head); keys/values of the sequence live in shared
// memory; one warp per query row.  Sequence element s of sequence q is token row base + s*stride.
//   time : base = q*n,  stride = 1   (q = band*B + b,  n = T)
//   freq : base = q,    stride = Mf  (q = b*T + t,     n = bands)
// ---------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_kernel(const float* __restrict__ qkvg, int ldq, const float* __restrict__ rcos,
                 const float* __restrict__ rsin, float* __restrict__ ao_hi, float* __restrict__ ao_lo, int n,
                 long long stride, int freq_mode, int heads) {
  extern __shared__ float sm[];
  const int n4 = (n + 3) & ~3;             // 16-byte aligned region sizes
  float* Ks = sm;                          // [n][65]
  float* Vs = Ks + (((size_t)n * 65 + 3) & ~(size_t)3);   // [n][64]
  float* Ps = Vs + (size_t)n * 64;         // [warps][n4]
  float* Qs = Ps + (size_t)ATT_WARPS * n4; // [warps][64]
  const int q = blockIdx.x, head = blockIdx.y;
  const int di = heads * DHEAD;
  const long long base = freq_mode ? (long long)q : (long long)q * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < n * 32; i += ATT_WARPS * 32) {
    const int s = i >> 5, p = i & 31;      // p = rotary pair index
    const float* row = qkvg + (base + (long long)s * stride) * ldq + head * DHEAD;
    const float2 kk = *reinterpret_cast<const float2*>(row + di + 2 * p);
    const float2 vv = *reinterpret_cast<const float2*>(row + 2 * di + 2 * p);
    const float2 c = *reinterpret_cast<const float2*>(rcos + s * DHEAD + 2 * p);
    const float2 sn = *reinterpret_cast<const float2*>(rsin + s * DHEAD + 2 * p);
    Ks[s * 65 + 2 * p] = kk.x * c.x + kk.y * sn.x;         // k*cos + rotate_half(k)*sin
    Ks[s * 65 + 2 * p + 1] = kk.y * c.y + kk.x * sn.y;
    Vs[s * 64 + 2 * p] = vv.x;
    Vs[s * 64 + 2 * p + 1] = vv.y;
  }
  __syncthreads();

  float* ps = Ps + (size_t)warp * n4;
  float* qs = Qs + warp * 64;
  for (int i = warp; i < n; i += ATT_WARPS) {
    const long long row = base + (long long)i * stride;
    const float* rp = qkvg + row * ldq;
    {
      const float2 qq = *reinterpret_cast<const float2*>(rp + head * DHEAD + 2 * lane);
      const float2 c = *reinterpret_cast<const float2*>(rcos + i * DHEAD + 2 * lane);
      const float2 sn = *reinterpret_cast<const float2*>(rsin + i * DHEAD + 2 * lane);
      qs[2 * lane] = qq.x * c.x + qq.y * sn.x;
      qs[2 * lane + 1] = qq.y * c.y + qq.x * sn.y;
    }
    __syncwarp();
    float qr[64];
#pragma unroll
    for (int d4 = 0; d4 < 16; ++d4) {
      const float4 v4 = *reinterpret_cast<const float4*>(qs + 4 * d4);
      qr[4 * d4] = v4.x; qr[4 * d4 + 1] = v4.y; qr[4 * d4 + 2] = v4.z; qr[4 * d4 + 3] = v4.w;
    }
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) {
      const float* kr = Ks + j * 65;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) a = fmaf(qr[d], kr[d], a);
      ps[j] = a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = expf(ps[j] - mx);
      ps[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < n; ++j) {
      const float p = ps[j];
      const float2 vv = *reinterpret_cast<const float2*>(Vs + j * 64 + 2 * lane);
      o0 = fmaf(p, vv.x, o0);
      o1 = fmaf(p, vv.y, o1);
    }
    const float gate = adn_sigmoid(__ldg(rp + 3 * di + head));
    const float sc = gate / sum;
    const long long o = row * di + head * DHEAD + 2 * lane;
    gtcrn::split_tf32_store(o0 * sc, ao_hi, ao_lo, o);
    gtcrn::split_tf32_store(o1 * sc, ao_hi, ao_lo, o + 1);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------
// attention2: same math as attention_kernel for sequences that fit in shared memory (n <= 160)
// organised as two register-tiled shared-memory GEMMs (4x4 outputs per thread, 16 FMA per
// 2 (scores) / 5 (PV) shared loads) around an in-place row softmax.
//   smem: Qt[64][np] | Kt[64][np] | V[n][64] | S[n][np+1] | rinv[n]      (np = n rounded up to 4)
// ---------------------------------------------------------------------------------
constexpr int ATT2_THREADS = 1024;

// helpers of the optional linear resampling either side of the model (Export_MelBandRoformer.py:631-644, :662-675)
__global__ void rs_scale_kernel(float* __restrict__ x, long long n, float sc) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[i] *= sc;
}
__global__ void rs_convert_kernel(const float* __restrict__ src, void* __restrict__ out, int out_dtype, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  if (out_dtype == ADN_I16) reinterpret_cast<int16_t*>(out)[i] = (int16_t)(int)fminf(fmaxf(v, -32768.0f), 32767.0f);
  else if (out_dtype == ADN_F32) reinterpret_cast<float*>(out)[i] = v;
  else reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
}

__host__ __device__ inline size_t att2_smem_floats(int n) {
  const int np = (n + 3) & ~3;
  size_t s_region = (size_t)n * (np + 1);
  const size_t tmp = (size_t)2 * n * 65;            // staging of rotated q,k aliases the S region
  if (tmp > s_region) s_region = tmp;
  return (size_t)2 * 64 * np + (size_t)n * 64 + s_region + (size_t)((n + 3) & ~3) + 4;
}

__global__ void __launch_bounds__(1024)
attention2_kernel(const float* __restrict__ qkvg, int ldq, const float* __restrict__ rcos,
                  const float* __restrict__ rsin, float* __restrict__ ao_hi, float* __restrict__ ao_lo, int n,
                  long long stride, int freq_mode, int heads) {
  extern __shared__ __align__(16) float sm2[];
  const int np = (n + 3) & ~3, ns = np + 1;
  float* Qt = sm2;                               // [64][np]
  float* Kt = Qt + 64 * np;                      // [64][np]
  float* Vs = Kt + 64 * np;                      // [n][64]
  float* S = Vs + (size_t)n * 64;                // [n][ns]   (first used as tq[n][65] | tk[n][65])
  size_t s_region = (size_t)n * ns;
  if ((size_t)2 * n * 65 > s_region) s_region = (size_t)2 * n * 65;
  float* rinv = S + ((s_region + 3) & ~(size_t)3);   // [n]  gate / softmax sum
  float* tq = S;
  float* tk = S + (size_t)n * 65;
  const int q = blockIdx.x, head = blockIdx.y, tid = threadIdx.x;
  const int NT = blockDim.x;
  const int di = heads * DHEAD;
  const long long base = freq_mode ? (long long)q : (long long)q * n;

  // phase 0a: coalesced row loads, rotary on q,k
  for (int i = tid; i < n * 32; i += NT) {
    const int s = i >> 5, p = i & 31;
    const float* row = qkvg + (base + (long long)s * stride) * ldq + head * DHEAD;
    const float2 qq = *reinterpret_cast<const float2*>(row + 2 * p);
    const float2 kk = *reinterpret_cast<const float2*>(row + di + 2 * p);
    const float2 vv = *reinterpret_cast<const float2*>(row + 2 * di + 2 * p);
    const float2 c = *reinterpret_cast<const float2*>(rcos + s * DHEAD + 2 * p);
    const float2 sn = *reinterpret_cast<const float2*>(rsin + s * DHEAD + 2 * p);
    tq[s * 65 + 2 * p] = qq.x * c.x + qq.y * sn.x;
    tq[s * 65 + 2 * p + 1] = qq.y * c.y + qq.x * sn.y;
    tk[s * 65 + 2 * p] = kk.x * c.x + kk.y * sn.x;
    tk[s * 65 + 2 * p + 1] = kk.y * c.y + kk.x * sn.y;
    *reinterpret_cast<float2*>(Vs + s * 64 + 2 * p) = vv;
  }
  __syncthreads();
  // phase 0b: transpose to [d][s] (zero the padded columns)
  for (int i = tid; i < 64 * np; i += NT) {
    const int d = i / np, s = i - d * np;
    Qt[i] = s < n ? tq[s * 65 + d] : 0.f;
    Kt[i] = s < n ? tk[s * 65 + d] : 0.f;
  }
  __syncthreads();

  // phase 1: S = Q K^T
  const int nt = np >> 2;
  for (int tile = tid; tile < nt * nt; tile += NT) {
    const int i0 = (tile / nt) * 4, j0 = (tile - (tile / nt) * nt) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float4 qa = *reinterpret_cast<const float4*>(Qt + k * np + i0);
      const float4 kb = *reinterpret_cast<const float4*>(Kt + k * np + j0);
      const float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(qv[a], kv[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (i0 + a < n && j0 + c < n) S[(i0 + a) * ns + j0 + c] = acc[a][c];
  }
  __syncthreads();

  // phase 2: row softmax (unnormalised exp in place; gate / sum kept per row)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < n; i += NT / 32) {
      float* sr = S + i * ns;
      float mx = -INFINITY;
      for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int j = lane; j < n; j += 32) {
        const float e = expf(sr[j] - mx);
        sr[j] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        const long long row = base + (long long)i * stride;
        rinv[i] = adn_sigmoid(__ldg(qkvg + row * ldq + 3 * di + head)) / sum;
      }
    }
  }
  __syncthreads();

  // phase 3: O = P V, scaled by gate/sum, written as tf32 planes
  for (int tile = tid; tile < nt * 16; tile += NT) {
    const int i0 = (tile >> 4) * 4, d0 = (tile & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    const float* s0 = S + (size_t)i0 * ns;
    const int r1 = (i0 + 1 < n) ? 1 : 0, r2 = (i0 + 2 < n) ? 2 : 0, r3 = (i0 + 3 < n) ? 3 : 0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const float4 v4 = *reinterpret_cast<const float4*>(Vs + j * 64 + d0);
      const float pv[4] = {s0[j], s0[r1 * ns + j], s0[r2 * ns + j], s0[r3 * ns + j]};
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(pv[a], vv[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + a;
      if (i >= n) continue;
      const float sc = rinv[i];
      const long long o = (base + (long long)i * stride) * di + head * DHEAD + d0;
      float h[4], l[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float x = acc[a][c] * sc;
        h[c] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        l[c] = x - h[c];
      }
      *reinterpret_cast<float4*>(ao_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(ao_lo + o) = make_float4(l[0], l[1], l[2], l[3]);
    }
  }
}

// ---------------------------------------------------------------------------------
// attention3: attention2 with both products on the warp-level tensor cores -- mma.sync m16n8k8 (tf32, fp32 accumulate) with
// the 3xTF32 split done in registers on fp32 fragments (hi = low 13 mantissa bits cleared, lo = x - hi; hi*hi + hi*lo + lo*hi).
//   smem: Qt[64][np] | Kt[64][np] | V[np][72] | S[n][np+1] | rinv[n]      np = n rounded up to 8 (mod 16)
// np = 8 or 24 (mod 32) makes the Q^T / K^T fragment loads conflict-free (bank = np * (lane & 3) + (lane >> 2)), the 72-float V
// rows do the same for the value fragments; key columns n .. np-1 of the softmax rows are zeroed so the padded K steps add nothing.
// A warp task is (16 query rows) x (16 keys) in the score product and (16 query rows) x (8 value columns) in the value product.
// ---------------------------------------------------------------------------------
__host__ __device__ inline int att3_np(int n) { return ((n + 7) / 16) * 16 + 8; }
__host__ __device__ inline size_t att3_smem_floats(int n) {
  const int np = att3_np(n);
  size_t s_region = (size_t)n * (np + 1);
  const size_t tmp = (size_t)2 * n * 65;
  if (tmp > s_region) s_region = tmp;
  return (size_t)2 * 64 * np + (size_t)np * 72 + s_region + (size_t)((n + 3) & ~3) + 8;
}
__device__ __forceinline__ void mbr_mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split3(float x, uint32_t& h, uint32_t& l) {
  h = __float_as_uint(x) & 0xFFFFE000u;
  l = __float_as_uint(x - __uint_as_float(h));
}

__global__ void __launch_bounds__(1024)
attention3_kernel(const float* __restrict__ qkvg, int ldq, const float* __restrict__ rcos,
                  const float* __restrict__ rsin, float* __restrict__ ao_hi, float* __restrict__ ao_lo, int n,
                  long long stride, int freq_mode, int heads) {
  extern __shared__ __align__(16) float sm3[];
  const int np = att3_np(n), ns = np + 1;
  float* Qt = sm3;                               // [64][np]
  float* Kt = Qt + 64 * np;                      // [64][np]
  float* Vs = Kt + 64 * np;                      // [np][72]
  float* S = Vs + (size_t)np * 72;               // [n][ns]   (first used as tq[n][65] | tk[n][65])
  size_t s_region = (size_t)n * ns;
  if ((size_t)2 * n * 65 > s_region) s_region = (size_t)2 * n * 65;
  float* rinv = S + ((s_region + 3) & ~(size_t)3);   // [n]  gate / softmax sum
  float* tq = S;
  float* tk = S + (size_t)n * 65;
  const int q = blockIdx.x, head = blockIdx.y, tid = threadIdx.x;
  const int NT = blockDim.x, warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
  const int gq = lane >> 2, tq4 = lane & 3;
  const int di = heads * DHEAD;
  const long long base = freq_mode ? (long long)q : (long long)q * n;

  // phase 0a: coalesced row loads, rotary on q,k
  for (int i = tid; i < np * 32; i += NT) {
    const int s = i >> 5, p = i & 31;
    if (s >= n) { *reinterpret_cast<float2*>(Vs + s * 72 + 2 * p) = make_float2(0.f, 0.f); continue; }
    const float* row = qkvg + (base + (long long)s * stride) * ldq + head * DHEAD;
    const float2 qq = *reinterpret_cast<const float2*>(row + 2 * p);
    const float2 kk = *reinterpret_cast<const float2*>(row + di + 2 * p);
    const float2 vv = *reinterpret_cast<const float2*>(row + 2 * di + 2 * p);
    const float2 c = *reinterpret_cast<const float2*>(rcos + s * DHEAD + 2 * p);
    const float2 sn = *reinterpret_cast<const float2*>(rsin + s * DHEAD + 2 * p);
    tq[s * 65 + 2 * p] = qq.x * c.x + qq.y * sn.x;
    tq[s * 65 + 2 * p + 1] = qq.y * c.y + qq.x * sn.y;
    tk[s * 65 + 2 * p] = kk.x * c.x + kk.y * sn.x;
    tk[s * 65 + 2 * p + 1] = kk.y * c.y + kk.x * sn.y;
    *reinterpret_cast<float2*>(Vs + s * 72 + 2 * p) = vv;
  }
  __syncthreads();
  // phase 0b: transpose to [d][s] (zero the padded columns)
  for (int i = tid; i < 64 * np; i += NT) {
    const int d = i / np, s = i - d * np;
    Qt[i] = s < n ? tq[s * 65 + d] : 0.f;
    Kt[i] = s < n ? tk[s * 65 + d] : 0.f;
  }
  __syncthreads();

  // phase 1: S = Q K^T
  const int MT = (n + 15) >> 4, NG = (np + 15) >> 4;
  for (int task = warp; task < MT * NG; task += NW) {
    const int i0 = (task / NG) * 16, j0 = (task % NG) * 16;
    const bool two = j0 + 8 < np;
    float acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll
    for (int k8 = 0; k8 < 64; k8 += 8) {
      const float* qa = Qt + (k8 + tq4) * np + i0 + gq;
      uint32_t ah[4], al[4];
      split3(qa[0], ah[0], al[0]); split3(qa[8], ah[1], al[1]);
      split3(qa[4 * np], ah[2], al[2]); split3(qa[4 * np + 8], ah[3], al[3]);
      const float* kb = Kt + (k8 + tq4) * np + j0 + gq;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        uint32_t bh0, bl0, bh1, bl1;
        split3(kb[8 * h], bh0, bl0); split3(kb[4 * np + 8 * h], bh1, bl1);
        mbr_mma_tf32(acc[h], al, bh0, bh1);
        mbr_mma_tf32(acc[h], ah, bl0, bl1);
        mbr_mma_tf32(acc[h], ah, bh0, bh1);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + gq + (r >> 1) * 8, j = j0 + 8 * h + 2 * tq4 + (r & 1);
        if (i < n && j < n) S[i * ns + j] = acc[h][r];
      }
  }
  __syncthreads();

  // phase 2: row softmax (unnormalised exp in place; gate / sum kept per row; padded key columns zeroed)
  for (int i = warp; i < n; i += NW) {
    float* sr = S + i * ns;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < np; j += 32) {
      const float e = j < n ? expf(sr[j] - mx) : 0.f;
      sr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      const long long row = base + (long long)i * stride;
      rinv[i] = adn_sigmoid(__ldg(qkvg + row * ldq + 3 * di + head)) / sum;
    }
  }
  __syncthreads();

  // phase 3: O = P V, scaled by gate / sum, written as tf32 planes
  for (int task = warp; task < MT * 8; task += NW) {
    const int i0 = (task >> 3) * 16, d0 = (task & 7) * 8;
    const int ra = min(i0 + gq, n - 1), rb = min(i0 + gq + 8, n - 1);
    const float* pa = S + (size_t)ra * ns + tq4;
    const float* pb = S + (size_t)rb * ns + tq4;
    const float* vb = Vs + tq4 * 72 + d0 + gq;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j8 = 0; j8 < np; j8 += 8) {
      uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
      split3(pa[j8], ah[0], al[0]); split3(pb[j8], ah[1], al[1]);
      split3(pa[j8 + 4], ah[2], al[2]); split3(pb[j8 + 4], ah[3], al[3]);
      split3(vb[j8 * 72], bh0, bl0); split3(vb[(j8 + 4) * 72], bh1, bl1);
      mbr_mma_tf32(acc, al, bh0, bh1);
      mbr_mma_tf32(acc, ah, bl0, bl1);
      mbr_mma_tf32(acc, ah, bh0, bh1);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int i = i0 + gq + hr * 8;
      if (i >= n) continue;
      const float sc = rinv[i];
      const long long o = (base + (long long)i * stride) * di + head * DHEAD + d0 + 2 * tq4;
      const float x0 = acc[2 * hr] * sc, x1 = acc[2 * hr + 1] * sc;
      const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
      *reinterpret_cast<float2*>(ao_hi + o) = make_float2(h0, h1);
      *reinterpret_cast<float2*>(ao_lo + o) = make_float2(x0 - h0, x1 - h1);
    }
  }
}

// ---------------------------------------------------------------------------------
// mask_apply: GLU of the mask-estimator output, scatter-add as a gather over the (<= 2)
// contributing bands of every (freq,chan) row (:615-618, averaging pre-folded into the weights),
// complex mask on the spectrum (:620-623), de-interleave into the per-channel ISTFT input.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_apply_kernel(const float* __restrict__ me3, const int* __restrict__ dst_src, const int* __restrict__ src_a,
                  const int* __restrict__ src_g, const float* __restrict__ spec, float* __restrict__ enh, int T,
                  int pad, long long me_ld) {
  const int tk = blockIdx.x, b = tk / T, t = tk - b * T;
  const float* mr = me3 + (long long)tk * me_ld;
  for (int fc = threadIdx.x; fc < FC; fc += 256) {
    float m_re = 0.f, m_im = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int s = __ldg(dst_src + fc * 2 + k);
      if (s >= 0) {
        const int ca = __ldg(src_a + s), cg = __ldg(src_g + s);
        m_re += __ldg(mr + ca) * adn_sigmoid(__ldg(mr + cg));          // F.glu: a * sigmoid(b)
        m_im += __ldg(mr + ca + 1) * adn_sigmoid(__ldg(mr + cg + 1));
      }
    }
    const int f = fc >> 1, c = fc & 1;
    const long long sp = ((long long)(b * CH + c) * T + t) * LD;
    const float re = __ldg(spec + sp + f), im = __ldg(spec + sp + FB + f);
    const long long o = ((long long)(b * CH + c) * (T + 2 * pad) + pad + t) * LD;
    enh[o + f] = re * m_re - im * m_im;
    enh[o + FB + f] = re * m_im + im * m_re;
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
  gtcrn::split_tf32_store(v, hi, lo, i);
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

int choose_bn(int N) {
  const int cands[3] = {256, 176, 128};
  int best = 128, best_pad = 1 << 30;
  for (int c : cands) {
    int pad = (N + c - 1) / c * c;
    if (pad < best_pad) { best_pad = pad; best = c; }
  }
  return best;
}

// One Linear layer's weights as tf32 planes + W tensor maps.
struct Lin {
  int N = 0, K = 0, n_pad = 0, k_pad = 0, bn = 0, batches = 1;
  float* planes = nullptr;       // hi | lo
  CUtensorMap w_hi, w_lo;
};

struct Gemm {                    // a fully planned GEMM launch
  tc::TcPlan plan;
  tc::TcArgs args;
};

class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int W = 0, T = 0, Lp = 0, pad = 4, R = 5;
  int W_in = 0, W_final = 0, in_sr = 44100, out_sr = 44100;   // I/O window at the in / out sample rates (== W at 44.1 kHz)
  bool rs_in = false, rs_out = false;
  double in_scale = 1.0, out_scale = 1.0;
  float *xr = nullptr, *yres = nullptr, *yout = nullptr;
  int D = 384, depth = 6, heads = 8, nb = 60, DI = 512, DQ = 1544, DHID = 1536, NSEL = 0, SD = 0;
  std::vector<int> din, off;     // per band input width / offset into SD
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;

  // device constant tables
  int *d_freq_idx = nullptr, *d_band_off = nullptr, *d_dst_src = nullptr, *d_src_a = nullptr, *d_src_g = nullptr;
  const float *d_fwd = nullptr, *d_norm = nullptr;
  float* d_ola = nullptr;
  const float *tcos = nullptr, *tsin = nullptr, *fcos = nullptr, *fsin = nullptr;

  std::vector<Lin> bs, me3;
  struct Layer { Lin in, out, ff1, ff2; const float *in_b, *ff1_b, *ff2_b, *out_g; };
  std::vector<Layer> layers;
  Lin me1, me2;
  std::vector<const float*> bs_b, me3_b;
  const float *me_b1 = nullptr, *me_b2 = nullptr;

  // workspace (sized for `planned` windows)
  int planned = 0;
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  float *xp = nullptr, *spec = nullptr, *xg = nullptr, *rs_bs = nullptr, *x = nullptr, *xpl = nullptr, *rn = nullptr;
  float *qkvg = nullptr, *ao = nullptr, *hpl = nullptr, *g1 = nullptr, *me3out = nullptr, *enh = nullptr;
  std::vector<Gemm> g_bs, g_me3;
  struct LayerG { Gemm in, out, ff1, ff2; };
  std::vector<LayerG> g_layers;
  Gemm g_me1, g_me2;
  int stop_after = 0, last_batch = 0;

  ~Model() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    auto fl = [](Lin& l) { if (l.planes) cudaFree(l.planes); };
    for (auto& l : bs) fl(l);
    for (auto& l : me3) fl(l);
    for (auto& L : layers) { fl(L.in); fl(L.out); fl(L.ff1); fl(L.ff2); }
    fl(me1); fl(me2);
    cudaFree(d_freq_idx); cudaFree(d_band_off); cudaFree(d_dst_src); cudaFree(d_src_a); cudaFree(d_src_g);
    cudaFree(d_ola);
  }

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

  bool make_lin(Lin& l, const float* src, int N, int K, int batches = 1) {
    l.N = N; l.K = K; l.batches = batches;
    l.bn = choose_bn(N);
    l.n_pad = round_up(N, l.bn);
    l.k_pad = round_up(K, 32);
    const long long plane = (long long)batches * l.n_pad * l.k_pad;
    if (cudaMalloc((void**)&l.planes, 2 * plane * sizeof(float)) != cudaSuccess) { err = "out of memory (weights)"; return false; }
    for (int b = 0; b < batches; ++b) {
      const long long tot = (long long)l.n_pad * l.k_pad;
      pad_split_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(src + (long long)b * N * K, l.planes + b * tot,
                                                               l.planes + plane + b * tot, N, K, l.k_pad, tot);
    }
    return tc::make_weight_map(&l.w_hi, l.planes, l.k_pad, l.n_pad, l.bn, err, batches) &&
           tc::make_weight_map(&l.w_lo, l.planes + plane, l.k_pad, l.n_pad, l.bn, err, batches);
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
    int nfft = 0, hop = 0, dh = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", W) || !geti("nfft", nfft) || !geti("hop_length", hop) || !geti("mbr_dim", D) ||
        !geti("mbr_depth", depth) || !geti("mbr_heads", heads) || !geti("mbr_dim_head", dh) ||
        !geti("mbr_num_bands", nb) || !gets("input_audio_dtype", sin) || !gets("output_audio_dtype", sout))
      return false;
    {
      auto opt = [&](const char* k, int& v) { auto it = meta.find(k); if (it != meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
      int model_sr = 44100;
      opt("in_sample_rate", in_sr); opt("out_sample_rate", out_sr); opt("model_sample_rate", model_sr);
      if (model_sr != 44100 || in_sr <= 0 || out_sr <= 0) { err = "mel_band_roformer runs at model_sample_rate 44100"; return false; }
    }
    W_in = W;
    rs_in = in_sr != 44100;
    rs_out = out_sr != 44100;
    in_scale = 44100.0 / (double)in_sr;            // INPUT_TO_MODEL_SCALE (:56)
    out_scale = (double)out_sr / 44100.0;          // MODEL_TO_OUTPUT_SCALE (:57)
    if (rs_in) W = (int)floor((double)W_in * in_scale);            // F.interpolate(scale_factor=...) output size
    W_final = rs_out ? (int)floor((double)W * out_scale) : W;
    if (nfft != NFFT || hop != HOP || dh != DHEAD || D % 32 || W % hop) {
      err = "mel_band_roformer needs nfft=2048, hop=441, dim_head=64, dim % 32 == 0, model-rate length % hop == 0";
      return false;
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = W / HOP + 1;
    Lp = round_up(W + NFFT, 4);
    DI = heads * DHEAD; DQ = 3 * DI + heads; DHID = 4 * D;
    if (T > 256) { err = "mel_band_roformer: windows longer than 255 hops are not supported yet (fold into 1.5 s windows)"; return false; }

    bool ok = true;
    // band layout tables
    auto itf = index.find("freq_indices"), itd = index.find("band_din");
    if (itf == index.end() || itd == index.end() || (int)itd->second.count != nb) { err = "missing freq_indices / band_din"; return false; }
    NSEL = (int)itf->second.count;
    std::vector<int> fidx(NSEL);
    for (int i = 0; i < NSEL; ++i) fidx[i] = (int)h_blob[itf->second.offset + i];
    din.resize(nb); off.resize(nb + 1);
    off[0] = 0;
    for (int i = 0; i < nb; ++i) { din[i] = (int)h_blob[itd->second.offset + i]; off[i + 1] = off[i] + din[i]; }
    SD = off[nb];
    if (SD != 2 * NSEL) { err = "band widths do not add up to the selected bins"; return false; }
    std::vector<int> dst_src(FC * 2, -1), src_a(NSEL), src_g(NSEL);
    {
      int band = 0;
      for (int s = 0; s < NSEL; ++s) {
        while (2 * s >= off[band + 1]) ++band;
        const int ls2 = 2 * s - off[band];
        src_a[s] = 2 * off[band] + ls2;
        src_g[s] = src_a[s] + din[band];
        const int fc = fidx[s];
        if (fc < 0 || fc >= FC) { err = "freq index out of range"; return false; }
        if (dst_src[fc * 2] < 0) dst_src[fc * 2] = s;
        else if (dst_src[fc * 2 + 1] < 0) dst_src[fc * 2 + 1] = s;
        else { err = "a frequency row belongs to more than two bands"; return false; }
      }
    }
    auto up = [&](int*& d, const std::vector<int>& h) {
      return cudaMalloc((void**)&d, h.size() * sizeof(int)) == cudaSuccess &&
             cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!up(d_freq_idx, fidx) || !up(d_band_off, off) || !up(d_dst_src, dst_src) || !up(d_src_a, src_a) || !up(d_src_g, src_g)) {
      err = "table upload failed";
      return false;
    }

    d_fwd = dptr("stft.fwd", (size_t)2050 * NFFT, ok);
    d_norm = dptr("istft.norm", (size_t)W, ok);
    tcos = dptr("rope.tcos", (size_t)T * DHEAD, ok);
    tsin = dptr("rope.tsin", (size_t)T * DHEAD, ok);
    fcos = dptr("rope.fcos", (size_t)nb * DHEAD, ok);
    fsin = dptr("rope.fsin", (size_t)nb * DHEAD, ok);
    if (!ok) return false;
    auto inv = index.find("istft.inv");
    if (inv == index.end() || inv->second.count != (size_t)2050 * NFFT) { err = "missing istft.inv"; return false; }
    {   // overlap-add weight for the FFMA ISTFT (hop 441 is not TMA-addressable: 1764 B rows)
      R = (NFFT + HOP - 1) / HOP; pad = R - 1;
      std::vector<float> w((size_t)HOP * R * LD, 0.f);
      const float* ib = h_blob + inv->second.offset;
      for (int n = 0; n < HOP; ++n)
        for (int q = 0; q < R; ++q) {
          int src = n + (R - 1 - q) * HOP;
          if (src >= NFFT) continue;
          for (int r = 0; r < 2050; ++r) w[(size_t)n * R * LD + (size_t)q * LD + r] = ib[(size_t)r * NFFT + src];
        }
      if (cudaMalloc((void**)&d_ola, w.size() * 4) != cudaSuccess ||
          cudaMemcpy(d_ola, w.data(), w.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { err = "ola upload failed"; return false; }
    }

    // weights -> tf32 planes
    bs.resize(nb); me3.resize(nb); bs_b.resize(nb); me3_b.resize(nb);
    for (int i = 0; i < nb && ok; ++i) {
      const float* w = dptr("bs_w." + std::to_string(i), (size_t)D * din[i], ok);
      bs_b[i] = dptr("bs_b." + std::to_string(i), D, ok);
      if (ok && !make_lin(bs[i], w, D, din[i])) return false;
      const float* w3 = dptr("me_w3." + std::to_string(i), (size_t)2 * din[i] * DHID, ok);
      me3_b[i] = dptr("me_b3." + std::to_string(i), (size_t)2 * din[i], ok);
      if (ok && !make_lin(me3[i], w3, 2 * din[i], DHID)) return false;
    }
    layers.resize(2 * depth);
    for (int l = 0; l < 2 * depth && ok; ++l) {
      const std::string p = "tf." + std::to_string(l);
      Layer& L = layers[l];
      const float* w;
      w = dptr(p + ".in_w", (size_t)DQ * D, ok);   if (ok && !make_lin(L.in, w, DQ, D)) return false;
      w = dptr(p + ".out_w", (size_t)D * DI, ok);  if (ok && !make_lin(L.out, w, D, DI)) return false;
      w = dptr(p + ".ff1_w", (size_t)DHID * D, ok); if (ok && !make_lin(L.ff1, w, DHID, D)) return false;
      w = dptr(p + ".ff2_w", (size_t)D * DHID, ok); if (ok && !make_lin(L.ff2, w, D, DHID)) return false;
      L.in_b = dptr(p + ".in_b", DQ, ok);
      L.ff1_b = dptr(p + ".ff1_b", DHID, ok);
      L.ff2_b = dptr(p + ".ff2_b", D, ok);
      L.out_g = dptr(p + ".out_g", D, ok);
    }
    if (ok) {
      const float* w1 = dptr("me_w1", (size_t)nb * DHID * D, ok);
      const float* w2 = dptr("me_w2", (size_t)nb * DHID * DHID, ok);
      me_b1 = dptr("me_b1", (size_t)nb * DHID, ok);
      me_b2 = dptr("me_b2", (size_t)nb * DHID, ok);
      if (ok && (!make_lin(me1, w1, DHID, D, nb) || !make_lin(me2, w2, DHID, DHID, nb))) return false;
    }
    if (!ok) return false;
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "weight split failed"; return false; }
    return true;
  }

  // ---- workspace + launch plans for `B` windows (dense layouts; rebuilt when B changes)
  bool alloc(float*& p, size_t nfloats, bool zero) {
    if (cudaMalloc((void**)&p, nfloats * sizeof(float)) != cudaSuccess) { err = "out of device memory (workspace)"; return false; }
    allocs.push_back(p);
    ws_bytes += nfloats * sizeof(float);
    if (zero) cudaMemset(p, 0, nfloats * sizeof(float));
    return true;
  }

  bool plan_gemm(Gemm& g, const float* a_planes, long long a_plane_stride, int K, int rows, long long row_stride,
                 int batches, long long batch_stride, const Lin& l) {
    const int bt = rows >= 128 ? 128 : rows;     // rows < 128: one partial tile per batch
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

  bool ensure(int B) {
    if (B == planned) return true;
    cudaDeviceSynchronize();
    free_ws();
    const long long Mf = (long long)B * T, M = (long long)nb * Mf;
    const int B2 = B * CH;
    if ((rs_in && !alloc(xr, (size_t)B2 * W, false)) ||
        (rs_out && (!alloc(yres, (size_t)B2 * W, false) || !alloc(yout, (size_t)B2 * W_final, false))))
      return false;
    if (!alloc(xp, (size_t)B2 * Lp, false) || !alloc(spec, (size_t)B2 * T * LD, true) ||
        !alloc(xg, (size_t)2 * Mf * SD, false) || !alloc(rs_bs, (size_t)M, false) || !alloc(x, (size_t)M * D, false) ||
        !alloc(xpl, (size_t)2 * M * D, false) || !alloc(rn, (size_t)M, false) || !alloc(qkvg, (size_t)M * DQ, false) ||
        !alloc(ao, (size_t)2 * M * DI, false) || !alloc(hpl, (size_t)2 * M * DHID, false) ||
        !alloc(g1, (size_t)2 * M * DHID, false) || !alloc(me3out, (size_t)Mf * 2 * SD, false) ||
        !alloc(enh, (size_t)B2 * (T + 2 * pad) * LD, true))
      return false;
    g_bs.assign(nb, Gemm{}); g_me3.assign(nb, Gemm{}); g_layers.assign(2 * depth, LayerG{});
    for (int i = 0; i < nb; ++i) {
      // band-split: A = columns [off_i, off_i+din_i) of the gathered planes, rows = tokens
      if (!plan_gemm(g_bs[i], xg + off[i], Mf * SD, din[i], (int)Mf, SD, 1, Mf * SD, bs[i])) return false;
      tc::TcArgs& a = g_bs[i].args;
      a.rowscale = rs_bs + (long long)i * Mf; a.bias = bs_b[i];
      a.C = x + (long long)i * Mf * D; a.Chi = xpl + (long long)i * Mf * D; a.Clo = xpl + M * D + (long long)i * Mf * D;
      a.ldc = D;
      // mask-estimator output layer of band i
      if (!plan_gemm(g_me3[i], hpl + (long long)i * Mf * DHID, M * DHID, DHID, (int)Mf, DHID, 1, Mf * DHID, me3[i])) return false;
      tc::TcArgs& c = g_me3[i].args;
      c.bias = me3_b[i]; c.C = me3out + 2 * off[i]; c.ldc = 2 * SD;
    }
    for (int l = 0; l < 2 * depth; ++l) {
      LayerG& G = g_layers[l];
      const Layer& L = layers[l];
      if (!plan_gemm(G.in, xpl, M * D, D, (int)M, D, 1, M * D, L.in)) return false;
      G.in.args.rowscale = rn; G.in.args.bias = L.in_b; G.in.args.C = qkvg; G.in.args.ldc = DQ;
      if (!plan_gemm(G.out, ao, M * DI, DI, (int)M, DI, 1, M * DI, L.out)) return false;
      G.out.args.resid = x; G.out.args.C = x; G.out.args.Chi = xpl; G.out.args.Clo = xpl + M * D; G.out.args.ldc = D;
      if (!plan_gemm(G.ff1, xpl, M * D, D, (int)M, D, 1, M * D, L.ff1)) return false;
      G.ff1.args.rowscale = rn; G.ff1.args.bias = L.ff1_b; G.ff1.args.act = tc::ACT_GELU;
      G.ff1.args.Chi = hpl; G.ff1.args.Clo = hpl + M * DHID; G.ff1.args.ldc = DHID;
      if (!plan_gemm(G.ff2, hpl, M * DHID, DHID, (int)M, DHID, 1, M * DHID, L.ff2)) return false;
      G.ff2.args.bias = L.ff2_b; G.ff2.args.resid = x; G.ff2.args.C = x; G.ff2.args.ldc = D;
    }
    // mask estimator layers 1/2: batched over bands (rows = tokens of one band, batch = band)
    if (!plan_gemm(g_me1, xpl, M * D, D, (int)Mf, D, nb, Mf * D, me1)) return false;
    g_me1.args.bias = me_b1; g_me1.args.bias_bstride = DHID; g_me1.args.act = tc::ACT_TANH;
    g_me1.args.Chi = g1; g_me1.args.Clo = g1 + M * DHID; g_me1.args.ldc = DHID;
    if (!plan_gemm(g_me2, g1, M * DHID, DHID, (int)Mf, DHID, nb, Mf * DHID, me2)) return false;
    g_me2.args.bias = me_b2; g_me2.args.bias_bstride = DHID; g_me2.args.act = tc::ACT_TANH;
    g_me2.args.Chi = hpl; g_me2.args.Clo = hpl + M * DHID; g_me2.args.ldc = DHID;
    // workspace memsets ran on the legacy default stream; runs use a non-blocking stream that does not order against it
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "workspace initialisation failed"; return false; }
    planned = B;
    return true;
  }

  // ---- ModelImpl
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);        // Export_MelBandRoformer.py:715
    in->dtype = in_dtype; in->channels = CH; in->length = W_in;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);   // :716
    out->dtype = out_dtype; out->channels = CH; out->length = W_final;
  }
  size_t workspace_bytes(int batch) override {
    const size_t Mf = (size_t)batch * T, M = (size_t)nb * Mf;
    size_t f = (size_t)batch * CH * Lp + (size_t)batch * CH * T * LD + 2 * Mf * SD + 2 * M + 3 * M * D + M * DQ +
               2 * M * DI + 4 * M * DHID + Mf * 2 * SD + (size_t)batch * CH * (T + 2 * pad) * LD;
    if (rs_in) f += (size_t)batch * CH * W;
    if (rs_out) f += (size_t)batch * CH * ((size_t)W + W_final);
    return f * sizeof(float);
  }
  int launches(int) override { return 3 + nb + 1 + 2 * depth * 7 + 2 + nb + 2; }
  void set_stop_after(int n) override { stop_after = n; }

#define MBR_TICK(name) do { ++n; if (tick) tick(tick_ctx, name); if (stop_after > 0 && n >= stop_after) return ADN_OK; } while (0)
#define MBR_GEMM(G, name) do { cudaError_t e_ = tc::launch((G).plan, (G).args, EPI_LIN, sms, st); \
    if (e_ != cudaSuccess) { err = std::string("gemm launch: ") + cudaGetErrorString(e_); return ADN_ERR_CUDA; } MBR_TICK(name); } while (0)

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    int n = 0;
    const long long Mf = (long long)B * T, M = (long long)nb * Mf;
    const int B2 = B * CH;
    // 1-2: conditioning + STFT (reference kernel pre-scaled by 1/32768 for int16 input, :327-328)
    const void* src = d_in;
    int src_dtype = in_dtype;
    if (rs_in) {                                   // F.interpolate on the raw samples, then the 1/32768 the STFT kernel carries (:327-328)
      if (adn_resample_linear(d_in, in_dtype, xr, B2, W_in, W, in_scale, st) != ADN_OK) { err = "input resampler launch failed"; return ADN_ERR_CUDA; }
      if (in_dtype == ADN_I16) rs_scale_kernel<<<(unsigned)(((long long)B2 * W + 255) / 256), 256, 0, st>>>(xr, (long long)B2 * W, 1.0f / 32768.0f);
      src = xr;
      src_dtype = ADN_F32;
    }
    gtcrn::launch_prep(src, src_dtype, xp, nullptr, nullptr, B2, W, Lp, NFFT / 2, 0, 1, st);
    MBR_TICK("prep");
    {
      GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.A = xp; g.a_sB = Lp; g.a_sT = HOP; g.a_t0 = 0; g.TM = T;
      g.W = d_fwd; g.ldw = NFFT; g.M = B2 * T; g.N = 2050; g.K = NFFT;
      g.C = spec; g.c_sB = (long long)T * LD; g.c_sT = LD; g.c_sN = 1;
      launch_gemm_ffma(g, EPI_STORE, st);
      MBR_TICK("stft_gemm");
    }
    gather_kernel<<<(unsigned)Mf, 256, (size_t)SD * sizeof(float), st>>>(spec, d_freq_idx, d_band_off, xg, xg + Mf * SD,
                                                                         rs_bs, T, (int)Mf, SD, nb);
    MBR_TICK("gather");
    for (int i = 0; i < nb; ++i) MBR_GEMM(g_bs[i], "bs_gemm");
    const unsigned rn_blocks = (unsigned)((M * 32 + 255) / 256);
    rownorm_kernel<<<rn_blocks, 256, 0, st>>>(x, rn, M, D);
    MBR_TICK("rownorm");
    for (int l = 0; l < 2 * depth; ++l) {
      LayerG& G = g_layers[l];
      const bool freq = l & 1;
      MBR_GEMM(G.in, "in_proj");
      {
        const int nseq = freq ? nb : T;
        const long long nq = freq ? Mf : (long long)nb * B;
        static unsigned long long cfg = 0;             // per device
        if (adn_first_use_on_device(cfg)) {
          cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          cudaFuncSetAttribute(attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
          cudaFuncSetAttribute(attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        }
        const size_t smem2 = att2_smem_floats(nseq) * sizeof(float);
        const size_t smem3 = att3_smem_floats(nseq) * sizeof(float);
        static const bool use_mma = !(getenv("ADN_MBR_MMA") && getenv("ADN_MBR_MMA")[0] == '0');
        if (use_mma && nseq > 96 && smem3 <= 220 * 1024) {     // (60-band sequences: the FFMA tiles measured faster, 7.7 vs 8.5 ms per step)
          const int att_threads = 1024;
          attention3_kernel<<<dim3((unsigned)nq, heads), att_threads, smem3, st>>>(
              qkvg, DQ, freq ? fcos : tcos, freq ? fsin : tsin, ao, ao + M * DI, nseq, freq ? Mf : 1, freq ? 1 : 0, heads);
        } else if (smem2 <= 220 * 1024) {
          const int att_threads = nseq > 96 ? 1024 : 512;     // small sequences: more CTAs per SM instead
          attention2_kernel<<<dim3((unsigned)nq, heads), att_threads, smem2, st>>>(
              qkvg, DQ, freq ? fcos : tcos, freq ? fsin : tsin, ao, ao + M * DI, nseq, freq ? Mf : 1, freq ? 1 : 0, heads);
        } else {
          const size_t smem = ((size_t)nseq * 65 + 4 + (size_t)nseq * 64 + (size_t)ATT_WARPS * (nseq + 4) + ATT_WARPS * 64) * sizeof(float);
          attention_kernel<<<dim3((unsigned)nq, heads), ATT_WARPS * 32, smem, st>>>(
              qkvg, DQ, freq ? fcos : tcos, freq ? fsin : tsin, ao, ao + M * DI, nseq, freq ? Mf : 1, freq ? 1 : 0, heads);
        }
        MBR_TICK(freq ? "attention_freq" : "attention_time");
      }
      MBR_GEMM(G.out, "out_proj");
      rownorm_kernel<<<rn_blocks, 256, 0, st>>>(x, rn, M, D);
      MBR_TICK("rownorm");
      MBR_GEMM(G.ff1, "ff1");
      MBR_GEMM(G.ff2, "ff2");
      renorm_kernel<<<rn_blocks, 256, 0, st>>>(x, layers[l].out_g, xpl, xpl + M * D, rn, M, D);
      MBR_TICK("renorm");
    }
    MBR_GEMM(g_me1, "me1");
    MBR_GEMM(g_me2, "me2");
    for (int i = 0; i < nb; ++i) MBR_GEMM(g_me3[i], "me3");
    mask_apply_kernel<<<(unsigned)Mf, 256, 0, st>>>(me3out, d_dst_src, d_src_a, d_src_g, spec, enh, T, pad,
                                                   (long long)2 * SD);
    MBR_TICK("mask_apply");
    {
      GemmArgs g;
      memset(&g, 0, sizeof(g));
      const int half = NFFT / 2;
      const int raw = NFFT + HOP * (T - 1);
      const int lo = half / HOP, hi = (raw - half - 1) / HOP;
      g.A = enh; g.a_sB = (long long)(T + 2 * pad) * LD; g.a_sT = LD; g.a_t0 = lo; g.TM = hi - lo + 1;
      g.W = d_ola; g.ldw = R * LD; g.M = B2 * g.TM; g.N = HOP; g.K = R * LD;
      g.norm = d_norm; g.norm_mul = 0; g.hop = HOP; g.shift = half; g.out_len = W; g.out_dtype = rs_out ? ADN_F32 : out_dtype; g.out = rs_out ? (void*)yres : d_out;
      launch_gemm_ffma(g, EPI_ISTFT, st);
      MBR_TICK("istft_gemm");
      if (rs_out) {                                // down-sampling before the x32767 PCM scale, up-sampling after it (:662-678)
        const long long nm = (long long)B2 * W, no = (long long)B2 * W_final;
        const bool pcm = out_dtype == ADN_I16;
        if (out_sr > 44100 && pcm) rs_scale_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(yres, nm, 32767.0f);
        if (adn_resample_linear(yres, ADN_F32, yout, B2, W, W_final, out_scale, st) != ADN_OK) { err = "output resampler launch failed"; return ADN_ERR_CUDA; }
        if (out_sr < 44100 && pcm) rs_scale_kernel<<<(unsigned)((no + 255) / 256), 256, 0, st>>>(yout, no, 32767.0f);
        rs_convert_kernel<<<(unsigned)((no + 255) / 256), 256, 0, st>>>(yout, d_out, out_dtype, no);
      }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mbr run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    const size_t B = last_batch;
    if (!B) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    const size_t Mf = B * T, M = (size_t)nb * Mf;
    std::map<std::string, std::pair<const float*, size_t>> tbl = {
        {"spec", {spec, B * CH * T * LD}}, {"x", {x, M * D}}, {"qkvg", {qkvg, M * DQ}}, {"rn", {rn, M}},
        {"me3out", {me3out, Mf * 2 * SD}}, {"enh", {enh, B * CH * (T + 2 * pad) * LD}}, {"rs_bs", {rs_bs, M}},
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

}  // namespace mbr

ModelImpl* mbr_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                      const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  mbr::Model* m = new mbr::Model();
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
