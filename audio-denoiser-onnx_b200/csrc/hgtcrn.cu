// H-GTCRN 16 kHz two-microphone denoiser behind the C ABI: model family "h_gtcrn".
// Reference: H-GTCRN/Export_H_GTCRN.py `H_GTCRN_CUSTOM.forward` (:952-1061).
//
//   hg_prep      cast, 1/32768, ONE DC mean over both microphones of a window (:957-967), reflect centre pad
//   stft         windowed-DFT GEMM (periodic hann, 512 / 256), rows = (window, microphone), frame-major (rows, T, 520)
//   hg_eps       per-window WPE floor 1e-3 * mean_f max_{m,t} |X|^2 (:699-700)
//   hg_wpe       one CTA per (window, bin): 36 x 36 weighted correlation of the delay bank (2 x 2 register blocks), 36 x 2 cross term, SIX fixed
//                conjugate-gradient steps per right-hand side, prediction subtracted (:637-753, :499-555)
//   hg_iva       one CTA per window, one thread per bin: ten AuxIVA sweeps (source activity from ALL bins through a block
//                reduction, weighted covariances, Cramer 2 x 2 solves, normalisation), projection back on microphone 0,
//                log-magnitudes and the lower-energy-first source order (:795-900, :1002-1015)
//   hg_enc_front ERB bands + SFE + en_convs.0 (18 -> 16) + en_convs.1 over [mic0 re, im, mic1 re, im, log|Y_sel|, log|Y_other|]
//   gtcrn::launch_backbone_core / launch_dec_tail   the GTCRN network and the complex ratio mask on microphone 0 (gtcrn.cu)
//   istft        overlap-add GEMM + divide by the overlap-added w^2
//   hg_out       x 32767, NaN -> 0 (a silent window makes the front end NaN, :1051-1056), clamp, truncate
//
// min(x, lo) style clamps are written as comparisons so that NaN propagates exactly as torch.clamp does.
#include "adn.h"
#include "common.cuh"
#include "gtcrn.cuh"
#include "model_impl.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

namespace hg {

using gtcrn::FB;
using gtcrn::SPEC_LD;
using gtcrn::ERB_F;
using gtcrn::E0_F;
using gtcrn::E1_F;
using gtcrn::FRAME16;
using gtcrn::FRAME_E0;

constexpr int TAPS = 18;     // Lg = int(0.3 * 16000 / 256)   (:613)
constexpr int DELAY = 2;     // :47
constexpr int NU = 2 * TAPS; // unknowns per bin
constexpr int CG_STEPS = 6;  // :50
constexpr int IVA_SWEEPS = 10;
constexpr float IVA_EPS = 1e-10f;

struct EncFrontHW {          // en_convs.0 (18 -> 16) / en_convs.1, BatchNorm folded (adn/hgtcrn_params.py)
  float w0[5][18][16];
  float b0[16];
  float w1[2][8][5][8];
  float b1[16];
  float a0, a1;
};

// ---------------------------------------------------------------------------------- prep
template <typename Tin>
__device__ __forceinline__ float ld_sample(const Tin* p, long long i);
template <> __device__ __forceinline__ float ld_sample<float>(const float* p, long long i) { return p[i]; }
template <> __device__ __forceinline__ float ld_sample<int16_t>(const int16_t* p, long long i) { return (float)p[i] * (1.0f / 32768.0f); }
template <> __device__ __forceinline__ float ld_sample<__half>(const __half* p, long long i) { return __half2float(p[i]); }

template <typename Tin>
__global__ void __launch_bounds__(256) prep_kernel(const Tin* __restrict__ in, float* __restrict__ xp, int L, int Lp, int half) {
  const int b = blockIdx.x;
  const Tin* x = in + (long long)b * 2 * L;
  __shared__ double red[8];
  __shared__ float mean_s;
  double s = 0.0;
  for (int i = threadIdx.x; i < 2 * L; i += 256) s += (double)ld_sample<Tin>(x, i);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    mean_s = (float)(t / (double)(2 * L));
  }
  __syncthreads();
  const float mean = mean_s;
  for (int i = threadIdx.x; i < 2 * Lp; i += 256) {
    const int m = i / Lp, p = i - m * Lp;
    int j = p - half;
    if (j < 0) j = -j;
    else if (j >= L) j = 2 * (L - 1) - j;
    float v = 0.f;
    if (j >= 0 && j < L) v = ld_sample<Tin>(x, (long long)m * L + j) - mean;
    xp[((long long)b * 2 + m) * Lp + p] = v;
  }
}

// ---------------------------------------------------------------------------------- WPE floor
__global__ void __launch_bounds__(288) eps_kernel(const float* __restrict__ spec, float* __restrict__ eps, int T) {
  const int b = blockIdx.x, f = threadIdx.x;
  __shared__ float part[9];
  float mx = 0.f;
  if (f < FB) {
    for (int m = 0; m < 2; ++m) {
      const float* row = spec + ((long long)b * 2 + m) * T * SPEC_LD;
      for (int t = 0; t < T; ++t) {
        const float re = row[(long long)t * SPEC_LD + f], im = row[(long long)t * SPEC_LD + FB + f];
        const float p = re * re + im * im;
        mx = (p > mx || p != p) ? p : mx;          // amax propagates NaN
      }
    }
  }
  float s = f < FB ? mx : 0.f;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 9; ++w) t += part[w];
    eps[b] = 1e-3f * (t / (float)FB);
  }
}

// ---------------------------------------------------------------------------------- WPE
// dynamic shared memory: X re / im (2 x T each), 1 / lambda (T)
constexpr int WPE_THREADS = 192;      // >= 171 = the 2 x 2 blocks of R's lower triangle: one block per thread, one round
__global__ void __launch_bounds__(WPE_THREADS) wpe_kernel(const float* __restrict__ spec, const float* __restrict__ epsb,
                                                   float* __restrict__ out, int T) {
  extern __shared__ float wsm[];
  float* xr = wsm;                 // [2][T]
  float* xi = wsm + 2 * T;         // [2][T]
  float* il = wsm + 4 * T;         // [T]
  __shared__ float Rr[NU][NU + 1], Ri[NU][NU + 1];
  __shared__ float vr[5][NU][2], vi[5][NU][2];     // P, x, r, p, Ap (re / im), [unknown][rhs column]
  __shared__ float red[NU][2];
  __shared__ float sc[4][2];                       // rr, pAp, rr_new per column
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float eps = epsb[b];
  const float* row0 = spec + ((long long)b * 2) * T * SPEC_LD;
  for (int i = tid; i < 2 * T; i += WPE_THREADS) {
    const int m = i / T, t = i - m * T;
    const float* r = row0 + ((long long)m * T + t) * SPEC_LD;
    xr[i] = r[f];
    xi[i] = r[FB + f];
  }
  __syncthreads();
  for (int t = tid; t < T; t += WPE_THREADS) {
    float p = ((xr[t] * xr[t] + xi[t] * xi[t]) + (xr[T + t] * xr[T + t] + xi[T + t] * xi[T + t])) * 0.5f;   // mean over the two microphones
    p = (p < eps) ? eps : p;
    il[t] = 1.0f / p;
  }
  __syncthreads();
  // unknown u = l * 2 + m is microphone m delayed by DELAY + l frames
  // R[i][j] = sum_t Xd_i conj(Xd_j) / lambda, + eps on the diagonal.  A work item is the 2 x 2 block of a tap pair (l >= l'):
  // its four entries share the eight delayed samples and the weight of every frame (9 shared-memory loads per 16 FMAs
  // instead of 20); the lower triangle is computed and mirrored, the summation order per entry is frame order.
  for (int e = tid; e < TAPS * (TAPS + 1) / 2; e += WPE_THREADS) {
    int bi = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    while (bi * (bi + 1) / 2 > e) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= e) ++bi;
    const int bj = e - bi * (bi + 1) / 2;
    const int si = DELAY + bi, sj = DELAY + bj;
    float re[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, im[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int t = si; t < T; ++t) {                 // si >= sj: both delayed samples exist from t = si on
      const float w = il[t];
      const float a_r[2] = {xr[t - si] * w, xr[T + t - si] * w}, a_i[2] = {xi[t - si] * w, xi[T + t - si] * w};
      const float b_r[2] = {xr[t - sj], xr[T + t - sj]}, b_i[2] = {xi[t - sj], xi[T + t - sj]};
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int mj = 0; mj < 2; ++mj) {
          re[mi][mj] = fmaf(a_r[mi], b_r[mj], fmaf(a_i[mi], b_i[mj], re[mi][mj]));
          im[mi][mj] = fmaf(a_i[mi], b_r[mj], fmaf(-a_r[mi], b_i[mj], im[mi][mj]));
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int mj = 0; mj < 2; ++mj) {
        const int i = 2 * bi + mi, j = 2 * bj + mj;
        if (i == j) { Rr[i][i] = re[mi][mj] + eps; Ri[i][i] = im[mi][mj]; }
        else if (i > j) { Rr[i][j] = re[mi][mj]; Ri[i][j] = im[mi][mj]; Rr[j][i] = re[mi][mj]; Ri[j][i] = -im[mi][mj]; }
      }
  }
  // P[i][c] = sum_t Xd_i conj(X_c) / lambda
  if (tid < NU * 2) {
    const int i = tid >> 1, c = tid & 1;
    const int si = DELAY + (i >> 1);
    const float* ar = xr + (i & 1) * T; const float* ai = xi + (i & 1) * T;
    const float* br = xr + c * T; const float* bi = xi + c * T;
    float re = 0.f, im = 0.f;
    for (int t = si; t < T; ++t) {
      const float w = il[t];
      const float a_r = ar[t - si] * w, a_i = ai[t - si] * w;
      re = fmaf(a_r, br[t], fmaf(a_i, bi[t], re));
      im = fmaf(a_i, br[t], fmaf(-a_r, bi[t], im));
    }
    vr[0][i][c] = re; vi[0][i][c] = im;          // P
    vr[1][i][c] = 0.f; vi[1][i][c] = 0.f;        // x
    vr[2][i][c] = re; vi[2][i][c] = im;          // r
    vr[3][i][c] = re; vi[3][i][c] = im;          // p
    red[i][c] = re * re + im * im;
  }
  __syncthreads();
  if (tid < 2) {
    float s = 0.f;
    for (int i = 0; i < NU; ++i) s += red[i][tid];
    sc[0][tid] = s + 1e-12f;
  }
  __syncthreads();
  for (int step = 0; step < CG_STEPS; ++step) {
    const int i = tid >> 1, c = tid & 1;
    float apr = 0.f, api = 0.f;
    if (tid < NU * 2) {
      for (int j = 0; j < NU; ++j) {
        const float rr_ = Rr[i][j], ri_ = Ri[i][j], pr = vr[3][j][c], pi = vi[3][j][c];
        apr = fmaf(rr_, pr, fmaf(-ri_, pi, apr));
        api = fmaf(rr_, pi, fmaf(ri_, pr, api));
      }
      vr[4][i][c] = apr; vi[4][i][c] = api;
      red[i][c] = vr[3][i][c] * apr + vi[3][i][c] * api;
    }
    __syncthreads();
    if (tid < 2) {
      float s = 0.f;
      for (int k = 0; k < NU; ++k) s += red[k][tid];
      sc[1][tid] = s + 1e-12f;
    }
    __syncthreads();
    if (tid < NU * 2) {
      const float alpha = sc[0][c] / sc[1][c];
      vr[1][i][c] = fmaf(alpha, vr[3][i][c], vr[1][i][c]);
      vi[1][i][c] = fmaf(alpha, vi[3][i][c], vi[1][i][c]);
      const float rr_ = fmaf(-alpha, apr, vr[2][i][c]), ri_ = fmaf(-alpha, api, vi[2][i][c]);
      vr[2][i][c] = rr_; vi[2][i][c] = ri_;
      red[i][c] = rr_ * rr_ + ri_ * ri_;
    }
    __syncthreads();
    if (tid < 2) {
      float s = 0.f;
      for (int k = 0; k < NU; ++k) s += red[k][tid];
      sc[2][tid] = s + 1e-12f;
    }
    __syncthreads();
    if (tid < NU * 2) {
      const float beta = sc[2][c] / sc[0][c];
      vr[3][i][c] = fmaf(beta, vr[3][i][c], vr[2][i][c]);
      vi[3][i][c] = fmaf(beta, vi[3][i][c], vi[2][i][c]);
    }
    __syncthreads();
    if (tid < 2) sc[0][tid] = sc[2][tid];
    __syncthreads();
  }
  // Y_m(t) = X_m(t) - sum_u conj(G[u][m]) Xd_u(t)
  for (int o = tid; o < 2 * T; o += WPE_THREADS) {
    const int m = o / T, t = o - m * T;
    float pr = 0.f, pi = 0.f;
    for (int u = 0; u < NU; ++u) {
      const int tt = t - DELAY - (u >> 1);
      if (tt < 0) break;
      const float gr = vr[1][u][m], gi = vi[1][u][m];
      const float dr = xr[(u & 1) * T + tt], di = xi[(u & 1) * T + tt];
      pr = fmaf(gr, dr, fmaf(gi, di, pr));
      pi = fmaf(gr, di, fmaf(-gi, dr, pi));
    }
    float* w = out + (((long long)b * 2 + m) * T + t) * SPEC_LD;
    w[f] = xr[o] - pr;
    w[FB + f] = xi[o] - pi;
  }
}

// ---------------------------------------------------------------------------------- AuxIVA + features
struct C2 { float r, i; };
__device__ __forceinline__ C2 cmul(C2 a, C2 b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ C2 cadd(C2 a, C2 b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ C2 csub(C2 a, C2 b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ C2 cconj(C2 a) { return {a.r, -a.i}; }

constexpr int IVA_THREADS = 288, IVA_WARPS = 9;

// dynamic shared memory: rinv [2][T], part [9][2][T]
__global__ void __launch_bounds__(IVA_THREADS) iva_kernel(const float* __restrict__ wpe, float* __restrict__ iva, float* __restrict__ logs,
                                                          int* __restrict__ swap, int T) {
  extern __shared__ float ism[];
  float* rinv = ism;                  // [2][T]
  float* part = ism + 2 * T;          // [9][2][T]
  __shared__ float epart[IVA_WARPS][2];
  const int b = blockIdx.x, f = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool on = f < FB;
  const float* x0 = wpe + ((long long)b * 2) * T * SPEC_LD + (on ? f : 0);
  const float* x1 = x0 + (long long)T * SPEC_LD;
  const float inv_T = 1.0f / (float)T;
  C2 W[2][2] = {{{1.f, 0.f}, {0.f, 0.f}}, {{0.f, 0.f}, {1.f, 0.f}}};
  for (int sweep = 0; sweep < IVA_SWEEPS; ++sweep) {
    // source activity r_m(t) = 2 sqrt(sum_f |Y_m|^2 + eps) from this sweep's starting W
    for (int t = 0; t < T; ++t) {
      float p0 = 0.f, p1 = 0.f;
      if (on) {
        const C2 a = {x0[(long long)t * SPEC_LD], x0[(long long)t * SPEC_LD + FB]};
        const C2 c = {x1[(long long)t * SPEC_LD], x1[(long long)t * SPEC_LD + FB]};
        const C2 y0 = cadd(cmul(W[0][0], a), cmul(W[0][1], c)), y1 = cadd(cmul(W[1][0], a), cmul(W[1][1], c));
        p0 = y0.r * y0.r + y0.i * y0.i;
        p1 = y1.r * y1.r + y1.i * y1.i;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, off);
        p1 += __shfl_xor_sync(0xffffffffu, p1, off);
      }
      if (lane == 0) { part[(warp * 2 + 0) * T + t] = p0; part[(warp * 2 + 1) * T + t] = p1; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * T; i += IVA_THREADS) {
      const int m = i / T, t = i - m * T;
      float s = 0.f;
      for (int w = 0; w < IVA_WARPS; ++w) s += part[(w * 2 + m) * T + t];
      rinv[i] = 1.0f / (2.0f * sqrtf(s + IVA_EPS));
    }
    __syncthreads();
    if (on) {
      // V_s = X diag(1 / r_s) X^H / T for both sources in one pass (V_s does not depend on W)
      float v00[2] = {0.f, 0.f}, v11[2] = {0.f, 0.f}, v01r[2] = {0.f, 0.f}, v01i[2] = {0.f, 0.f};
      for (int t = 0; t < T; ++t) {
        const float ar = x0[(long long)t * SPEC_LD], ai = x0[(long long)t * SPEC_LD + FB];
        const float cr = x1[(long long)t * SPEC_LD], ci = x1[(long long)t * SPEC_LD + FB];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const float w = rinv[s * T + t];
          const float war = ar * w, wai = ai * w, wcr = cr * w, wci = ci * w;
          v00[s] = fmaf(war, ar, fmaf(wai, ai, v00[s]));
          v11[s] = fmaf(wcr, cr, fmaf(wci, ci, v11[s]));
          v01r[s] = fmaf(war, cr, fmaf(wai, ci, v01r[s]));
          v01i[s] = fmaf(wai, cr, fmaf(-war, ci, v01i[s]));
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const C2 V00 = {v00[s] * inv_T, 0.f}, V11 = {v11[s] * inv_T, 0.f};
        const C2 V01 = {v01r[s] * inv_T, v01i[s] * inv_T}, V10 = {V01.r, -V01.i};
        // A = W V + eps I
        C2 A00 = cadd(cmul(W[0][0], V00), cmul(W[0][1], V10)), A01 = cadd(cmul(W[0][0], V01), cmul(W[0][1], V11));
        C2 A10 = cadd(cmul(W[1][0], V00), cmul(W[1][1], V10)), A11 = cadd(cmul(W[1][0], V01), cmul(W[1][1], V11));
        A00.r += IVA_EPS; A11.r += IVA_EPS;
        const C2 det = csub(cmul(A00, A11), cmul(A01, A10));
        const float k = 1.0f / ((det.r * det.r + det.i * det.i) + 1e-12f);
        const C2 inv = {det.r * k, -det.i * k};
        C2 w0, w1;                               // A w = e_s (Cramer)
        if (s == 0) { w0 = cmul(A11, inv); w1 = cmul(C2{-A10.r, -A10.i}, inv); }
        else { w0 = cmul(C2{-A01.r, -A01.i}, inv); w1 = cmul(A00, inv); }
        const C2 Vw0 = cadd(cmul(V00, w0), cmul(V01, w1)), Vw1 = cadd(cmul(V10, w0), cmul(V11, w1));
        float den = (w0.r * Vw0.r + w0.i * Vw0.i) + (w1.r * Vw1.r + w1.i * Vw1.i);
        den = (den < 0.f) ? 0.f : den;
        const float sc = rsqrtf(den + IVA_EPS);
        W[s][0] = {w0.r * sc, -w0.i * sc};
        W[s][1] = {w1.r * sc, -w1.i * sc};
      }
    }
    __syncthreads();                             // rinv / part are rewritten by the next sweep
  }
  // projection back on microphone 0, source powers, log-magnitudes
  float e0 = 0.f, e1 = 0.f;
  if (on) {
    C2 num[2] = {{0.f, 0.f}, {0.f, 0.f}};
    float den[2] = {0.f, 0.f};
    for (int t = 0; t < T; ++t) {
      const C2 a = {x0[(long long)t * SPEC_LD], x0[(long long)t * SPEC_LD + FB]};
      const C2 c = {x1[(long long)t * SPEC_LD], x1[(long long)t * SPEC_LD + FB]};
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const C2 y = cadd(cmul(W[m][0], a), cmul(W[m][1], c));
        num[m].r = fmaf(a.r, y.r, fmaf(a.i, y.i, num[m].r));        // conj(ref) * Y
        num[m].i = fmaf(a.r, y.i, fmaf(-a.i, y.r, num[m].i));
        den[m] = fmaf(y.r, y.r, fmaf(y.i, y.i, den[m]));
      }
    }
    C2 cc[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const bool ok = den[m] > 0.f;
      const float id = 1.0f / (ok ? den[m] : 1.0f);
      cc[m] = ok ? C2{num[m].r * id, num[m].i * id} : C2{1.f, 0.f};
    }
    for (int t = 0; t < T; ++t) {
      const C2 a = {x0[(long long)t * SPEC_LD], x0[(long long)t * SPEC_LD + FB]};
      const C2 c = {x1[(long long)t * SPEC_LD], x1[(long long)t * SPEC_LD + FB]};
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const C2 y = cadd(cmul(W[m][0], a), cmul(W[m][1], c));
        const float yr = cc[m].r * y.r + cc[m].i * y.i, yi = cc[m].r * y.i - cc[m].i * y.r;     // conj(c) * Y  (:895-898)
        float* o = iva + (((long long)b * 2 + m) * T + t) * SPEC_LD;
        o[f] = yr; o[FB + f] = yi;
        float p = yr * yr + yi * yi;
        if (m == 0) e0 += p; else e1 += p;
        p = (p < 1e-24f) ? 1e-24f : p;
        logs[(((long long)b * 2 + m) * T + t) * FB + f] = 0.5f * log10f(p);
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    e0 += __shfl_xor_sync(0xffffffffu, e0, off);
    e1 += __shfl_xor_sync(0xffffffffu, e1, off);
  }
  if (lane == 0) { epart[warp][0] = e0; epart[warp][1] = e1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int w = 0; w < IVA_WARPS; ++w) { s0 += epart[w][0]; s1 += epart[w][1]; }
    swap[b] = (s0 < s1) ? 0 : 1;                 // :1006-1015: the lower-energy source goes first
  }
}

// ---------------------------------------------------------------------------------- encoder front
constexpr int EF_FR = 2, EF_THREADS = 288;

__global__ void __launch_bounds__(EF_THREADS)
enc_front_kernel(const __grid_constant__ EncFrontHW w, const gtcrn::ErbW erb, const float* __restrict__ spec,
                 const float* __restrict__ logs, const int* __restrict__ swap, float* __restrict__ e0, float* __restrict__ e1,
                 int T, int nframes) {
  __shared__ float ch[EF_FR][6][FB];                  // the six feature channels of a frame
  __shared__ float fe[EF_FR][6][ERB_F + 8];           // ERB features, 4 zeros each side
  __shared__ float e0s[EF_FR][16][E0_F + 4];          // en_convs.0 output, 2 zeros each side
  const int tid = threadIdx.x;
  const long long f0 = (long long)blockIdx.x * EF_FR;
  for (int i = tid; i < EF_FR * 6 * FB; i += EF_THREADS) {
    const int fr = i / (6 * FB), r = i - fr * (6 * FB);
    const int c = r / FB, f = r - c * FB;
    const long long fg = f0 + fr;
    float v = 0.f;
    if (fg < nframes) {
      const long long b = fg / T, t = fg - b * T;
      if (c < 4) v = __ldg(spec + ((b * 2 + (c >> 1)) * T + t) * SPEC_LD + (c & 1) * FB + f);
      else {
        const int src = (c - 4) ^ __ldg(swap + b);
        v = __ldg(logs + ((b * 2 + src) * T + t) * FB + f);
      }
    }
    ch[fr][c][f] = v;
  }
  for (int i = tid; i < EF_FR * 6 * (ERB_F + 8); i += EF_THREADS) (&fe[0][0][0])[i] = 0.f;
  for (int i = tid; i < EF_FR * 16 * (E0_F + 4); i += EF_THREADS) (&e0s[0][0][0])[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < EF_FR * 6 * 65; i += EF_THREADS) {
    const int fr = i / (6 * 65), r = i - fr * (6 * 65);
    const int c = r / 65, f = r - c * 65;
    fe[fr][c][4 + f] = ch[fr][c][f];
  }
  for (int i = tid; i < EF_FR * 6 * 64; i += EF_THREADS) {
    const int fr = i / (6 * 64), r = i - fr * (6 * 64);
    const int c = r >> 6, j = r & 63;
    const int lo = (int)__ldg(erb.bm_lo + j), hi = (int)__ldg(erb.bm_hi + j);
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(ch[fr][c][65 + k], __ldg(erb.bm + k * 64 + j), acc);
    fe[fr][c][4 + 65 + j] = acc;
  }
  __syncthreads();
  // en_convs.0: Conv2d(18 -> 16, (1,5), stride 2, pad 2) over SFE(k = 3) of the six ERB channels
  for (int i = tid; i < EF_FR * E0_F; i += EF_THREADS) {
    const int fr = i / E0_F, g = i - fr * E0_F;
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = w.b0[o];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int p = 2 * g + k - 2;
      const bool pv = (p >= 0) && (p < ERB_F);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float v = pv ? fe[fr][c][4 + p + s - 1] : 0.f;
#pragma unroll
          for (int o = 0; o < 16; ++o) acc[o] = fmaf(w.w0[k][c * 3 + s][o], v, acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) e0s[fr][o][2 + g] = adn_prelu(acc[o], w.a0);
  }
  __syncthreads();
  for (int i = tid; i < EF_FR * FRAME_E0; i += EF_THREADS) {
    const int fr = i / FRAME_E0, r = i - fr * FRAME_E0;
    const int o = r / E0_F, g = r - o * E0_F;
    const long long fg = f0 + fr;
    if (fg < nframes) e0[fg * FRAME_E0 + r] = e0s[fr][o][2 + g];
  }
  // en_convs.1: Conv2d(16 -> 16, (1,5), stride 2, pad 2, groups 2)
  for (int i = tid; i < EF_FR * 2 * E1_F; i += EF_THREADS) {
    const int fr = i / (2 * E1_F), r = i - fr * (2 * E1_F);
    const int grp = r / E1_F, g = r - grp * E1_F;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = w.b1[grp * 8 + o];
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float v = e0s[fr][grp * 8 + ci][2 * g + k];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = fmaf(w.w1[grp][ci][k][o], v, acc[o]);
      }
    const long long fg = f0 + fr;
    if (fg < nframes) {
#pragma unroll
      for (int o = 0; o < 8; ++o) e1[fg * FRAME16 + (grp * 8 + o) * E1_F + g] = adn_prelu(acc[o], w.a1);
    }
  }
}

// ---------------------------------------------------------------------------------- output rule (:1042-1061)
template <typename Tout>
__global__ void __launch_bounds__(256) out_kernel(const float* __restrict__ y, Tout* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float v = y[i];
  if (sizeof(Tout) == 2 && !std::is_same<Tout, __half>::value) v *= 32767.0f;
  if (v != v) v = 0.f;
  if (std::is_same<Tout, int16_t>::value) {
    v = v < -32768.0f ? -32768.0f : (v > 32767.0f ? 32767.0f : v);
    ((int16_t*)out)[i] = (int16_t)__float2int_rz(v);
  } else if (std::is_same<Tout, __half>::value) {
    ((__half*)out)[i] = __float2half_rn(v);
  } else {
    ((float*)out)[i] = v;
  }
}

// ==================================================================================
class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, T = 0, Lp = 0, pad = 0, enh_slack = 0;
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;
  gtcrn::Weights W;
  EncFrontHW efw;
  gtcrn::Buffers buf{};
  adn_stft* stft = nullptr;
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  int cap = 0;
  float *xp = nullptr, *spec = nullptr, *eps = nullptr, *wpe = nullptr, *iva = nullptr, *logs = nullptr, *wave = nullptr;
  int* swap = nullptr;
  int stop_after = 0, last_launches = 0, last_batch = 0;

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
    ws_bytes = 0;
    cap = 0;
  }
  float* dalloc(size_t n, bool zero = false) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) { err = "h_gtcrn: out of device memory for the workspace"; return nullptr; }
    allocs.push_back(p);
    ws_bytes += n * sizeof(float);
    if (zero) cudaMemset(p, 0, n * sizeof(float));
    return (float*)p;
  }
  bool init(const std::map<std::string, std::string>& meta, const float* h_blob) {
    auto gets = [&](const char* k, std::string& v) {
      auto it = meta.find(k);
      if (it == meta.end() || it->second.empty()) { err = std::string("Required metadata key ") + k + " is missing."; return false; }
      v = it->second;
      return true;
    };
    auto geti = [&](const char* k, int& v) { std::string s; if (!gets(k, s)) return false; v = atoi(s.c_str()); return true; };
    int nfft = 0, hop = 0, chans = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", L) || !geti("nfft", nfft) || !geti("hop_length", hop) || !geti("input_channels", chans) ||
        !gets("input_audio_dtype", sin) || !gets("output_audio_dtype", sout))
      return false;
    if (nfft != gtcrn::NFFT || hop != gtcrn::HOP || chans != 2 || L < nfft || L % hop) {
      err = "h_gtcrn needs nfft=512, hop_length=256, input_channels=2 and a window of k * 256 >= 512 samples";
      return false;
    }
    {
      auto opt = [&](const char* k, double dflt) { auto it = meta.find(k); return (it != meta.end() && !it->second.empty()) ? atof(it->second.c_str()) : dflt; };
      if ((int)opt("in_sample_rate", 16000) != 16000 || (int)opt("out_sample_rate", 16000) != 16000) { err = "h_gtcrn runs at 16 kHz I/O only"; return false; }
      if ((int)opt("wpe_delay", DELAY) != DELAY || (int)opt("wpe_iter", 1) != 1 || (int)opt("iva_iter", IVA_SWEEPS) != IVA_SWEEPS ||
          (int)opt("cg_solve_iter", CG_STEPS) != CG_STEPS || (int)(opt("wpe_rt60", 0.3) * 16000.0 / 256.0) != TAPS) {
        err = "h_gtcrn front end is built for wpe_rt60=0.3, wpe_delay=2, wpe_iter=1, cg_solve_iter=6, iva_iter=10";
        return false;
      }
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "input/output_audio_dtype must be F32, F16 or INT16"; return false; }
    T = L / hop + 1;
    err.clear();
    bool ok = true;
    auto dp = [&](const std::string& name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || (expect && it->second.count != expect)) {
        if (ok) err = "weight blob: tensor '" + name + "' missing or wrong size";
        ok = false;
        return nullptr;
      }
      return d_blob + it->second.offset;
    };
    auto ls = [&](const std::string& name, void* dst, size_t bytes) {
      auto it = index.find(name);
      if (it == index.end() || it->second.count != bytes / sizeof(float)) {
        if (ok) err = "weight blob: tensor '" + name + "' missing or not " + std::to_string(bytes / sizeof(float)) + " floats";
        ok = false;
        return;
      }
      memcpy(dst, h_blob + it->second.offset, bytes);
    };
    auto gru = [&](const std::string& p, int I, int H) {
      gtcrn::GruPtrs g;
      g.w_ih = dp(p + ".w_ih", (size_t)3 * H * I); g.w_hh = dp(p + ".w_hh", (size_t)3 * H * H);
      g.b_ih = dp(p + ".b_ih", (size_t)3 * H); g.b_hh = dp(p + ".b_hh", (size_t)3 * H);
      return g;
    };
    ls("enc_front_h", &efw, sizeof(efw));
    ls("dec_tail", &W.dec_tail, sizeof(W.dec_tail));
    for (int i = 0; i < 3; ++i) {
      const std::string si = std::to_string(i);
      ls("enc_gt." + si, &W.enc_gt[i], sizeof(gtcrn::GTW));
      ls("dec_gt." + si, &W.dec_gt[i], sizeof(gtcrn::GTW));
      W.enc_tra[i].gru = gru("enc_tra." + si, 8, 16);
      W.enc_tra[i].fc_w = dp("enc_tra." + si + ".fc_w", 128); W.enc_tra[i].fc_b = dp("enc_tra." + si + ".fc_b", 8);
      W.dec_tra[i].gru = gru("dec_tra." + si, 8, 16);
      W.dec_tra[i].fc_w = dp("dec_tra." + si + ".fc_w", 128); W.dec_tra[i].fc_b = dp("dec_tra." + si + ".fc_b", 8);
    }
    for (int i = 0; i < 2; ++i) {
      const std::string p = "dp." + std::to_string(i);
      for (int g = 0; g < 2; ++g) {
        for (int d = 0; d < 2; ++d) W.dp[i].intra[g][d] = gru(p + ".intra." + std::to_string(g) + "." + std::to_string(d), 8, 4);
        W.dp[i].inter[g] = gru(p + ".inter." + std::to_string(g), 8, 8);
      }
      W.dp[i].intra_fc_w = dp(p + ".intra_fc_w", 256); W.dp[i].intra_fc_b = dp(p + ".intra_fc_b", 16);
      W.dp[i].intra_ln_w = dp(p + ".intra_ln_w", 528); W.dp[i].intra_ln_b = dp(p + ".intra_ln_b", 528);
      W.dp[i].inter_fc_w = dp(p + ".inter_fc_w", 256); W.dp[i].inter_fc_b = dp(p + ".inter_fc_b", 16);
      W.dp[i].inter_ln_w = dp(p + ".inter_ln_w", 528); W.dp[i].inter_ln_b = dp(p + ".inter_ln_b", 528);
    }
    W.erb.bm = dp("erb.bm", 192 * 64); W.erb.bm_lo = dp("erb.bm_lo", 64); W.erb.bm_hi = dp("erb.bm_hi", 64);
    W.erb.bs = dp("erb.bs", 64 * 192); W.erb.bs_lo = dp("erb.bs_lo", 192); W.erb.bs_hi = dp("erb.bs_hi", 192);
    if (!ok) return false;
    auto host = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || it->second.count != expect) { err = std::string("weight blob: tensor '") + name + "' missing or wrong size"; return nullptr; }
      return h_blob + it->second.offset;
    };
    const float* fwd = host("stft.fwd", (size_t)514 * 512);
    const float* inv = host("istft.inv", (size_t)514 * 512);
    const float* nrm = host("istft.norm", (size_t)L);
    if (!fwd || !inv || !nrm) return false;
    adn_stft_geom g;
    memset(&g, 0, sizeof(g));
    g.nfft = nfft; g.hop = hop; g.center = 1; g.pad_reflect = 1; g.norm_multiply = 0;
    if (adn_stft_create(&stft, &g, fwd, inv, nrm, T, device) != ADN_OK) { err = std::string("h_gtcrn: ") + adn_last_error(nullptr); return false; }
    Lp = adn_stft_padded_len(stft, L);
    pad = adn_stft_pad_frames(stft);
    if (adn_stft_ld(stft) != SPEC_LD) { err = "h_gtcrn: unexpected spectrum row stride"; return false; }
    { const char* e = getenv("ADN_STFT_TC"); if (!(e && e[0] == '0')) enh_slack = adn_stft_enable_tc(stft, sms); }   // STFT / ISTFT on tcgen05 (3xTF32)
    return true;
  }
  bool ensure(int B) {
    if (B <= cap) return true;
    cudaDeviceSynchronize();
    free_ws();
    const size_t b = (size_t)B, t = (size_t)T;
    bool ok = (xp = dalloc(b * 2 * Lp)) && (spec = dalloc(b * 2 * t * SPEC_LD, true)) && (eps = dalloc(b)) &&
              (wpe = dalloc(b * 2 * t * SPEC_LD, true)) && (iva = dalloc(b * 2 * t * SPEC_LD, true)) && (logs = dalloc(b * 2 * t * FB)) &&
              (swap = (int*)dalloc(b)) && (wave = dalloc(b * L));
    if (!ok) return false;
    ok = (buf.e0 = dalloc(b * t * FRAME_E0));
    buf.e[0] = nullptr;
    for (int i = 1; i <= 4 && ok; ++i) ok = (buf.e[i] = dalloc(b * t * FRAME16));
    ok = ok && (buf.h1 = dalloc(b * t * 8 * E1_F)) && (buf.zt = dalloc(b * t * 8)) && (buf.at = dalloc(b * t * 8)) &&
         (buf.tgi = dalloc(b * t * 48)) && (buf.thid = dalloc(b * t * 16)) && (buf.gi = dalloc(b * t * 3 * FRAME16)) &&
         (buf.xa = dalloc(b * t * FRAME16)) && (buf.xb = dalloc(b * t * FRAME16)) && (buf.inter = dalloc(b * t * FRAME16)) &&
         (buf.enh = dalloc(b * (t + 2 * pad) * SPEC_LD + enh_slack, true));
    if (!ok) return false;
    buf.xp = xp; buf.spec = spec;
    buf.xp_hi = buf.xp_lo = buf.enh_hi = buf.enh_lo = nullptr;
    cudaDeviceSynchronize();                       // the zero-fills ran on the legacy default stream
    cap = B;
    return true;
  }
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);        // Export_H_GTCRN.py:1151
    in->dtype = in_dtype; in->channels = 2; in->length = L;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);   // :1152
    out->dtype = out_dtype; out->channels = 1; out->length = L;
  }
  size_t workspace_bytes(int batch) override {
    const size_t b = (size_t)batch, t = (size_t)T;
    size_t f = b * 2 * Lp + 3 * b * 2 * t * SPEC_LD + b * 2 * t * FB + 2 * b + b * L;
    f += b * t * FRAME_E0 + 4 * b * t * FRAME16 + b * t * (8 * E1_F + 8 + 8 + 48 + 16) + b * t * 3 * FRAME16 + 3 * b * t * FRAME16 +
         b * (t + 2 * pad) * SPEC_LD;
    return f * sizeof(float);
  }
  int launches(int) override { return 6 + 23 + 3; }
  void set_stop_after(int n) override { stop_after = n; }

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    int n = 0;
#define HG_TICK(name) do { ++n; if (tick) tick(tick_ctx, name); } while (0)
    if (in_dtype == ADN_I16) prep_kernel<int16_t><<<B, 256, 0, st>>>((const int16_t*)d_in, xp, L, Lp, gtcrn::NFFT / 2);
    else if (in_dtype == ADN_F16) prep_kernel<__half><<<B, 256, 0, st>>>((const __half*)d_in, xp, L, Lp, gtcrn::NFFT / 2);
    else prep_kernel<float><<<B, 256, 0, st>>>((const float*)d_in, xp, L, Lp, gtcrn::NFFT / 2);
    HG_TICK("hg_prep");
    if (adn_stft_forward_fm(stft, xp, spec, 2 * B, T, Lp, st) != ADN_OK) { err = "h_gtcrn: stft launch failed"; return ADN_ERR_CUDA; }
    HG_TICK("stft");
    eps_kernel<<<B, 288, 0, st>>>(spec, eps, T);
    HG_TICK("hg_eps");
    {
      static unsigned long long configured = 0;
      const size_t sm = (size_t)5 * T * sizeof(float);
      if (sm > 48 * 1024 - 24 * 1024 && adn_first_use_on_device(configured))
        cudaFuncSetAttribute(wpe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      wpe_kernel<<<dim3(FB, B), WPE_THREADS, sm, st>>>(spec, eps, wpe, T);
      HG_TICK("hg_wpe");
    }
    {
      static unsigned long long configured = 0;
      const size_t sm = (size_t)(2 + 2 * IVA_WARPS) * T * sizeof(float);
      if (sm > 40 * 1024 && adn_first_use_on_device(configured))
        cudaFuncSetAttribute(iva_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      iva_kernel<<<B, IVA_THREADS, sm, st>>>(wpe, iva, logs, swap, T);
      HG_TICK("hg_iva");
    }
    const int nframes = B * T;
    enc_front_kernel<<<(nframes + EF_FR - 1) / EF_FR, EF_THREADS, 0, st>>>(efw, W.erb, spec, logs, swap, buf.e0, buf.e[1], T, nframes);
    HG_TICK("hg_enc_front");
    gtcrn::Dims d{B, L, Lp, T};
    float* cur = nullptr;
    n = gtcrn::launch_backbone_core(W, buf, d, st, tick, tick_ctx, 0, n, &cur);
    gtcrn::launch_dec_tail(W, buf, cur, spec, (long long)2 * T * SPEC_LD, d, pad, st);
    HG_TICK("dec_tail");
    if (adn_stft_inverse_fm(stft, buf.enh, wave, B, T, st) != ADN_OK) { err = "h_gtcrn: istft launch failed"; return ADN_ERR_CUDA; }
    HG_TICK("istft");
    const long long tot = (long long)B * L;
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    if (out_dtype == ADN_I16) out_kernel<int16_t><<<blocks, 256, 0, st>>>(wave, (int16_t*)d_out, tot);
    else if (out_dtype == ADN_F16) out_kernel<__half><<<blocks, 256, 0, st>>>(wave, (__half*)d_out, tot);
    else out_kernel<float><<<blocks, 256, 0, st>>>(wave, (float*)d_out, tot);
    HG_TICK("hg_out");
#undef HG_TICK
    last_launches = n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("h_gtcrn run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    if (!last_batch) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    const size_t b = (size_t)last_batch, t = (size_t)T;
    const float* src = nullptr;
    size_t nel = 0;
    float tmp = 0.f;
    if (!strcmp(name, "launches")) { tmp = (float)last_launches; nel = 1; }
    else if (!strcmp(name, "spec")) { src = spec; nel = b * 2 * t * SPEC_LD; }
    else if (!strcmp(name, "wpe")) { src = wpe; nel = b * 2 * t * SPEC_LD; }
    else if (!strcmp(name, "iva")) { src = iva; nel = b * 2 * t * SPEC_LD; }
    else if (!strcmp(name, "logs")) { src = logs; nel = b * 2 * t * FB; }
    else if (!strcmp(name, "eps")) { src = eps; nel = b; }
    else if (!strcmp(name, "swap")) { src = (const float*)swap; nel = b; }      // int32 bit patterns
    else if (!strcmp(name, "e0")) { src = buf.e0; nel = b * t * FRAME_E0; }
    else if (!strcmp(name, "e1")) { src = buf.e[1]; nel = b * t * FRAME16; }
    else if (!strcmp(name, "e4")) { src = buf.e[4]; nel = b * t * FRAME16; }
    else if (!strcmp(name, "enh")) { src = buf.enh; nel = b * (t + 2 * pad) * SPEC_LD; }
    else { err = std::string("adn_debug_read: unknown tensor '") + name + "'"; return ADN_ERR_INVALID; }
    if (actual) *actual = nel;
    if (!h_dst) return ADN_OK;
    const size_t nc = count < nel ? count : nel;
    if (!src) { if (nc) h_dst[0] = tmp; return ADN_OK; }
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    if (cudaMemcpy(h_dst, src, nc * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { err = "adn_debug_read: copy failed"; return ADN_ERR_CUDA; }
    return ADN_OK;
  }
};

}  // namespace hg

ModelImpl* hgtcrn_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                         const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  hg::Model* m = new hg::Model();
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
