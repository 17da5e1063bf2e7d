// MossFormer2-SS-16K (two-speaker separation): two-stage RMS normalisation, learned Conv1d encoder,
// 24 x (FLASH attention block + dilated gated FSMN block), speaker-stacked mask tail, ConvTranspose1d
// decoder and per-speaker RMS restore (reference MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:403-662).
//
// Layout: activations are token-major fp32, row m = window*n + frame (n = (L-16)/8 + 1 encoder frames).
// A window spans G = ceil(n/256) FLASH groups; everything the attention touches lives in a group-padded
// layout (row = window*Tg + frame, Tg = 256*G) whose pad rows stay zero, which is exactly the reference's
// zero padding of keys and values (:482-493).  All dense contractions run on the tcgen05 3xTF32 GEMM:
//     S_g  = relu(Qq_g Kq_g^T)^2          chunk = (window, group)      A = quad_q   W = quad_k (per chunk)
//     KV^T = [v|u]^T Kl                   chunk = window, K = Tg       A = [v|u]^T  W = lin_k^T
//     O_g  = [S_g | Ql_g] [[v|u]_g ; KV]  chunk = (window, group)      A = [S | lin_q] (K = 256 + 128)
//                                                                      W = [v|u]^T at K offset 256*g, then KV^T
// The linear branch is global over the window (:500-504; 1/n is folded into lin_k, :250-251); its product with
// lin_q rides along the quadratic value product as 128 extra K columns, so O is written exactly once.
//
// Kernel <-> reference map (shared FLASH / FSMN kernels: mf2_kernels.cuh):
//   adn_two_stage_rms   norm_audio (:403-423)                       [ends.cu]
//   enc_kernel          Conv1d(1->512, k16, s8) + ReLU (:588-591), GroupNorm partial sums
//   encnorm_kernel      GroupNorm(1) statistics (:598), operand planes of the folded 1x1 conv (:599), emb_pos (:600)
//   mem1 / mem2         dilated dense depthwise memory convs (:531-537), InstanceNorm partial sums
//   inorm_stats_kernel  InstanceNorm statistics over the window (:538-541)
//   fsmn_out_kernel     InstanceNorm + PReLU (:538-542), xu + mem, gate (:545-547), norm2 (:549)
//   dec_kernel          mask x encoder output (:619-620), ConvTranspose1d frame products (:621-624)
//   ola_out_kernel      overlap-add, RMS restore (:627-632), output conversion (:649-657)
#include "mf2_kernels.cuh"

namespace mf2 {

constexpr int ENC_K = 16, ENC_S = 8, SPK = 2, GROUP = 256;
constexpr int SQ = GROUP + QK;              // row of the attention A operand: [S (256 keys of the group) | lin_q (128)]
constexpr int ENC_TT = 32;                  // frames per encoder CTA
constexpr int MEM_TT = 64;                  // frames per memory-conv CTA
constexpr int MEM1_ROWS = MEM_TT + 2 * MEMH;        // dilation 1: halo 19 each side
constexpr int MEM2_ROWS = MEM_TT + 4 * MEMH;        // dilation 2: halo 38 each side

__device__ __forceinline__ double block_sum_d(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// CTA = 32 frames of one window, thread = 2 channels.  x_enc (token-major) and the CTA's sum / sum of squares.
static __global__ void __launch_bounds__(256)
enc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
           float* __restrict__ xenc, double* __restrict__ part, int L, int n) {
  __shared__ float xs[ENC_TT * ENC_S + ENC_K];
  __shared__ double red[8];
  const int b = blockIdx.y, t0 = blockIdx.x * ENC_TT, tid = threadIdx.x;
  const float* xb = x + (long long)b * L;
  for (int i = tid; i < ENC_TT * ENC_S + ENC_K; i += 256) {
    const int s = t0 * ENC_S + i;
    xs[i] = s < L ? __ldg(xb + s) : 0.f;
  }
  float wk[2][ENC_K], bb[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = tid + j * 256;
#pragma unroll
    for (int q = 0; q < ENC_K / 4; ++q) {
      const float4 v = ld4(w + c * ENC_K + 4 * q);
      wk[j][4 * q] = v.x; wk[j][4 * q + 1] = v.y; wk[j][4 * q + 2] = v.z; wk[j][4 * q + 3] = v.w;
    }
    bb[j] = __ldg(bias + c);
  }
  __syncthreads();
  double su = 0.0, sq = 0.0;
  for (int f = 0; f < ENC_TT; ++f) {
    const int t = t0 + f;
    if (t >= n) break;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < ENC_K; ++k) acc += wk[j][k] * xs[f * ENC_S + k];
      acc = fmaxf(acc + bb[j], 0.f);
      xenc[((long long)b * n + t) * D + tid + j * 256] = acc;
      su += (double)acc;
      sq += (double)acc * acc;
    }
  }
  su = block_sum_d(su, red);
  sq = block_sum_d(sq, red);
  if (tid == 0) {
    double* p = part + ((long long)b * gridDim.x + blockIdx.x) * 2;
    p[0] = su; p[1] = sq;
  }
}

// CTA = 8 tokens of one window: GroupNorm(1, 512) over (512, n) with the affine folded into the next 1x1 conv
// (:222-228) -> operand planes; z is seeded with the position table (the GEMM accumulates onto it).
static __global__ void __launch_bounds__(256)
encnorm_kernel(const float* __restrict__ xenc, const double* __restrict__ part, int tiles, const float* __restrict__ emb,
               float* __restrict__ phi, float* __restrict__ plo, float* __restrict__ z, int n) {
  __shared__ float stat[2];
  const int b = blockIdx.y, tid = threadIdx.x;
  if (tid == 0) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < tiles; ++i) { a += part[((long long)b * tiles + i) * 2]; q += part[((long long)b * tiles + i) * 2 + 1]; }
    const double cnt = (double)n * D, mean = a / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + 1e-8));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (int i = tid; i < 8 * D / 4; i += 256) {
    const int t = blockIdx.x * 8 + i / (D / 4), c = (i % (D / 4)) * 4;
    if (t >= n) break;
    const long long o = ((long long)b * n + t) * D + c;
    const float4 v = ld4(xenc + o);
    split4(make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd), phi, plo, o);
    st4(z + o, ld4(emb + (long long)t * D + c));
  }
}

// Memory conv 1 (:532-537, j = 0): depthwise k = 39, dilation 1, zero padding 19.  CTA = 64 frames of one window,
// thread = channel (own smem column, so no barrier).  Raw output + the CTA's per-channel sum / sum of squares.
static __global__ void __launch_bounds__(256)
mem1_kernel(const float* __restrict__ xp, const float* __restrict__ taps, float* __restrict__ out,
            float* __restrict__ part, int T) {
  extern __shared__ float msm[];
  float (*tile)[FI] = reinterpret_cast<float (*)[FI]>(msm);
  const int t0 = blockIdx.x * MEM_TT, b = blockIdx.y, c = threadIdx.x;
  const long long base = (long long)b * T;
  for (int r = 0; r < MEM1_ROWS; ++r) {
    const int t = t0 + r - MEMH;
    tile[r][c] = (t >= 0 && t < T) ? __ldg(xp + (base + t) * FI + c) : 0.f;
  }
  float k[MEMK];
#pragma unroll
  for (int i = 0; i < MEMK; ++i) k[i] = __ldg(taps + i * FI + c);
  float su = 0.f, sq = 0.f;
  for (int tt = 0; tt < MEM_TT; tt += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < MEMK + 3; ++i) {
      const float v = tile[tt + i][c];
      if (i < MEMK) a0 += k[i] * v;
      if (i >= 1 && i < MEMK + 1) a1 += k[i - 1] * v;
      if (i >= 2 && i < MEMK + 2) a2 += k[i - 2] * v;
      if (i >= 3) a3 += k[i - 3] * v;
    }
    const float a[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tt + j;
      if (t < T) {
        out[(base + t) * FI + c] = a[j];
        su += a[j];
        sq += a[j] * a[j];
      }
    }
  }
  float* p = part + (((long long)b * gridDim.x + blockIdx.x) * FI + c) * 2;
  p[0] = su; p[1] = sq;
}

// InstanceNorm statistics of one window: fixed-order reduction of the per-CTA partial sums (deterministic).
static __global__ void __launch_bounds__(256)
inorm_stats_kernel(const float* __restrict__ part, int tiles, float* __restrict__ stats, int T) {
  const int b = blockIdx.x, c = threadIdx.x;
  double a = 0.0, q = 0.0;
  for (int i = 0; i < tiles; ++i) {
    const float* p = part + (((long long)b * tiles + i) * FI + c) * 2;
    a += (double)p[0];
    q += (double)p[1];
  }
  const double mean = a / T;
  double var = q / T - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[((long long)b * FI + c) * 2] = (float)mean;
  stats[((long long)b * FI + c) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
}

// Memory conv 2 (:532-537, j = 1): input = cat(PReLU(IN(conv1)), xp) (512 channels), groups = 256 (output c reads
// input channels 2c, 2c+1), k = 39, dilation 2, zero padding 38.  blockIdx.z = 2*half + sub: half 0 -> outputs 0..127
// from the normalised conv-1 branch (normalised while the tile is loaded), half 1 -> outputs 128..255 from xp; sub picks
// 64 of those outputs (128 input columns).  Thread = (output channel, quarter of the 64 frames).
// taps: (2, 39, 256) = [input slot][tap][output channel].
constexpr int MEM2_CH = 64;
static __global__ void __launch_bounds__(256, 2)
mem2_kernel(const float* __restrict__ m1, const float* __restrict__ stats1, const float* __restrict__ nw,
            const float* __restrict__ nb, const float* __restrict__ alpha, const float* __restrict__ xp,
            const float* __restrict__ taps, float* __restrict__ out, float* __restrict__ part, int T) {
  extern __shared__ float msm[];
  float (*tile)[2 * MEM2_CH] = reinterpret_cast<float (*)[2 * MEM2_CH]>(msm);   // [MEM2_ROWS][128 input channels]
  __shared__ float red[4][2][MEM2_CH];
  const int t0 = blockIdx.x * MEM_TT, b = blockIdx.y, half = blockIdx.z >> 1, sub = blockIdx.z & 1, tid = threadIdx.x;
  const long long base = (long long)b * T;
  {
    const int col = tid & (2 * MEM2_CH - 1), ch = sub * 2 * MEM2_CH + col;          // input channel of this branch
    float g = 1.f, sh = 0.f, a = 1.f;
    const float* src = xp;
    if (half == 0) {
      const float mean = stats1[((long long)b * FI + ch) * 2], rstd = stats1[((long long)b * FI + ch) * 2 + 1];
      g = __ldg(nw + ch) * rstd; sh = __ldg(nb + ch) - mean * g; a = __ldg(alpha + ch);
      src = m1;
    }
    for (int r = tid >> 7; r < MEM2_ROWS; r += 2) {
      const int t = t0 + r - 2 * MEMH;
      float v = 0.f;                                                        // zero padding applies AFTER norm + PReLU
      if (t >= 0 && t < T) v = adn_prelu(__ldg(src + (base + t) * FI + ch) * g + sh, a);
      tile[r][col] = v;
    }
  }
  const int j = tid & (MEM2_CH - 1), th = tid >> 6, c = half * 128 + sub * MEM2_CH + j;
  float k0[MEMK], k1[MEMK];
#pragma unroll
  for (int i = 0; i < MEMK; ++i) { k0[i] = __ldg(taps + i * FI + c); k1[i] = __ldg(taps + (MEMK + i) * FI + c); }
  __syncthreads();
  float su = 0.f, sq = 0.f;
  for (int tt = th * (MEM_TT / 4); tt < (th + 1) * (MEM_TT / 4); tt += 8) {
    // 8 outputs per pass: output tt + q reads rows tt + q + 2i, so the four even (odd) outputs share the 42 rows
    // tt (+1) + 2r -- one 64-bit shared load feeds up to 8 FMAs
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.f;
#pragma unroll
    for (int par = 0; par < 2; ++par) {
#pragma unroll
      for (int r = 0; r < MEMK + 3; ++r) {
        const float2 v = *reinterpret_cast<const float2*>(&tile[tt + par + 2 * r][2 * j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = r - e;
          if (i >= 0 && i < MEMK) a[2 * e + par] += k0[i] * v.x + k1[i] * v.y;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = t0 + tt + q;
      if (t < T) {
        out[(base + t) * FI + c] = a[q];
        su += a[q];
        sq += a[q] * a[q];
      }
    }
  }
  red[th][0][j] = su;
  red[th][1][j] = sq;
  __syncthreads();
  if (th == 0) {
    float* p = part + (((long long)b * gridDim.x + blockIdx.x) * FI + c) * 2;
    p[0] = (red[0][0][j] + red[1][0][j]) + (red[2][0][j] + red[3][0][j]);
    p[1] = (red[0][1][j] + red[1][1][j]) + (red[2][1][j] + red[3][1][j]);
  }
}

// One warp per token: mem = PReLU(IN(conv2)); xu += mem; y = xv*xu + g_in; norm2 -> operand planes of conv2.
static __global__ void __launch_bounds__(256)
fsmn_out_kernel(const float* __restrict__ m2, const float* __restrict__ stats2, const float* __restrict__ nw,
                const float* __restrict__ nb, const float* __restrict__ alpha, const float* __restrict__ uv,
                const float* __restrict__ gin, const float* __restrict__ w, const float* __restrict__ bvec,
                float* __restrict__ yhi, float* __restrict__ ylo, long long M, int T) {
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int lane = threadIdx.x & 31;
  const long long b = m / T;
  float v[8];
#pragma unroll
  for (int hq = 0; hq < 2; ++hq) {
    const int c = hq * 128 + lane * 4;
    const float4 r = ld4(m2 + m * FI + c), xu = ld4(uv + m * (2 * FI) + c), xv = ld4(uv + m * (2 * FI) + FI + c);
    const float4 g = ld4(gin + m * FI + c), gw = ld4(nw + c), gb = ld4(nb + c), al = ld4(alpha + c);
    const float4 s0 = ld4(stats2 + (b * FI + c) * 2), s1 = ld4(stats2 + (b * FI + c) * 2 + 4);   // (mean, rstd) x 4
    const float rr[4] = {r.x, r.y, r.z, r.w}, mu[4] = {s0.x, s0.z, s1.x, s1.z}, rs[4] = {s0.y, s0.w, s1.y, s1.w};
    const float gg[4] = {gw.x, gw.y, gw.z, gw.w}, bb[4] = {gb.x, gb.y, gb.z, gb.w}, aa[4] = {al.x, al.y, al.z, al.w};
    const float uu[4] = {xu.x, xu.y, xu.z, xu.w}, vv[4] = {xv.x, xv.y, xv.z, xv.w}, gi[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float mem = adn_prelu((rr[i] - mu[i]) * rs[i] * gg[i] + bb[i], aa[i]);
      v[hq * 4 + i] = vv[i] * (uu[i] + mem) + gi[i];
    }
  }
  float mean, rstd;
  ln256(v, mean, rstd);
  const float4 w0 = ld4(w + lane * 4), w1 = ld4(w + 128 + lane * 4), b0 = ld4(bvec + lane * 4), b1 = ld4(bvec + 128 + lane * 4);
  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
  split4(make_float4(v[0], v[1], v[2], v[3]), yhi, ylo, m * FI + lane * 4);
  split4(make_float4(v[4], v[5], v[6], v[7]), yhi, ylo, m * FI + 128 + lane * 4);
}

// One warp per 2 tokens x 2 speakers: sep = x_enc * mask, then the 16 ConvTranspose1d taps of each frame.  The four
// rows share every decoder-weight load (register tile 4 rows x 16 taps per lane; lanes split the 512 channels).
// dec_w: (512, 16).  fo: (window, speaker, frame, 16).
constexpr int DEC_TOK = 2;
static __global__ void __launch_bounds__(256)
dec_kernel(const float* __restrict__ xenc, const float* __restrict__ mask, const float* __restrict__ dw,
           float* __restrict__ fo, long long M, int n) {
  const long long m0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * DEC_TOK;
  if (m0 >= M) return;
  const int lane = threadIdx.x & 31;
  float acc[DEC_TOK * SPK][ENC_K];
#pragma unroll
  for (int r = 0; r < DEC_TOK * SPK; ++r)
#pragma unroll
    for (int k = 0; k < ENC_K; ++k) acc[r][k] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    float p[DEC_TOK * SPK][4];
#pragma unroll
    for (int tk = 0; tk < DEC_TOK; ++tk) {
      const long long m = m0 + tk < M ? m0 + tk : M - 1;
      const float4 e = ld4(xenc + m * D + c);
#pragma unroll
      for (int sp = 0; sp < SPK; ++sp) {
        const float4 mk = ld4(mask + (m * SPK + sp) * D + c);
        p[tk * SPK + sp][0] = e.x * mk.x; p[tk * SPK + sp][1] = e.y * mk.y;
        p[tk * SPK + sp][2] = e.z * mk.z; p[tk * SPK + sp][3] = e.w * mk.w;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int k4 = 0; k4 < ENC_K / 4; ++k4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(dw + (c + q) * ENC_K) + k4);
#pragma unroll
        for (int r = 0; r < DEC_TOK * SPK; ++r) {
          acc[r][4 * k4] += p[r][q] * wv.x; acc[r][4 * k4 + 1] += p[r][q] * wv.y;
          acc[r][4 * k4 + 2] += p[r][q] * wv.z; acc[r][4 * k4 + 3] += p[r][q] * wv.w;
        }
      }
    }
  }
  // 64 lane-partial sums -> lane l keeps (row l / 16, tap l % 16) of each half: fold the 32 lanes pairwise so that
  // every shuffle halves the number of live values (63 shuffles instead of 64 x 5)
  float v[DEC_TOK * SPK * ENC_K];
#pragma unroll
  for (int r = 0; r < DEC_TOK * SPK; ++r)
#pragma unroll
    for (int k = 0; k < ENC_K; ++k) v[r * ENC_K + k] = acc[r][k];
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int half = (DEC_TOK * SPK * ENC_K) >> (step + 1);          // 32, 16, 8, 4, 2 values kept
    const bool up = (lane >> step) & 1;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float mine = up ? v[j + half] : v[j], other = up ? v[j] : v[j + half];
      v[j] = mine + __shfl_xor_sync(0xffffffffu, other, 1 << step);
    }
  }
  // lane bits (b0..b4) selected value index bit (5 - step): lane l now holds values idx0 = bitrev-style index, 2 values
  int idx = 0;
#pragma unroll
  for (int step = 0; step < 5; ++step) idx |= ((lane >> step) & 1) << (5 - step);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int e = idx + j, r = e / ENC_K, k = e % ENC_K;
    const long long m = m0 + r / SPK;
    if (m < M) {
      const long long b = m / n;
      const int t = (int)(m - b * n);
      fo[((b * SPK + (r % SPK)) * n + t) * ENC_K + k] = v[j];
    }
  }
}

// One CTA per (window, speaker): overlap-add of the frame products (+ bias), RMS of the separated waveform,
// gain = rms_in / rms_out (0 for a silent output, :631), output conversion (:649-657).
__device__ __forceinline__ void ss_store(void* out, long long i, float v, int out_dtype) {
  if (out_dtype == ADN_I16) {
    const int q = max(-32768, min(32767, (int)fminf(fmaxf(v, -2147483648.f), 2147483520.f)));
    reinterpret_cast<int16_t*>(out)[i] = (int16_t)q;
  } else if (out_dtype == ADN_F32) {
    reinterpret_cast<float*>(out)[i] = v * (1.0f / 32768.0f);
  } else {
    reinterpret_cast<__half*>(out)[i] = __float2half_rn(v * (1.0f / 32768.0f));
  }
}

// Output conversion behind the output resampler (:649-657): rows = (window, speaker) of the resampled, gain-restored signal.
static __global__ void __launch_bounds__(256)
ss_convert_kernel(const float* __restrict__ src, void* __restrict__ out0, void* __restrict__ out1, int out_dtype, int len) {
  const int b = blockIdx.y / SPK, s = blockIdx.y % SPK;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= len) return;
  ss_store(s == 0 ? out0 : out1, (long long)b * len + i, src[(long long)blockIdx.y * len + i], out_dtype);
}

// out_dtype < 0: keep the gain-restored fp32 signal in `wav` (int16 scale) for the output resampler instead of converting.
static __global__ void __launch_bounds__(512)
ola_out_kernel(const float* __restrict__ fo, const float* __restrict__ dbias, const float* __restrict__ rms_in,
               float* __restrict__ wav, void* __restrict__ out0, void* __restrict__ out1, int out_dtype, int n, int Lout) {
  __shared__ double red[16];
  const int b = blockIdx.x / SPK, s = blockIdx.x % SPK, tid = threadIdx.x;
  const float* f = fo + (long long)blockIdx.x * n * ENC_K;
  float* wv = wav + (long long)blockIdx.x * Lout;
  const float bias = __ldg(dbias);
  double sq = 0.0;
  for (int i = tid; i < Lout; i += 512) {
    const int t = i / ENC_S, j = i - t * ENC_S;
    float v = 0.f;
    if (t < n) v += f[t * ENC_K + j];
    if (t >= 1) v += f[(t - 1) * ENC_K + ENC_S + j];
    v += bias;
    wv[i] = v;
    sq += (double)(v * v);
  }
  sq = block_sum_d(sq, red);
  const float rms_out = sqrtf((float)(sq / (double)Lout));
  const float g = rms_out > 0.f ? __ldg(rms_in + b) / rms_out : 0.f;
  void* out = s == 0 ? out0 : out1;
  const long long ob = (long long)b * Lout;
  for (int i = tid; i < Lout; i += 512) {
    const float v = wv[i] * g;
    if (out_dtype < 0) wv[i] = v;
    else ss_store(out, ob + i, v, out_dtype);
  }
}

class SsModel : public Base {
 public:
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, Lout = 0, T = 0, G = 1, Tg = 256, layers = 24;
  int L_in = 0, L_final = 0;      // window length at the input rate / output length at the output rate (== L, Lout at 16 kHz)
  bool rs_in = false, rs_out = false;
  float *xr = nullptr, *wres = nullptr;
  int enc_tiles = 0, mem_tiles = 0;

  const float *enc_w = nullptr, *enc_b = nullptr, *dec_w = nullptr, *dec_b = nullptr, *front_b = nullptr, *emb = nullptr;
  const float *rcos = nullptr, *rsin = nullptr, *mm_w = nullptr, *mm_b = nullptr, *in_w = nullptr, *in_b = nullptr;
  const float *prelu_a = nullptr, *gate_b = nullptr;
  struct Layer {
    Lin in, out, c1, uv, ul, up, c2;
    const float *in_b, *in_c, *gamma, *beta, *out_b, *out_c, *c1_b, *c1_a, *n1_w, *n1_b, *uv_b, *uv_c, *ul_b;
    const float *mem0_c, *mem0_nw, *mem0_nb, *mem0_a, *mem1_c, *mem1_nw, *mem1_nb, *mem1_a, *n2_w, *n2_b, *c2_b;
  };
  std::vector<Layer> lw;
  Lin front, gate, maskl;

  float *xn = nullptr, *rms_in = nullptr, *xenc = nullptr, *z = nullptr, *h = nullptr, *xs = nullptr, *rs = nullptr;
  float *proj = nullptr, *vu = nullptr, *vuT = nullptr, *qq = nullptr, *qk = nullptr, *lkT = nullptr;
  float *spl = nullptr, *kvT = nullptr, *att = nullptr, *gated = nullptr, *rs2 = nullptr, *y = nullptr, *hpl = nullptr;
  float *c1y = nullptr, *gin = nullptr, *xnp = nullptr, *uvp = nullptr, *uv = nullptr, *xupl = nullptr, *f1 = nullptr;
  float *xp2 = nullptr, *m1 = nullptr, *m2 = nullptr, *part = nullptr, *stats = nullptr, *yn = nullptr, *hn = nullptr;
  float *tpl = nullptr, *gbuf = nullptr, *tg = nullptr, *mask = nullptr, *fo = nullptr, *wav = nullptr;
  double* encpart = nullptr;
  Lin a_qk, a_vuT, a_lkT, a_kvT;
  Gemm g_front, g_qk, g_pv, g_kv, g_gate, g_mask;
  struct LayerG { Gemm in, out, c1, uv, ul, up, c2; };
  std::vector<LayerG> lg;
  int stop_after = 0, last_batch = 0;

  ~SsModel() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    auto fl = [](Lin& l) { if (l.planes) cudaFree(l.planes); };
    for (auto& w : lw) { fl(w.in); fl(w.out); fl(w.c1); fl(w.uv); fl(w.ul); fl(w.up); fl(w.c2); }
    fl(front); fl(gate); fl(maskl);
  }

  bool init(const std::map<std::string, std::string>& meta) {
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
    int stride = 0, sources = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", L) || !geti("enc_stride", stride) || !geti("output_sources", sources) ||
        !geti("mf2_layers", layers) || !gets("input_audio_dtype", sin) || !gets("output_audio_dtype", sout))
      return false;
    // optional linear resampling either side of the model (:564-579, :633-648): input_audio_length is at in_sample_rate
    int in_sr = 16000, out_sr = 16000, model_sr = 16000;
    {
      auto opt = [&](const char* k, int& v) { auto it = meta.find(k); if (it != meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
      opt("in_sample_rate", in_sr); opt("out_sample_rate", out_sr); opt("model_sample_rate", model_sr);
    }
    if (model_sr != 16000 || in_sr <= 0 || out_sr <= 0) { err = "mossformer2_ss runs at model_sample_rate 16000"; return false; }
    L_in = L;
    rs_in = in_sr != model_sr;
    rs_out = out_sr != model_sr;
    if (rs_in) L = (int)llround((double)L_in * model_sr / in_sr);          // MODEL_AUDIO_LENGTH (:36)
    if (stride != ENC_S || sources != SPK || L < ENC_K) {
      err = "mossformer2_ss needs enc_stride=8, output_sources=2 and a window of at least 16 model-rate samples";
      return false;
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = (L - ENC_K) / ENC_S + 1;
    Lout = (T - 1) * ENC_S + ENC_K;
    L_final = rs_out ? (int)llround((double)L_in * out_sr / in_sr) : Lout;   // OUTPUT_AUDIO_LENGTH (:37)
    G = (T + GROUP - 1) / GROUP;
    Tg = G * GROUP;
    enc_tiles = (T + ENC_TT - 1) / ENC_TT;
    mem_tiles = (T + MEM_TT - 1) / MEM_TT;

    bool ok = true;
    enc_w = dptr("enc_w", (size_t)D * ENC_K, ok); enc_b = dptr("enc_b", D, ok);
    dec_w = dptr("dec_w", (size_t)D * ENC_K, ok); dec_b = dptr("dec_b", 1, ok);
    front_b = dptr("front_b", D, ok);
    emb = dptr("emb_pos", (size_t)T * D, ok);
    rcos = dptr("rot_cos", (size_t)T * ROT, ok); rsin = dptr("rot_sin", (size_t)T * ROT, ok);
    mm_w = dptr("mm_norm.w", D, ok); mm_b = dptr("mm_norm.b", D, ok);
    in_w = dptr("intra_norm.w", D, ok); in_b = dptr("intra_norm.b", D, ok);
    prelu_a = dptr("prelu_a", 1, ok);
    gate_b = dptr("gate_b", (size_t)SPK * 2 * D, ok);
    if (!ok) return false;
    const float* w;
    w = dptr("front_w", (size_t)D * D, ok);              if (ok && !make_lin(front, w, D, D)) return false;
    w = dptr("gate_w", (size_t)SPK * 2 * D * D, ok);     if (ok && !make_lin(gate, w, SPK * 2 * D, D)) return false;
    w = dptr("mask_w", (size_t)D * D, ok);               if (ok && !make_lin(maskl, w, D, D)) return false;
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
      Y.ul_b = dptr(p + "ul_b", FI, ok);
      Y.mem0_c = dptr(p + "mem0_c", (size_t)MEMK * FI, ok); Y.mem1_c = dptr(p + "mem1_c", (size_t)2 * MEMK * FI, ok);
      Y.mem0_nw = dptr(p + "mem0_nw", FI, ok); Y.mem0_nb = dptr(p + "mem0_nb", FI, ok); Y.mem0_a = dptr(p + "mem0_a", FI, ok);
      Y.mem1_nw = dptr(p + "mem1_nw", FI, ok); Y.mem1_nb = dptr(p + "mem1_nb", FI, ok); Y.mem1_a = dptr(p + "mem1_a", FI, ok);
      Y.n2_w = dptr(p + "n2_w", FI, ok); Y.n2_b = dptr(p + "n2_b", FI, ok);
      Y.c2_b = dptr(p + "c2_b", D, ok);
    }
    if (!ok) return false;
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "weight split failed"; return false; }
    return true;
  }

  size_t floats_needed(size_t B) const {
    const size_t M = B * T, Mg = B * Tg;
    return B * L + B + (rs_in ? B * L : 0) + (rs_out ? B * SPK * L_final : 0) + 3 * M * D + 2 * M * D + M + M * PROJ + M * VU2 + 2 * B * VU2 * Tg + 4 * Mg * QK + 2 * B * QK * Tg +
           2 * Mg * SQ + 2 * B * VU2 * QK + Mg * VU2 + 2 * M * VU + M + M * D + 2 * M * D + 2 * M * FI + 2 * M * FI +
           2 * M * 2 * FI + 4 * M * FI + 3 * M * FI + 2 * B * mem_tiles * FI * 2 + 2 * B * FI * 2 + 2 * M * FI + M * D +
           2 * M * D + M * SPK * 2 * D + 2 * M * SPK * D + M * SPK * D + M * SPK * ENC_K + B * SPK * Lout + B * enc_tiles * 4;
  }

  bool ensure(int B) {
    if (B == planned) return true;
    cudaDeviceSynchronize();
    free_ws();
    const size_t M = (size_t)B * T, Mg = (size_t)B * Tg;
    float* ep = nullptr;
    if ((rs_in && !alloc(xr, (size_t)B * L, false)) || (rs_out && !alloc(wres, (size_t)B * SPK * L_final, false))) return false;
    if (!alloc(xn, (size_t)B * L, false) || !alloc(rms_in, B, false) || !alloc(xenc, M * D, false) || !alloc(z, M * D, false) ||
        !alloc(h, M * D, false) || !alloc(xs, 2 * M * D, false) || !alloc(rs, M, false) || !alloc(proj, M * PROJ, false) ||
        !alloc(vu, M * VU2, false) || !alloc(vuT, 2 * (size_t)B * VU2 * Tg, true) || !alloc(qq, 2 * Mg * QK, true) ||
        !alloc(qk, 2 * Mg * QK, true) || !alloc(lkT, 2 * (size_t)B * QK * Tg, true) ||
        !alloc(spl, 2 * Mg * SQ, true) || !alloc(kvT, 2 * (size_t)B * VU2 * QK, false) || !alloc(att, Mg * VU2, false) ||
        !alloc(gated, 2 * M * VU, false) || !alloc(rs2, M, false) || !alloc(y, M * D, false) || !alloc(hpl, 2 * M * D, false) ||
        !alloc(c1y, M * FI, false) || !alloc(gin, M * FI, false) || !alloc(xnp, 2 * M * FI, false) ||
        !alloc(uvp, M * 2 * FI, false) || !alloc(uv, M * 2 * FI, false) || !alloc(xupl, 2 * M * FI, false) ||
        !alloc(f1, 2 * M * FI, false) || !alloc(xp2, M * FI, false) || !alloc(m1, M * FI, false) || !alloc(m2, M * FI, false) ||
        !alloc(part, 2 * (size_t)B * mem_tiles * FI * 2, false) || !alloc(stats, 2 * (size_t)B * FI * 2, false) ||
        !alloc(yn, 2 * M * FI, false) || !alloc(hn, M * D, false) || !alloc(tpl, 2 * M * D, false) ||
        !alloc(gbuf, M * SPK * 2 * D, false) || !alloc(tg, 2 * M * SPK * D, false) || !alloc(mask, M * SPK * D, false) ||
        !alloc(fo, M * SPK * ENC_K, false) || !alloc(wav, (size_t)B * SPK * Lout, false) ||
        !alloc(ep, (size_t)B * enc_tiles * 4, false))
      return false;
    encpart = reinterpret_cast<double*>(ep);

    // folded front 1x1 conv: accumulates onto the position table seeded in z
    if (!plan_gemm(g_front, xs, (long long)(M * D), D, (int)M, D, 1, (long long)(M * D), front)) return false;
    g_front.args.bias = front_b; g_front.args.resid = z; g_front.args.C = z; g_front.args.ldc = D;

    // attention operands that live in activations
    const int BG = B * G;
    if (!make_act_lin(a_qk, qk, qk + Mg * QK, GROUP, GROUP, QK, QK, 256, BG) ||
        !make_act_lin(a_vuT, vuT, vuT + (size_t)B * VU2 * Tg, VU2, VU2, GROUP, Tg, 256, B) ||
        !make_act_lin(a_lkT, lkT, lkT + (size_t)B * QK * Tg, QK, QK, Tg, Tg, 128, B) ||
        !make_act_lin(a_kvT, kvT, kvT + (size_t)B * VU2 * QK, VU2, VU2, QK, QK, 256, B))
      return false;
    if (!plan_gemm(g_qk, qq, (long long)(Mg * QK), QK, GROUP, QK, BG, (long long)GROUP * QK, a_qk)) return false;
    g_qk.args.w_batched = 1; g_qk.args.act = tc::ACT_RELU2; g_qk.args.Chi = spl; g_qk.args.Clo = spl + Mg * SQ; g_qk.args.ldc = SQ;
    if (!plan_gemm(g_pv, spl, (long long)(Mg * SQ), SQ, GROUP, SQ, BG, (long long)GROUP * SQ, a_vuT)) return false;
    g_pv.args.w_batched = 1; g_pv.args.w_group = G; g_pv.args.w_kstep = GROUP; g_pv.args.C = att; g_pv.args.ldc = VU2;
    g_pv.args.K = SQ; g_pv.args.k_split = GROUP;
    g_pv.plan.map_w2_hi = a_kvT.w_hi; g_pv.plan.map_w2_lo = a_kvT.w_lo;
    if (!plan_gemm(g_kv, vuT, (long long)B * VU2 * Tg, Tg, VU2, Tg, B, (long long)VU2 * Tg, a_lkT)) return false;
    g_kv.args.w_batched = 1; g_kv.args.Chi = kvT; g_kv.args.Clo = kvT + (size_t)B * VU2 * QK; g_kv.args.ldc = QK;

    lg.assign(layers, LayerG{});
    const long long Ml = (long long)M;
    for (int i = 0; i < layers; ++i) {
      LayerG& Gm = lg[i];
      const Layer& Y = lw[i];
      if (!plan_gemm(Gm.in, xs, Ml * D, D, (int)M, D, 1, Ml * D, Y.in)) return false;
      Gm.in.args.rowscale = rs; Gm.in.args.bias = Y.in_b; Gm.in.args.act = tc::ACT_SILU; Gm.in.args.C = proj; Gm.in.args.ldc = PROJ;
      if (!plan_gemm(Gm.out, gated, Ml * VU, VU, (int)M, VU, 1, Ml * VU, Y.out)) return false;
      Gm.out.args.rowscale = rs2; Gm.out.args.bias = Y.out_b; Gm.out.args.act = tc::ACT_SILU; Gm.out.args.C = y; Gm.out.args.ldc = D;
      if (!plan_gemm(Gm.c1, hpl, Ml * D, D, (int)M, D, 1, Ml * D, Y.c1)) return false;
      Gm.c1.args.bias = Y.c1_b; Gm.c1.args.act = tc::ACT_PRELU; Gm.c1.args.act_param = Y.c1_a; Gm.c1.args.C = c1y; Gm.c1.args.ldc = FI;
      if (!plan_gemm(Gm.uv, xnp, Ml * FI, FI, (int)M, FI, 1, Ml * FI, Y.uv)) return false;
      Gm.uv.args.bias = Y.uv_b; Gm.uv.args.act = tc::ACT_SILU; Gm.uv.args.C = uvp; Gm.uv.args.ldc = 2 * FI;
      if (!plan_gemm(Gm.ul, xupl, Ml * FI, FI, (int)M, FI, 1, Ml * FI, Y.ul)) return false;
      Gm.ul.args.bias = Y.ul_b; Gm.ul.args.act = tc::ACT_RELU; Gm.ul.args.Chi = f1; Gm.ul.args.Clo = f1 + M * FI; Gm.ul.args.ldc = FI;
      if (!plan_gemm(Gm.up, f1, Ml * FI, FI, (int)M, FI, 1, Ml * FI, Y.up)) return false;
      Gm.up.args.C = xp2; Gm.up.args.ldc = FI;
      if (!plan_gemm(Gm.c2, yn, Ml * FI, FI, (int)M, FI, 1, Ml * FI, Y.c2)) return false;
      Gm.c2.args.bias = Y.c2_b; Gm.c2.args.resid = h; Gm.c2.args.C = h; Gm.c2.args.ldc = D;
    }
    if (!plan_gemm(g_gate, tpl, Ml * D, D, (int)M, D, 1, Ml * D, gate)) return false;
    g_gate.args.bias = gate_b; g_gate.args.C = gbuf; g_gate.args.ldc = SPK * 2 * D;
    // rows = (token, speaker): the speaker-stacked gate output viewed as (M*SPK, 1024)
    if (!plan_gemm(g_mask, tg, Ml * SPK * D, D, (int)(M * SPK), D, 1, Ml * SPK * D, maskl)) return false;
    g_mask.args.act = tc::ACT_RELU; g_mask.args.C = mask; g_mask.args.ldc = D;
    // workspace memsets ran on the legacy default stream; runs use a non-blocking stream that does not order against it
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "workspace initialisation failed"; return false; }
    planned = B;
    return true;
  }

  // ---- ModelImpl
  int n_outputs() override { return SPK; }
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    strncpy(in->name, "mix_audio", sizeof(in->name) - 1);              // Export_MossFormer2_SS_16K.py:689
    in->dtype = in_dtype; in->channels = 1; in->length = L_in;
    for (int s = 0; s < SPK; ++s) {
      memset(out + s, 0, sizeof(*out));
      snprintf(out[s].name, sizeof(out[s].name), "separated_%d", s);   // :690
      out[s].dtype = out_dtype; out[s].channels = 1; out[s].length = L_final;
    }
  }
  size_t workspace_bytes(int batch) override { return floats_needed((size_t)batch) * sizeof(float); }
  int launches(int) override { return 4 + layers * 21 + 6 + (rs_in ? 1 : 0) + (rs_out ? 2 : 0); }
  void set_stop_after(int n) override { stop_after = n; }

#define SS_TICK(name) do { ++n; if (tick) tick(tick_ctx, name); if (stop_after > 0 && n >= stop_after) return ADN_OK; } while (0)
#define SS_GEMM(Gx, name) do { cudaError_t e_ = tc::launch((Gx).plan, (Gx).args, EPI_LIN, sms, st); \
    if (e_ != cudaSuccess) { err = std::string("gemm launch (") + name + "): " + cudaGetErrorString(e_); return ADN_ERR_CUDA; } SS_TICK(name); } while (0)

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    (void)d_in; (void)d_out; (void)B; (void)st;
    err = "mossformer2_ss has two outputs: use the d_outs array";
    return ADN_ERR_INVALID;
  }

  adn_status run_multi(const void* d_in, void* const* d_outs, int B, cudaStream_t st) override {
    if (!d_outs[0] || !d_outs[1]) { err = "mossformer2_ss: two output buffers are required"; return ADN_ERR_INVALID; }
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    int n = 0;
    const long long M = (long long)B * T, Mg = (long long)B * Tg;
    const unsigned wtok = (unsigned)((M + 7) / 8);
    static unsigned long long cfg = 0;             // per device
    if (adn_first_use_on_device(cfg)) {
      cudaFuncSetAttribute(mem1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEM1_ROWS * FI * 4);
      cudaFuncSetAttribute(mem2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEM2_ROWS * 2 * MEM2_CH * 4);
    }

    const void* src = d_in;
    int src_dtype = in_dtype;
    if (rs_in) {                                   // F.interpolate(size=MODEL_AUDIO_LENGTH, mode='linear') (:564-579)
      if (adn_resample_linear(d_in, in_dtype, xr, B, L_in, L, 0.0, st) != ADN_OK) { err = "input resampler launch failed"; return ADN_ERR_CUDA; }
      SS_TICK("resample_in");
      src = xr;
      src_dtype = ADN_F32;
    }
    if (adn_two_stage_rms(src, src_dtype, 0.05623413251903491f, 1e-6f, xn, rms_in, B, L, st) != ADN_OK) {
      err = "two-stage RMS launch failed";
      return ADN_ERR_CUDA;
    }
    SS_TICK("norm_audio");
    enc_kernel<<<dim3(enc_tiles, B), 256, 0, st>>>(xn, enc_w, enc_b, xenc, encpart, L, T);
    SS_TICK("encoder");
    encnorm_kernel<<<dim3((T + 7) / 8, B), 256, 0, st>>>(xenc, encpart, enc_tiles, emb, xs, xs + M * D, z, T);
    SS_TICK("encnorm");
    SS_GEMM(g_front, "front_gemm");

    for (int i = 0; i < layers; ++i) {
      LayerG& Gm = lg[i];
      const Layer& Y = lw[i];
      const float* hin = i == 0 ? z : h;
      shiftnorm_kernel<<<wtok, 256, 0, st>>>(hin, xs, xs + M * D, rs, M, T, 1);
      SS_TICK("shiftnorm");
      SS_GEMM(Gm.in, "fl_in");
      // (the TMA-fed tile form, dwconv_in_tma_kernel, wins for one-group windows -- MossFormer2-SE 7.96 -> 6.84 ms -- but
      // loses here, 1.31 -> 1.51 ms: with Tg = 2048 every 256-byte piece of [v|u]^T lands in its own 8 KB-strided row)
      dwconv_in_kernel<<<dim3(PROJ / 32 / DWI_WARPS, B, (T + DW_SEG - 1) / DW_SEG), DWI_WARPS * 32, 0, st>>>(
          proj, Y.in_c, Y.gamma, Y.beta, rcos, rsin, vu, vuT, vuT + (size_t)B * VU2 * Tg, qq, qq + Mg * QK, spl + GROUP, spl + Mg * SQ + GROUP,
          qk, qk + Mg * QK, nullptr, nullptr, lkT, lkT + (size_t)B * QK * Tg, T, Tg, Tg, Tg, SQ);
      SS_TICK("dwconv_in");
      SS_GEMM(g_qk, "att_qk");
      SS_GEMM(g_kv, "att_kv");
      SS_GEMM(g_pv, "att_pv");
      gate_kernel<<<wtok, 256, 0, st>>>(att, vu, gated, gated + M * VU, rs2, M, T, Tg, 1);
      SS_TICK("gate");
      SS_GEMM(Gm.out, "fl_out");
      dwconv_kernel<<<dim3(D / 32 / DWI_WARPS, B, (T + DW_SEG - 1) / DW_SEG), DWI_WARPS * 32, 0, st>>>(y, Y.out_c, hin, h, hpl, hpl + M * D, D, T, D);
      SS_TICK("dwconv_out");
      SS_GEMM(Gm.c1, "fsmn_conv1");
      ln2_kernel<<<wtok, 256, 0, st>>>(c1y, Y.n1_w, Y.n1_b, gin, xnp, xnp + M * FI, M);
      SS_TICK("ln2");
      SS_GEMM(Gm.uv, "fsmn_uv");
      dwconv_kernel<<<dim3(2 * FI / 32 / DWI_WARPS, B, (T + DW_SEG - 1) / DW_SEG), DWI_WARPS * 32, 0, st>>>(uvp, Y.uv_c, nullptr, uv, xupl, xupl + M * FI, FI, T, 2 * FI);
      SS_TICK("dwconv_uv");
      SS_GEMM(Gm.ul, "fsmn_linear");
      SS_GEMM(Gm.up, "fsmn_project");
      float* part2 = part + (size_t)B * mem_tiles * FI * 2;
      float* stats2 = stats + (size_t)B * FI * 2;
      mem1_kernel<<<dim3(mem_tiles, B), 256, MEM1_ROWS * FI * sizeof(float), st>>>(xp2, Y.mem0_c, m1, part, T);
      SS_TICK("fsmn_mem1");
      inorm_stats_kernel<<<B, 256, 0, st>>>(part, mem_tiles, stats, T);
      SS_TICK("fsmn_stats1");
      mem2_kernel<<<dim3(mem_tiles, B, 4), 256, MEM2_ROWS * 2 * MEM2_CH * sizeof(float), st>>>(m1, stats, Y.mem0_nw, Y.mem0_nb, Y.mem0_a, xp2,
                                                                                    Y.mem1_c, m2, part2, T);
      SS_TICK("fsmn_mem2");
      inorm_stats_kernel<<<B, 256, 0, st>>>(part2, mem_tiles, stats2, T);
      SS_TICK("fsmn_stats2");
      fsmn_out_kernel<<<wtok, 256, 0, st>>>(m2, stats2, Y.mem1_nw, Y.mem1_nb, Y.mem1_a, uv, gin, Y.n2_w, Y.n2_b, yn, yn + M * FI, M, T);
      SS_TICK("fsmn_out");
      SS_GEMM(Gm.c2, "fsmn_conv2");
    }

    tail_norm_kernel<<<B, 512, 0, st>>>(layers ? h : z, z, mm_w, mm_b, in_w, in_b, prelu_a, hn, tpl, tpl + M * D, T);
    SS_TICK("tail_norm");
    SS_GEMM(g_gate, "tail_gate_gemm");
    tail_gate_kernel<<<(unsigned)((M * SPK * (D / 4) + 255) / 256), 256, 0, st>>>(gbuf, tg, tg + M * SPK * D, M * SPK);
    SS_TICK("tail_gate");
    SS_GEMM(g_mask, "mask_gemm");
    dec_kernel<<<(unsigned)((M + 8 * DEC_TOK - 1) / (8 * DEC_TOK)), 256, 0, st>>>(xenc, mask, dec_w, fo, M, T);
    SS_TICK("decoder");
    ola_out_kernel<<<B * SPK, 512, 0, st>>>(fo, dec_b, rms_in, wav, d_outs[0], d_outs[1], rs_out ? -1 : out_dtype, T, Lout);
    SS_TICK("ola_out");
    if (rs_out) {                                  // F.interpolate(size=OUTPUT_AUDIO_LENGTH) on the gain-restored rows (:633-648)
      if (adn_resample_linear(wav, ADN_F32, wres, B * SPK, Lout, L_final, 0.0, st) != ADN_OK) { err = "output resampler launch failed"; return ADN_ERR_CUDA; }
      SS_TICK("resample_out");
      ss_convert_kernel<<<dim3((L_final + 255) / 256, B * SPK), 256, 0, st>>>(wres, d_outs[0], d_outs[1], out_dtype, L_final);
      SS_TICK("convert_out");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mf2ss run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    const size_t B = last_batch;
    if (!B) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    const size_t M = B * T, Mg = B * Tg;
    std::map<std::string, std::pair<const float*, size_t>> tbl = {
        {"x", {xn, B * L}}, {"x_enc", {xenc, M * D}}, {"z", {z, M * D}}, {"h", {h, M * D}}, {"proj", {proj, M * PROJ}},
        {"vu", {vu, M * VU2}}, {"att", {att, Mg * VU2}}, {"y", {y, M * D}}, {"c1y", {c1y, M * FI}}, {"gin", {gin, M * FI}},
        {"uv", {uv, M * 2 * FI}}, {"xp2", {xp2, M * FI}}, {"m1", {m1, M * FI}}, {"m2", {m2, M * FI}}, {"gate", {gbuf, M * SPK * 2 * D}},
        {"mask", {mask, M * SPK * D}}, {"wav", {wav, B * SPK * Lout}}, {"rs", {rs, M}}, {"rs2", {rs2, M}}, {"rms_in", {rms_in, B}},
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

ModelImpl* mf2ss_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  (void)h_blob;
  mf2::SsModel* m = new mf2::SsModel();
  m->device = device;
  m->sms = sms;
  m->d_blob = d_blob;
  m->index = index;
  if (!m->init(meta)) {
    err = m->err;
    delete m;
    return nullptr;
  }
  return m;
}
