// GTCRN backbone kernels (fp32, CUDA cores; channel counts are 2..24 so there is no
// GEMM-shaped work here -- the dense DFT GEMMs live in gemm_*.cu).
//
// Layout: every activation is frame-major (B, T, C, F): one frame's (C,F) block is
// contiguous, so per-frame kernels stage it with fully coalesced loads and the causal
// time-dilated taps of the GTConv blocks fetch whole contiguous frames.
//
// Reference map (GTCRN/Export_GTCRN.py):
//   enc_front_kernel  : forward_packed magnitude :592-596, ERB.bm :99-102, SFE :131-141,
//                       en_convs.0/.1 :488-489,499-501
//   gt_main_kernel    : GTConvBlock.forward :303-320 (SFE, point_conv1, depth_conv, point_conv2)
//   tra_apply_kernel  : TRA :152-156 + channel shuffle :324 (+ decoder skip add :524-528)
//   dp_intra_kernel   : DPGRNN intra path :471-475 (+ previous inter LayerNorm :481)
//   dp_inter_kernel   : DPGRNN inter path :478-480
//   ln_res_kernel     : final inter LayerNorm + residual :481 (+ skip add :524)
//   dec_tail_kernel   : de_convs.3/.4 :515-516,527-528, ERB.bs :104-107, complex mask :585-590
#include "adn.h"
#include "gtcrn.cuh"

namespace gtcrn {

// =================================================================================
// prep: cast + scale + DC removal + centre padding.  One CTA per chunk.
// =================================================================================
template <typename Tin>
__device__ __forceinline__ float load_sample(const Tin* p, long long i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, long long i) {
  return (float)p[i] * (1.0f / 32768.0f);   // Export_GTCRN.py:645-646
}
template <>
__device__ __forceinline__ float load_sample<__half>(const __half* p, long long i) {
  return __half2float(p[i]);
}

template <typename Tin>
__global__ void __launch_bounds__(256) prep_kernel(const Tin* __restrict__ in, float* __restrict__ xp,
                                                   float* __restrict__ hi, float* __restrict__ lo,
                                                   int L, int Lp, int half, int remove_dc, int reflect) {
  const int b = blockIdx.x;
  const Tin* x = in + (long long)b * L;
  float* o = xp + (long long)b * Lp;
  __shared__ double red[8];
  __shared__ float mean_s;
  float mean = 0.f;
  if (remove_dc) {
    double s = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) s += (double)load_sample<Tin>(x, i);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      mean_s = (float)(t / (double)L);
    }
    __syncthreads();
    mean = mean_s;
  }
  for (int i = threadIdx.x; i < Lp; i += blockDim.x) {
    int j = i - half;        // index into the unpadded signal
    float v = 0.f;
    if (j >= 0 && j < L) {
      v = load_sample<Tin>(x, j) - mean;
    } else if (reflect) {
      int jj = j < 0 ? -j : 2 * (L - 1) - j;   // reflect without repeating the edge sample
      if (jj >= 0 && jj < L) v = load_sample<Tin>(x, jj) - mean;
    }
    o[i] = v;
    if (hi) split_tf32_store(v, hi, lo, (long long)b * Lp + i);
  }
}

void launch_prep(const void* in, int in_dtype, float* xp, float* hi, float* lo, int B, int L, int Lp, int half,
                 int remove_dc, int reflect, cudaStream_t st) {
  if (in_dtype == ADN_F32)
    prep_kernel<float><<<B, 256, 0, st>>>((const float*)in, xp, hi, lo, L, Lp, half, remove_dc, reflect);
  else if (in_dtype == ADN_I16)
    prep_kernel<int16_t><<<B, 256, 0, st>>>((const int16_t*)in, xp, hi, lo, L, Lp, half, remove_dc, reflect);
  else
    prep_kernel<__half><<<B, 256, 0, st>>>((const __half*)in, xp, hi, lo, L, Lp, half, remove_dc, reflect);
}

// =================================================================================
// enc_front: spectrum frame -> |X|, ERB, SFE, en_convs.0, en_convs.1.  Kernel height is 1
// for all of these, so frames are independent.  FR frames per CTA.
// =================================================================================
constexpr int EF_FR = 4;
constexpr int EF_THREADS = 288;

__global__ void __launch_bounds__(EF_THREADS)
enc_front_kernel(const __grid_constant__ EncFrontW w, const ErbW erb, const float* __restrict__ spec,
                 float* __restrict__ e0, float* __restrict__ e1, int nframes) {
  __shared__ float sp[EF_FR][SPEC_LD];            // packed spectrum frame
  __shared__ float fe[EF_FR][3][ERB_F + 8];       // ERB features, 4 zeros each side
  __shared__ float e0s[EF_FR][16][E0_F + 4];      // en_convs.0 output, 2 zeros each side

  const int tid = threadIdx.x;
  const long long f0 = (long long)blockIdx.x * EF_FR;

  for (int i = tid; i < EF_FR * SPEC_LD; i += EF_THREADS) {
    int fr = i / SPEC_LD, c = i - fr * SPEC_LD;
    long long fg = f0 + fr;
    sp[fr][c] = (fg < nframes) ? __ldg(spec + fg * SPEC_LD + c) : 0.f;
  }
  for (int i = tid; i < EF_FR * 3 * (ERB_F + 8); i += EF_THREADS) (&fe[0][0][0])[i] = 0.f;
  for (int i = tid; i < EF_FR * 16 * (E0_F + 4); i += EF_THREADS) (&e0s[0][0][0])[i] = 0.f;
  __syncthreads();

  // low 65 bins pass through, |X| = sqrt(re^2+im^2+1e-12) (:594-595)
  for (int i = tid; i < EF_FR * 65; i += EF_THREADS) {
    int fr = i / 65, f = i - fr * 65;
    float re = sp[fr][f], im = sp[fr][FB + f];
    fe[fr][0][4 + f] = sqrtf(re * re + im * im + 1e-12f);
    fe[fr][1][4 + f] = re;
    fe[fr][2][4 + f] = im;
  }
  // 192 high bins -> 64 ERB bands, ascending-bin accumulation over the nonzero range
  for (int i = tid; i < EF_FR * 3 * 64; i += EF_THREADS) {
    int fr = i / 192, r = i - fr * 192;
    int c = r >> 6, j = r & 63;
    int lo = (int)__ldg(erb.bm_lo + j), hi = (int)__ldg(erb.bm_hi + j);
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) {
      int f = 65 + k;
      float v;
      if (c == 0) {
        float re = sp[fr][f], im = sp[fr][FB + f];
        v = sqrtf(re * re + im * im + 1e-12f);
      } else {
        v = sp[fr][(c == 1 ? 0 : FB) + f];
      }
      acc = fmaf(v, __ldg(erb.bm + k * 64 + j), acc);
    }
    fe[fr][c][4 + 65 + j] = acc;
  }
  __syncthreads();

  // en_convs.0: Conv2d(9->16,(1,5),stride 2,pad 2) over SFE(k=3) of the 3 ERB channels.
  // SFE channel c*3+s at position p is fe[c][p+s-1]; positions p outside [0,129) are the
  // conv's own zero padding.
  for (int i = tid; i < EF_FR * E0_F; i += EF_THREADS) {
    int fr = i / E0_F, g = i - fr * E0_F;
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = w.b0[o];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      int p = 2 * g + k - 2;
      bool pv = (p >= 0) && (p < ERB_F);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          float v = pv ? fe[fr][c][4 + p + s - 1] : 0.f;
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
    int fr = i / FRAME_E0, r = i - fr * FRAME_E0;
    int o = r / E0_F, g = r - o * E0_F;
    long long fg = f0 + fr;
    if (fg < nframes) e0[fg * FRAME_E0 + r] = e0s[fr][o][2 + g];
  }

  // en_convs.1: Conv2d(16->16,(1,5),stride 2,pad 2,groups 2).  item = (frame, group, g)
  for (int i = tid; i < EF_FR * 2 * E1_F; i += EF_THREADS) {
    int fr = i / (2 * E1_F), r = i - fr * (2 * E1_F);
    int grp = r / E1_F, g = r - grp * E1_F;
    float acc[8];
    if (grp == 0) {
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = w.b1[o];
#pragma unroll
      for (int ci = 0; ci < 8; ++ci)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          float v = e0s[fr][ci][2 * g + k];
#pragma unroll
          for (int o = 0; o < 8; ++o) acc[o] = fmaf(w.w1[0][ci][k][o], v, acc[o]);
        }
    } else {
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = w.b1[8 + o];
#pragma unroll
      for (int ci = 0; ci < 8; ++ci)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          float v = e0s[fr][8 + ci][2 * g + k];
#pragma unroll
          for (int o = 0; o < 8; ++o) acc[o] = fmaf(w.w1[1][ci][k][o], v, acc[o]);
        }
    }
    long long fg = f0 + fr;
    if (fg < nframes) {
#pragma unroll
      for (int o = 0; o < 8; ++o) e1[fg * FRAME16 + (grp * 8 + o) * E1_F + g] = adn_prelu(acc[o], w.a1);
    }
  }
}

// =================================================================================
// gt_main: GTConvBlock up to point_conv2 (+ the TRA energy z_t).  Frames with the same
// residue t mod d form an undilated causal sequence, so a CTA owns KT consecutive steps of
// one residue class (+2 recomputed halo frames) of one chunk.  Each thread computes TWO
// frames at one frequency bin, so every weight fetched from the constant bank feeds two
// FFMAs and the depthwise taps of adjacent steps share their shared-memory loads.
// =================================================================================
template <int KT>
struct GtCfg {
  static constexpr int FL = KT + 2;                       // frames in the tile (2 halo)
  static constexpr int P1 = FL / 2, P2 = KT / 2;          // frame pairs in phase 1 / 2
  static constexpr int THREADS = ((P1 * E1_F + 31) / 32) * 32;
  static constexpr int XS = FL * 8 * (E1_F + 2);          // floats
  static constexpr int HS = FL * 16 * (E1_F + 2);
  static constexpr size_t SMEM = (size_t)(XS + HS) * sizeof(float);
  static_assert(KT % 2 == 0, "pairs");
  static_assert(KT * 8 * E1_F <= XS, "h1 staging aliases the x tile");
};

template <int KT>
__global__ void __launch_bounds__(GtCfg<KT>::THREADS)
gt_main_kernel(const __grid_constant__ GTW w, const float* __restrict__ xin, float* __restrict__ h1,
               float* __restrict__ zt, int T, int dil) {
  using C = GtCfg<KT>;
  constexpr int FP = E1_F + 2;
  extern __shared__ __align__(16) float gt_smem[];
  float (*xs)[8][FP] = reinterpret_cast<float (*)[8][FP]>(gt_smem);             // [FL][8][35]
  float (*hs)[16][FP] = reinterpret_cast<float (*)[16][FP]>(gt_smem + C::XS);   // [FL][16][35]
  float (*h1s)[8][E1_F] = reinterpret_cast<float (*)[8][E1_F]>(gt_smem);        // [KT][8][33], aliases xs

  const int tid = threadIdx.x;
  const int r = blockIdx.y, b = blockIdx.z;
  const int nk = (T - r + dil - 1) / dil;      // frames in this residue class
  const int k0 = blockIdx.x * KT;
  if (r >= T || k0 >= nk) return;
  const float* xb = xin + (long long)b * T * FRAME16;

  // x1 = first 8 channels = first 264 contiguous floats of every frame
  for (int i = tid; i < C::FL * 264; i += C::THREADS) {
    const int fl = i / 264, rem = i - fl * 264;
    const int c = rem / E1_F, f = rem - c * E1_F;
    const int k = k0 - 2 + fl;
    float v = 0.f;
    if (k >= 0 && k < nk) v = __ldg(xb + (long long)(r + k * dil) * FRAME16 + rem);
    xs[fl][c][f + 1] = v;
  }
  for (int i = tid; i < C::FL * 8; i += C::THREADS) {
    xs[i >> 3][i & 7][0] = 0.f;
    xs[i >> 3][i & 7][FP - 1] = 0.f;
  }
  for (int i = tid; i < C::FL * 16; i += C::THREADS) {
    hs[i >> 4][i & 15][0] = 0.f;
    hs[i >> 4][i & 15][FP - 1] = 0.f;
  }
  __syncthreads();

  // phase 1: SFE(k=3) + point_conv1 (24->16) + PReLU for frames (2p, 2p+1).  Frames before the
  // chunk start are the causal zero padding of the depthwise conv's INPUT (:314-318): exact zeros.
  if (tid < C::P1 * E1_F) {
    const int p = tid / E1_F, f = tid - p * E1_F;
    const int fl0 = 2 * p, fl1 = fl0 + 1;
    const int ka = k0 - 2 + fl0, kb = ka + 1;
    float in0[24], in1[24];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        in0[c * 3 + s] = xs[fl0][c][f + s];
        in1[c * 3 + s] = xs[fl1][c][f + s];
      }
    const bool va = (ka >= 0 && ka < nk), vb = (kb >= 0 && kb < nk);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
      float a0 = w.b1[o], a1 = w.b1[o];
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        a0 = fmaf(w.w1[o][i], in0[i], a0);
        a1 = fmaf(w.w1[o][i], in1[i], a1);
      }
      hs[fl0][o][f + 1] = va ? adn_prelu(a0, w.a1) : 0.f;
      hs[fl1][o][f + 1] = vb ? adn_prelu(a1, w.a1) : 0.f;
    }
  }
  __syncthreads();

  // phase 2: depthwise (3,3) dilated causal conv + PReLU + point_conv2 (16->8) for steps (2q, 2q+1)
  if (tid < C::P2 * E1_F) {
    const int q = tid / E1_F, f = tid - q * E1_F;
    const int kl = 2 * q;
    float acc0[8], acc1[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { acc0[o] = w.b2[o]; acc1[o] = w.b2[o]; }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float v[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) v[j][kf] = hs[kl + j][c][f + kf];
      float d0 = w.bd[c], d1 = w.bd[c];
#pragma unroll
      for (int kt = 0; kt < 3; ++kt)
#pragma unroll
        for (int kf = 0; kf < 3; ++kf) {
          d0 = fmaf(w.wd[c][kt][kf], v[kt][kf], d0);
          d1 = fmaf(w.wd[c][kt][kf], v[kt + 1][kf], d1);
        }
      d0 = adn_prelu(d0, w.ad);
      d1 = adn_prelu(d1, w.ad);
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        acc0[o] = fmaf(w.w2[c][o], d0, acc0[o]);
        acc1[o] = fmaf(w.w2[c][o], d1, acc1[o]);
      }
    }
    const int ka = k0 + kl, kb = ka + 1;
    if (ka < nk) {
      float* ho = h1 + ((long long)b * T + (r + ka * dil)) * (8 * E1_F);
#pragma unroll
      for (int o = 0; o < 8; ++o) { ho[o * E1_F + f] = acc0[o]; h1s[kl][o][f] = acc0[o]; }
    }
    if (kb < nk) {
      float* ho = h1 + ((long long)b * T + (r + kb * dil)) * (8 * E1_F);
#pragma unroll
      for (int o = 0; o < 8; ++o) { ho[o * E1_F + f] = acc1[o]; h1s[kl + 1][o][f] = acc1[o]; }
    }
  }
  __syncthreads();

  // phase 3: z_t[c] = mean_f h1^2 (TRA input, :154)
  if (tid < KT * 8) {
    const int kl = tid >> 3, c = tid & 7;
    const int k = k0 + kl;
    float s = 0.f;
    if (k < nk) {
#pragma unroll
      for (int f = 0; f < E1_F; ++f) s = fmaf(h1s[kl][c][f], h1s[kl][c][f], s);
      s = s / (float)E1_F;
      zt[((long long)b * T + (r + k * dil)) * 8 + c] = s;
    }
  }
}

template <int KT>
static void launch_gt_main_t(const GTW& w, const float* xin, float* h1, float* zt, int B, int T, int dil,
                             cudaStream_t st) {
  using C = GtCfg<KT>;
  static unsigned long long configured = 0;      // per device
  if (adn_first_use_on_device(configured))
    cudaFuncSetAttribute(gt_main_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  const int nkmax = (T + dil - 1) / dil;
  dim3 grid((nkmax + KT - 1) / KT, dil, B);
  gt_main_kernel<KT><<<grid, C::THREADS, C::SMEM, st>>>(w, xin, h1, zt, T, dil);
}

static void launch_gt_main(const GTW& w, const float* xin, float* h1, float* zt, int B, int T, int dil,
                           cudaStream_t st) {
  // pick the tile length that wastes the fewest steps of the residue classes
  const int nk = (T + dil - 1) / dil;
  const int waste16 = (nk + 15) / 16 * 16 - nk, waste14 = (nk + 13) / 14 * 14 - nk;
  if (waste14 < waste16) launch_gt_main_t<14>(w, xin, h1, zt, B, T, dil, st);
  else launch_gt_main_t<16>(w, xin, h1, zt, B, T, dil, st);
}

// =================================================================================
// dec_tail: de_convs.3 (ConvT 16->16,(1,5),s2,g2)+PReLU, +e0, de_convs.4 (ConvT 16->2)+tanh,
// ERB.bs, complex ratio mask on the noisy spectrum -> enhanced spectrum frame.
// Transposed conv (stride 2, pad 2): out[g] += in[i]*w[k] with g = 2i + k - 2.
// =================================================================================
// Transposed-conv taps with compile-time group / parity so every weight is a constant-bank
// operand.  Even outputs take kernel taps 0,2,4, odd outputs taps 1,3.
template <int GRP, int PAR>
__device__ __forceinline__ void deconv3_taps(const DecTailW& w, const float (*x)[E1_F + 2], int g,
                                             float (&acc)[8]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = w.b3[GRP * 8 + o];
#pragma unroll
  for (int k = PAR; k < 5; k += 2) {
    int i = (g + 2 - k) >> 1;      // -1..33 -> column i+1 (0 and 34 are zeros)
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
      float v = x[GRP * 8 + ci][i + 1];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = fmaf(w.w3[GRP * 8 + ci][k][o], v, acc[o]);
    }
  }
}

template <int PAR>
__device__ __forceinline__ void deconv4_taps(const DecTailW& w, const float (*y)[E0_F + 2], int g,
                                             float (&a)[2]) {
  a[0] = w.b4[0];
  a[1] = w.b4[1];
#pragma unroll
  for (int k = PAR; k < 5; k += 2) {
    int i = (g + 2 - k) >> 1;      // -1..65 -> column i+1 (0 and 66 are zeros)
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
      float v = y[ci][i + 1];
      a[0] = fmaf(w.w4[ci][k][0], v, a[0]);
      a[1] = fmaf(w.w4[ci][k][1], v, a[1]);
    }
  }
}

constexpr int DT_FR = 4;
constexpr int DT_THREADS = 288;

__global__ void __launch_bounds__(DT_THREADS)
dec_tail_kernel(const __grid_constant__ DecTailW w, const ErbW erb, const float* __restrict__ xin,
                const float* __restrict__ e0, const float* __restrict__ spec, long long spec_chunk_stride,
                float* __restrict__ enh, float* __restrict__ enh_hi, float* __restrict__ enh_lo, int T, int nframes,
                int pad_frames) {
  __shared__ float xs[DT_FR][16][E1_F + 2];      // zero column each side
  __shared__ float ys[DT_FR][16][E0_F + 2];      // d3 + e0, zero column each side
  __shared__ float ms[DT_FR][2][ERB_F];
  __shared__ float mf[DT_FR][2][FB];

  const int tid = threadIdx.x;
  const long long f0 = (long long)blockIdx.x * DT_FR;

  for (int i = tid; i < DT_FR * 16 * (E1_F + 2); i += DT_THREADS) {
    int fr = i / (16 * (E1_F + 2)), rem = i - fr * (16 * (E1_F + 2));
    int c = rem / (E1_F + 2), fp = rem - c * (E1_F + 2);
    long long fg = f0 + fr;
    float v = 0.f;
    if (fg < nframes && fp >= 1 && fp <= E1_F) v = __ldg(xin + fg * FRAME16 + c * E1_F + fp - 1);
    xs[fr][c][fp] = v;
  }
  for (int i = tid; i < DT_FR * 16; i += DT_THREADS) {
    ys[i / 16][i % 16][0] = 0.f;
    ys[i / 16][i % 16][E0_F + 1] = 0.f;
  }
  __syncthreads();

  // de_convs.3: item = (frame, group, g in 0..64); 8 outputs per item
  for (int it = tid; it < DT_FR * 2 * E0_F; it += DT_THREADS) {
    int fr = it / (2 * E0_F), rem = it - fr * (2 * E0_F);
    int grp = rem / E0_F, g = rem - grp * E0_F;
    long long fg = f0 + fr;
    float acc[8];
    if (grp == 0) {
      if (g & 1) deconv3_taps<0, 1>(w, xs[fr], g, acc);
      else deconv3_taps<0, 0>(w, xs[fr], g, acc);
    } else {
      if (g & 1) deconv3_taps<1, 1>(w, xs[fr], g, acc);
      else deconv3_taps<1, 0>(w, xs[fr], g, acc);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float e0v = (fg < nframes) ? __ldg(e0 + fg * FRAME_E0 + (grp * 8 + o) * E0_F + g) : 0.f;
      ys[fr][grp * 8 + o][g + 1] = adn_prelu(acc[o], w.a3) + e0v;
    }
  }
  __syncthreads();

  // de_convs.4 + tanh: item = (frame, g in 0..128), 2 outputs
  for (int it = tid; it < DT_FR * ERB_F; it += DT_THREADS) {
    int fr = it / ERB_F, g = it - fr * ERB_F;
    float a[2];
    if (g & 1) deconv4_taps<1>(w, ys[fr], g, a);
    else deconv4_taps<0>(w, ys[fr], g, a);
    ms[fr][0][g] = tanhf(a[0]);
    ms[fr][1][g] = tanhf(a[1]);
  }
  __syncthreads();

  // ERB.bs: 65 low bins pass through, 64 bands -> 192 high bins
  for (int it = tid; it < DT_FR * 2 * FB; it += DT_THREADS) {
    int fr = it / (2 * FB), rem = it - fr * (2 * FB);
    int c = rem / FB, f = rem - c * FB;
    float v;
    if (f < 65) {
      v = ms[fr][c][f];
    } else {
      int i = f - 65;
      int lo = (int)__ldg(erb.bs_lo + i), hi = (int)__ldg(erb.bs_hi + i);
      v = 0.f;
      for (int j = lo; j < hi; ++j) v = fmaf(ms[fr][c][65 + j], __ldg(erb.bs + j * 192 + i), v);
    }
    mf[fr][c][f] = v;
  }
  __syncthreads();

  // complex ratio mask (:585-590) -> enhanced spectrum frame (inside the zero-framed buffer)
  for (int it = tid; it < DT_FR * FB; it += DT_THREADS) {
    int fr = it / FB, f = it - fr * FB;
    long long fg = f0 + fr;
    if (fg >= nframes) continue;
    long long b = fg / T, t = fg - b * T;
    const float* sp = spec + b * spec_chunk_stride + t * SPEC_LD;     // the masked spectrum: chunk b's frames (H-GTCRN: microphone 0's rows)
    float re = __ldg(sp + f), im = __ldg(sp + FB + f);
    float m0 = mf[fr][0][f], m1 = mf[fr][1][f];
    const long long base = (b * (T + 2 * pad_frames) + pad_frames + t) * SPEC_LD;
    const float er = re * m0 - im * m1, ei = im * m0 + re * m1;
    enh[base + f] = er;
    enh[base + FB + f] = ei;
    if (enh_hi) {
      split_tf32_store(er, enh_hi, enh_lo, base + f);
      split_tf32_store(ei, enh_hi, enh_lo, base + FB + f);
    }
  }
}

// =================================================================================
// launcher
// =================================================================================
#define TICK(name) do { ++n; if (tick) tick(tick_ctx, name); if (stop_after > 0 && n >= stop_after) return n; } while (0)

int launch_backbone_core(const Weights& w, const Buffers& buf, const Dims& d, cudaStream_t st, TickFn tick, void* tick_ctx,
                         int stop_after, int n, float** last) {
  const int B = d.B, T = d.T;
  const int nframes = B * T;
  const int dil_enc[3] = {1, 2, 5};
  for (int i = 0; i < 3; ++i) {
    int dl = dil_enc[i];
    launch_gt_main(w.enc_gt[i], buf.e[i + 1], buf.h1, buf.zt, B, T, dl, st);
    TICK("gt_main");
    launch_tra_gru(w.enc_tra[i], buf.zt, buf.tgi, buf.thid, buf.at, B, T, st);
    TICK("tra_gru");
    launch_tra_apply(buf.at, buf.h1, buf.e[i + 1], nullptr, buf.e[i + 2], B, T, st);
    TICK("tra_apply");
  }

  // DPGRNN x2: x = e4
  launch_dp_intra(w.dp[0], buf.e[4], nullptr, nullptr, buf.xa, buf.gi, nframes, st);
  TICK("dp_intra");
  launch_dp_inter(w.dp[0], buf.gi, buf.inter, B, T, st);
  TICK("dp_inter");
  launch_dp_intra(w.dp[1], buf.xa, buf.inter, &w.dp[0], buf.xb, buf.gi, nframes, st);
  TICK("dp_intra");
  launch_dp_inter(w.dp[1], buf.gi, buf.inter, B, T, st);
  TICK("dp_inter");
  // decoder input 0 = dp2 output + e4
  launch_ln_res(w.dp[1], buf.xb, buf.inter, buf.e[4], buf.xa, nframes, st);
  TICK("ln_res");

  const int dil_dec[3] = {5, 2, 1};
  float* cur = buf.xa;
  float* nxt = buf.xb;
  for (int i = 0; i < 3; ++i) {
    int dl = dil_dec[i];
    launch_gt_main(w.dec_gt[i], cur, buf.h1, buf.zt, B, T, dl, st);
    TICK("gt_main");
    // next stage input = this block's output + encoder skip (e3, e2, e1)
    launch_tra_gru(w.dec_tra[i], buf.zt, buf.tgi, buf.thid, buf.at, B, T, st);
    TICK("tra_gru");
    launch_tra_apply(buf.at, buf.h1, cur, buf.e[3 - i], nxt, B, T, st);
    TICK("tra_apply");
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  *last = cur;
  return n;
}

void launch_dec_tail(const Weights& w, const Buffers& buf, const float* xin, const float* spec, long long spec_chunk_stride,
                     const Dims& d, int enh_pad_frames, cudaStream_t st) {
  const int nframes = d.B * d.T;
  dec_tail_kernel<<<(nframes + DT_FR - 1) / DT_FR, DT_THREADS, 0, st>>>(w.dec_tail, w.erb, xin, buf.e0, spec, spec_chunk_stride,
                                                                      buf.enh, buf.enh_hi, buf.enh_lo, d.T, nframes,
                                                                      enh_pad_frames);
}

int launch_backbone(const Weights& w, const Buffers& buf, const Dims& d, int enh_pad_frames,
                    cudaStream_t st, TickFn tick, void* tick_ctx, int stop_after) {
  int n = 0;
  const int nframes = d.B * d.T;
  enc_front_kernel<<<(nframes + EF_FR - 1) / EF_FR, EF_THREADS, 0, st>>>(w.enc_front, w.erb, buf.spec,
                                                                       buf.e0, buf.e[1], nframes);
  TICK("enc_front");
  float* cur = nullptr;
  n = launch_backbone_core(w, buf, d, st, tick, tick_ctx, stop_after, n, &cur);
  if (!cur) return n;                       // stopped inside the core
  launch_dec_tail(w, buf, cur, buf.spec, (long long)d.T * SPEC_LD, d, enh_pad_frames, st);
  TICK("dec_tail");
  return n;
}

}  // namespace gtcrn
