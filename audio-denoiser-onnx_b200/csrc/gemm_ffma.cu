// Row-gather fp32 GEMM on CUDA cores (FFMA): the exact-fp32 path for the reference's
// "DFT as a convolution" (STFT: Conv1d, GTCRN/STFT_Process.py:316; ISTFT:
// ConvTranspose1d + overlap-add, :328-336).  Frames are never materialised: row m of A
// is a window into the (padded) waveform / spectrum buffer.
#include "adn.h"
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <int EPI>
__global__ void __launch_bounds__(256) gemm_rows_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // each thread stages 4 A and 4 W elements per K-chunk: element e = tid + i*256,
  // row = e / 16, kk = e % 16 (16 consecutive lanes read 64 contiguous bytes)
  const float* a_ptr[4];
  const float* w_ptr[4];
  bool a_ok[4], w_ok[4];
  const int kk = tid & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = (tid >> 4) + i * 16;
    int m = m0 + row;
    a_ok[i] = m < g.M;
    int mm = a_ok[i] ? m : 0;
    int b = mm / g.TM, t = mm - b * g.TM;
    a_ptr[i] = g.A + (long long)b * g.a_sB + (long long)(t + g.a_t0) * g.a_sT + kk;
    int n = n0 + row;
    w_ok[i] = n < g.N;
    w_ptr[i] = g.W + (long long)(w_ok[i] ? n : 0) * g.ldw + kk;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rw[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool kin = (k0 + kk) < g.K;
      ra[i] = (a_ok[i] && kin) ? __ldg(a_ptr[i] + k0) : 0.f;
      rw[i] = (w_ok[i] && kin) ? __ldg(w_ptr[i] + k0) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row = (tid >> 4) + i * 16;
      As[buf][kk][row] = ra[i];
      Ws[buf][kk][row] = rw[i];
    }
  };

  const int nchunks = (g.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) gload((c + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float a[4] = {av.x, av.y, av.z, av.w};
      float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (c + 1 < nchunks) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    int b = m / g.TM, t = m - b * g.TM;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_STORE) {
        g.C[(long long)b * g.c_sB + (long long)t * g.c_sT + (long long)n * g.c_sN] = v;
      } else {
        int s = (t + g.a_t0) * g.hop + n - g.shift;
        if (s < 0 || s >= g.out_len) continue;
        float nv = __ldg(g.norm + s);
        v = g.norm_mul ? v * nv : v / nv;
        long long o = (long long)b * g.out_len + s;
        if (g.out_dtype == ADN_F32) {
          reinterpret_cast<float*>(g.out)[o] = v;
        } else if (g.out_dtype == ADN_I16) {
          // *32767, clamp, truncate toward zero (Export_GTCRN.py:680-690)
          float q = fminf(fmaxf(v * 32767.0f, -32768.0f), 32767.0f);
          reinterpret_cast<int16_t*>(g.out)[o] = (int16_t)(int)q;
        } else {
          reinterpret_cast<__half*>(g.out)[o] = __float2half_rn(v);
        }
      }
    }
  }
}

}  // namespace

void launch_gemm_ffma(const GemmArgs& g, int epi, cudaStream_t st) {
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
  if (epi == EPI_STORE)
    gemm_rows_kernel<EPI_STORE><<<grid, 256, 0, st>>>(g);
  else
    gemm_rows_kernel<EPI_ISTFT><<<grid, 256, 0, st>>>(g);
}
