// Shared helpers for libadn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>

#define ADN_WARP 32

struct AdnError {
  std::string msg;
};

#define ADN_CUDA_TRY(expr, errstr)                                                   \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      (errstr) = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
      return ADN_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// cudaFuncSetAttribute is per device: true the first time `mask` (one static per call site) sees the current device.
inline bool adn_first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// Run-time workspaces are raw cudaMalloc blocks; a captured CUDA graph (adn_run, api.cu) holds their addresses.  Every path that
// frees one calls adn_note_free(); adn_run drops the graphs captured before the epoch changed.
void adn_note_free();
unsigned long long adn_alloc_epoch();

__device__ __forceinline__ float adn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float adn_prelu(float x, float a) { return x >= 0.f ? x : a * x; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------
// Row-gather GEMM  C[m,n] = sum_k A(m,k) * W[n,k]
//   m = b*TM + t,  A(m,k) = A[b*a_sB + (t + a_t0)*a_sT + k]   (k contiguous)
// With a_sT < K the rows overlap: that is the STFT framing (a_sT = hop, K = nfft) and
// the ISTFT overlap-add (rows = runs of R consecutive spectrum frames).
// ---------------------------------------------------------------------------------
enum { EPI_STORE = 0, EPI_ISTFT = 1, EPI_LIN = 2 };

struct GemmArgs {
  const float* A;
  long long a_sB;
  int a_sT;
  int a_t0;
  int TM;
  const float* W;   // [N][ldw]
  int ldw;
  int M, N, K;
  // EPI_STORE: C[b*c_sB + t*c_sT + n*c_sN] = acc
  float* C;
  long long c_sB;
  long long c_sT;
  long long c_sN;
  // EPI_ISTFT: s = (t + a_t0)*hop + n - shift; if 0<=s<out_len:
  //   v = acc (/ or *) norm[s]; out[b*out_len + s] = convert(v)
  const float* norm;
  int norm_mul;
  int hop;
  int shift;
  int out_len;
  int out_dtype;    // ADN_F32 / ADN_I16 / ADN_F16
  void* out;
};

void launch_gemm_ffma(const GemmArgs& g, int epi, cudaStream_t st);
