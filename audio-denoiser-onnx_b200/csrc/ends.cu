// Stand-alone waveform-conditioning / spectral-feature / recombine / output-conditioning operators for
// the families whose backbone is not built here (ZipEnhancer, MossFormerGAN-SE-16K, MossFormer2-SS-16K)
// and the linear resampler every wrapper shares.  Together with
// adn_stft_forward / adn_stft_inverse they are the complete front and back ends around those backbones.
// All are single-pass, HBM-bound kernels; one CTA per waveform row where a per-row statistic is needed.
#include "adn.h"
#include "common.cuh"

#include <cfloat>
#include <mutex>
#include <string>

void adn_internal_set_error(const std::string& s);   // api.cu

namespace ends {

template <typename T>
__device__ __forceinline__ float ld(const T* p, long long i);
template <>
__device__ __forceinline__ float ld<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ld<int16_t>(const int16_t* p, long long i) { return (float)p[i]; }
template <>
__device__ __forceinline__ float ld<__half>(const __half* p, long long i) { return __half2float(p[i]); }

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// F.interpolate(mode='linear', align_corners=False): src = max(r*(i+0.5)-0.5, 0) with ONE rounding
// (ATen is FMA-contracted; ulp(16000) = 1e-3 makes this visible), i0 = floor, lambda = src - i0.
template <typename T>
__global__ void __launch_bounds__(256)
resample_kernel(const T* __restrict__ in, float* __restrict__ out, int len_in, int len_out, float r) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= len_out) return;
  const long long row = blockIdx.y;
  float src = fmaxf(fmaf(r, (float)i + 0.5f, -0.5f), 0.f);
  int i0 = min((int)floorf(src), len_in - 1);
  const float lam = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  const int i1 = i0 + (i0 < len_in - 1 ? 1 : 0);
  const T* x = in + row * len_in;
  out[row * len_out + i] = (1.0f - lam) * ld<T>(x, i0) + lam * ld<T>(x, i1);
}

// y = (x*pre) / sqrt(mean((x*pre)^2) + eps), optionally extended to len_out by wrapping the head
// (Export_ZipEnhancer.py:839-840; MossFormerGAN .../Export_MossFormer_SE.py:564-568).
template <typename T>
__global__ void __launch_bounds__(256)
rms_normalize_kernel(const T* __restrict__ in, float pre, float eps, float* __restrict__ out, float* __restrict__ norm,
                     int len, int len_out) {
  __shared__ double red[8];
  const long long row = blockIdx.x;
  const T* x = in + row * len;
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) { const float v = ld<T>(x, i) * pre; s += (double)v * v; }
  s = block_sum(s, red);
  const float nf = sqrtf((float)(s / (double)len) + eps);
  if (threadIdx.x == 0) norm[row] = nf;
  float* o = out + row * len_out;
  for (int i = threadIdx.x; i < len_out; i += 256) o[i] = (ld<T>(x, i < len ? i : i - len) * pre) / nf;
}

// MossFormer2_SS norm_audio (:403-423): whole-window RMS gain, then RMS of the above-average-power samples.
template <typename T>
__global__ void __launch_bounds__(256)
two_stage_rms_kernel(const T* __restrict__ in, float target, float eps, float* __restrict__ out,
                     float* __restrict__ rms_in, int len) {
  __shared__ double red[8];
  const long long row = blockIdx.x;
  const T* x = in + row * len;
  const float inv = 1.0f / 32768.0f;
  double s = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) { const float v = ld<T>(x, i) * inv; s += (double)(v * v); }
  s = block_sum(s, red);
  const float avg = (float)(s / (double)len);
  const float rms = sqrtf(avg);
  const float s1 = target / (rms + eps);
  double hs = 0.0, hc = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) {
    const float v = ld<T>(x, i) * inv, p = v * v;
    if (p > avg) { hs += (double)p; hc += 1.0; }
  }
  hs = block_sum(hs, red);
  hc = block_sum(hc, red);
  const float high = sqrtf((float)(hs / (hc < 1.0 ? 1.0 : hc)));
  const float s2 = target / (high * s1 + eps);
  if (threadIdx.x == 0) {
    const float g = s1 * s2, undo = 1.0f / (g + eps);
    rms_in[row] = rms * g * undo * 32767.0f;
  }
  float* o = out + row * len;
  for (int i = threadIdx.x; i < len; i += 256) o[i] = ((ld<T>(x, i) * inv) * s1) * s2;
}

// spec (B,2F,T) -> feature map (B,C,T,F), 32x32 transposing tiles.
//   ZipEnhancer (:843-850): [ (re^2+im^2+1e-9)^0.15 , atan2(im, re+1e-5) ]
//   MossFormerGAN (:578-586): [ p^0.15 , re*s , im*s ], s = max(p, FLT_MIN)^(0.15-0.5); keep = (re*s, im*s) as (B,2,F,T)
__global__ void __launch_bounds__(256)
spec_features_kernel(int family, const float* __restrict__ spec, float* __restrict__ feat, float* __restrict__ keep,
                     int F, int T) {
  __shared__ float tile[3][32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int C = family == ADN_FAMILY_ZIPENHANCER ? 2 : 3;
  const float* sb = spec + (long long)b * 2 * F * T;
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r, t = t0 + tx;
    if (f < F && t < T) {
      const float re = sb[(long long)f * T + t], im = sb[(long long)(F + f) * T + t];
      if (family == ADN_FAMILY_ZIPENHANCER) {
        tile[0][r][tx] = powf(re * re + im * im + 1e-9f, 0.15f);
        tile[1][r][tx] = atan2f(im, re + 1e-5f);
      } else {
        const float p = re * re + im * im;
        const float s = powf(fmaxf(p, FLT_MIN), 0.15f - 0.5f);
        tile[0][r][tx] = powf(p, 0.15f);
        tile[1][r][tx] = re * s;
        tile[2][r][tx] = im * s;
        if (keep) {
          keep[((long long)b * 2 * F + f) * T + t] = re * s;
          keep[((long long)b * 2 * F + F + f) * T + t] = im * s;
        }
      }
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, f = f0 + tx;
    if (t < T && f < F)
      for (int c = 0; c < C; ++c) feat[(((long long)b * C + c) * T + t) * F + f] = tile[c][tx][r];
  }
}

//   ZipEnhancer (:882-892): a = mx (B,1,T,F), bb = phase_ri (B,2,T,F) -> relu(mx)^(1/0.3) * unit(phase) (zero-phase guard)
//   MossFormerGAN (:863-868): a = mask (B,F,T), bb = complex_out (B,2,F,T), keep = compressed spectrum -> decompress
__global__ void __launch_bounds__(256)
spec_recombine_kernel(int family, const float* __restrict__ a, const float* __restrict__ bb,
                      const float* __restrict__ keep, float* __restrict__ spec, int F, int T) {
  const int b = blockIdx.z;
  float* ob = spec + (long long)b * 2 * F * T;
  if (family == ADN_FAMILY_ZIPENHANCER) {
    __shared__ float tile[3][32][33];
    const int f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
      const int t = t0 + r, f = f0 + tx;
      if (t < T && f < F) {
        tile[0][r][tx] = a[((long long)b * T + t) * F + f];
        tile[1][r][tx] = bb[(((long long)b * 2 + 0) * T + t) * F + f];
        tile[2][r][tx] = bb[(((long long)b * 2 + 1) * T + t) * F + f];
      }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int f = f0 + r, t = t0 + tx;
      if (f < F && t < T) {
        const float mag = powf(fmaxf(tile[0][tx][r], 0.f), 1.0f / 0.3f);
        float pr = tile[1][tx][r], pi = tile[2][tx][r];
        float nrm = sqrtf(pr * pr + pi * pi);
        if (!(nrm > 0.f)) { pr = 1.f; pi = 0.f; nrm = 1.f; }
        const float g = mag / nrm;
        ob[(long long)f * T + t] = pr * g;
        ob[(long long)(F + f) * T + t] = pi * g;
      }
    }
  } else {
    const long long n = (long long)F * T;
    const long long i = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
    if (i >= n) return;
    const float m = a[(long long)b * n + i];
    const float fr = m * keep[(long long)b * 2 * n + i] + bb[(long long)b * 2 * n + i];
    const float fi = m * keep[(long long)b * 2 * n + n + i] + bb[(long long)b * 2 * n + n + i];
    const float fac = powf(fr * fr + fi * fi, (0.5f / 0.3f) - 0.5f);
    ob[i] = fr * fac;
    ob[n + i] = fi * fac;
  }
}

// wave (rows, len_src) -> out (rows, len): trim, per-row gain, family output rule.
//   ZipEnhancer (:899-926): x*nf; int16: NaN->0, clamp, truncate; float: nan_to_num(0, 32767, -32768) / 32768
//   MossFormerGAN (:880-897): x*nf; int16: clamp, truncate; float: / 32768
//   MossFormer2-SS (:627-660): gain = rms_out > 0 ? rms_in / rms_out : 0; int16: int32 staging; float: / 32768
template <typename TO>
__device__ __forceinline__ void store_out(TO* o, long long i, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* o, long long i, float v) { o[i] = v; }
template <>
__device__ __forceinline__ void store_out<__half>(__half* o, long long i, float v) { o[i] = __float2half_rn(v); }

__global__ void __launch_bounds__(256)
condition_output_kernel(int family, const float* __restrict__ wave, int len_src, const float* __restrict__ gain,
                        int gain_group, void* __restrict__ out, int out_dtype, int len) {
  __shared__ double red[8];
  const long long row = blockIdx.x;
  const float* x = wave + row * len_src;
  float g = gain[row / gain_group];
  if (family == ADN_FAMILY_MOSSFORMER2_SS) {
    double s = 0.0;
    for (int i = threadIdx.x; i < len; i += 256) s += (double)(x[i] * x[i]);
    s = block_sum(s, red);
    const float rms_out = sqrtf((float)(s / (double)len));
    g = rms_out > 0.f ? g / rms_out : 0.f;
  }
  const float inv = 1.0f / 32768.0f;
  for (int i = threadIdx.x; i < len; i += 256) {
    float v = x[i] * g;
    const long long o = row * len + i;
    if (out_dtype == ADN_I16) {
      int q;
      if (family == ADN_FAMILY_MOSSFORMER2_SS) {
        q = max(-32768, min(32767, (int)fminf(fmaxf(v, -2147483648.f), 2147483520.f)));
      } else {
        if (family == ADN_FAMILY_ZIPENHANCER && isnan(v)) v = 0.f;
        q = (int)fminf(fmaxf(v, -32768.f), 32767.f);
      }
      reinterpret_cast<int16_t*>(out)[o] = (int16_t)q;
    } else {
      if (family == ADN_FAMILY_ZIPENHANCER) {
        if (isnan(v)) v = 0.f;
        else if (isinf(v)) v = v > 0.f ? 32767.f : -32768.f;
      }
      v *= inv;
      if (out_dtype == ADN_F32) store_out<float>(reinterpret_cast<float*>(out), o, v);
      else store_out<__half>(reinterpret_cast<__half*>(out), o, v);
    }
  }
}

adn_status fail(const char* what) {
  adn_internal_set_error(what);
  return ADN_ERR_INVALID;
}
adn_status done(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    adn_internal_set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return ADN_ERR_CUDA;
  }
  return ADN_OK;
}

}  // namespace ends

extern "C" {

adn_status adn_resample_linear(const void* d_in, int32_t in_dtype, float* d_out, int32_t rows, int32_t len_in,
                               int32_t len_out, double scale_factor, void* stream) {
  if (!d_in || !d_out || rows <= 0 || len_in <= 0 || len_out <= 0) return ends::fail("adn_resample_linear: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // scale_factor given (GTCRN): r = 1/scale_factor; sizes only (the other wrappers): r = L_in / L_out
  const float r = scale_factor > 0.0 ? (float)(1.0 / scale_factor) : (float)len_in / (float)len_out;
  dim3 grid((len_out + 255) / 256, rows);
  if (in_dtype == ADN_F32) ends::resample_kernel<float><<<grid, 256, 0, st>>>((const float*)d_in, d_out, len_in, len_out, r);
  else if (in_dtype == ADN_I16) ends::resample_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)d_in, d_out, len_in, len_out, r);
  else if (in_dtype == ADN_F16) ends::resample_kernel<__half><<<grid, 256, 0, st>>>((const __half*)d_in, d_out, len_in, len_out, r);
  else return ends::fail("adn_resample_linear: bad dtype");
  return ends::done("adn_resample_linear");
}

adn_status adn_rms_normalize(const void* d_in, int32_t in_dtype, float pre_scale, float eps, float* d_out, float* d_norm,
                             int32_t rows, int32_t len, int32_t len_out, void* stream) {
  if (!d_in || !d_out || !d_norm || rows <= 0 || len <= 0 || len_out < len || len_out > 2 * len)
    return ends::fail("adn_rms_normalize: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == ADN_F32) ends::rms_normalize_kernel<float><<<rows, 256, 0, st>>>((const float*)d_in, pre_scale, eps, d_out, d_norm, len, len_out);
  else if (in_dtype == ADN_I16) ends::rms_normalize_kernel<int16_t><<<rows, 256, 0, st>>>((const int16_t*)d_in, pre_scale, eps, d_out, d_norm, len, len_out);
  else if (in_dtype == ADN_F16) ends::rms_normalize_kernel<__half><<<rows, 256, 0, st>>>((const __half*)d_in, pre_scale, eps, d_out, d_norm, len, len_out);
  else return ends::fail("adn_rms_normalize: bad dtype");
  return ends::done("adn_rms_normalize");
}

adn_status adn_two_stage_rms(const void* d_in, int32_t in_dtype, float target, float eps, float* d_out, float* d_rms_in,
                             int32_t rows, int32_t len, void* stream) {
  if (!d_in || !d_out || !d_rms_in || rows <= 0 || len <= 0) return ends::fail("adn_two_stage_rms: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == ADN_F32) ends::two_stage_rms_kernel<float><<<rows, 256, 0, st>>>((const float*)d_in, target, eps, d_out, d_rms_in, len);
  else if (in_dtype == ADN_I16) ends::two_stage_rms_kernel<int16_t><<<rows, 256, 0, st>>>((const int16_t*)d_in, target, eps, d_out, d_rms_in, len);
  else if (in_dtype == ADN_F16) ends::two_stage_rms_kernel<__half><<<rows, 256, 0, st>>>((const __half*)d_in, target, eps, d_out, d_rms_in, len);
  else return ends::fail("adn_two_stage_rms: bad dtype");
  return ends::done("adn_two_stage_rms");
}

adn_status adn_spec_features(int32_t family, const float* d_spec, float* d_feat, float* d_keep, int32_t batch,
                             int32_t fbins, int32_t frames, void* stream) {
  if ((family != ADN_FAMILY_ZIPENHANCER && family != ADN_FAMILY_MOSSFORMERGAN) || !d_spec || !d_feat || batch <= 0 ||
      fbins <= 0 || frames <= 0 || (family == ADN_FAMILY_MOSSFORMERGAN && !d_keep))
    return ends::fail("adn_spec_features: bad argument");
  dim3 grid((frames + 31) / 32, (fbins + 31) / 32, batch);
  ends::spec_features_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(family, d_spec, d_feat, d_keep, fbins, frames);
  return ends::done("adn_spec_features");
}

adn_status adn_spec_recombine(int32_t family, const float* d_a, const float* d_b, const float* d_keep, float* d_spec,
                              int32_t batch, int32_t fbins, int32_t frames, void* stream) {
  if ((family != ADN_FAMILY_ZIPENHANCER && family != ADN_FAMILY_MOSSFORMERGAN) || !d_a || !d_b || !d_spec || batch <= 0 ||
      fbins <= 0 || frames <= 0 || (family == ADN_FAMILY_MOSSFORMERGAN && !d_keep))
    return ends::fail("adn_spec_recombine: bad argument");
  dim3 grid((frames + 31) / 32, (fbins + 31) / 32, batch);
  if (family == ADN_FAMILY_MOSSFORMERGAN) grid = dim3((unsigned)(((long long)fbins * frames + 255) / 256), 1, batch);
  ends::spec_recombine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(family, d_a, d_b, d_keep, d_spec, fbins, frames);
  return ends::done("adn_spec_recombine");
}

adn_status adn_condition_output(int32_t family, const float* d_wave, int32_t len_src, const float* d_gain,
                                int32_t gain_group, void* d_out, int32_t out_dtype, int32_t rows, int32_t len,
                                void* stream) {
  if (family < ADN_FAMILY_ZIPENHANCER || family > ADN_FAMILY_MOSSFORMER2_SS || !d_wave || !d_gain || !d_out || rows <= 0 ||
      len <= 0 || len_src < len || gain_group <= 0 || (out_dtype != ADN_F32 && out_dtype != ADN_I16 && out_dtype != ADN_F16))
    return ends::fail("adn_condition_output: bad argument");
  ends::condition_output_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(family, d_wave, len_src, d_gain, gain_group, d_out,
                                                                        out_dtype, len);
  return ends::done("adn_condition_output");
}

}  // extern "C"
