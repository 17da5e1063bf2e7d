// CUDA executor of the functor launch sequences (csrc/mfgan_ops.cuh, csrc/dfsmn_ops.cuh): one grid per functor, the
// contractions on the tiled GEMM of csrc/mfgan_gemm.cuh; per-launch tick (CUDA-event timing) and stage dumps.
#pragma once
#include "mfgan_gemm.cuh"

#include "model_impl.h"

#include <map>
#include <string>
#include <vector>

namespace gan {

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
    launch_gemm(g, st);
    ++launches;
    if (tick) tick(tick_ctx, name);
  }
  void run(long long n, const Linear& f) { run_gemm(n, f); }
  void run(long long n, const SimLocal& f) { run_gemm(n, f); }
  void run(long long n, const SimCross& f) { run_gemm(n, f); }
  void run(long long n, const LinKV& f) { run_gemm(n, f); }
  void run(long long n, const Att& f) { run_gemm(n, f); }
  void run(long long n, const GateConvT& f) { run_gemm(n, f); }
  void run(long long n, const TaScores& f) { run_gemm(n, f); }
  void run(long long n, const TaAV& f) { run_gemm(n, f); }
  void run(long long n, const Conv2d& f) {
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
