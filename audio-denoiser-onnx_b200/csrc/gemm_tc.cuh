// tcgen05 / TMA row-gather GEMM (3xTF32): host interface.  See gemm_tc.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string>

namespace tc {

struct TcArgs {
  int m_tiles;           // number of 128-row tiles
  int bb, bt;            // TMA box: bb chunks x bt rows per tile (bb*bt <= 128)
  int tiles_per_chunk;   // when bb == 1: ceil(TM / 128)
  int t0;                // first row (per chunk) in the A tensor map
  int B, TM;             // chunks (exclusive upper bound of the chunk index), valid rows per chunk
  int b_off;             // first chunk index of this launch
  int w_batched;         // 1: weights differ per chunk (W map has a chunk dimension)
  int w_group;           // > 1: `w_group` consecutive chunks share one W batch and walk its K axis:
  int w_kstep;           //   chunk c reads W batch c / w_group at K offset (c % w_group) * w_kstep
  int k_split;           // > 0: K is a concatenation; columns >= k_split come from the second W operand (plan.map_w2_*,
                         //   batch c / w_group, K offset 0): C = A[:, :k_split] W1^T + A[:, k_split:] W2^T in one pass
  int N, K;
  // EPI_STORE: C[b*c_sB + t*c_sT + n]
  float* C;
  long long c_sB, c_sT;
  // EPI_ISTFT (see GemmArgs in common.cuh)
  const float* norm;
  int norm_mul, hop, shift, out_len, out_dtype;
  void* out;
  // EPI_LIN: row m = b*TM + t;  v = act(acc*rowscale[m] + bias[n]) (+ resid[m*ldc+n]);
  //          C[m*ldc+n] = v (fp32, optional; may alias resid) and/or Chi/Clo (tf32 planes, optional)
  const float* rowscale;
  const float* bias;
  long long bias_bstride; // bias of chunk b starts at bias + b*bias_bstride (batched weights)
  const float* resid;
  float* Chi;            // with Clo == nullptr: ONE bf16 plane (2-byte elements at the same element offsets)
  float* Clo;
  long long ldc;
  int act;               // ACT_*
  const float* act_param; // ACT_PRELU: device scalar slope
  // conv-as-GEMM (ZipEnhancer dense blocks): taps > 0 turns K into a concatenation of `taps` channel windows of tap_kb K blocks
  // each; K block kb reads A-map columns tap_k0 + (kb % tap_kb)*32 of row t + tap_shift[kb / tap_kb] (rows outside the
  // chunk read as zeros: the causal / sub-band zero padding of the conv).  The W operand keeps its plain K axis.
  int taps, tap_kb, tap_k0;
  int tap_shift[6];
  // EPI_LIN: after the residual, v = resid2[m*ldc+n] + (v - resid2[m*ldc+n]) * colscale[n]   (Zipformer2 BypassModule)
  const float* resid2;
  const float* colscale;
  // fp32-A mode extras (MossFormerGAN / DFSMN operators on this GEMM):
  const float* a_rowstat;  // (mean, rstd) per output row m: the converter warps normalise the A row while splitting it (LayerNorm on load)
  int tap_w;               // taps > 0 on an UN-padded map of row width tap_w: rows whose column + tap_df[tap] leaves [0, tap_w) are zeroed
  int tap_df[6];           //   by the converter (the conv's zero padding along the row axis)
  const float* act_vec;    // ACT_PRELU_VEC: slope per output column
  int probe;             // diagnostics (ADN_TC_PROBE): 1 = converters do not convert, 2 = epilogue does not load / store, 4 = W loaded once per CTA
  int i16_mode;          // EPI_ISTFT int16 output: 0 = x*32767, clamp, truncate (GTCRN, Export_GTCRN.py:680-693)
                         //   1 = clamp(x,-1,32767/32768)*32768, truncate (MossFormer2_SE_48K/Export_MossFormer_SE.py:499-504)
};

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_TANH = 2, ACT_SILU = 3, ACT_RELU = 4, ACT_RELU2 = 5, ACT_PRELU = 6,
       ACT_SWOOSH_L = 7, ACT_SWOOSH_R = 8, ACT_PRELU_VEC = 9, ACT_SIGMOID = 10 };   // softplus(x - 4 | 1) - 0.08 x (Export_ZipEnhancer.py:131-140)

struct TcPlan {
  CUtensorMap map_a_hi, map_a_lo, map_w_hi, map_w_lo;
  CUtensorMap map_w2_hi, map_w2_lo;   // only read when args.k_split > 0
  CUtensorMap map_c;                  // a_f32 only: the fp32 output, written by TMA stores (make_store_map)
  int bn = 0;            // N tile: 64 | 128 | 176 | 256
  bool a_f32 = false;    // map_a_hi is over fp32 activations; the kernel splits them into tf32 hi / lo tiles itself (EPI_LIN only)
  bool bf16 = false;     // operands are single bf16 planes (maps built with bf16 = true); EPI_LIN only, bn 128 | 256
};

// A planes: element (b, t, k) at base[b*batch_stride + t*row_stride + k]; rows may overlap.
bool make_row_map(CUtensorMap* map, const float* base, int k_extent, int rows, long long row_stride, int batches,
                  long long batch_stride, int box_rows, int box_batches, std::string& err, bool bf16 = false);
// W planes: (n_pad, k_pad) row-major, zero padded.
bool make_weight_map(CUtensorMap* map, const float* base, int k_pad, int n_pad, int box_n, std::string& err,
                     int batches = 1, bool bf16 = false);

// Plain (un-swizzled) 3-D tile map over fp32 rows: element (b, r, c) at base[b*batch_stride + r*row_stride + c]; boxes of
// box_rows x box_cols; out-of-range rows / columns (negative start coordinates included) read as zeros.
bool make_tile_map(CUtensorMap* map, const float* base, int cols, int rows, long long row_stride, int batches,
                   long long batch_stride, int box_cols, int box_rows, std::string& err);

// fp32 output of an a_f32 plan: element (b, r, n) at base[b*batch_stride + r*row_stride + n]; rows >= `rows` and columns >= `cols`
// are never written.
bool make_store_map(CUtensorMap* map, float* base, int cols, int rows, long long row_stride, int batches, long long batch_stride,
                    std::string& err);

cudaError_t launch(const TcPlan& p, const TcArgs& a, int epi, int sms, cudaStream_t st);

void split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st);

}  // namespace tc
