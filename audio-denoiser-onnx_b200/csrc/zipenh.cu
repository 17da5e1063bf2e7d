// ZipEnhancer 16 kHz (BASELINE configs[1]) behind the C ABI: model family `zipenhancer`.
// Reference: ZipEnhancer/Export_ZipEnhancer.py `ZipEnhancer.forward` (:818-927).
//   RMS norm (:839-840) -> STFT 400/100 hann (:841) -> compressed magnitude + phase (:843-844)                [ends.cu operators]
//   -> DenseEncoder, 4 dual-path Zipformer2 encoders, mask / phase decoders                                 [zipenh_ops.cuh sequence]
//   -> magnitude decompress x unit phase (:882-892) -> ISTFT x 1/sum w^2 (:893) -> x norm factor, output rule (:899-926) [ends.cu]
// The sequence's LinOps (every Linear, the (2,3) dilated causal convs, the stride-2 and sub-pixel convs) run on the tcgen05
// 3xTF32 GEMM (gemm_tc.cu: TMA-fed fp32 activations split into tf32 hi / lo tiles in shared memory, TMEM accumulators), planned once per batch size; attention weights, the two attention
// value products, the gated depthwise conv and the final BiasNorm are cooperative kernels below; the remaining element-wise
// functors run one thread per output.
#include "zipenh_ops.cuh"

#include "common.cuh"
#include "gemm_tc.cuh"
#include "model_impl.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace zip {

template <class F>
__global__ void __launch_bounds__(256) op_kernel(long long n, F f) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) f(i);
}

template <class F> struct OpName { static const char* get() { return "zip_op"; } };
#define ZIP_OP_NAME(T, s) template <> struct OpName<T> { static const char* get() { return s; } }
ZIP_OP_NAME(FeatConv, "zip_feat_conv");
ZIP_OP_NAME(InPart, "zip_in_part");
ZIP_OP_NAME(InFin, "zip_in_fin");
ZIP_OP_NAME(InApply, "zip_in_apply");
ZIP_OP_NAME(PadCopy, "zip_pad_copy");
ZIP_OP_NAME(AttnW, "zip_attn_w");
ZIP_OP_NAME(SaApply, "zip_sa_apply");
ZIP_OP_NAME(NlApply, "zip_nl_apply");
ZIP_OP_NAME(GluDwConv, "zip_glu_dwconv");
ZIP_OP_NAME(NormBypass, "zip_norm_bypass");
ZIP_OP_NAME(Down, "zip_down");
ZIP_OP_NAME(UpCombine, "zip_up_combine");
ZIP_OP_NAME(Head, "zip_head");

// ------------------------------------------------------------------------------------------------ attention weights
// One CTA = one (sequence, head): q | p rows, k^T and the head's relative-position table staged in shared memory; one warp
// = four query rows at a time, lanes over the keys; softmax in registers; rows written coalesced.
template <int JJ, int NR>
__global__ void __launch_bounds__(256) attn_w_kernel(const float* __restrict__ ap, SeqMap sm, const float* __restrict__ pos,
                                                    float* __restrict__ aw) {
  extern __shared__ float sh[];
  const int S = sm.S, SP = (S + 31) & ~31, RP = 2 * S;
  float* kT = sh;                  // [QD][SP]
  float* qp = kT + QD * SP;        // [S][16]: q[12] | p[4]
  float* R = qp + S * 16;          // [PD][RP]
  const long long n = blockIdx.x / HEADS;
  const int h = blockIdx.x % HEADS;
  for (int idx = threadIdx.x; idx < S * 7; idx += 256) {
    const int s = idx / 7, e = idx - s * 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(ap + sm.tok(n, s) * AP + h * HB) + e);
    if (e < 3) *reinterpret_cast<float4*>(qp + s * 16 + 4 * e) = v;
    else if (e < 6) {
      const int d = 4 * (e - 3);
      kT[(d + 0) * SP + s] = v.x; kT[(d + 1) * SP + s] = v.y; kT[(d + 2) * SP + s] = v.z; kT[(d + 3) * SP + s] = v.w;
    } else *reinterpret_cast<float4*>(qp + s * 16 + 12) = v;
  }
  const int RL = 2 * S - 1;
  for (int idx = threadIdx.x; idx < PD * RL; idx += 256) {
    const int d = idx / RL, r = idx - d * RL;
    R[d * RP + r] = __ldg(pos + (long long)h * PD * RL + idx);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i0 = warp * NR; i0 < S; i0 += 8 * NR) {
    float q[NR][16];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int ii = min(i0 + r, S - 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 v = *reinterpret_cast<const float4*>(qp + ii * 16 + 4 * e);
        q[r][4 * e] = v.x; q[r][4 * e + 1] = v.y; q[r][4 * e + 2] = v.z; q[r][4 * e + 3] = v.w;
      }
    }
    float sc[NR][JJ];
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        float k[QD];
#pragma unroll
        for (int d = 0; d < QD; ++d) k[d] = kT[d * SP + j];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int ii = min(i0 + r, S - 1);
          float a = 0.f;
#pragma unroll
          for (int d = 0; d < QD; ++d) a += q[r][d] * k[d];
          float e = 0.f;
#pragma unroll
          for (int d = 0; d < PD; ++d) e += q[r][QD + d] * R[d * RP + (S - 1 - ii + j)];
          sc[r][jj] = a + e;
        }
      } else {
#pragma unroll
        for (int r = 0; r < NR; ++r) sc[r][jj] = -INFINITY;
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float mx = sc[r][0];
#pragma unroll
      for (int jj = 1; jj < JJ; ++jj) mx = fmaxf(mx, sc[r][jj]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < JJ; ++jj) { sc[r][jj] = expf(sc[r][jj] - mx); sum += sc[r][jj]; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      if (i0 + r < S) {
        const int ld = aw_ld(S);
        float* row = aw + ((n * HEADS + h) * S + i0 + r) * (long long)ld;
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
          const int j = lane + 32 * jj;
          if (j < S) row[j] = sc[r][jj] * inv;
          else if (j < ld) row[j] = 0.f;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention value products
// 48 partial sums per lane -> lane pair (2p, 2p+1) holds the complete sums base .. base+2 in v[0..2]
template <int MASK, int N, int SZ>
__device__ __forceinline__ void fold_step(float (&v)[SZ], int lane) {
  const bool up = lane & MASK;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float send = up ? v[k] : v[k + N / 2];
    const float keep = up ? v[k + N / 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, MASK);
  }
}
// SZ = 48: lane pair (2p, 2p+1) ends with the complete sums base .. base+2 in v[0..2]; SZ = 24: lane quad (4p .. 4p+3) does
__device__ __forceinline__ int fold48(float (&v)[48], int lane) {
  fold_step<16, 48, 48>(v, lane);
  fold_step<8, 24, 48>(v, lane);
  fold_step<4, 12, 48>(v, lane);
  fold_step<2, 6, 48>(v, lane);
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 1);
  return ((lane & 16) ? 24 : 0) + ((lane & 8) ? 12 : 0) + ((lane & 4) ? 6 : 0) + ((lane & 2) ? 3 : 0);
}
__device__ __forceinline__ int fold24(float (&v)[24], int lane) {
  fold_step<16, 24, 24>(v, lane);
  fold_step<8, 12, 24>(v, lane);
  fold_step<4, 6, 24>(v, lane);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 2);
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 1);
  }
  return ((lane & 16) ? 12 : 0) + ((lane & 8) ? 6 : 0) + ((lane & 4) ? 3 : 0);
}

constexpr int VST = 52;            // value-row stride in shared memory: conflict-free 128-bit loads at lane stride 52 floats
// One CTA = one sequence; the 48-wide value rows staged in shared memory (NonlinAttention: x_mid * tanh(s) formed while
// staging).  One warp = one task, lanes over the keys, 48 accumulators per lane, folded across the warp at the end:
//   SelfAttention (NL = false): task = (head, four query rows) -> 4 x 12 outputs;
//   NonlinAttention (NL = true): task = one query row -> 48 outputs of head 0, times the y gate.
template <bool NL, int JJ>
__global__ void __launch_bounds__(256) attn_apply_kernel(const float* __restrict__ aw, SeqMap sm, const float* __restrict__ src,
                                                        float* __restrict__ out) {
  extern __shared__ float sh[];
  const int S = sm.S;
  const long long n = blockIdx.x;
  for (int idx = threadIdx.x; idx < S * 12; idx += 256) {
    const int s = idx / 12, e = idx - s * 12;
    float4 v;
    if (NL) {
      const float4* pj = reinterpret_cast<const float4*>(src + sm.tok(n, s) * (3 * NH));
      const float4 g = __ldg(pj + e), m = __ldg(pj + NH / 4 + e);
      v = make_float4(m.x * tanhf(g.x), m.y * tanhf(g.y), m.z * tanhf(g.z), m.w * tanhf(g.w));
    } else {
      v = __ldg(reinterpret_cast<const float4*>(src + sm.tok(n, s) * SV) + e);
    }
    *reinterpret_cast<float4*>(sh + s * VST + 4 * e) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NRS = 4;                         // SelfAttention rows per task (2 rows / 24 accumulators raised occupancy but measured slower: 13.4 -> 14.8 ms)
  constexpr int NACC = NL ? 48 : NRS * 12;
  const int ntask = NL ? S : HEADS * ((S + NRS - 1) / NRS);
  for (int task = warp; task < ntask; task += 8) {
    const int h = NL ? 0 : task % HEADS;
    const int i0 = NL ? task : (task / HEADS) * NRS;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const int ld = aw_ld(S);
    const float* arow = aw + ((n * HEADS + h) * S) * (long long)ld;
#pragma unroll
    for (int jj = 0; jj < JJ; ++jj) {
      const int j = lane + 32 * jj;
      if (j < S) {
        if (NL) {
          const float a = __ldg(arow + (long long)i0 * ld + j);
#pragma unroll
          for (int e = 0; e < 12; ++e) {
            const float4 v = *reinterpret_cast<const float4*>(sh + j * VST + 4 * e);
            acc[4 * e] += a * v.x; acc[4 * e + 1] += a * v.y; acc[4 * e + 2] += a * v.z; acc[4 * e + 3] += a * v.w;
          }
        } else {
          float v[12];
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            const float4 t = *reinterpret_cast<const float4*>(sh + j * VST + h * VD + 4 * e);
            v[4 * e] = t.x; v[4 * e + 1] = t.y; v[4 * e + 2] = t.z; v[4 * e + 3] = t.w;
          }
#pragma unroll
          for (int r = 0; r < NRS; ++r) {
            const float a = __ldg(arow + (long long)min(i0 + r, S - 1) * ld + j);
#pragma unroll
            for (int c = 0; c < 12; ++c) acc[r * 12 + c] += a * v[c];
          }
        }
      }
    }
    int base;
    bool writer;
    if constexpr (NACC == 48) { base = fold48(acc, lane); writer = !(lane & 1); }
    else { base = fold24(acc, lane); writer = !(lane & 3); }
    if (writer) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int idx = base + k;
        int i, c;
        if (NL) { i = i0; c = idx; }
        else { i = i0 + idx / 12; c = h * VD + idx % 12; }
        if (i < S) {
          const long long t = sm.tok(n, i);
          float v = acc[k];
          if (NL) v *= __ldg(src + t * (3 * NH) + 2 * NH + c);
          out[t * SV + c] = v;
        }
      }
    }
  }
}

// Tensor-core variant of the two value products: out = AW (S x S, fp32 in HBM) . V (S x 8 per n-tile) with mma.sync m16n8k8 on
// tf32 operands, 3xTF32 (hi.hi + lo.hi + hi.lo) for fp32-class accuracy.  The products are bound by streaming the attention weights,
// not by math, so the legacy warp-level MMA is the right tool: the A fragments are loaded straight from global memory in the
// fragment layout (rows g / g+8, columns t / t+4 of the 16 x 8 block) and split in registers; V sits in shared memory as tf32 hi / lo
// planes with a row stride of 56 floats (bank = 24 t + g: conflict-free B-fragment loads).  About a quarter of the instructions of
// the FFMA kernel above (no 48-value warp fold).  One CTA = one sequence, one warp task = (head, 16 query rows) for SelfAttention
// (two 8-column tiles: 12 of 16 columns used) or (16 query rows, half of the 48 channels) for NonlinAttention.
constexpr int VSM = 56;
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool NL>
__global__ void __launch_bounds__(256, 3) attn_apply_mma_kernel(const float* __restrict__ aw, SeqMap sm, const float* __restrict__ src,
                                                            float* __restrict__ out) {
  extern __shared__ float sh[];
  const int S = sm.S, SP16 = (S + 15) & ~15, ld = aw_ld(S);
  float* vsm = sh;                    // [SP16][VSM] fp32 values, rows permuted inside every block of 16 (see below); split into tf32
                                      // hi / lo after the fragment load: ONE plane halves the shared memory (78.8 -> 39.4 KB at
                                      // S = 161), i.e. twice the resident CTAs for a kernel that ncu shows waiting on its global
                                      // loads at 25 % occupancy (63 % of the stall samples on the long scoreboard)
  const long long n = blockIdx.x;
  for (int idx = threadIdx.x; idx < SP16 * (VSM / 4); idx += 256) {
    const int s = idx / (VSM / 4), e = idx - s * (VSM / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < S && e < 12) {
      if (NL) {
        const float4* pj = reinterpret_cast<const float4*>(src + sm.tok(n, s) * (3 * NH));
        const float4 g = __ldg(pj + e), m = __ldg(pj + NH / 4 + e);
        v = make_float4(m.x * tanhf(g.x), m.y * tanhf(g.y), m.z * tanhf(g.z), m.w * tanhf(g.w));
      } else {
        v = __ldg(reinterpret_cast<const float4*>(src + sm.tok(n, s) * SV) + e);
      }
    }
    // key j = 16 b + 4 t + e lives in row 16 b + 4 e + t: lane t of an MMA reads keys 4t .. 4t+3 of a 16-key block (one 16-byte
    // load of the weights), and this placement keeps its four B-fragment loads on bank 24 t + g
    const int ps = (s & ~15) + ((s & 3) << 2) + ((s >> 2) & 3);
    *reinterpret_cast<float4*>(vsm + ps * VSM + 4 * e) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int row_tiles = (S + 15) / 16;
  constexpr int NT = NL ? 3 : 2;                 // 8-column tiles per task
  const int ntask = (NL ? 2 : HEADS) * row_tiles;
  for (int task = warp; task < ntask; task += 8) {
    const int part = task % (NL ? 2 : HEADS);    // head (SelfAttention) or channel half (NonlinAttention)
    const int i0 = (task / (NL ? 2 : HEADS)) * 16;
    const int h = NL ? 0 : part;
    const int c0 = NL ? part * 24 : part * VD;   // first channel of this task
    const float* a_lo_row = aw + ((n * HEADS + h) * S + min(i0 + g, S - 1)) * (long long)ld;
    const float* a_hi_row = aw + ((n * HEADS + h) * S + min(i0 + g + 8, S - 1)) * (long long)ld;
    float acc[NT][4];
#pragma unroll
    for (int q = 0; q < NT; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
#pragma unroll 2
    for (int k16 = 0; k16 < SP16; k16 += 16) {
      // the sum over keys may take them in any order: MMA step s of this block uses keys 4t + 2s (k index t) and 4t + 2s + 1
      // (k index t + 4), so each lane's share of the weights is ONE aligned 16-byte load per row (pad columns hold zeros)
      const bool in = k16 + 4 * t < ld;
      const float4 r0 = in ? __ldg(reinterpret_cast<const float4*>(a_lo_row + k16) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 r1 = in ? __ldg(reinterpret_cast<const float4*>(a_hi_row + k16) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float av[2][4] = {{r0.x, r1.x, r0.y, r1.y}, {r0.z, r1.z, r0.w, r1.w}};
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float hh, ll;
          split_tf32(av[st][e], hh, ll);
          ah[e] = __float_as_uint(hh); al[e] = __float_as_uint(ll);
        }
        const int p0 = (k16 + 8 * st + t) * VSM, p1 = p0 + 4 * VSM;      // rows of keys 4t + 2 st and 4t + 2 st + 1
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          const int col = c0 + 8 * q + g;
          float h0, l0, h1, l1;
          split_tf32(vsm[p0 + col], h0, l0);
          split_tf32(vsm[p1 + col], h1, l1);
          const uint32_t bh0 = __float_as_uint(h0), bh1 = __float_as_uint(h1), bl0 = __float_as_uint(l0), bl1 = __float_as_uint(l1);
          mma_tf32(acc[q], al, bh0, bh1);
          mma_tf32(acc[q], ah, bl0, bl1);
          mma_tf32(acc[q], ah, bh0, bh1);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NT; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + g + (e >> 1) * 8;
        const int cc = 8 * q + 2 * t + (e & 1);          // column within the task
        if (i >= S || (!NL && cc >= VD)) continue;
        const int c = c0 + cc;
        const long long tk = sm.tok(n, i);
        float v = acc[q][e];
        if (NL) v *= __ldg(src + tk * (3 * NH) + 2 * NH + c);
        out[tk * SV + c] = v;
      }
  }
}

// ------------------------------------------------------------------------------------------------ gated depthwise conv
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + ex2_ftz(x * -1.4426950408889634f)); }
// softplus(x - o) - 0.08 x = max(y, 0) + ln 2 * lg2(1 + 2^(-|y| log2 e)) - 0.08 x: the log argument is in (1, 2], abs error ~1e-7
__device__ __forceinline__ float fast_swoosh(float x, float o) {
  const float y = x - o;
  return fmaf(lg2_ftz(1.0f + ex2_ftz(fabsf(y) * -1.4426950408889634f)), 0.69314718055994531f, fmaxf(y, 0.f)) - 0.08f * x;
}

// One thread = one (sequence, channel) strip walking the sequence: u = x_mid * sigmoid(gate) computed once per position and
// kept in a 15-deep register ring, so the fused projection is read once and only the tf32 planes of the SwooshR output are
// written.  Adjacent threads = adjacent channels (coalesced rows of 64 floats).
__global__ void __launch_bounds__(256) glu_dwconv_kernel(const float* __restrict__ cp, SeqMap sm, long long nseq,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c = (int)(idx % C);
  const long long n = idx / C;
  if (n >= nseq) return;
  const int S = sm.S;
  float wk[DWK], ring[DWK];
#pragma unroll
  for (int k = 0; k < DWK; ++k) wk[k] = __ldg(w + c * DWK + k);
  const float bias = __ldg(b + c);
  // (fast_sigmoid / fast_swoosh: the approximations the GEMM epilogues use; the accurate expf / log1pf / IEEE division made
  // this kernel instruction-bound -- ncu: 61 % SM throughput at 22 % of HBM bandwidth)
  const long long tok0 = sm.tok(n, 0);
  auto u_at = [&](int sj) -> float {
    if (sj < 0 || sj >= S) return 0.f;
    const float* pj = cp + (tok0 + (long long)sj * sm.sS) * (2 * C);
    return __ldg(pj + c) * fast_sigmoid(__ldg(pj + C + c));
  };
  // slot (sj + 7) % 15 holds u[sj]; at step s the taps k read slots (s + k) % 15
#pragma unroll
  for (int k = 0; k < DWK; ++k) ring[k] = u_at(k - DWK / 2);
  for (int s0 = 0; s0 < S; s0 += DWK) {
#pragma unroll
    for (int d = 0; d < DWK; ++d) {
      const int s = s0 + d;
      if (s < S) {
        const float nxt = u_at(s + DWK / 2 + 1);
        float acc = bias;
#pragma unroll
        for (int k = 0; k < DWK; ++k) acc += wk[k] * ring[(d + k) % DWK];
        out[(tok0 + (long long)s * sm.sS) * C + c] = fast_swoosh(acc, 1.0f);
        ring[d] = nxt;
      }
    }
  }
}

// final BiasNorm + bypasses: half a warp per token row (float4 per lane)
__global__ void __launch_bounds__(256) norm_bypass_kernel(const float* __restrict__ x, float* __restrict__ x0, long long M,
                                                         const float* __restrict__ nbias, const float* __restrict__ nscale,
                                                         const float* __restrict__ rscale) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long row = idx >> 4;
  const int l = (int)(idx & 15);
  const long long rr = row < M ? row : M - 1;
  const float4 xv = *reinterpret_cast<const float4*>(x + rr * C + 4 * l);
  const float4 nb = __ldg(reinterpret_cast<const float4*>(nbias) + l);
  float ss = (xv.x - nb.x) * (xv.x - nb.x) + (xv.y - nb.y) * (xv.y - nb.y) + (xv.z - nb.z) * (xv.z - nb.z) + (xv.w - nb.w) * (xv.w - nb.w);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (row >= M) return;
  const float nrm = sqrtf(ss);
  const float4 ns = __ldg(reinterpret_cast<const float4*>(nscale) + l), rs = __ldg(reinterpret_cast<const float4*>(rscale) + l);
  const float4 o0 = *reinterpret_cast<const float4*>(x0 + row * C + 4 * l);
  *reinterpret_cast<float4*>(x0 + row * C + 4 * l) = make_float4((xv.x / nrm) * ns.x + o0.x * rs.x, (xv.y / nrm) * ns.y + o0.y * rs.y,
                                                                  (xv.z / nrm) * ns.z + o0.z * rs.z, (xv.w / nrm) * ns.w + o0.w * rs.w);
}

// InstanceNorm2d apply + PReLU: 16 threads per destination pixel, one float4 of channels each; pad columns of the destination
// grid are written as zeros.  POOL == 2 is the sub-pixel shuffle of the two decoders (output column 2 k' + u takes channel
// 2 c + u of source column k': the four channels of a thread are the even or the odd floats of eight consecutive ones).
template <int POOL>
__global__ void __launch_bounds__(256) in_apply_kernel(InApply f, int npix) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int c4 = idx & 15, pix = idx >> 4;
  if (pix >= npix) return;
  const int fd = pix % f.Wd, bt = pix / f.Wd, bb = bt / f.T;
  const int k = fd - f.dst_lo;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k >= 0 && k < f.nout) {
    float4 x;
    if (POOL == 1) {
      x = *reinterpret_cast<const float4*>(f.raw + ((long long)bt * f.Ws + f.src_lo + k) * f.ld + 4 * c4);
    } else {
      const float4* src = reinterpret_cast<const float4*>(f.raw + ((long long)bt * f.Ws + f.src_lo + (k >> 1)) * f.ld + 8 * c4);
      const float4 lo = src[0], hi = src[1];
      x = (k & 1) ? make_float4(lo.y, lo.w, hi.y, hi.w) : make_float4(lo.x, lo.z, hi.x, hi.z);
    }
    const float4* st = reinterpret_cast<const float4*>(f.stat + 2 * (bb * C + 4 * c4));
    const float4 s0 = st[0], s1 = st[1];           // (mean, rstd) x 4 channels
    const float4 w = __ldg(reinterpret_cast<const float4*>(f.w) + c4), b = __ldg(reinterpret_cast<const float4*>(f.b) + c4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(f.slope) + c4);
    v.x = (x.x - s0.x) * s0.y * w.x + b.x; v.y = (x.y - s0.z) * s0.w * w.y + b.y;
    v.z = (x.z - s1.x) * s1.y * w.z + b.z; v.w = (x.w - s1.z) * s1.w * w.w + b.w;
    v.x = v.x >= 0.f ? v.x : a.x * v.x; v.y = v.y >= 0.f ? v.y : a.y * v.y;
    v.z = v.z >= 0.f ? v.z : a.z * v.z; v.w = v.w >= 0.f ? v.w : a.w * v.w;
  }
  *reinterpret_cast<float4*>(f.of + (long long)pix * f.ldd + f.coff + 4 * c4) = v;
}

// InstanceNorm2d statistics, second stage: one WARP per (window, channel), lanes over the frames' double partial sums, butterfly
// reduction (fixed order: deterministic).  The functor walks the T frames with one thread: 4 096 threads in all, 89 us per launch.
__global__ void __launch_bounds__(256) in_fin_kernel(InFin f, long long n) {
  const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int cn = f.ld / f.pool;
  const int c = (int)(i % cn);
  const long long b = i / cn;
  double s = 0.0, s2 = 0.0;
  for (int t = lane; t < f.T; t += 32)
    for (int u = 0; u < f.pool; ++u) {
      const double2 v = *reinterpret_cast<const double2*>(f.part + 2 * ((b * f.T + t) * f.ld + c * f.pool + u));
      s += v.x; s2 += v.y;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if (lane == 0) {
    const double cnt = (double)f.T * f.nvalid * f.pool, mu = s / cnt;
    double var = s2 / cnt - mu * mu;
    var = var > 0.0 ? var : 0.0;
    f.stat[2 * i] = (float)mu;
    f.stat[2 * i + 1] = (float)(1.0 / sqrt(var + (double)IN_EPS));
  }
}

// FeatConv / UpCombine functors, four channels per thread (float4 rows instead of one 4-byte access per thread).
__global__ void __launch_bounds__(256) feat_conv_kernel(FeatConv f, long long n4) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  const int c4 = (int)(idx & 15);
  long long p = idx >> 4;
  const int fb = (int)(p % FB); p /= FB;
  const int t = (int)(p % f.T);
  const long long bb = p / f.T;
  const float* x = f.feat + (bb * 2 * f.T + t) * FB + fb;
  const float x0 = x[0], x1 = x[(long long)f.T * FB];
  const float4 wa = __ldg(reinterpret_cast<const float4*>(f.w) + 2 * c4), wb = __ldg(reinterpret_cast<const float4*>(f.w) + 2 * c4 + 1);
  const float4 b = __ldg(reinterpret_cast<const float4*>(f.b) + c4);
  *reinterpret_cast<float4*>(f.raw + ((bb * f.T + t) * FPE + fb + 1) * C + 4 * c4) =
      make_float4(b.x + wa.x * x0 + wa.y * x1, b.y + wa.z * x0 + wa.w * x1, b.z + wb.x * x0 + wb.y * x1, b.w + wb.z * x0 + wb.w * x1);
}
__global__ void __launch_bounds__(256) up_combine_kernel(UpCombine f, long long n4) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= n4) return;
  const int c4 = (int)(idx & 15);
  long long p = idx >> 4;
  const int fq = (int)(p % f.F); p /= f.F;
  const int t = (int)(p % f.T);
  const long long bb = p / f.T;
  const float4 y = *reinterpret_cast<const float4*>(f.y + ((bb * f.Td + t / f.ds) * f.Fd + fq / f.ds) * C + 4 * c4);
  const float4 sc = __ldg(reinterpret_cast<const float4*>(f.scale) + c4), rs = __ldg(reinterpret_cast<const float4*>(f.rscale) + c4);
  float4* o = reinterpret_cast<float4*>(f.x0) + idx;
  const float4 x = *o;
  *o = make_float4(x.x * rs.x + (y.x * sc.x), x.y * rs.y + (y.y * sc.y), x.z * rs.z + (y.z * sc.z), x.w * rs.w + (y.w * sc.w));
}

// decoder heads: one thread = one (window, frame, bin) with both taps' 2 x 64 inputs read as float4 and all outputs of the head
__global__ void __launch_bounds__(256) head_kernel(Head f, int npix) {
  __shared__ float4 ws[2 * 2 * C / 4];
  for (int i = threadIdx.x; i < f.nout * 2 * C / 4; i += 256) ws[i] = __ldg(reinterpret_cast<const float4*>(f.w) + i);
  __syncthreads();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= npix) return;
  const int fb = idx % FB, bt = idx / FB, t = bt % f.T, bb = bt / f.T;
  const float4* x = reinterpret_cast<const float4*>(f.up + ((long long)bt * FU + fb) * C);
  float acc[2] = {__ldg(f.b), f.nout > 1 ? __ldg(f.b + 1) : 0.f};
#pragma unroll 8
  for (int k = 0; k < 2 * C / 4; ++k) {
    const float4 v = x[k];
#pragma unroll
    for (int o = 0; o < 2; ++o)
      if (o < f.nout) {
        const float4 w = ws[o * (2 * C / 4) + k];
        acc[o] += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
      }
  }
  for (int o = 0; o < f.nout; ++o) f.out[(((long long)bb * f.nout + o) * f.T + t) * FB + fb] = acc[o];
}

__global__ void pad_split_w_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) split_tf32(src[i], hi[i], lo[i]);
}

// ------------------------------------------------------------------------------------------------ executor
struct PlanEntry {
  bool valid = false;
  LinOp key;
  tc::TcPlan plan;
  tc::TcArgs args;
};

static bool same_op(const LinOp& a, const LinOp& b) {
  bool s = a.a == b.a && a.a_sB == b.a_sB && a.a_sR == b.a_sR && a.a_r0 == b.a_r0 && a.a_ke == b.a_ke &&
           a.a_rows == b.a_rows && a.chunks == b.chunks && a.rows == b.rows && a.K == b.K && a.N == b.N && a.taps == b.taps &&
           a.tap_c == b.tap_c && a.a_k0 == b.a_k0 && a.W.w == b.W.w && a.W.b == b.W.b && a.act == b.act && a.resid == b.resid &&
           a.resid2 == b.resid2 && a.colscale == b.colscale && a.Cf == b.Cf && a.ldc == b.ldc;
  for (int i = 0; s && i < 6; ++i) s = a.tap_shift[i] == b.tap_shift[i];
  return s;
}

struct WPlanes { float *hi = nullptr, *lo = nullptr; };

struct CudaExec {
  cudaStream_t st = nullptr;
  int sms = 148;
  int launches = 0, gemms = 0;
  bool functors_only = false;       // ADN_ZIP_FUNCTORS=1: run the plain functors instead of the cooperative kernels (debugging)
  ImplTickFn tick = nullptr;
  void* tick_ctx = nullptr;
  bool capture = false;
  std::map<std::string, std::vector<float>>* dumps = nullptr;
  std::vector<PlanEntry>* plans = nullptr;
  std::map<const float*, WPlanes>* wplanes = nullptr;
  std::string* err = nullptr;
  bool failed = false;

  void done(const char* name) {
    ++launches;
    if (tick) tick(tick_ctx, name);
  }
  template <class F>
  void run_functor(long long n, const F& f) {
    if (n <= 0) return;
    op_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f);
    done(OpName<F>::get());
  }
  template <class F>
  void run(long long n, const F& f) { run_functor(n, f); }

  static int jj_of(int S) { return S <= 64 ? 2 : S <= 128 ? 4 : S <= 192 ? 6 : S <= 256 ? 8 : 0; }

  void run(long long n, const AttnW& f) {
    const int S = f.sm.S, jj = jj_of(S);
    if (functors_only || !jj) { run_functor(n, f); return; }
    const long long nseq = n / ((long long)HEADS * S);
    const size_t smem = ((size_t)QD * ((S + 31) & ~31) + 16 * S + (size_t)PD * 2 * S) * sizeof(float);
    const unsigned grid = (unsigned)(nseq * HEADS);
    static const int nr = getenv("ADN_ZIP_AW_ROWS") ? atoi(getenv("ADN_ZIP_AW_ROWS")) : 2;   // 2 rows per pass: 64-80 registers, 3-4 CTAs per SM (9.8 -> 7.7 ms per step vs 4 rows)
    if (nr == 2) {
      if (jj == 2) attn_w_kernel<2, 2><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
      else if (jj == 4) attn_w_kernel<4, 2><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
      else if (jj == 6) attn_w_kernel<6, 2><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
      else attn_w_kernel<8, 2><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
    } else {
      if (jj == 2) attn_w_kernel<2, 4><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
      else if (jj == 4) attn_w_kernel<4, 4><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
      else if (jj == 6) attn_w_kernel<6, 4><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
      else attn_w_kernel<8, 4><<<grid, 256, smem, st>>>(f.ap, f.sm, f.pos, f.aw);
    }
    done("zip_attn_w");
  }
  template <bool NL>
  void apply(long long nseq, const SeqMap& sm, const float* aw, const float* src, float* out, int jj) {
    static const bool use_mma = !(getenv("ADN_ZIP_APPLY") && !strcmp(getenv("ADN_ZIP_APPLY"), "ffma"));
    if (use_mma) {
      const size_t smem_m = (size_t)((sm.S + 15) & ~15) * VSM * sizeof(float);       // <= 256 * 56 * 4 = 57 344 B
      static unsigned long long conf_m = 0;
      auto km = attn_apply_mma_kernel<NL>;
      if (adn_first_use_on_device(conf_m)) cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * VSM * 4);
      km<<<(unsigned)nseq, 256, smem_m, st>>>(aw, sm, src, out);
      return;
    }
    const size_t smem = (size_t)sm.S * VST * sizeof(float);     // <= 256 * 52 * 4 = 53 248 B
    static unsigned long long configured[4] = {0, 0, 0, 0};
#define ZIP_APPLY(J, slot)                                                                                                    \
  {                                                                                                                           \
    auto k = attn_apply_kernel<NL, J>;                                                                                        \
    if (adn_first_use_on_device(configured[slot])) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * VST * 4); \
    k<<<(unsigned)nseq, 256, smem, st>>>(aw, sm, src, out);                                                              \
  }
    if (jj == 2) ZIP_APPLY(2, 0) else if (jj == 4) ZIP_APPLY(4, 1) else if (jj == 6) ZIP_APPLY(6, 2) else ZIP_APPLY(8, 3)
#undef ZIP_APPLY
  }
  void run(long long n, const SaApply& f) {
    const int S = f.sm.S, jj = jj_of(S);
    if (functors_only || !jj) { run_functor(n, f); return; }
    apply<false>(n / ((long long)SV * S), f.sm, f.aw, f.v, f.out, jj);
    done("zip_sa_apply");
  }
  void run(long long n, const NlApply& f) {
    const int S = f.sm.S, jj = jj_of(S);
    if (functors_only || !jj) { run_functor(n, f); return; }
    apply<true>(n / ((long long)NH * S), f.sm, f.aw, f.np, f.out, jj);
    done("zip_nl_apply");
  }
  void run(long long n, const GluDwConv& f) {
    if (functors_only) { run_functor(n, f); return; }
    const long long nseq = n / ((long long)C * f.sm.S);
    glu_dwconv_kernel<<<(unsigned)((nseq * C + 255) / 256), 256, 0, st>>>(f.cp, f.sm, nseq, f.w, f.b, f.out);
    done("zip_glu_dwconv");
  }
  void run(long long n, const InFin& f) {
    if (functors_only) { run_functor(n, f); return; }
    in_fin_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(f, n);
    done("zip_in_fin");
  }
  void run(long long n, const InApply& f) {
    if (functors_only || f.pool > 2 || (f.ld & 3)) { run_functor(n, f); return; }
    const long long npix = n / C;
    if (f.pool == 1) in_apply_kernel<1><<<(unsigned)((npix * 16 + 255) / 256), 256, 0, st>>>(f, (int)npix);
    else in_apply_kernel<2><<<(unsigned)((npix * 16 + 255) / 256), 256, 0, st>>>(f, (int)npix);
    done("zip_in_apply");
  }
  void run(long long n, const Head& f) {
    if (functors_only) { run_functor(n, f); return; }
    const long long npix = n / f.nout;
    head_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(f, (int)npix);
    done("zip_head");
  }
  void run(long long n, const FeatConv& f) {
    if (functors_only) { run_functor(n, f); return; }
    feat_conv_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, n / 4);
    done("zip_feat_conv");
  }
  void run(long long n, const UpCombine& f) {
    if (functors_only) { run_functor(n, f); return; }
    up_combine_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f, n / 4);
    done("zip_up_combine");
  }
  void run(long long n, const NormBypass& f) {
    if (functors_only) { run_functor(n, f); return; }
    const long long M = n / C;
    norm_bypass_kernel<<<(unsigned)((M * 16 + 255) / 256), 256, 0, st>>>(f.x, f.x0, M, f.nbias, f.nscale, f.rscale);
    done("zip_norm_bypass");
  }

  bool build(PlanEntry& e, const LinOp& g) {
    auto it = wplanes->find(g.W.w);
    if (it == wplanes->end()) { *err = "zipenhancer: LinOp weight without operand planes"; return false; }
    // (no 176-wide tile here: the TMA-store epilogue writes 32-column boxes, so N tiles must be multiples of 32 columns)
    const int bn = g.N <= 64 ? 64 : g.N <= 128 ? 128 : 256;
    const int bt = g.rows >= 128 ? 128 : g.rows;
    e.plan = tc::TcPlan{};
    e.plan.bn = bn;
    e.plan.a_f32 = true;                     // fp32 activations, split into tf32 hi / lo tiles inside the GEMM
    const int batches = g.chunks;
    const long long sB = g.a_sB ? g.a_sB : (long long)g.a_rows * g.a_sR;
    if (!tc::make_row_map(&e.plan.map_a_hi, g.a, g.a_ke, g.a_rows, g.a_sR, batches, sB, bt, 1, *err) ||
        !tc::make_weight_map(&e.plan.map_w_hi, it->second.hi, g.W.k_pad, g.W.n_pad, bn, *err) ||
        !tc::make_weight_map(&e.plan.map_w_lo, it->second.lo, g.W.k_pad, g.W.n_pad, bn, *err))
      return false;
    if (!tc::make_store_map(&e.plan.map_c, g.Cf, g.N, g.rows, g.ldc, g.chunks, (long long)g.rows * g.ldc, *err)) return false;
    e.plan.map_a_lo = e.plan.map_a_hi;
    e.plan.map_w2_hi = e.plan.map_w_hi;
    e.plan.map_w2_lo = e.plan.map_w_lo;
    tc::TcArgs& a = e.args;
    a = tc::TcArgs{};
    a.bb = 1; a.bt = bt; a.tiles_per_chunk = (g.rows + 127) / 128; a.t0 = g.a_r0;
    a.B = g.chunks; a.TM = g.rows; a.N = g.N; a.K = g.K;
    a.m_tiles = g.chunks * a.tiles_per_chunk;
    a.taps = g.taps;
    if (g.taps > 0) {
      a.tap_kb = g.tap_c / 32; a.tap_k0 = g.a_k0;
      for (int i = 0; i < 6; ++i) a.tap_shift[i] = g.tap_shift[i];
    }
    a.C = g.Cf; a.ldc = g.ldc;
    a.bias = g.W.b; a.resid = g.resid; a.resid2 = g.resid2; a.colscale = g.colscale; a.act = g.act;
    { const char* pr = getenv("ADN_TC_PROBE"); a.probe = pr ? atoi(pr) : 0; }
    e.key = g;
    e.valid = true;
    return true;
  }
  void gemm(const LinOp& g, const char* name) {
    if (failed) return;
    const size_t idx = (size_t)gemms++;
    if (plans->size() <= idx) plans->resize(idx + 1);
    PlanEntry& e = (*plans)[idx];
    if (!e.valid || !same_op(e.key, g)) {
      if (!build(e, g)) { failed = true; return; }
    }
    cudaError_t ce = tc::launch(e.plan, e.args, EPI_LIN, sms, st);
    if (ce != cudaSuccess) { *err = std::string("zipenhancer gemm launch: ") + cudaGetErrorString(ce); failed = true; return; }
    done(name);
  }
  void mark(const char* name, const float* p, long long count) {
    if (!capture || !dumps) return;
    std::vector<float>& v = (*dumps)[name];
    v.resize((size_t)count);
    cudaStreamSynchronize(st);
    cudaMemcpy(v.data(), p, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost);
  }
  void mark_strided(const char* name, const float* src, long long pixels, int ld, int coff, int width) {
    if (!capture || !dumps) return;
    std::vector<float> all((size_t)pixels * ld);
    cudaStreamSynchronize(st);
    cudaMemcpy(all.data(), src, all.size() * sizeof(float), cudaMemcpyDeviceToHost);
    std::vector<float>& v = (*dumps)[name];
    v.resize((size_t)pixels * width);
    for (long long p = 0; p < pixels; ++p)
      for (int c = 0; c < width; ++c) v[p * width + c] = all[p * ld + coff + c];
  }
};

// ------------------------------------------------------------------------------------------------ model
class Model : public ModelImpl {
 public:
  int device = 0, sms = 148;
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, T = 0, Lout = 0;
  int ds[NENC] = {1, 2, 2, 1};
  float* d_blob = nullptr;
  std::map<std::string, TensorRef> index;
  Weights W;
  Workspace ws;
  adn_stft* stft = nullptr;
  std::vector<void*> allocs, wallocs;
  std::map<const float*, WPlanes> wplanes;
  std::map<int, std::vector<PlanEntry>> plans;      // per windows-in-pass
  int cap = 0, cap_sub = 0;
  float *xn = nullptr, *nf = nullptr, *spec = nullptr, *feat = nullptr, *mx = nullptr, *ri = nullptr, *spec2 = nullptr, *wave = nullptr;
  int stop_after = 0, last_launches = 0, last_batch = 0;
  bool functors_only = false;
  std::map<std::string, std::vector<float>> dumps;
  static constexpr int SUB = 64;    // windows per pass of the backbone (bounds the workspace: ~0.3 GB per window)

  ~Model() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    for (void* p : wallocs) cudaFree(p);
    if (stft) adn_stft_destroy(stft);
  }
  void free_ws() {
    adn_note_free();
    for (void* p : allocs) cudaFree(p);
    allocs.clear();
    plans.clear();
    cap = cap_sub = 0;
  }
  float* dalloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(float)) != cudaSuccess) { err = "zipenhancer: out of device memory for the workspace"; return nullptr; }
    allocs.push_back(p);
    return (float*)p;
  }
  bool add_planes(const LinW& w) {
    if (!w.w || wplanes.count(w.w)) return true;
    const long long n = (long long)w.n_pad * w.k_pad;
    float* p = nullptr;
    if (cudaMalloc((void**)&p, 2 * n * sizeof(float)) != cudaSuccess) { err = "zipenhancer: out of device memory (weights)"; return false; }
    wallocs.push_back(p);
    pad_split_w_kernel<<<(unsigned)((n + 255) / 256), 256>>>(w.w, p, p + n, n);
    wplanes[w.w] = WPlanes{p, p + n};
    return true;
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
    int nfft = 0, hop = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", L) || !geti("nfft", nfft) || !geti("hop_length", hop) || !gets("input_audio_dtype", sin) ||
        !gets("output_audio_dtype", sout))
      return false;
    if (nfft != 400 || hop != 100 || L < 400) { err = "zipenhancer needs nfft=400, hop_length=100, input_audio_length >= 400"; return false; }
    {   // the dimensions compiled into zipenh_ops.cuh; resampled I/O is not built for this family
      struct { const char* k; int v; } dims[] = {{"zip_channels", C}, {"zip_heads", HEADS}, {"zip_query_head_dim", QD}, {"zip_pos_head_dim", PD},
                                                 {"zip_value_head_dim", VD}, {"zip_ff_dim", FF2}, {"zip_conv_kernel", DWK}};
      for (auto& d : dims) {
        auto it = meta.find(d.k);
        if (it != meta.end() && !it->second.empty() && atoi(it->second.c_str()) != d.v) {
          err = std::string("zipenhancer: ") + d.k + " = " + it->second + " differs from the compiled value " + std::to_string(d.v);
          return false;
        }
      }
      for (const char* k : {"in_sample_rate", "out_sample_rate", "model_sample_rate"}) {
        auto it = meta.find(k);
        if (it != meta.end() && !it->second.empty() && atoi(it->second.c_str()) != 16000) { err = "zipenhancer runs at 16 kHz in, model and out sample rates"; return false; }
      }
      auto it = meta.find("zip_downsample");
      if (it != meta.end() && !it->second.empty()) {
        int k = 0;
        const char* p = it->second.c_str();
        while (*p && k < NENC) { ds[k++] = atoi(p); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
        if (k != NENC) { err = "zipenhancer: zip_downsample needs four factors"; return false; }
      }
      for (int k = 0; k < NENC; ++k)
        if (ds[k] < 1 || ds[k] > 2) { err = "zipenhancer: down-sampling factors must be 1 or 2"; return false; }
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = L / hop + 1;
    Lout = hop * (T - 1);
    auto lk = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || (expect && it->second.count != expect)) {
        if (err.empty()) err = std::string("weight blob: tensor '") + name + "' missing or wrong size";
        return nullptr;
      }
      return d_blob + it->second.offset;
    };
    err.clear();
    if (!bind(W, T, ds, lk)) { if (err.empty()) err = "zipenhancer: weight binding failed"; return false; }
    auto dense = [&](const DenseW& d) { for (int i = 0; i < DEPTH; ++i) if (!add_planes(d.conv[i])) return false; return true; };
    bool ok = dense(W.enc_dense) && dense(W.mask_dense) && dense(W.phase_dense) && add_planes(W.c2) && add_planes(W.mask_up) && add_planes(W.phase_up);
    for (int k = 0; ok && k < NENC; ++k)
      for (int dir = 0; ok && dir < 2; ++dir) {
        const LayerW& l = dir ? W.enc[k].t : W.enc[k].f;
        const LinW* all[] = {&l.attn_in, &l.ff1_in, &l.ff1_out, &l.nl_in, &l.nl_out, &l.sa1_in, &l.sa1_out, &l.cv1_in, &l.cv1_out,
                             &l.ff2_in, &l.ff2_out, &l.sa2_in, &l.sa2_out, &l.cv2_in, &l.cv2_out, &l.ff3_in, &l.ff3_out};
        for (const LinW* w : all) ok = ok && add_planes(*w);
      }
    if (!ok) return false;
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "zipenhancer: weight operand split failed"; return false; }
    auto host = [&](const char* name, size_t expect) -> const float* {
      auto it = index.find(name);
      if (it == index.end() || it->second.count != expect) { err = std::string("weight blob: tensor '") + name + "' missing or wrong size"; return nullptr; }
      return h_blob + it->second.offset;
    };
    const float* fwd = host("stft.fwd", (size_t)2 * FB * 400);
    const float* inv = host("stft.inv", (size_t)2 * FB * 400);
    const float* nrm = host("stft.norm", (size_t)Lout);
    if (!fwd || !inv || !nrm) return false;
    adn_stft_geom g;
    memset(&g, 0, sizeof(g));
    g.nfft = 400; g.hop = 100; g.center = 1; g.pad_reflect = 1; g.norm_multiply = 1;      // ZipEnhancer/STFT_Process.py:243-248, :294-295
    if (adn_stft_create(&stft, &g, fwd, inv, nrm, T, device) != ADN_OK) { err = std::string("zipenhancer: ") + adn_last_error(nullptr); return false; }
    const char* fo = getenv("ADN_ZIP_FUNCTORS");
    functors_only = fo && fo[0] == '1';
    return true;
  }
  bool ensure(int B) {
    if (B <= cap) return true;
    cudaDeviceSynchronize();
    free_ws();
    const int sub = B < SUB ? B : SUB;
    auto a = [&](size_t n) { return dalloc(n); };
    if (!alloc_ws(ws, sub, T, a)) return false;
    const size_t b = (size_t)B;
    if (!(xn = dalloc(b * L)) || !(nf = dalloc(b)) || !(spec = dalloc(b * 2 * FB * T)) || !(feat = dalloc(b * 2 * T * FB)) ||
        !(mx = dalloc(b * T * FB)) || !(ri = dalloc(b * 2 * T * FB)) || !(spec2 = dalloc(b * 2 * FB * T)) || !(wave = dalloc(b * Lout)))
      return false;
    cap = B;
    cap_sub = sub;
    return true;
  }
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);          // Export_ZipEnhancer.py:964-965
    in->dtype = in_dtype; in->channels = 1; in->length = L;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);
    out->dtype = out_dtype; out->channels = 1; out->length = Lout;
  }
  size_t workspace_bytes(int batch) override {
    const int sub = batch < SUB ? batch : SUB;
    return (ws_floats(sub, T) + (size_t)batch * (L + 1 + 9 * FB * T + Lout)) * sizeof(float);
  }
  int per_pass() const {
    int n = 1 + 3 + DEPTH * 4 + 4 + 1 + 2 * (DEPTH * 4 + 4 + 1);      // encoder, pad copy, two decoders
    for (int k = 0; k < NENC; ++k) n += 2 * 24 + (ds[k] > 1 ? 2 : 0);
    return n;
  }
  int launches(int batch) override { return 6 + ((batch + SUB - 1) / SUB) * per_pass(); }
  void set_stop_after(int n) override { stop_after = n; }

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    auto chk = [&](adn_status s, const char* what) {
      if (s != ADN_OK) err = std::string("zipenhancer ") + what + ": " + adn_last_error(nullptr);
      return s == ADN_OK;
    };
    auto tk = [&](const char* name) { if (tick) tick(tick_ctx, name); };
    // F32 / F16 inputs are in [-1, 1]: x 32768 (:820-821), folded into the RMS pass
    if (!chk(adn_rms_normalize(d_in, in_dtype, in_dtype == ADN_I16 ? 1.0f : 32768.0f, 1e-6f, xn, nf, B, L, L, st), "rms_normalize")) return ADN_ERR_CUDA;
    tk("rms_normalize");
    if (!chk(adn_stft_forward(stft, xn, spec, B, L, st), "stft")) return ADN_ERR_CUDA;
    tk("stft");
    if (!chk(adn_spec_features(ADN_FAMILY_ZIPENHANCER, spec, feat, nullptr, B, FB, T, st), "spec_features")) return ADN_ERR_CUDA;
    tk("spec_features");
    CudaExec ex;
    ex.st = st; ex.sms = sms; ex.tick = tick; ex.tick_ctx = tick_ctx; ex.functors_only = functors_only;
    ex.capture = stop_after != 0; ex.dumps = &dumps; ex.wplanes = &wplanes; ex.err = &err;
    if (ex.capture) {
      dumps.clear();
      ex.mark("feat", feat, (long long)B * 2 * T * FB);      // the phase branch-cut decisions of this run (see the oracle's zipenh_forward)
    }
    for (int b0 = 0; b0 < B; b0 += cap_sub) {
      const int nb = B - b0 < cap_sub ? B - b0 : cap_sub;
      ex.plans = &plans[nb];
      ex.gemms = 0;
      forward(ex, ws, W, feat + (size_t)b0 * 2 * T * FB, mx + (size_t)b0 * T * FB, ri + (size_t)b0 * 2 * T * FB, nb, T);
      if (ex.failed) return ADN_ERR_CUDA;
      ex.capture = false;                          // stage dumps cover the first pass only
    }
    last_launches = ex.launches + 6;
    if (!chk(adn_spec_recombine(ADN_FAMILY_ZIPENHANCER, mx, ri, nullptr, spec2, B, FB, T, st), "spec_recombine")) return ADN_ERR_CUDA;
    tk("spec_recombine");
    if (!chk(adn_stft_inverse(stft, spec2, wave, B, T, st), "istft")) return ADN_ERR_CUDA;
    tk("istft");
    if (!chk(adn_condition_output(ADN_FAMILY_ZIPENHANCER, wave, Lout, nf, 1, d_out, out_dtype, B, Lout, st), "condition_output")) return ADN_ERR_CUDA;
    tk("condition_output");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("zipenhancer run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    if (!last_batch) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    if (!strcmp(name, "launches")) {
      if (actual) *actual = 1;
      if (h_dst && count) h_dst[0] = (float)last_launches;
      return ADN_OK;
    }
    auto it = dumps.find(name);
    if (it == dumps.end()) { err = std::string("adn_debug_read: unknown tensor '") + name + "' (stage dumps need adn_debug_stop_after(m, -1) before the run)"; return ADN_ERR_INVALID; }
    if (actual) *actual = it->second.size();
    if (!h_dst) return ADN_OK;
    const size_t nc = count < it->second.size() ? count : it->second.size();
    memcpy(h_dst, it->second.data(), nc * sizeof(float));
    return ADN_OK;
  }
};

}  // namespace zip

ModelImpl* zipenh_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                         const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  zip::Model* m = new zip::Model();
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
