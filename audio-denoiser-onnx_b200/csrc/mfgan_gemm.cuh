// Batched strided FFMA GEMM for the MossFormerGAN-SE-16K operators that are contractions (csrc/mfgan_ops.cuh:
// Linear, SimLocal, SimCross, LinKV, Att, Conv2d, GateConvT, TaScores, TaAV).  `translate()` maps each functor to one or more GemmOp (pure index
// arithmetic, host code, checked on the CPU by tests/harness/mfgan_host.cpp through `gemm_ref`); the CUDA executor runs
// GemmOps on `gemm_kernel` (64 x 64 x 16 shared-memory tiles, 4 x 4 outputs per thread, fp32 FFMA) instead of the
// one-output-per-thread functor.
#pragma once
#include "mfgan_ops.cuh"

#include <stdint.h>
#include <stdlib.h>

namespace gan {

enum { GEPI_ACT = 0, GEPI_ACC = 1, GEPI_RELU2 = 2 };

struct ConvGeom { int on, T, Fin, Fout, KT, KF, dil, sf, pf, Cin, ldi; };
struct GateGeom { int on, S, Q, ldu, ldv; const float* iv; };

// C[b](m, n) (op)= sum_k A[b](m, k) * B[b](k, n);  batch b = b1 * nb2 + b2, every operand with its own element strides
struct GemmOp {
  const float* A; long long a_b1, a_b2, a_m, a_k;
  const float* B; long long b_b1, b_b2, b_k, b_n;
  float* C; long long c_b1, c_b2, c_m, c_n;
  int batch, nb2, M, N, K;
  const float* stat;          // per-row (mean, rstd) applied to A on load (batch == 1)
  const float* bias; int act; const float* slope;
  int epi; float scale; int zero_diag;
  ConvGeom cv;                // cv.on: A is the im2col view of a channel-last map (k = (kt, kf, ci))
  GateGeom gt;                // gt.on: A(m = (n, q), k = (tap, ci)) = iu[n, q - tap, ci] * iv[n, q - tap, ci] (gate + ConvTranspose1d)
  // two-level K / N walks: offset(k) = (k / kin) * k1 + (k % kin) * k-stride (kin == 0: single level); same for n
  int a_kin, b_kin, b_nin, c_nin;
  long long a_k1, b_k1, b_n1, c_n1;
};

GAN_HD long long gemm_boff(long long s1, long long s2, int nb2, int b) { return (long long)(b / nb2) * s1 + (long long)(b % nb2) * s2; }

GAN_HD float gemm_load_a(const GemmOp& g, long long boff, int m, int k) {
  if (m >= g.M || k >= g.K) return 0.f;
  if (g.cv.on) {
    const ConvGeom& c = g.cv;
    const int tap = k / c.Cin, ci = k - tap * c.Cin, kt = tap / c.KF, kf = tap - kt * c.KF;
    const int fo = m % c.Fout; const long long bt = m / c.Fout; const int t = (int)(bt % c.T); const long long b = bt / c.T;
    const int ti = t - (c.KT - 1 - kt) * c.dil, fi = fo * c.sf + kf - c.pf;
    if (ti < 0 || fi < 0 || fi >= c.Fin) return 0.f;
    return g.A[((b * c.T + ti) * c.Fin + fi) * c.ldi + ci];
  }
  if (g.gt.on) {
    const int tap = k / UV, ci = k - tap * UV, q = m % g.gt.Q, s = q - tap;
    if (s < 0 || s >= g.gt.S) return 0.f;
    const long long row = (long long)(m / g.gt.Q) * g.gt.S + s;
    return g.gt.iv[row * g.gt.ldv + ci] * g.A[row * g.gt.ldu + ci];
  }
  const long long ko = g.a_kin ? (long long)(k / g.a_kin) * g.a_k1 + (long long)(k % g.a_kin) * g.a_k : (long long)k * g.a_k;
  float v = g.A[boff + (long long)m * g.a_m + ko];
  if (g.stat) v = (v - g.stat[2 * m]) * g.stat[2 * m + 1];
  return v;
}
GAN_HD float gemm_load_b(const GemmOp& g, long long boff, int k, int n) {
  if (k >= g.K || n >= g.N) return 0.f;
  const long long ko = g.b_kin ? (long long)(k / g.b_kin) * g.b_k1 + (long long)(k % g.b_kin) * g.b_k : (long long)k * g.b_k;
  const long long no = g.b_nin ? (long long)(n / g.b_nin) * g.b_n1 + (long long)(n % g.b_nin) * g.b_n : (long long)n * g.b_n;
  return g.B[boff + ko + no];
}
GAN_HD void gemm_store(const GemmOp& g, long long boff, int m, int n, float acc) {
  const long long no = g.c_nin ? (long long)(n / g.c_nin) * g.c_n1 + (long long)(n % g.c_nin) * g.c_n : (long long)n * g.c_n;
  float* c = g.C + boff + (long long)m * g.c_m + no;
  if (g.epi == GEPI_ACC) { *c += acc; return; }
  if (g.epi == GEPI_RELU2) {
    acc *= g.scale;
    acc = acc > 0.f ? acc : 0.f;
    *c = (g.zero_diag && m == n) ? 0.f : acc * acc;
    return;
  }
  if (g.bias) acc += g.bias[n];
  *c = act(acc, g.act, g.slope, n);
}

inline GemmOp gemm_blank() {
  GemmOp g;
  g.A = nullptr; g.a_b1 = g.a_b2 = g.a_m = g.a_k = 0;
  g.B = nullptr; g.b_b1 = g.b_b2 = g.b_k = g.b_n = 0;
  g.C = nullptr; g.c_b1 = g.c_b2 = g.c_m = g.c_n = 0;
  g.batch = 1; g.nb2 = 1; g.M = g.N = g.K = 0;
  g.stat = nullptr; g.bias = nullptr; g.act = ACT_NONE; g.slope = nullptr;
  g.epi = GEPI_ACT; g.scale = 1.f; g.zero_diag = 0;
  g.cv = ConvGeom{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  g.gt = GateGeom{0, 0, 0, 0, 0, nullptr};
  g.a_kin = g.b_kin = g.b_nin = g.c_nin = 0;
  g.a_k1 = g.b_k1 = g.b_n1 = g.c_n1 = 0;
  return g;
}

// ---- functor -> GemmOp(s); `count` is the functor's output count (what Exec::run receives)
inline int translate(const Linear& f, long long count, GemmOp* out) {
  GemmOp g = gemm_blank();
  g.A = f.in; g.a_m = f.ldi; g.a_k = 1;
  g.B = f.Wt; g.b_k = f.N; g.b_n = 1;
  g.C = f.out; g.c_m = f.ldo; g.c_n = 1;
  g.M = (int)(count / f.N); g.N = f.N; g.K = f.K;
  g.stat = f.stat; g.bias = f.bias; g.act = f.a; g.slope = f.slope;
  out[0] = g;
  return 1;
}
inline int translate(const SimLocal& f, long long count, GemmOp* out) {
  const int Q = f.Q;
  GemmOp g = gemm_blank();
  g.batch = (int)(count / ((long long)Q * Q));
  g.A = f.heads; g.a_b1 = (long long)Q * 4 * QK; g.a_m = 4 * QK; g.a_k = 1;
  g.B = f.heads + 2 * QK; g.b_b1 = (long long)Q * 4 * QK; g.b_k = 1; g.b_n = 4 * QK;
  g.C = f.A; g.c_b1 = (long long)Q * Q; g.c_m = Q; g.c_n = 1;
  g.M = Q; g.N = Q; g.K = QK;
  g.epi = GEPI_RELU2;
  out[0] = g;
  return 1;
}
inline int translate(const SimCross& f, long long count, GemmOp* out) {
  const int Q = f.Q, BT = f.BT;
  const int Bw = (int)(count / ((long long)Q * BT * BT));
  GemmOp g = gemm_blank();
  g.batch = Bw * Q; g.nb2 = Q;
  g.A = f.heads; g.a_b1 = (long long)BT * Q * 4 * QK; g.a_b2 = 4 * QK; g.a_m = (long long)Q * 4 * QK; g.a_k = 1;
  g.B = f.heads + 2 * QK; g.b_b1 = g.a_b1; g.b_b2 = 4 * QK; g.b_k = 1; g.b_n = (long long)Q * 4 * QK;
  g.C = f.Ac; g.c_b1 = (long long)Q * BT * BT; g.c_b2 = (long long)BT * BT; g.c_m = BT; g.c_n = 1;
  g.M = BT; g.N = BT; g.K = QK;
  g.epi = GEPI_RELU2; g.scale = f.scale; g.zero_diag = 1;
  out[0] = g;
  return 1;
}
inline int translate(const LinKV& f, long long count, GemmOp* out) {
  const int Q = f.Q;
  GemmOp g = gemm_blank();
  g.batch = (int)(count / ((long long)QK * HID));
  g.A = f.heads + 3 * QK; g.a_b1 = (long long)Q * 4 * QK; g.a_m = 1; g.a_k = 4 * QK;
  g.B = f.huv; g.b_b1 = (long long)Q * HUV; g.b_k = HUV; g.b_n = 1;
  g.C = f.kv; g.c_b1 = (long long)QK * HID; g.c_m = HID; g.c_n = 1;
  g.M = QK; g.N = HID; g.K = Q;
  out[0] = g;
  return 1;
}
inline int translate(const Att& f, long long count, GemmOp* out) {
  const int Q = f.Q, BT = f.BT;
  const int Nseq = (int)(count / ((long long)Q * HID)), Bw = Nseq / BT;
  GemmOp g = gemm_blank();                       // local: A[n] (Q x Q) . hs[n] (Q x 256)
  g.batch = Nseq;
  g.A = f.A; g.a_b1 = (long long)Q * Q; g.a_m = Q; g.a_k = 1;
  g.B = f.huv; g.b_b1 = (long long)Q * HUV; g.b_k = HUV; g.b_n = 1;
  g.C = f.att; g.c_b1 = (long long)Q * HID; g.c_m = HID; g.c_n = 1;
  g.M = Q; g.N = HID; g.K = Q;
  out[0] = g;
  g = gemm_blank();                              // cross-token: Ac[b, q] (BT x BT) . hs[b, :, q] (BT x 256), accumulated
  g.batch = Bw * Q; g.nb2 = Q;
  g.A = f.Ac; g.a_b1 = (long long)Q * BT * BT; g.a_b2 = (long long)BT * BT; g.a_m = BT; g.a_k = 1;
  g.B = f.huv; g.b_b1 = (long long)BT * Q * HUV; g.b_b2 = HUV; g.b_k = (long long)Q * HUV; g.b_n = 1;
  g.C = f.att; g.c_b1 = (long long)BT * Q * HID; g.c_b2 = HID; g.c_m = (long long)Q * HID; g.c_n = 1;
  g.M = BT; g.N = HID; g.K = BT;
  g.epi = GEPI_ACC;
  out[1] = g;
  g = gemm_blank();                              // linear: lin_q[n] (Q x 128) . kv[n] (128 x 256), accumulated
  g.batch = Nseq;
  g.A = f.heads + QK; g.a_b1 = (long long)Q * 4 * QK; g.a_m = 4 * QK; g.a_k = 1;
  g.B = f.kv; g.b_b1 = (long long)QK * HID; g.b_k = HID; g.b_n = 1;
  g.C = f.att; g.c_b1 = (long long)Q * HID; g.c_m = HID; g.c_n = 1;
  g.M = Q; g.N = HID; g.K = QK;
  g.epi = GEPI_ACC;
  out[2] = g;
  return 3;
}
inline int translate(const Conv2d& f, long long count, GemmOp* out) {
  GemmOp g = gemm_blank();
  g.A = f.in;
  g.cv = ConvGeom{1, f.T, f.Fin, f.Fout, f.KT, f.KF, f.dil, f.sf, f.pf, f.Cin, f.ldi};
  g.B = f.W; g.b_k = f.Cout; g.b_n = 1;
  g.C = f.out; g.c_m = f.ldo; g.c_n = 1;
  g.M = (int)(count / f.Cout); g.N = f.Cout; g.K = f.KT * f.KF * f.Cin;
  g.bias = f.bias;
  out[0] = g;
  return 1;
}

inline int translate(const GateConvT& f, long long count, GemmOp* out) {
  GemmOp g = gemm_blank();
  g.A = f.iu;
  g.gt = GateGeom{1, f.S, f.Q, f.ldu, f.ldv, f.iv};
  g.B = f.w; g.b_k = C; g.b_n = 1;
  g.C = f.out; g.c_m = C; g.c_n = 1;
  g.M = (int)(count / C); g.N = C; g.K = KS * UV;
  g.bias = f.b;
  out[0] = g;
  return 1;
}
inline int translate(const TaScores& f, long long count, GemmOp* out) {
  const int T = f.T, Fw = f.Fw;
  GemmOp g = gemm_blank();
  g.batch = (int)(count / ((long long)T * T)); g.nb2 = HEADS;
  g.A = f.qkv; g.a_b1 = (long long)T * Fw * QKV; g.a_b2 = AE; g.a_m = (long long)Fw * QKV; g.a_k = 1; g.a_kin = AE; g.a_k1 = QKV;
  g.B = f.qkv + HEADS * AE; g.b_b1 = g.a_b1; g.b_b2 = AE; g.b_n = (long long)Fw * QKV; g.b_k = 1; g.b_kin = AE; g.b_k1 = QKV;
  g.C = f.a; g.c_b1 = (long long)HEADS * T * T; g.c_b2 = (long long)T * T; g.c_m = T; g.c_n = 1;
  g.M = T; g.N = T; g.K = Fw * AE;
  out[0] = g;
  return 1;
}
inline int translate(const TaAV& f, long long count, GemmOp* out) {
  const int T = f.T, Fw = f.Fw;
  GemmOp g = gemm_blank();
  g.batch = (int)(count / ((long long)T * Fw * C)) * HEADS; g.nb2 = HEADS;
  g.A = f.a; g.a_b1 = (long long)HEADS * T * T; g.a_b2 = (long long)T * T; g.a_m = T; g.a_k = 1;
  g.B = f.qkv + 2 * HEADS * AE; g.b_b1 = (long long)T * Fw * QKV; g.b_b2 = VC; g.b_k = (long long)Fw * QKV; g.b_n = 1; g.b_nin = VC; g.b_n1 = QKV;
  g.C = f.out; g.c_b1 = (long long)T * Fw * C; g.c_b2 = VC; g.c_m = (long long)Fw * C; g.c_n = 1; g.c_nin = VC; g.c_n1 = C;
  g.M = T; g.N = Fw * VC; g.K = T;
  out[0] = g;
  return 1;
}

// plain-loop semantics of a GemmOp (host harness only)
inline void gemm_ref(const GemmOp& g) {
#pragma omp parallel for schedule(static)
  for (long long bm = 0; bm < (long long)g.batch * g.M; ++bm) {
    const int b = (int)(bm / g.M), m = (int)(bm % g.M);
    const long long ao = gemm_boff(g.a_b1, g.a_b2, g.nb2, b), bo = gemm_boff(g.b_b1, g.b_b2, g.nb2, b),
                    co = gemm_boff(g.c_b1, g.c_b2, g.nb2, b);
    for (int n = 0; n < g.N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < g.K; ++k) acc += gemm_load_a(g, ao, m, k) * gemm_load_b(g, bo, k, n);
      gemm_store(g, co, m, n, acc);
    }
  }
}

#if defined(__CUDACC__)
constexpr int GT = 64, GK = 16, GLD = GT + 4;

static __global__ void __launch_bounds__(256) gemm_kernel(const GemmOp g, const int tiles_n) {
  __shared__ __align__(16) float As[2][GK][GLD];
  __shared__ __align__(16) float Bs[2][GK][GLD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.y;
  const int m0 = (blockIdx.x / tiles_n) * GT, n0 = (blockIdx.x % tiles_n) * GT;
  const long long ao = gemm_boff(g.a_b1, g.a_b2, g.nb2, b), bo = gemm_boff(g.b_b1, g.b_b2, g.nb2, b),
                  co = gemm_boff(g.c_b1, g.c_b2, g.nb2, b);
  const bool a_kfast = g.cv.on || g.gt.on || g.a_k == 1, b_nfast = g.b_n == 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + 256 * i;
    if (a_kfast) { ak[i] = e & (GK - 1); am[i] = e >> 4; } else { am[i] = e & (GT - 1); ak[i] = e >> 6; }
    if (b_nfast) { bn[i] = e & (GT - 1); bk[i] = e >> 6; } else { bk[i] = e & (GK - 1); bn[i] = e >> 4; }
  }
  // im2col view: the (window, frame, sub-band) of each of this thread's A rows is fixed over the K loop
  long long cbt[4];
  int ct[4], cf[4];
  if (g.cv.on) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + am[i];
      const int fo = m % g.cv.Fout;
      const long long bt = m / g.cv.Fout;
      ct[i] = m < g.M ? (int)(bt % g.cv.T) : -(1 << 28);       // rows past M read as zeros (frame < 0)
      cbt[i] = (bt / g.cv.T) * g.cv.T;
      cf[i] = fo * g.cv.sf - g.cv.pf;
    }
  }
  float ra[4], rb[4];
  auto gload = [&](int k0) {
    if (g.cv.on) {                                 // Cin % GK == 0: one tap per K tile
      const int tap = k0 / g.cv.Cin, ci0 = k0 - tap * g.cv.Cin, kt = tap / g.cv.KF, kf = tap - kt * g.cv.KF;
      const int dt = (g.cv.KT - 1 - kt) * g.cv.dil;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ti = ct[i] - dt, fi = cf[i] + kf;
        float v = 0.f;
        if (ti >= 0 && fi >= 0 && fi < g.cv.Fin && k0 + ak[i] < g.K)
          v = g.A[((cbt[i] + ti) * g.cv.Fin + fi) * g.cv.ldi + ci0 + ak[i]];
        ra[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = gemm_load_a(g, ao, m0 + am[i], k0 + ak[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) rb[i] = gemm_load_b(g, bo, k0 + bk[i], n0 + bn[i]);
  };
  auto sstore = [&](int st) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[st][ak[i]][am[i]] = ra[i]; Bs[st][bk[i]][bn[i]] = rb[i]; }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int st = 0;
  for (int k0 = 0; k0 < g.K; k0 += GK) {
    const bool next = k0 + GK < g.K;
    if (next) gload(k0 + GK);                      // global loads of the next tile fly under this tile's FMAs
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (next) sstore(st ^ 1);
    __syncthreads();
    st ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.N) gemm_store(g, co, m, n, acc[i][j]);
    }
  }
}

// The same tile on the warp-level tensor-core path: mma.sync m16n8k8 (tf32 operands, fp32 accumulate) with the 3xTF32 split
// (hi * hi + hi * lo + lo * hi) so the products keep fp32-class accuracy.  Operands arrive through the same generic strided
// loads (im2col / gate views included) into fp32 tiles; fragments are split in registers (hi = low 13 mantissa bits cleared,
// lo = x - hi): the shared-memory pipe, not the ALU, is what this kernel runs out of (ncu: l1tex 62-71 %).  8 warps = 4 (m) x 2 (n); a warp owns 16 x 32 outputs = four n8
// fragments; the leading dimension 72 makes every fragment load conflict-free (bank = 8 * (lane & 3) + (lane >> 2)).
// For the batched per-sequence products whose strided / transposed operands the TMA row maps of gemm_tc.cu cannot address.
constexpr int MLD = GT + 8;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// SIMPLE: plain single-level strides on both operands (every product but the im2col / gate / two-level views): the row base
// pointers are formed once, a K step costs one predicated load per element (ncu on the generic path: 800 instructions per warp
// and K tile, 24 of them MMAs; long-scoreboard stalls at 24 warps per SM)
template <bool SIMPLE>
static __global__ void __launch_bounds__(256, 3) gemm_mma_kernel(const GemmOp g, const int tiles_n) {
  // tile layouts follow the operand's fast axis so that BOTH the staging stores and the fragment loads are conflict-free:
  // [k][m] with leading dimension 72 when consecutive lanes hold consecutive m (or n), [m][k] with leading dimension 20 when
  // they hold consecutive k (fragment load bank = 20 * (lane >> 2) + (lane & 3) mod 32: all distinct)
  constexpr int TSZ = GT * (GK + 4);
  __shared__ __align__(16) float As[2][TSZ], Bs[2][TSZ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1, gq = lane >> 2, tq = lane & 3;
  const int b = blockIdx.y;
  const int m0 = (blockIdx.x / tiles_n) * GT, n0 = (blockIdx.x % tiles_n) * GT;
  const long long ao = gemm_boff(g.a_b1, g.a_b2, g.nb2, b), bo = gemm_boff(g.b_b1, g.b_b2, g.nb2, b),
                  co = gemm_boff(g.c_b1, g.c_b2, g.nb2, b);
  const bool a_kfast = g.cv.on || g.gt.on || g.a_k == 1, b_nfast = g.b_n == 1;
  const int sAm = a_kfast ? GK + 4 : 1, sAk = a_kfast ? 1 : MLD, sBn = b_nfast ? 1 : GK + 4, sBk = b_nfast ? MLD : 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + 256 * i;
    if (a_kfast) { ak[i] = e & (GK - 1); am[i] = e >> 4; } else { am[i] = e & (GT - 1); ak[i] = e >> 6; }
    if (b_nfast) { bn[i] = e & (GT - 1); bk[i] = e >> 6; } else { bk[i] = e & (GK - 1); bn[i] = e >> 4; }
  }
  long long cbt[4];
  int ct[4], cf[4];
  const float* pa[4];
  const float* pb[4];
  float smean[4], srstd[4];
  if (SIMPLE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + am[i], n = n0 + bn[i];
      pa[i] = m < g.M ? g.A + ao + (long long)m * g.a_m + (long long)ak[i] * g.a_k : nullptr;
      pb[i] = n < g.N ? g.B + bo + (long long)n * g.b_n + (long long)bk[i] * g.b_k : nullptr;
      smean[i] = 0.f; srstd[i] = 1.f;
      if (g.stat && m < g.M) { smean[i] = g.stat[2 * m]; srstd[i] = g.stat[2 * m + 1]; }
    }
  } else if (g.cv.on) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + am[i];
      const int fo = m % g.cv.Fout;
      const long long bt = m / g.cv.Fout;
      ct[i] = m < g.M ? (int)(bt % g.cv.T) : -(1 << 28);
      cbt[i] = (bt / g.cv.T) * g.cv.T;
      cf[i] = fo * g.cv.sf - g.cv.pf;
    }
  }
  // 16-byte operand loads where the fast axis is contiguous and aligned: the A tile (64 m x 16 k) is one float4 along k per
  // thread (m = tid / 4), the B tile (16 k x 64 n) one float4 along n per thread (k = tid / 16); the thread's four staging
  // slots then are those four consecutive elements
  bool a_vec = false, b_vec = false;
  if (SIMPLE) {
    a_vec = a_kfast && !g.stat && g.a_m % 4 == 0 && g.K % 4 == 0 && ((((uintptr_t)(g.A + ao)) & 15) == 0);
    b_vec = b_nfast && g.b_k % 4 == 0 && g.N % 4 == 0 && ((((uintptr_t)(g.B + bo)) & 15) == 0);
    if (a_vec) {
      const int m = m0 + (tid >> 2), k4 = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) { am[i] = tid >> 2; ak[i] = k4 + i; }
      pa[0] = m < g.M ? g.A + ao + (long long)m * g.a_m + k4 : nullptr;
    }
    if (b_vec) {
      const int n4 = (tid & 15) * 4, k = tid >> 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) { bn[i] = n4 + i; bk[i] = k; }
      pb[0] = n0 + n4 < g.N ? g.B + bo + (long long)k * g.b_k + n0 + n4 : nullptr;
    }
  }
  float ra[4], rb[4];
  auto gload = [&](int k0) {
    if (SIMPLE) {
      const long long oa = (long long)k0 * g.a_k, ob = (long long)k0 * g.b_k;
      if (a_vec) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pa[0] && k0 + ak[0] < g.K) v = *reinterpret_cast<const float4*>(pa[0] + oa);
        ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ra[i] = (pa[i] && k0 + ak[i] < g.K) ? (pa[i][oa] - smean[i]) * srstd[i] : 0.f;
      }
      if (b_vec) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pb[0] && k0 + bk[0] < g.K) v = *reinterpret_cast<const float4*>(pb[0] + ob);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) rb[i] = (pb[i] && k0 + bk[i] < g.K) ? pb[i][ob] : 0.f;
      }
      return;
    }
    if (g.cv.on) {
      const int tap = k0 / g.cv.Cin, ci0 = k0 - tap * g.cv.Cin, kt = tap / g.cv.KF, kf = tap - kt * g.cv.KF;
      const int dt = (g.cv.KT - 1 - kt) * g.cv.dil;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ti = ct[i] - dt, fi = cf[i] + kf;
        float v = 0.f;
        if (ti >= 0 && fi >= 0 && fi < g.cv.Fin && k0 + ak[i] < g.K)
          v = g.A[((cbt[i] + ti) * g.cv.Fin + fi) * g.cv.ldi + ci0 + ak[i]];
        ra[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = gemm_load_a(g, ao, m0 + am[i], k0 + ak[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) rb[i] = gemm_load_b(g, bo, k0 + bk[i], n0 + bn[i]);
  };
  auto sstore = [&](int st) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[st][am[i] * sAm + ak[i] * sAk] = ra[i];
      Bs[st][bn[i] * sBn + bk[i] * sBk] = rb[i];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int st = 0;
  for (int k0 = 0; k0 < g.K; k0 += GK) {
    const bool next = k0 + GK < g.K;
    if (next) gload(k0 + GK);
#pragma unroll
    for (int k8 = 0; k8 < GK; k8 += 8) {
      // fp32 fragments from shared memory, split in registers: hi = x with the 13 low mantissa bits cleared, lo = x - hi
      uint32_t ah[4], al[4];
      const int a00 = (wm * 16 + gq) * sAm + (k8 + tq) * sAk, a10 = a00 + 8 * sAm, a01 = a00 + 4 * sAk, a11 = a10 + 4 * sAk;
      const float ax[4] = {As[st][a00], As[st][a10], As[st][a01], As[st][a11]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = __float_as_uint(ax[i]) & 0xFFFFE000u;
        al[i] = __float_as_uint(ax[i] - __uint_as_float(ah[i]));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int b0i = (wn * 32 + nt * 8 + gq) * sBn + (k8 + tq) * sBk, b1i = b0i + 4 * sBk;
        const float bx0 = Bs[st][b0i], bx1 = Bs[st][b1i];
        const uint32_t bh0 = __float_as_uint(bx0) & 0xFFFFE000u, bh1 = __float_as_uint(bx1) & 0xFFFFE000u;
        const uint32_t bl0 = __float_as_uint(bx0 - __uint_as_float(bh0)), bl1 = __float_as_uint(bx1 - __uint_as_float(bh1));
        mma_tf32(acc[nt], al, bh0, bh1);
        mma_tf32(acc[nt], ah, bl0, bl1);
        mma_tf32(acc[nt], ah, bh0, bh1);
      }
    }
    if (next) sstore(st ^ 1);
    __syncthreads();
    st ^= 1;
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + wm * 16 + gq + (r >> 1) * 8, n = n0 + wn * 32 + nt * 8 + 2 * tq + (r & 1);
      if (m < g.M && n < g.N) gemm_store(g, co, m, n, acc[nt][r]);
    }
  }
}

// Plain-stride products without load-time normalisation: operands go global -> shared with cp.async (4-byte copies, 16-byte
// where the fast axis is contiguous and aligned; out-of-range elements are zero-filled by a zero source size) through a
// four-stage ring, so three K tiles of loads are in flight while one is multiplied.  (ncu on the register-prefetch kernel: a third
// of all stall samples sat on the first use of the next tile's loads -- one tile of prefetch does not cover L2 latency when
// a K tile is only 24 MMAs per warp.)
constexpr int NST = 4;
// (`safe`: any valid global address; it is what a zero-sized copy names)
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok, const float* safe) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = ok ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(ok ? src : safe), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool ok, const float* safe) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(ok ? src : safe), "r"(n) : "memory");
}

static __global__ void __launch_bounds__(256, 3) gemm_mma_async_kernel(const GemmOp g, const int tiles_n) {
  constexpr int TSZ = GT * (GK + 4);
  __shared__ __align__(16) float As[NST][TSZ], Bs[NST][TSZ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1, gq = lane >> 2, tq = lane & 3;
  const int b = blockIdx.y;
  const int m0 = (blockIdx.x / tiles_n) * GT, n0 = (blockIdx.x % tiles_n) * GT;
  const long long ao = gemm_boff(g.a_b1, g.a_b2, g.nb2, b), bo = gemm_boff(g.b_b1, g.b_b2, g.nb2, b),
                  co = gemm_boff(g.c_b1, g.c_b2, g.nb2, b);
  const bool a_kfast = g.a_k == 1, b_nfast = g.b_n == 1;
  const int sAm = a_kfast ? GK + 4 : 1, sAk = a_kfast ? 1 : MLD, sBn = b_nfast ? 1 : GK + 4, sBk = b_nfast ? MLD : 1;
  const bool a_vec = a_kfast && g.a_m % 4 == 0 && g.K % 4 == 0 && ((((uintptr_t)(g.A + ao)) & 15) == 0);
  const bool b_vec = b_nfast && g.b_k % 4 == 0 && g.N % 4 == 0 && ((((uintptr_t)(g.B + bo)) & 15) == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // this thread's copies: four scalars or one 16-byte vector per operand and K tile
  const float* pa[4];
  const float* pb[4];
  int ia[4], ib[4], ka[4], kb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + 256 * i;
    int am, ak, bn, bk;
    if (a_vec) { am = tid >> 2; ak = (tid & 3) * 4; }
    else if (a_kfast) { ak = e & (GK - 1); am = e >> 4; } else { am = e & (GT - 1); ak = e >> 6; }
    if (b_vec) { bn = (tid & 15) * 4; bk = tid >> 4; }
    else if (b_nfast) { bn = e & (GT - 1); bk = e >> 6; } else { bk = e & (GK - 1); bn = e >> 4; }
    const int m = m0 + am, n = n0 + bn;
    pa[i] = m < g.M ? g.A + ao + (long long)m * g.a_m + (long long)ak * g.a_k : nullptr;
    pb[i] = n < g.N ? g.B + bo + (long long)n * g.b_n + (long long)bk * g.b_k : nullptr;
    ia[i] = am * sAm + ak * sAk; ib[i] = bn * sBn + bk * sBk; ka[i] = ak; kb[i] = bk;
  }
  auto issue = [&](int k0, int st) {
    if (k0 < g.K) {
      const long long oa = (long long)k0 * g.a_k, ob = (long long)k0 * g.b_k;
      if (a_vec) cp_async16(&As[st][ia[0]], pa[0] + oa, pa[0] && k0 + ka[0] < g.K, g.A);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async4(&As[st][ia[i]], pa[i] + oa, pa[i] && k0 + ka[i] < g.K, g.A);
      }
      if (b_vec) cp_async16(&Bs[st][ib[0]], pb[0] + ob, pb[0] && k0 + kb[0] < g.K, g.B);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async4(&Bs[st][ib[i]], pb[i] + ob, pb[i] && k0 + kb[i] < g.K, g.B);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int s = 0; s < NST - 1; ++s) issue(s * GK, s);
  int st = 0;
  for (int k0 = 0; k0 < g.K; k0 += GK) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
    __syncthreads();                               // tile k0 has landed for every thread; the stage refilled below was read last iteration
    issue(k0 + (NST - 1) * GK, (st + NST - 1) % NST);
#pragma unroll
    for (int k8 = 0; k8 < GK; k8 += 8) {
      uint32_t ah[4], al[4];
      const int a00 = (wm * 16 + gq) * sAm + (k8 + tq) * sAk, a10 = a00 + 8 * sAm, a01 = a00 + 4 * sAk, a11 = a10 + 4 * sAk;
      const float ax[4] = {As[st][a00], As[st][a10], As[st][a01], As[st][a11]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = __float_as_uint(ax[i]) & 0xFFFFE000u;
        al[i] = __float_as_uint(ax[i] - __uint_as_float(ah[i]));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int b0i = (wn * 32 + nt * 8 + gq) * sBn + (k8 + tq) * sBk, b1i = b0i + 4 * sBk;
        const float bx0 = Bs[st][b0i], bx1 = Bs[st][b1i];
        const uint32_t bh0 = __float_as_uint(bx0) & 0xFFFFE000u, bh1 = __float_as_uint(bx1) & 0xFFFFE000u;
        const uint32_t bl0 = __float_as_uint(bx0 - __uint_as_float(bh0)), bl1 = __float_as_uint(bx1 - __uint_as_float(bh1));
        mma_tf32(acc[nt], al, bh0, bh1);
        mma_tf32(acc[nt], ah, bl0, bl1);
        mma_tf32(acc[nt], ah, bh0, bh1);
      }
    }
    st = (st + 1) % NST;
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + wm * 16 + gq + (r >> 1) * 8, n = n0 + wn * 32 + nt * 8 + 2 * tq + (r & 1);
      if (m < g.M && n < g.N) gemm_store(g, co, m, n, acc[nt][r]);
    }
  }
}

inline bool gemm_use_mma() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ADN_GAN_MMA"); on = !(e && e[0] == '0'); }
  return on != 0;
}

inline void launch_gemm(const GemmOp& g, cudaStream_t st) {
  const int tm = (g.M + GT - 1) / GT, tn = (g.N + GT - 1) / GT;
  dim3 grid((unsigned)(tm * tn), (unsigned)g.batch);
  const bool simple = !g.cv.on && !g.gt.on && !g.a_kin && !g.b_kin && !g.b_nin;
  if (!gemm_use_mma()) gemm_kernel<<<grid, 256, 0, st>>>(g, tn);
  else if (simple && !g.stat) gemm_mma_async_kernel<<<grid, 256, 0, st>>>(g, tn);
  else if (simple) gemm_mma_kernel<true><<<grid, 256, 0, st>>>(g, tn);
  else gemm_mma_kernel<false><<<grid, 256, 0, st>>>(g, tn);
}
#endif

}  // namespace gan
