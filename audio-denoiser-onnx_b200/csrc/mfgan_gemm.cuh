// Batched strided FFMA GEMM for the MossFormerGAN-SE-16K operators that are contractions (csrc/mfgan_ops.cuh:
// Linear, SimLocal, SimCross, LinKV, Att, Conv2d, GateConvT, TaScores, TaAV).  `translate()` maps each functor to one or more GemmOp (pure index
// arithmetic, host code, checked on the CPU by tests/harness/mfgan_host.cpp through `gemm_ref`); the CUDA executor runs
// GemmOps on `gemm_kernel` (64 x 64 x 16 shared-memory tiles, 4 x 4 outputs per thread, fp32 FFMA) instead of the
// one-output-per-thread functor.
#pragma once
#include "mfgan_ops.cuh"

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

inline void launch_gemm(const GemmOp& g, cudaStream_t st) {
  const int tm = (g.M + GT - 1) / GT, tn = (g.N + GT - 1) / GT;
  dim3 grid((unsigned)(tm * tn), (unsigned)g.batch);
  gemm_kernel<<<grid, 256, 0, st>>>(g, tn);
}
#endif

}  // namespace gan
