// MossFormer2-SE-48K: Kaldi-fbank + STFT frontend, 24 x (FLASH attention block + gated FSMN block),
// mask tail and ISTFT (reference MossFormer2_SE_48K/Export_MossFormer_SE.py:312-507).
//
// Layout: activations are token-major fp32, row m = window*T + frame.  Every dense contraction
// (frontend DFT, 7 linear layers per block, the four attention products, mask tail, ISTFT
// overlap-add) runs on the tcgen05 3xTF32 GEMM (gemm_tc.cu); producers write the tf32 hi/lo
// operand planes directly.  The attention products are batched per window with the window's own
// keys / values as the "weight" operand:
//     S1  = Ql Kl^T                    A = lin_q  (T x 128)      W = lin_k    (T x 128)
//     S   = relu(Qq Kq^T)^2 + S1       A = quad_q (T x 128)      W = quad_k   (T x 128)
//     O   = S [v|u]                    A = S      (T x Tp)       W = [v|u]^T  (2048 x Tp)
// A window is one FLASH group (T <= 256 frames), so the reference's linear branch
// Ql (Kl^T [v|u]) (:422-427) is re-associated to (Ql Kl^T) [v|u] and shares the value product with
// the quadratic branch: T <= 256 makes the T x T form the cheaper one, and it removes two GEMMs and
// the 2048 x 128 per-window KV round trip.  The depthwise-conv kernel that follows the input
// projection emits [v|u]^T through a shared-memory transpose.  Zero-padded keys contribute exactly
// nothing (:417-420) and are never stored.
//
// Kernel <-> reference map:
//   feat_kernel       power spectrum, mel filterbank, log (:337-341)
//   featnorm_kernel   deltas (:304-310, :345-347), GroupNorm `norm` (:348), position table (:350)
//   shiftnorm_kernel  token shift + ScaleNorm denominator (:393-397)
//   dwconv_in_kernel  ConvModule residual (:399-400), OffsetScale + rotary (:404-409)
//   gate_kernel       (att_u*v)*sigmoid(att_v*u) (:431-432) + ScaleNorm denominator (:435)
//   dwconv_kernel     ConvModule residual of to_out (:437-438) / to_u||to_v (:451-452)
//   ln2_kernel        norm1 + affine-free LayerNorm (:446-449)
//   fsmn_mem_kernel   UniDeepFsmn memory conv (:459-462), gate (:464), norm2 (:467)
//   tail_norm_kernel  final LayerNorm, GroupNorm, skip, PReLU (:474-482)
//   tail_gate_kernel  tanh * sigmoid (:484-485)
//   mask_apply_kernel mask x STFT rows into the zero-framed ISTFT operand (:487)
#include "mf2_kernels.cuh"

namespace mf2 {

// ---------------------------------------------------------------------------------
// One CTA per frame: Kaldi power spectrum -> 60 mel bands -> log (+ int16-domain offset).
__global__ void __launch_bounds__(256)
feat_kernel(const float* __restrict__ fr, const float* __restrict__ banks, const int* __restrict__ mel_lo,
            const int* __restrict__ mel_hi, float* __restrict__ mel, float floor_v, float log_off) {
  __shared__ float pw[KB + 3];
  const long long m = blockIdx.x;
  const float* row = fr + m * FRONT;
  for (int i = threadIdx.x; i < KB; i += 256) {
    const float re = __ldg(row + i), im = __ldg(row + KB + i);
    pw[i] = re * re + im * im;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < NM; j += 8) {
    const int lo = mel_lo[j], hi = mel_hi[j];
    float acc = 0.f;
    for (int i = lo + lane; i < hi; i += 32) acc += __ldg(banks + j * KB + i) * pw[i];
    acc = warp_sum(acc);
    if (lane == 0) mel[m * NM + j] = logf(fmaxf(acc, floor_v)) + log_off;
  }
}

// One CTA per window: deltas, delta-deltas, GroupNorm(1, 180) over (180, T), operand planes for the
// 1x1 encoder GEMM; also seeds z with the position table (the GEMM adds onto it).
__global__ void __launch_bounds__(256)
featnorm_kernel(const float* __restrict__ mel, const float* __restrict__ gw, const float* __restrict__ gb,
                const float* __restrict__ emb, float* __restrict__ fhi, float* __restrict__ flo,
                float* __restrict__ z, int T) {
  extern __shared__ float sm[];
  float* s0 = sm;                 // [T][60] log-mel
  float* s1 = sm + T * NM;        // [T][60] delta
  __shared__ double red[2][8];
  __shared__ float stat[2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = mel + (long long)b * T * NM;
  for (int i = tid; i < T * NM; i += 256) s0[i] = __ldg(src + i);
  __syncthreads();
  auto delta = [&](const float* s, int t, int j) {
    float acc = 0.f;
#pragma unroll
    for (int k = -2; k <= 2; ++k) {
      int tt = t + k;
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      acc += ((float)k * 0.1f) * s[tt * NM + j];
    }
    return acc;
  };
  for (int i = tid; i < T * NM; i += 256) s1[i] = delta(s0, i / NM, i % NM);
  __syncthreads();
  double su = 0.0, sq = 0.0;
  for (int i = tid; i < T * NM; i += 256) {
    const float a = s0[i], d1 = s1[i], d2 = delta(s1, i / NM, i % NM);
    su += (double)a + (double)d1 + (double)d2;
    sq += (double)a * a + (double)d1 * d1 + (double)d2 * d2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { su += __shfl_xor_sync(0xffffffffu, su, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
  if ((tid & 31) == 0) { red[0][tid >> 5] = su; red[1][tid >> 5] = sq; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; q += red[1][w]; }
    const double n = (double)T * FEAT, mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + 1e-8));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (int i = tid; i < T * FEAT; i += 256) {
    const int t = i / FEAT, c = i - t * FEAT;
    const int part = c / NM, j = c - part * NM;
    const float v = part == 0 ? s0[t * NM + j] : (part == 1 ? s1[t * NM + j] : delta(s1, t, j));
    const float y = (v - mean) * rstd * __ldg(gw + c) + __ldg(gb + c);
    split_tf32_store(y, fhi, flo, ((long long)b * T + t) * FEATP + c);
  }
  float* zb = z + (long long)b * T * D;
  for (int i = tid; i < T * D / 4; i += 256) st4(zb + 4 * i, __ldg(reinterpret_cast<const float4*>(emb) + i));
}

// CTA = (8 frames, window), thread = channel: memory conv (k = 39, zero padded) on the projected branch, gate with xv,
// residual g_in, norm2 -> operand planes of conv2.  Each thread keeps its own shared-memory column (no barrier before
// the taps) and produces 4 outputs per pass so that 42 shared loads feed 156 FMAs (8 frames per CTA: with 121-frame
// windows more, smaller CTAs beat a smaller halo: 32-frame tiles measured 2.1 -> 3.2 ms).
constexpr int FM_TOK = 8;
__global__ void __launch_bounds__(256)
fsmn_mem_kernel(const float* __restrict__ xp, const float* __restrict__ uv, const float* __restrict__ gin,
                const float* __restrict__ taps, const float* __restrict__ w, const float* __restrict__ bvec,
                float* __restrict__ yhi, float* __restrict__ ylo, int T) {
  extern __shared__ float fsm[];
  float (*tile)[FI] = reinterpret_cast<float (*)[FI]>(fsm);                              // [FM_TOK + 2*MEMH][FI]
  float (*ys)[FI] = reinterpret_cast<float (*)[FI]>(fsm + (FM_TOK + 2 * MEMH) * FI);       // [FM_TOK][FI]
  const int t0 = blockIdx.x * FM_TOK, b = blockIdx.y, c = threadIdx.x;
  const long long base = (long long)b * T;
  for (int r = 0; r < FM_TOK + 2 * MEMH; ++r) {
    const int t = t0 + r - MEMH;
    tile[r][c] = (t >= 0 && t < T) ? __ldg(xp + (base + t) * FI + c) : 0.f;
  }
  float k[MEMK];
#pragma unroll
  for (int i = 0; i < MEMK; ++i) k[i] = __ldg(taps + i * FI + c);
  // own column only: no barrier needed between the tile fill and the taps
  for (int tt = 0; tt < FM_TOK; tt += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < MEMK + 3; ++i) {
      const float v = tile[tt + i][c];
      if (i < MEMK) a0 += k[i] * v;
      if (i >= 1 && i < MEMK + 1) a1 += k[i - 1] * v;
      if (i >= 2 && i < MEMK + 2) a2 += k[i - 2] * v;
      if (i >= 3) a3 += k[i - 3] * v;
    }
    const float conv[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tt + j;
      float y = 0.f;
      if (t < T) {
        const long long m = base + t;
        const float xu = __ldg(uv + m * (2 * FI) + c) + (tile[tt + j + MEMH][c] + conv[j]);
        y = __ldg(uv + m * (2 * FI) + FI + c) * xu + __ldg(gin + m * FI + c);
      }
      ys[tt + j][c] = y;
    }
  }
  __syncthreads();
  const int warp = c >> 5, lane = c & 31;
  const float4 w0 = ld4(w + lane * 4), w1 = ld4(w + 128 + lane * 4), b0 = ld4(bvec + lane * 4), b1 = ld4(bvec + 128 + lane * 4);
  const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  for (int tok = warp; tok < FM_TOK; tok += 8) {
    const int t = t0 + tok;
    if (t >= T) break;
    float v[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = ys[tok][lane * 4 + i]; v[4 + i] = ys[tok][128 + lane * 4 + i]; }
    float mean, rstd;
    ln256(v, mean, rstd);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
    const long long m = base + t;
    split4(make_float4(v[0], v[1], v[2], v[3]), yhi, ylo, m * FI + lane * 4);
    split4(make_float4(v[4], v[5], v[6], v[7]), yhi, ylo, m * FI + 128 + lane * 4);
  }
}

// In-place x *= s (the int16 PCM scale of a resampled input; 2^-15 commutes exactly with the linear interpolation).
__global__ void scale_kernel(float* __restrict__ x, long long n, float s) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[i] *= s;
}

// Output conversion behind the output resampler (:499-507): int16 = clamp(x, -1, 32767/32768) * 32768 through int32.
__global__ void se_convert_kernel(const float* __restrict__ src, void* __restrict__ out, int out_dtype, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  if (out_dtype == ADN_I16)
    reinterpret_cast<int16_t*>(out)[i] = (int16_t)max(-32768, min(32767, (int)(fminf(fmaxf(v, -1.0f), 32767.0f / 32768.0f) * 32768.0f)));
  else if (out_dtype == ADN_F32) reinterpret_cast<float*>(out)[i] = v;
  else reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
}

// One CTA per frame: real mask on both STFT row blocks, written into the zero-framed ISTFT operand.
__global__ void __launch_bounds__(256)
mask_apply_kernel(const float* __restrict__ fr, const float* __restrict__ mask, float* __restrict__ ehi,
                  float* __restrict__ elo, int T) {
  const long long m = blockIdx.x;
  const int b = (int)(m / T), t = (int)(m - (long long)b * T);
  const float* st = fr + m * FRONT + KROWS;
  const float* mk = mask + m * BINSP;
  const long long o = ((long long)b * (T + 2 * PADF) + PADF + t) * SPEC_LD;
  for (int r = threadIdx.x; r < SROWS; r += 256) {
    const int f = r >= BINS ? r - BINS : r;
    split_tf32_store(__ldg(st + r) * __ldg(mk + f), ehi, elo, o + r);
  }
}
class Model : public Base {
 public:
  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, Lp = 0, T = 0, Tp = 0, T4 = 0, Tn = 0, layers = 24;
  int L_in = 0, L_final = 0;     // window length at the input rate / output length at the output rate (== L at 48 kHz)
  bool rs_in = false, rs_out = false;
  float *xr = nullptr, *yres = nullptr, *yout = nullptr;

  int *d_mel_lo = nullptr, *d_mel_hi = nullptr;
  const float *banks = nullptr, *norm_w = nullptr, *norm_b = nullptr, *emb = nullptr, *rcos = nullptr, *rsin = nullptr;
  const float *mm_w = nullptr, *mm_b = nullptr, *in_w = nullptr, *in_b = nullptr, *prelu_a = nullptr, *gate_b = nullptr;
  const float* d_norm = nullptr;
  struct Layer {
    Lin in, out, c1, uv, ul, up, c2;
    const float *in_b, *in_c, *gamma, *beta, *out_b, *out_c, *c1_b, *c1_a, *n1_w, *n1_b, *uv_b, *uv_c, *ul_b, *mem_c,
        *n2_w, *n2_b, *c2_b;
  };
  std::vector<Layer> lw;
  Lin front, enc, gate, dec, ola;

  float *xp = nullptr, *xpl = nullptr, *fr = nullptr, *mel = nullptr, *featpl = nullptr, *z = nullptr, *h = nullptr;
  float *xs = nullptr, *rs = nullptr, *proj = nullptr, *vu = nullptr, *vuT = nullptr, *qq = nullptr, *lq = nullptr;
  float *qk = nullptr, *lk = nullptr, *s1 = nullptr, *ppl = nullptr, *att = nullptr, *gated = nullptr, *rs2 = nullptr;
  float *y = nullptr, *hpl = nullptr, *c1y = nullptr, *gin = nullptr, *xn = nullptr, *uvp = nullptr, *uv = nullptr;
  float *xupl = nullptr, *f1 = nullptr, *xp2 = nullptr, *yn = nullptr, *hn = nullptr, *tpl = nullptr, *gbuf = nullptr;
  float *tg = nullptr, *mask = nullptr, *enh = nullptr;
  size_t enh_plane = 0;
  Lin a_qk, a_lk, a_vuT;                   // per-window activation operands (W side)
  Gemm g_front, g_enc, g_lk, g_qk, g_pv, g_gate, g_dec, g_istft;
  struct LayerG { Gemm in, out, c1, uv, ul, up, c2; };
  std::vector<LayerG> lg;
  CUtensorMap map_proj;          // tile map over `proj` for the TMA-fed dwconv_in
  int stop_after = 0, last_batch = 0;
  bool bf = false;               // metadata matmul_dtype == BF16: the 24 layers' GEMMs run on bf16 operands (BASELINE configs[2])
  float* LO(float* p) const { return bf ? nullptr : p; }      // lo plane of an operand, absent in bf16 mode
  float* LOA(float* hi, float* lo) const { return af ? hi : LO(lo); }   // A-side operand: fp32-A mode stores fp32 in the hi buffer only

  ~Model() override {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    free_ws();
    auto fl = [](Lin& l) { if (l.planes) cudaFree(l.planes); };
    for (auto& w : lw) { fl(w.in); fl(w.out); fl(w.c1); fl(w.uv); fl(w.ul); fl(w.up); fl(w.c2); }
    fl(front); fl(enc); fl(gate); fl(dec); fl(ola);
    cudaFree(d_mel_lo); cudaFree(d_mel_hi);
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
    int nfft = 0, hop = 0, nmels = 0;
    std::string sin, sout;
    if (!geti("input_audio_length", L) || !geti("nfft", nfft) || !geti("hop_length", hop) || !geti("n_mels", nmels) ||
        !geti("mf2_layers", layers) || !gets("input_audio_dtype", sin) || !gets("output_audio_dtype", sout))
      return false;
    // optional linear resampling either side of the model (:318-325, :491-498): input_audio_length is at in_sample_rate
    int in_sr = 48000, out_sr = 48000, model_sr = 48000;
    {
      auto opt = [&](const char* k, int& v) { auto it = meta.find(k); if (it != meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
      opt("in_sample_rate", in_sr); opt("out_sample_rate", out_sr); opt("model_sample_rate", model_sr);
    }
    if (model_sr != 48000 || in_sr <= 0 || out_sr <= 0) { err = "mossformer2_se runs at model_sample_rate 48000"; return false; }
    L_in = L;
    rs_in = in_sr != model_sr;
    rs_out = out_sr != model_sr;
    if (rs_in) L = (int)llround((double)L_in * model_sr / in_sr);          // MODEL_AUDIO_LENGTH (:48)
    L_final = rs_out ? (int)llround((double)L_in * out_sr / in_sr) : L;    // OUTPUT_AUDIO_LENGTH (:49)
    if (nfft != NFFT || hop != HOP || nmels != NM || L < NFFT || (L - NFFT) % HOP) {
      err = "mossformer2_se needs nfft=1920, hop=384, n_mels=60 and a model-rate window of 1920 + k*384 samples";
      return false;
    }
    {
      auto it = meta.find("matmul_dtype");
      if (it != meta.end() && !it->second.empty() && it->second != "F32") {
        if (it->second != "BF16") { err = "matmul_dtype must be F32 (3xTF32, default) or BF16"; return false; }
        bf = true;
      }
      // 3xTF32 mode: fp32 activations + in-kernel operand split + TMA-store epilogues (ADN_MF2_AF=0: the round-1 operand planes)
      const char* e = getenv("ADN_MF2_AF");
      af = !bf && !(e && e[0] == '0');
    }
    auto pdt = [&](const std::string& s, int& o) { if (s == "F32") o = ADN_F32; else if (s == "INT16") o = ADN_I16; else if (s == "F16") o = ADN_F16; else return false; return true; };
    if (!pdt(sin, in_dtype) || !pdt(sout, out_dtype)) { err = "bad audio dtype"; return false; }
    T = (L - NFFT) / HOP + 1;
    if (T > 256) { err = "mossformer2_se: windows longer than one FLASH group (256 frames) are not supported; fold the audio"; return false; }
    Lp = round_up(L, 4);
    Tp = round_up(T, 32);
    T4 = round_up(T, 4);
    Tn = T4 <= 128 ? 128 : 256;

    bool ok = true;
    banks = dptr("mel_banks", (size_t)NM * KB, ok);
    norm_w = dptr("norm.w", FEAT, ok); norm_b = dptr("norm.b", FEAT, ok);
    emb = dptr("emb_pos", (size_t)T * D, ok);
    rcos = dptr("rot_cos", (size_t)T * ROT, ok); rsin = dptr("rot_sin", (size_t)T * ROT, ok);
    mm_w = dptr("mm_norm.w", D, ok); mm_b = dptr("mm_norm.b", D, ok);
    in_w = dptr("intra_norm.w", D, ok); in_b = dptr("intra_norm.b", D, ok);
    prelu_a = dptr("prelu_a", 1, ok);
    gate_b = dptr("gate_b", 2 * D, ok);
    d_norm = dptr("istft.norm", (size_t)L, ok);
    if (!ok) return false;
    {   // non-zero span of every mel filter
      auto it = index.find("mel_banks");
      const float* bk = h_blob + it->second.offset;
      std::vector<int> lo(NM), hi(NM);
      for (int j = 0; j < NM; ++j) {
        int a = KB, b = 0;
        for (int i = 0; i < KB; ++i)
          if (bk[j * KB + i] != 0.f) { if (i < a) a = i; b = i + 1; }
        if (a > b) a = b = 0;
        lo[j] = a; hi[j] = b;
      }
      if (cudaMalloc((void**)&d_mel_lo, NM * 4) != cudaSuccess || cudaMalloc((void**)&d_mel_hi, NM * 4) != cudaSuccess ||
          cudaMemcpy(d_mel_lo, lo.data(), NM * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(d_mel_hi, hi.data(), NM * 4, cudaMemcpyHostToDevice) != cudaSuccess) { err = "mel table upload failed"; return false; }
    }
    const float* w;
    w = dptr("frontend", (size_t)FRONT * NFFT, ok); if (ok && !make_lin(front, w, FRONT, NFFT)) return false;
    w = dptr("enc.w", (size_t)D * FEAT, ok);        if (ok && !make_lin(enc, w, D, FEAT)) return false;
    w = dptr("gate_w", (size_t)2 * D * D, ok);      if (ok && !make_lin(gate, w, 2 * D, D)) return false;
    w = dptr("dec_w", (size_t)BINS * D, ok);        if (ok && !make_lin(dec, w, BINS, D)) return false;
    lw.resize(layers);
    for (int i = 0; i < layers && ok; ++i) {
      const std::string p = "L" + std::to_string(i) + ".";
      Layer& Y = lw[i];
      w = dptr(p + "in_w", (size_t)PROJ * D, ok);  if (ok && !make_lin(Y.in, w, PROJ, D, bf)) return false;
      w = dptr(p + "out_w", (size_t)D * VU, ok);   if (ok && !make_lin(Y.out, w, D, VU, bf)) return false;
      w = dptr(p + "c1_w", (size_t)FI * D, ok);    if (ok && !make_lin(Y.c1, w, FI, D, bf)) return false;
      w = dptr(p + "uv_w", (size_t)2 * FI * FI, ok); if (ok && !make_lin(Y.uv, w, 2 * FI, FI, bf)) return false;
      w = dptr(p + "ul_w", (size_t)FI * FI, ok);   if (ok && !make_lin(Y.ul, w, FI, FI, bf)) return false;
      w = dptr(p + "up_w", (size_t)FI * FI, ok);   if (ok && !make_lin(Y.up, w, FI, FI, bf)) return false;
      w = dptr(p + "c2_w", (size_t)D * FI, ok);    if (ok && !make_lin(Y.c2, w, D, FI, bf)) return false;
      Y.in_b = dptr(p + "in_b", PROJ, ok); Y.in_c = dptr(p + "in_c", (size_t)DW * PROJ, ok);
      Y.gamma = dptr(p + "qk_gamma", 4 * QK, ok); Y.beta = dptr(p + "qk_beta", 4 * QK, ok);
      Y.out_b = dptr(p + "out_b", D, ok); Y.out_c = dptr(p + "out_c", (size_t)DW * D, ok);
      Y.c1_b = dptr(p + "c1_b", FI, ok); Y.c1_a = dptr(p + "c1_a", 1, ok);
      Y.n1_w = dptr(p + "n1_w", FI, ok); Y.n1_b = dptr(p + "n1_b", FI, ok);
      Y.uv_b = dptr(p + "uv_b", 2 * FI, ok); Y.uv_c = dptr(p + "uv_c", (size_t)DW * 2 * FI, ok);
      Y.ul_b = dptr(p + "ul_b", FI, ok); Y.mem_c = dptr(p + "mem_c", (size_t)MEMK * FI, ok);
      Y.n2_w = dptr(p + "n2_w", FI, ok); Y.n2_b = dptr(p + "n2_b", FI, ok);
      Y.c2_b = dptr(p + "c2_b", D, ok);
    }
    if (!ok) return false;
    {   // overlap-add weight: raw hop-block j = sum over the R frames that cover it (see api.cu build_ola_weight)
      auto inv = index.find("istft.inv");
      if (inv == index.end() || inv->second.count != (size_t)SROWS * NFFT) { err = "missing istft.inv"; return false; }
      const float* ib = h_blob + inv->second.offset;
      const int K = R_OLA * SPEC_LD;
      std::vector<float> wv((size_t)HOP * K, 0.f);
      for (int n = 0; n < HOP; ++n)
        for (int q = 0; q < R_OLA; ++q) {
          const int src = n + (R_OLA - 1 - q) * HOP;
          if (src >= NFFT) continue;
          for (int r = 0; r < SROWS; ++r) wv[(size_t)n * K + (size_t)q * SPEC_LD + r] = ib[(size_t)r * NFFT + src];
        }
      float* tmp = nullptr;
      if (cudaMalloc((void**)&tmp, wv.size() * 4) != cudaSuccess ||
          cudaMemcpy(tmp, wv.data(), wv.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { err = "ola upload failed"; return false; }
      const bool r = make_lin(ola, tmp, HOP, K);
      cudaDeviceSynchronize();
      cudaFree(tmp);
      if (!r) return false;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "weight split failed"; return false; }
    return true;
  }

  size_t floats_needed(size_t B) const {
    const size_t M = B * T;
    return (rs_in ? B * L : 0) + (rs_out ? B * (L + L_final) : 0) + B * Lp * 3 + M * FRONT + M * NM + 2 * M * FEATP + 2 * M * D + 2 * M * D + 2 * M + M * PROJ + M * VU2 +
           2 * B * VU2 * Tp + 4 * M * QK + 4 * B * Tn * QK + 3 * M * Tp + M * VU2 +
           2 * M * VU + M * D + 2 * M * D + 2 * M * FI + 2 * M * FI + 2 * M * D + 2 * M * FI + 2 * M * FI + M * FI +
           2 * M * FI + M * D + 2 * M * D + M * 2 * D + 2 * M * D + M * BINSP + 2 * (B * (T + 2 * PADF) * SPEC_LD + 9664);
  }

  bool ensure(int B) {
    if (B == planned) return true;
    cudaDeviceSynchronize();
    free_ws();
    const long long M = (long long)B * T;
    const size_t xplane = (size_t)B * Lp;
    if ((rs_in && !alloc(xr, (size_t)B * L, false)) ||
        (rs_out && (!alloc(yres, (size_t)B * L, false) || !alloc(yout, (size_t)B * L_final, false))))
      return false;
    enh_plane = (size_t)B * (T + 2 * PADF) * SPEC_LD + ola.k_pad;       // + slack for the K overrun of the last block
    if (!alloc(xp, xplane, false) || !alloc(xpl, 2 * xplane, false) || !alloc(fr, (size_t)M * FRONT, false) ||
        !alloc(mel, (size_t)M * NM, false) || !alloc(featpl, 2 * (size_t)M * FEATP, true) || !alloc(z, (size_t)M * D, false) ||
        !alloc(h, (size_t)M * D, false) || !alloc(xs, 2 * (size_t)M * D, false) || !alloc(rs, (size_t)M, false) ||
        !alloc(proj, (size_t)M * PROJ, false) || !alloc(vu, (size_t)M * VU2, false) ||
        !alloc(vuT, 2 * (size_t)B * VU2 * Tp, true) || !alloc(qq, 2 * (size_t)M * QK, false) ||
        !alloc(lq, 2 * (size_t)M * QK, false) || !alloc(qk, 2 * (size_t)B * Tn * QK, true) ||
        !alloc(lk, 2 * (size_t)B * Tn * QK, true) || !alloc(s1, (size_t)M * Tp, true) ||
        !alloc(ppl, 2 * (size_t)M * Tp, true) || !alloc(att, (size_t)M * VU2, false) ||
        !alloc(gated, 2 * (size_t)M * VU, false) || !alloc(rs2, (size_t)M, false) || !alloc(y, (size_t)M * D, false) ||
        !alloc(hpl, 2 * (size_t)M * D, false) || !alloc(c1y, (size_t)M * FI, false) || !alloc(gin, (size_t)M * FI, false) ||
        !alloc(xn, 2 * (size_t)M * FI, false) || !alloc(uvp, (size_t)M * 2 * FI, false) || !alloc(uv, (size_t)M * 2 * FI, false) ||
        !alloc(xupl, 2 * (size_t)M * FI, false) || !alloc(f1, 2 * (size_t)M * FI, false) || !alloc(xp2, (size_t)M * FI, false) ||
        !alloc(yn, 2 * (size_t)M * FI, false) || !alloc(hn, (size_t)M * D, false) || !alloc(tpl, 2 * (size_t)M * D, false) ||
        !alloc(gbuf, (size_t)M * 2 * D, false) || !alloc(tg, 2 * (size_t)M * D, false) || !alloc(mask, (size_t)M * BINSP, false) ||
        !alloc(enh, 2 * enh_plane, true))
      return false;

    // frontend: rows = frames (stride hop) of the raw window
    if (!plan_gemm(g_front, af ? xp : xpl, (long long)xplane, NFFT, T, HOP, B, Lp, front)) return false;
    g_front.args.C = fr; g_front.args.ldc = FRONT;
    if (!plan_gemm(g_enc, featpl, M * FEATP, FEATP, (int)M, FEATP, 1, M * FEATP, enc)) return false;
    g_enc.args.K = FEATP; g_enc.args.resid = z; g_enc.args.C = z; g_enc.args.ldc = D;

    // attention operands that live in activations
    const int bn_qk = Tn;                    // 128 or 256 keys per tile
    if (!make_act_lin(a_qk, qk, LO(qk + (size_t)B * Tn * QK), T4, Tn, QK, QK, bn_qk, B) ||
        !make_act_lin(a_lk, lk, LO(lk + (size_t)B * Tn * QK), T4, Tn, QK, QK, bn_qk, B) ||
        !make_act_lin(a_vuT, vuT, LO(vuT + (size_t)B * VU2 * Tp), VU2, VU2, Tp, Tp, 256, B))
      return false;
    if (!plan_gemm(g_lk, lq, M * QK, QK, T, QK, B, (long long)T * QK, a_lk)) return false;
    g_lk.args.C = s1; g_lk.args.ldc = Tp;
    if (!plan_gemm(g_qk, qq, M * QK, QK, T, QK, B, (long long)T * QK, a_qk)) return false;
    g_qk.args.act = tc::ACT_RELU2; g_qk.args.resid = s1; g_qk.args.Chi = ppl; g_qk.args.Clo = LOA(ppl, ppl + M * Tp); g_qk.args.ldc = Tp;
    if (!plan_gemm(g_pv, ppl, M * Tp, Tp, T, Tp, B, (long long)T * Tp, a_vuT)) return false;
    g_pv.args.C = att; g_pv.args.ldc = VU2;

    lg.assign(layers, LayerG{});
    for (int i = 0; i < layers; ++i) {
      LayerG& G = lg[i];
      const Layer& Y = lw[i];
      if (!plan_gemm(G.in, xs, M * D, D, (int)M, D, 1, M * D, Y.in)) return false;
      G.in.args.rowscale = rs; G.in.args.bias = Y.in_b; G.in.args.act = tc::ACT_SILU; G.in.args.C = proj; G.in.args.ldc = PROJ;
      if (!plan_gemm(G.out, gated, M * VU, VU, (int)M, VU, 1, M * VU, Y.out)) return false;
      G.out.args.rowscale = rs2; G.out.args.bias = Y.out_b; G.out.args.act = tc::ACT_SILU; G.out.args.C = y; G.out.args.ldc = D;
      if (!plan_gemm(G.c1, af ? h : hpl, M * D, D, (int)M, D, 1, M * D, Y.c1)) return false;
      G.c1.args.bias = Y.c1_b; G.c1.args.act = tc::ACT_PRELU; G.c1.args.act_param = Y.c1_a; G.c1.args.C = c1y; G.c1.args.ldc = FI;
      if (!plan_gemm(G.uv, xn, M * FI, FI, (int)M, FI, 1, M * FI, Y.uv)) return false;
      G.uv.args.bias = Y.uv_b; G.uv.args.act = tc::ACT_SILU; G.uv.args.C = uvp; G.uv.args.ldc = 2 * FI;
      if (!plan_gemm(G.ul, af ? uv : xupl, M * FI, FI, (int)M, af ? 2 * FI : FI, 1, af ? M * 2 * FI : M * FI, Y.ul)) return false;
      G.ul.args.bias = Y.ul_b; G.ul.args.act = tc::ACT_RELU; G.ul.args.Chi = f1; G.ul.args.Clo = LOA(f1, f1 + M * FI); G.ul.args.ldc = FI;
      if (!plan_gemm(G.up, f1, M * FI, FI, (int)M, FI, 1, M * FI, Y.up)) return false;
      G.up.args.C = xp2; G.up.args.ldc = FI;
      if (!plan_gemm(G.c2, yn, M * FI, FI, (int)M, FI, 1, M * FI, Y.c2)) return false;
      G.c2.args.bias = Y.c2_b; G.c2.args.resid = h; G.c2.args.C = h; G.c2.args.ldc = D;
    }
    if (!plan_gemm(g_gate, tpl, M * D, D, (int)M, D, 1, M * D, gate)) return false;
    g_gate.args.bias = gate_b; g_gate.args.C = gbuf; g_gate.args.ldc = 2 * D;
    if (!plan_gemm(g_dec, tg, M * D, D, (int)M, D, 1, M * D, dec)) return false;
    g_dec.args.N = BINSP; g_dec.args.act = tc::ACT_RELU; g_dec.args.C = mask; g_dec.args.ldc = BINSP;

    {   // fp32-A plans: output tensor maps (TMA-store epilogue)
      std::vector<Gemm*> all = {&g_front, &g_enc, &g_lk, &g_qk, &g_pv, &g_gate, &g_dec};
      for (auto& G : lg) for (Gemm* q : {&G.in, &G.out, &G.c1, &G.uv, &G.ul, &G.up, &G.c2}) all.push_back(q);
      for (Gemm* q : all) if (!finish_af(*q)) return false;
    }
    {   // inverse: rows = raw hop blocks, each a run of R consecutive zero-framed spectrum frames
      const int rows = T + 2 * PADF;
      const int TM = T + R_OLA - 1;          // blocks 0 .. (raw-1)/hop
      if (!plan_gemm(g_istft, enh, (long long)enh_plane, ola.k_pad, rows, SPEC_LD, B, (long long)rows * SPEC_LD, ola, false)) return false;
      tc::TcArgs& c = g_istft.args;
      const int bt = TM >= 128 ? 128 : TM;
      c.bt = bt; c.tiles_per_chunk = (TM + 127) / 128; c.m_tiles = B * c.tiles_per_chunk; c.TM = TM; c.t0 = 0;
      c.N = HOP; c.K = R_OLA * SPEC_LD;
      c.norm = d_norm; c.norm_mul = 0; c.hop = HOP; c.shift = 0; c.out_len = L; c.out_dtype = rs_out ? ADN_F32 : out_dtype; c.i16_mode = 1;
      // the A box must match bt rows
      if (!tc::make_row_map(&g_istft.plan.map_a_hi, enh, ola.k_pad, rows, SPEC_LD, B, (long long)rows * SPEC_LD, bt, 1, err) ||
          !tc::make_row_map(&g_istft.plan.map_a_lo, enh + enh_plane, ola.k_pad, rows, SPEC_LD, B, (long long)rows * SPEC_LD, bt, 1, err))
        return false;
    }
    if (!tc::make_tile_map(&map_proj, proj, PROJ, T, PROJ, B, (long long)T * PROJ, DP_C, DP_ROWS, err)) return false;
    // workspace memsets ran on the legacy default stream; runs use a non-blocking stream that does not order against it
    if (cudaDeviceSynchronize() != cudaSuccess) { err = "workspace initialisation failed"; return false; }
    planned = B;
    return true;
  }

  // ---- ModelImpl
  void io_info(adn_tensor_info* in, adn_tensor_info* out) override {
    memset(in, 0, sizeof(*in));
    memset(out, 0, sizeof(*out));
    strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);        // Export_MossFormer_SE.py:537
    in->dtype = in_dtype; in->channels = 1; in->length = L_in;
    strncpy(out->name, "denoised_audio", sizeof(out->name) - 1);   // :538
    out->dtype = out_dtype; out->channels = 1; out->length = L_final;
  }
  size_t workspace_bytes(int batch) override { return floats_needed((size_t)batch) * sizeof(float); }
  int launches(int) override { return 5 + layers * 17 + 6 + (rs_in ? (in_dtype == ADN_I16 ? 2 : 1) : 0) + (rs_out ? 2 : 0); }
  void set_stop_after(int n) override { stop_after = n; }

#define MF_TICK(name) do { ++n; if (tick) tick(tick_ctx, name); if (stop_after > 0 && n >= stop_after) return ADN_OK; } while (0)
#define MF_GEMM(G, epi, name) do { cudaError_t e_ = tc::launch((G).plan, (G).args, epi, sms, st); \
    if (e_ != cudaSuccess) { err = std::string("gemm launch (") + name + "): " + cudaGetErrorString(e_); return ADN_ERR_CUDA; } MF_TICK(name); } while (0)

  adn_status run(const void* d_in, void* d_out, int B, cudaStream_t st) override {
    if (!ensure(B)) return ADN_ERR_CUDA;
    last_batch = B;
    int n = 0;
    const long long M = (long long)B * T;
    const unsigned wtok = (unsigned)((M + 7) / 8);
    static unsigned long long cfg = 0;             // per device
    if (adn_first_use_on_device(cfg)) {
      cudaFuncSetAttribute(dwconv_in_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP_SMEM);
      cudaFuncSetAttribute(featnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 2 * NM * 4);
      cudaFuncSetAttribute(fsmn_mem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * FM_TOK + 2 * MEMH) * FI * 4);
    }

    // 1-3: cast (+1/32768 for int16, :315-317), fused Kaldi||STFT frontend (:335), log-mel (:337-341)
    const void* src = d_in;
    int src_dtype = in_dtype;
    if (rs_in) {                                   // x * 1/32768 (int16), then F.interpolate(size=MODEL_AUDIO_LENGTH) (:315-325)
      if (adn_resample_linear(d_in, in_dtype, xr, B, L_in, L, 0.0, st) != ADN_OK) { err = "input resampler launch failed"; return ADN_ERR_CUDA; }
      MF_TICK("resample_in");
      if (in_dtype == ADN_I16) {
        scale_kernel<<<(unsigned)(((long long)B * L + 255) / 256), 256, 0, st>>>(xr, (long long)B * L, 1.0f / 32768.0f);
        MF_TICK("pcm_scale");
      }
      src = xr;
      src_dtype = ADN_F32;
    }
    gtcrn::launch_prep(src, src_dtype, xp, xpl, xpl + (size_t)B * Lp, B, L, Lp, 0, 0, 0, st);
    MF_TICK("prep");
    MF_GEMM(g_front, EPI_LIN, "frontend_gemm");
    feat_kernel<<<(unsigned)M, 256, 0, st>>>(fr, banks, d_mel_lo, d_mel_hi, mel, 1.1920929e-07f * (1.0f / 32768.0f) * (1.0f / 32768.0f),
                                            20.794415416798357f);
    MF_TICK("feat");
    featnorm_kernel<<<B, 256, (size_t)T * 2 * NM * sizeof(float), st>>>(mel, norm_w, norm_b, emb, featpl, af ? featpl : featpl + M * FEATP, z, T);
    MF_TICK("featnorm");
    MF_GEMM(g_enc, EPI_LIN, "enc_gemm");

    for (int i = 0; i < layers; ++i) {
      LayerG& G = lg[i];
      const Layer& Y = lw[i];
      const float* hin = i == 0 ? z : h;
      shiftnorm_kernel<<<wtok, 256, 0, st>>>(hin, xs, LOA(xs, xs + M * D), rs, M, T, 0);
      MF_TICK("shiftnorm");
      MF_GEMM(G.in, EPI_LIN, "fl_in");
      dwconv_in_tma_kernel<<<dim3(PROJ / DP_C, B, (T + DP_TILES * DP_F - 1) / (DP_TILES * DP_F)), 256, DP_SMEM, st>>>(
          map_proj, Y.in_c, Y.gamma, Y.beta, rcos, rsin, vu, vuT, LO(vuT + (size_t)B * VU2 * Tp), qq, LOA(qq, qq + M * QK), lq, LOA(lq, lq + M * QK),
          qk, LO(qk + (size_t)B * Tn * QK), lk, LO(lk + (size_t)B * Tn * QK), nullptr, nullptr, T, Tp, Tn, T, QK);
      MF_TICK("dwconv_in");
      MF_GEMM(g_lk, EPI_LIN, "att_lk");
      MF_GEMM(g_qk, EPI_LIN, "att_qk");
      MF_GEMM(g_pv, EPI_LIN, "att_pv");
      gate_kernel<<<wtok, 256, 0, st>>>(att, vu, gated, LOA(gated, gated + M * VU), rs2, M, T, T, 0);
      MF_TICK("gate");
      MF_GEMM(G.out, EPI_LIN, "fl_out");
      dwconv_kernel<<<dim3(D / 32 / DWI_WARPS, B, (T + DW_SEG - 1) / DW_SEG), DWI_WARPS * 32, 0, st>>>(y, Y.out_c, hin, h, af ? nullptr : hpl, LO(hpl + M * D), D, T, D);      // fp32-A: fsmn_conv1 reads h itself
      MF_TICK("dwconv_out");
      MF_GEMM(G.c1, EPI_LIN, "fsmn_conv1");
      ln2_kernel<<<wtok, 256, 0, st>>>(c1y, Y.n1_w, Y.n1_b, gin, xn, LOA(xn, xn + M * FI), M);
      MF_TICK("ln2");
      MF_GEMM(G.uv, EPI_LIN, "fsmn_uv");
      dwconv_kernel<<<dim3(2 * FI / 32 / DWI_WARPS, B, (T + DW_SEG - 1) / DW_SEG), DWI_WARPS * 32, 0, st>>>(uvp, Y.uv_c, nullptr, uv, af ? nullptr : xupl, LO(xupl + M * FI), FI, T, 2 * FI);   // fp32-A: fsmn_linear reads uv's first FI columns
      MF_TICK("dwconv_uv");
      MF_GEMM(G.ul, EPI_LIN, "fsmn_linear");
      MF_GEMM(G.up, EPI_LIN, "fsmn_project");
      fsmn_mem_kernel<<<dim3((T + FM_TOK - 1) / FM_TOK, B), 256, (2 * FM_TOK + 2 * MEMH) * FI * sizeof(float), st>>>(xp2, uv, gin, Y.mem_c, Y.n2_w, Y.n2_b, yn, LOA(yn, yn + M * FI), T);
      MF_TICK("fsmn_mem");
      MF_GEMM(G.c2, EPI_LIN, "fsmn_conv2");
    }

    tail_norm_kernel<<<B, 512, 0, st>>>(layers ? h : z, z, mm_w, mm_b, in_w, in_b, prelu_a, hn, tpl, af ? tpl : tpl + M * D, T);
    MF_TICK("tail_norm");
    MF_GEMM(g_gate, EPI_LIN, "tail_gate_gemm");
    tail_gate_kernel<<<(unsigned)((M * (D / 4) + 255) / 256), 256, 0, st>>>(gbuf, tg, af ? tg : tg + M * D, M);
    MF_TICK("tail_gate");
    MF_GEMM(g_dec, EPI_LIN, "mask_gemm");
    mask_apply_kernel<<<(unsigned)M, 256, 0, st>>>(fr, mask, enh, enh + enh_plane, T);
    MF_TICK("mask_apply");
    g_istft.args.out = rs_out ? (void*)yres : d_out;
    MF_GEMM(g_istft, EPI_ISTFT, "istft_gemm");
    if (rs_out) {                                  // F.interpolate(size=OUTPUT_AUDIO_LENGTH), then the output rule (:491-507)
      if (adn_resample_linear(yres, ADN_F32, yout, B, L, L_final, 0.0, st) != ADN_OK) { err = "output resampler launch failed"; return ADN_ERR_CUDA; }
      MF_TICK("resample_out");
      se_convert_kernel<<<(unsigned)(((long long)B * L_final + 255) / 256), 256, 0, st>>>(yout, d_out, out_dtype, (long long)B * L_final);
      MF_TICK("convert_out");
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mf2se run: ") + cudaGetErrorString(e); return ADN_ERR_CUDA; }
    return ADN_OK;
  }

  adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) override {
    const size_t B = last_batch;
    if (!B) { err = "adn_debug_read: no run yet"; return ADN_ERR_INVALID; }
    const size_t M = B * T;
    std::map<std::string, std::pair<const float*, size_t>> tbl = {
        {"fr", {fr, M * FRONT}}, {"mel", {mel, M * NM}}, {"z", {z, M * D}}, {"h", {h, M * D}}, {"proj", {proj, M * PROJ}},
        {"vu", {vu, M * VU2}}, {"att", {att, M * VU2}}, {"y", {y, M * D}}, {"c1y", {c1y, M * FI}}, {"gin", {gin, M * FI}},
        {"uv", {uv, M * 2 * FI}}, {"xp2", {xp2, M * FI}}, {"gate", {gbuf, M * 2 * D}}, {"mask", {mask, M * BINSP}},
        {"rs", {rs, M}}, {"rs2", {rs2, M}},
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

ModelImpl* mf2se_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err) {
  mf2::Model* m = new mf2::Model();
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
