// GTCRN backbone on sm_100a: parameter structs + launcher interface.
// Follows reference GTCRN/Export_GTCRN.py:55-693 (see DESIGN.md for the kernel map).
#pragma once
#include "common.cuh"

namespace gtcrn {

constexpr int NFFT = 512, HOP = 256, FB = 257;     // Export_GTCRN.py:37-39
constexpr int SPEC_LD = 520;                       // frame stride of the packed spectrum (514 used)
constexpr int ERB_F = 129, E0_F = 65, E1_F = 33;   // Export_GTCRN.py:488-489 (stride-2 convs)
constexpr int FRAME16 = 16 * E1_F;                 // 528 floats: one (16,33) frame
constexpr int FRAME_E0 = 16 * E0_F;                // 1040 floats: one (16,65) frame

// Weights that every thread of a warp reads at the same index are passed BY VALUE as
// kernel parameters: they land in the constant bank and feed FFMA directly as c[0][..]
// operands once the loops are unrolled.
struct EncFrontW {          // en_convs.0 / en_convs.1, BN folded (Export_GTCRN.py:171-194)
  float w0[5][9][16];       // [k][ci][o]: the unrolled inner loop (o) walks contiguous constants (LDCU.128)
  float b0[16];
  float w1[2][8][5][8];     // [group][ci][k][o_local]; groups=2: out o uses inputs (o/8)*8 .. +8
  float b1[16];
  float a0, a1;             // PReLU slopes
};

struct GTW {                // one GTConvBlock, BN folded; deconv blocks are stored in the
  float w1[16][24];         // equivalent causal-conv form (flipped taps), see gtcrn_params.py
  float b1[16];
  float wd[16][3][3];       // [c][kt][kf], kt=0 is the oldest frame (t-2d)
  float bd[16];
  float w2[16][8];          // [c][o]
  float b2[8];
  float a1, ad;
};

struct DecTailW {           // de_convs.3 (groups=2) / de_convs.4, BN folded
  float w3[16][5][8];       // [ci][k][o_local]
  float b3[16];
  float w4[16][5][2];       // [ci][k][o]
  float b4[2];
  float a3;
};

struct GruPtrs {            // PyTorch layout: w_ih (3H,I), w_hh (3H,H), b_ih (3H), b_hh (3H)
  const float* w_ih;
  const float* w_hh;
  const float* b_ih;
  const float* b_hh;
};

struct TraW {               // TRA: GRU(8->16) + Linear(16->8) (Export_GTCRN.py:144-156)
  GruPtrs gru;
  const float* fc_w;        // (8,16)
  const float* fc_b;        // (8)
};

struct DpW {                // DPGRNN (Export_GTCRN.py:431-481); LN tables stored [c][f]
  GruPtrs intra[2][2];      // [group][direction]
  const float* intra_fc_w;  // (16,16)
  const float* intra_fc_b;
  const float* intra_ln_w;  // (16,33)
  const float* intra_ln_b;
  GruPtrs inter[2];         // [group]
  const float* inter_fc_w;
  const float* inter_fc_b;
  const float* inter_ln_w;
  const float* inter_ln_b;
};

struct ErbW {               // ERB.bm / ERB.bs (Export_GTCRN.py:99-107), dense + nonzero ranges
  const float* bm;          // (192,64)
  const float* bm_lo;       // (64) first nonzero input bin per band (stored as float)
  const float* bm_hi;       // (64) one past the last
  const float* bs;          // (64,192)
  const float* bs_lo;       // (192)
  const float* bs_hi;       // (192)
};

struct Weights {
  EncFrontW enc_front;
  GTW enc_gt[3];
  GTW dec_gt[3];
  TraW enc_tra[3];
  TraW dec_tra[3];
  DpW dp[2];
  DecTailW dec_tail;
  ErbW erb;
};

struct Buffers {            // all fp32, frame-major: (B, T, C, F) with F innermost
  float* xp;                // (B, Lp) conditioned + centre-padded waveform
  float* spec;              // (B, T, 520)   [Re 0..256 | Im 0..256 | pad]
  float* e0;                // (B, T, 16, 65)
  float* e[5];              // e[1..4]: (B, T, 16, 33) encoder skips; e[0] unused
  float* h1;                // (B, T, 8, 33) GTConvBlock output before TRA
  float* zt;                // (B, T, 8)   TRA energies
  float* at;                // (B, T, 8)   TRA gates
  float* tgi;               // (B, T, 48)  TRA GRU input projections (scratch of tra_gru)
  float* thid;              // (B, T, 16)  TRA GRU hidden states     (scratch of tra_gru)
  float* gi;                // (B, T, 3, 33, 16) inter-GRU input projections
  float* xp_hi;             // tf32 hi/lo planes of xp / enh for the tensor-core GEMMs (may be null)
  float* xp_lo;
  float* enh_hi;
  float* enh_lo;
  float* xa;                // (B, T, 16, 33) ping
  float* xb;                // (B, T, 16, 33) pong
  float* inter;             // (B, T, 33, 16) inter-path GRU output h (pre-Linear/LN)
  float* enh;               // (B, T + 2*(R-1), 520) enhanced spectrum, zero frames around
};

struct Dims {
  int B, L, Lp, T;
};

// Launches the GTCRN spectrum->spectrum stages (everything between STFT and ISTFT).
// Returns the number of kernels launched.  `tick` (may be null) is called after every
// launch with the kernel's name for per-kernel event timing.
typedef void (*TickFn)(void* ctx, const char* name);
// `stop_after` > 0 returns after that many launches (diagnostics: lets tests inspect
// ping-pong buffers mid-pipeline).
int launch_backbone(const Weights& w, const Buffers& buf, const Dims& d, int enh_pad_frames,
                    cudaStream_t st, TickFn tick, void* tick_ctx, int stop_after);

// The pieces of launch_backbone, for H-GTCRN (csrc/hgtcrn.cu: its own six-channel encoder front, the same network between
// en_convs.1 and the band synthesis, the complex ratio mask applied to microphone 0's rows of a two-microphone spectrum).
// launch_backbone_core continues the launch count from `n` and returns the decoder's last activation in *last (left null
// when `stop_after` ended the sequence early).
int launch_backbone_core(const Weights& w, const Buffers& buf, const Dims& d, cudaStream_t st, TickFn tick, void* tick_ctx,
                         int stop_after, int n, float** last);
void launch_dec_tail(const Weights& w, const Buffers& buf, const float* xin, const float* spec, long long spec_chunk_stride,
                     const Dims& d, int enh_pad_frames, cudaStream_t st);

// Input conditioning (cast, 1/32768, DC removal, centre pad): Export_GTCRN.py:637-647 +
// STFT_Process.py:305-309.
// hi/lo (nullable): additionally emit the 3xTF32 operand planes.
void launch_prep(const void* in, int in_dtype, float* xp, float* hi, float* lo, int B, int L, int Lp, int half,
                 int remove_dc, int reflect, cudaStream_t st);

// recurrent stages (gtcrn_rnn.cu)
void launch_tra_gru(const TraW& w, const float* zt, float* tgi, float* hbuf, float* at, int B, int T,
                    cudaStream_t st);
void launch_tra_apply(const float* at, const float* h1, const float* xin, const float* skip, float* out, int B,
                      int T, cudaStream_t st);
void launch_dp_intra(const DpW& w, const float* a, const float* hprev, const DpW* prev, float* out, float* gi,
                     int nframes, cudaStream_t st);
void launch_dp_inter(const DpW& w, const float* gi, float* hout, int B, int T, cudaStream_t st);
void launch_ln_res(const DpW& w, const float* a, const float* hin, const float* skip, float* out, int nframes,
                   cudaStream_t st);

__device__ __forceinline__ void split_tf32_store(float v, float* hi, float* lo, long long i) {
  const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = v - h;
}

}  // namespace gtcrn
