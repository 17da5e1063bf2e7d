/*
 * adn.h -- C ABI of libadn.so: B200-native speech-denoising inference path.
 *
 * The reference (DakeQQ/Audio-Denoiser-ONNX) has no FFI of its own: its run-time
 * boundary is the slice of the onnxruntime Python API used by Inference_*_ONNX.py
 * (SURVEY.md 8b).  Each entry point below names the reference call it replaces.
 * The ctypes binding lives in audio-denoiser-onnx_b200/adn/_lib.py and the
 * onnxruntime-shaped shim on top of it in audio-denoiser-onnx_b200/adn/ort_shim.py.
 *
 * Conventions: plain pointers and sizes only; every function returns an adn_status
 * (0 == ADN_OK) unless stated otherwise; the failing call's message is available from
 * adn_last_error().  A handle is bound to one CUDA device and is not re-entrant
 * (the reference runs one in-flight run per session, Inference_GTCRN_ONNX.py:209-210).
 * The caller owns all I/O buffers; the library owns weights and workspace.
 * There is no CPU fallback: every entry point fails with ADN_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef ADN_H_
#define ADN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t adn_status;
enum {
  ADN_OK = 0,
  ADN_ERR_INVALID = 1,   /* bad argument / unknown family / missing tensor or key */
  ADN_ERR_CUDA = 2,      /* CUDA runtime error, message has the cudaError string  */
  ADN_ERR_UNSUPPORTED = 3
};

/* audio sample types == metadata values of input_audio_dtype / output_audio_dtype
 * (reference Export_GTCRN.py:47-48: 'F16' | 'F32' | 'INT16'). */
enum { ADN_F32 = 0, ADN_I16 = 1, ADN_F16 = 2 };

typedef struct adn_model adn_model;

/* One named fp32 tensor inside the flat weight blob (offset/count in floats). */
typedef struct {
  const char* name;
  uint64_t offset;
  uint64_t count;
} adn_tensor_entry;

/* Model descriptor: the string key/value metadata the reference stamps into the ONNX
 * model (audio_onnx_metadata.py:161-204; e.g. model_family, input_audio_length,
 * input_audio_dtype, nfft, hop_length, ...) plus the tensor index of the blob. */
typedef struct {
  const char* const* keys;
  const char* const* values;
  int32_t n_kv;
  const adn_tensor_entry* tensors;
  int32_t n_tensors;
} adn_desc;

/* I/O description == what session.get_inputs()/get_outputs() report
 * (Inference_GTCRN_ONNX.py:262-267,276-277): name, dtype, (1, channels, length). */
typedef struct {
  char name[32];
  int32_t dtype;      /* ADN_F32 | ADN_I16 | ADN_F16 */
  int32_t channels;
  int32_t length;     /* samples per chunk */
} adn_tensor_info;

/* Replaces onnxruntime.InferenceSession(path, ...) (Inference_GTCRN_ONNX.py:213-214,237):
 * builds a model of desc["model_family"] (gtcrn | zipenhancer | mel_band_roformer | mossformer2_se |
 * mossformer2_ss | mossformergan_se | dfsmn | ulunas | h_gtcrn)
 * on CUDA device `device_id` from a host blob of `nfloats` fp32 values.  The keys are the reference's
 * metadata keys (audio_onnx_metadata.py:115-205) plus, for mossformer2_se, the optional
 * "matmul_dtype" = F32 (default: 3xTF32 tensor-core GEMMs, fp32-class) | BF16 (the layers' GEMMs on bf16
 * operands -- BASELINE.json configs[2] "bf16 matmuls").  On failure *out is NULL and adn_last_error(NULL) has the reason.
 * Window limits: `input_audio_length` is ONE window of the model (the reference's static export length; its in-graph batch
 * fold is the leading batch dimension here).  mel_band_roformer accepts windows of at most 255 hops (112 455 samples, 2.55 s
 * at 44.1 kHz: the reference's 1.5 s fold window, Mel_Band_Roformer/Stereo/Export_MelBandRoformer.py:46-50) -- a longer
 * un-folded static length (Inference_MelBandRoformer_ONNX.py:301-313) is rejected at adn_create and must be folded on the
 * host (adn.chunker.denoise does); zipenhancer sequences (frames, sub-bands) of more than 256 positions run on the
 * slower one-thread-per-row attention path; h_gtcrn (two microphones in, one channel out: H-GTCRN/Export_H_GTCRN.py:1145-1152,
 * input_channels = 2) takes windows of k * 256 >= 512 samples at 16 kHz and the reference's front-end constants (wpe_rt60 0.3,
 * wpe_delay 2, wpe_iter 1, cg_solve_iter 6, iva_iter 10); every other family takes any window its STFT geometry allows. */
adn_status adn_create(adn_model** out, const adn_desc* desc, const float* weights,
                      size_t nfloats, int device_id);

/* Replaces session.get_inputs()/get_outputs()/_inputs_meta (…:262-267,276-277).  `outs` must have
 * room for 4 entries; *n_out receives how many the model has: 1 (`denoised_audio`) for every family except
 * mossformer2_ss, which reports 2 (`separated_0`, `separated_1`;
 * MossFormer2_SS_16K/Export_MossFormer2_SS_16K.py:689-690, Inference_MossFormer_SS_ONNX.py:312-317). */
adn_status adn_io_info(const adn_model* m, adn_tensor_info* in, adn_tensor_info* outs,
                       int32_t* n_out);

/* Replaces session.run_with_iobinding(binding) for a batch of `batch` independent
 * (1,C,L) chunks resident on the device (…:209-210, 314-317).  d_in is (batch,C,L)
 * contiguous in the input dtype, d_outs[i] (i < n_out of adn_io_info) is (batch,C_out,L_out) with the channel count
 * adn_io_info reports for that output (h_gtcrn: C = 2, C_out = 1).
 * Asynchronous on `stream` (a cudaStream_t passed as void*).  On a non-default stream the launch sequence of a
 * (buffers, batch) combination seen before is replayed as one CUDA graph (captured on its second run; environment
 * ADN_GRAPHS=0 disables this); the legacy default stream always runs the kernels one by one.  Families: gtcrn,
 * zipenhancer, mel_band_roformer, mossformer2_se, mossformer2_ss, mossformergan_se, dfsmn, ulunas, h_gtcrn. */
adn_status adn_run(adn_model* m, const void* d_in, void* const* d_outs, int32_t batch,
                   void* stream);

/* Same call with HOST buffers: copies in (pinned staging), runs, copies out and
 * synchronises -- the exact contract of process_segment() (…:314-317) where the
 * OrtValues live on the CPU. */
adn_status adn_run_host(adn_model* m, const void* h_in, void* const* h_outs, int32_t batch);

/* Bytes of device workspace the library holds for `batch` chunks. */
size_t adn_workspace_bytes(const adn_model* m, int32_t batch);

/* Number of CUDA kernels one adn_run() of `batch` chunks launches. */
int32_t adn_launches_per_run(const adn_model* m, int32_t batch);

/* Test/diagnostic hook: copy a named intermediate of the LAST adn_run() into a host
 * fp32 buffer of `count` floats; *actual receives the tensor's true element count. */
adn_status adn_debug_read(adn_model* m, const char* name, float* h_dst, size_t count,
                          size_t* actual);

/* Test/diagnostic hook: make adn_run() return after its first `n_launches` kernels
 * (0 = run everything) so ping-pong workspace buffers can be read mid-pipeline. */
adn_status adn_debug_stop_after(adn_model* m, int32_t n_launches);

/* Device time of the last adn_run() broken down per kernel (ms, CUDA events on the
 * run stream); only recorded when adn_set_profiling(m, 1).  names[i] points into
 * library-owned storage. */
adn_status adn_set_profiling(adn_model* m, int32_t enabled);
adn_status adn_last_kernel_times(adn_model* m, const char** names, float* ms, int32_t cap,
                                 int32_t* n);

const char* adn_last_error(const adn_model* m);
void adn_destroy(adn_model* m);

/* ---- stand-alone DSP operators (reference STFT_Process forward variants) ---------- */

/* Geometry of an STFT/ISTFT pair (metadata keys nfft / hop_length / center_pad /
 * pad_mode; STFT_Process.__init__, GTCRN/STFT_Process.py:144-211). */
typedef struct {
  int32_t nfft;
  int32_t hop;
  int32_t center;        /* 1: pad nfft/2 both sides */
  int32_t pad_reflect;   /* 1: reflect, 0: zeros */
  int32_t norm_multiply; /* 1: multiply by reciprocal table (ZipEnhancer), 0: divide */
} adn_stft_geom;

typedef struct adn_stft adn_stft;

/* fwd_basis: (2F, nfft) windowed DFT rows == STFT_Process.stft_kernel;
 * inv_basis: (2F, nfft) == STFT_Process.inverse_kernel; win_norm: (L_out) overlap-added
 * w^2 (or its reciprocal when norm_multiply) for `n_frames` frames.  All host fp32. */
adn_status adn_stft_create(adn_stft** out, const adn_stft_geom* g, const float* fwd_basis,
                           const float* inv_basis, const float* win_norm, int32_t n_frames,
                           int device_id);
/* _stft_B_packed_forward: d_x (batch, L) fp32 -> d_spec (batch, 2F, T) fp32. */
adn_status adn_stft_forward(adn_stft* s, const float* d_x, float* d_spec, int32_t batch,
                            int32_t length, void* stream);
/* _istft_B_packed_forward: d_spec (batch, 2F, T) -> d_y (batch, L_out) fp32. */
adn_status adn_stft_inverse(adn_stft* s, const float* d_spec, float* d_y, int32_t batch,
                            int32_t n_frames, void* stream);
void adn_stft_destroy(adn_stft* s);

/* ---- stand-alone conditioning / feature / recombine / output operators ------------------------
 * The wrapper-forward steps either side of the ZipEnhancer, MossFormerGAN-SE-16K and MossFormer2-SS-16K
 * backbones (the zipenhancer and mossformergan_se models run these operators internally) and the linear resampler every wrapper shares.  With adn_stft_forward / adn_stft_inverse they form the complete front and back ends
 * around those backbones.  All pointers are DEVICE pointers, rows are contiguous, `stream` is a
 * cudaStream_t (NULL = default stream); errors go to adn_last_error(NULL). */
enum { ADN_FAMILY_ZIPENHANCER = 1, ADN_FAMILY_MOSSFORMERGAN = 2, ADN_FAMILY_MOSSFORMER2_SS = 3 };

/* torch.nn.functional.interpolate(mode='linear', align_corners=False) (GTCRN/Export_GTCRN.py:638-654,
 * ZipEnhancer/Export_ZipEnhancer.py:826-832, :906-912): d_in (rows, len_in) of in_dtype -> d_out
 * (rows, len_out) fp32.  scale_factor > 0: coordinate scale 1/scale_factor (the scale_factor= form);
 * otherwise len_in/len_out (the size= form). */
adn_status adn_resample_linear(const void* d_in, int32_t in_dtype, float* d_out, int32_t rows,
                               int32_t len_in, int32_t len_out, double scale_factor, void* stream);
/* Per-window RMS normalisation (Export_ZipEnhancer.py:839-840; MossFormerGAN_SE_16K/Export_MossFormer_SE.py
 * :564-568): y = (x*pre_scale) / sqrt(mean((x*pre_scale)^2) + eps); d_norm (rows) receives the factor.
 * len_out > len appends the head of the window (MossFormerGAN's wrap-around pad to a hop multiple). */
adn_status adn_rms_normalize(const void* d_in, int32_t in_dtype, float pre_scale, float eps, float* d_out,
                             float* d_norm, int32_t rows, int32_t len, int32_t len_out, void* stream);
/* MossFormer2_SS_16K norm_audio (Export_MossFormer2_SS_16K.py:403-423): two-stage RMS normalisation of raw
 * PCM amplitude; d_rms_in (rows) receives the gain-restore reference. */
adn_status adn_two_stage_rms(const void* d_in, int32_t in_dtype, float target, float eps, float* d_out,
                             float* d_rms_in, int32_t rows, int32_t len, void* stream);
/* Packed spectrum (batch, 2F, T) -> backbone input (batch, C, T, F).  ZIPENHANCER (:843-850): C = 2,
 * [(re^2+im^2+1e-9)^0.15, atan2(im, re+1e-5)].  MOSSFORMERGAN (:578-586): C = 3, power-law compressed
 * [magnitude, re, im]; d_keep (batch, 2, F, T) receives the compressed complex spectrum for recombine. */
adn_status adn_spec_features(int32_t family, const float* d_spec, float* d_feat, float* d_keep, int32_t batch,
                             int32_t fbins, int32_t frames, void* stream);
/* Backbone outputs -> packed spectrum (batch, 2F, T) for adn_stft_inverse.  ZIPENHANCER (:882-892):
 * d_a = mask-decoder output (batch,1,T,F), d_b = rectangular phase (batch,2,T,F).  MOSSFORMERGAN (:863-868):
 * d_a = mask (batch,F,T), d_b = complex branch (batch,2,F,T), d_keep from adn_spec_features. */
adn_status adn_spec_recombine(int32_t family, const float* d_a, const float* d_b, const float* d_keep,
                              float* d_spec, int32_t batch, int32_t fbins, int32_t frames, void* stream);
/* ISTFT output (rows, len_src) -> model output (rows, len) of out_dtype: trim, gain, family rule
 * (Export_ZipEnhancer.py:899-926; MossFormerGAN :880-897; MossFormer2_SS :627-660).  d_gain holds one value per
 * `gain_group` consecutive rows: the norm factor (ZIPENHANCER, MOSSFORMERGAN) or rms_in (MOSSFORMER2_SS, rows =
 * windows x speakers, gain_group = speakers). */
adn_status adn_condition_output(int32_t family, const float* d_wave, int32_t len_src, const float* d_gain,
                                int32_t gain_group, void* d_out, int32_t out_dtype, int32_t rows, int32_t len,
                                void* stream);

/* Library/build identification: returns e.g. "adn 0.1 sm_100a". */
const char* adn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ADN_H_ */
