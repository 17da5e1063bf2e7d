// Internal interface between the C ABI (api.cu) and per-family model implementations.
#pragma once
#include "adn.h"

#include <cuda_runtime.h>
#include <map>
#include <string>

struct TensorRef {
  uint64_t offset, count;
};

typedef void (*ImplTickFn)(void* ctx, const char* name);

struct ModelImpl {
  std::string err;
  ImplTickFn tick = nullptr;   // called after every launch (per-kernel event timing), may be null
  void* tick_ctx = nullptr;
  virtual ~ModelImpl() {}
  virtual void io_info(adn_tensor_info* in, adn_tensor_info* out) = 0;
  virtual adn_status run(const void* d_in, void* d_out, int batch, cudaStream_t st) = 0;
  // families with several outputs (MossFormer2-SS: one waveform per speaker) override these two;
  // io_info() then fills n_outputs() consecutive entries of `out`
  virtual int n_outputs() { return 1; }
  virtual adn_status run_multi(const void* d_in, void* const* d_outs, int batch, cudaStream_t st) {
    return run(d_in, d_outs[0], batch, st);
  }
  virtual size_t workspace_bytes(int batch) = 0;
  virtual int launches(int batch) = 0;
  virtual adn_status debug_read(const char* name, float* h_dst, size_t count, size_t* actual) = 0;
  virtual void set_stop_after(int n) = 0;
};

// Mel-Band-Roformer (stereo): csrc/mbr.cu
ModelImpl* mbr_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                      const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// MossFormer2-SE-48K: csrc/mf2se.cu
ModelImpl* mf2se_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// MossFormer2-SS-16K: csrc/mf2ss.cu
ModelImpl* mf2ss_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// MossFormerGAN-SE-16K: csrc/mfgan.cu
ModelImpl* mfgan_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// DFSMN 48 kHz: csrc/dfsmn.cu
ModelImpl* dfsmn_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                        const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// UL-UNAS 16 kHz: csrc/ulunas.cu
ModelImpl* ulunas_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                         const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// ZipEnhancer 16 kHz: csrc/zipenh.cu
ModelImpl* zipenh_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                         const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// H-GTCRN 16 kHz stereo (WPE + AuxIVA front end, GTCRN_IVA network): csrc/hgtcrn.cu
ModelImpl* hgtcrn_create(const std::map<std::string, std::string>& meta, const std::map<std::string, TensorRef>& index,
                         const float* h_blob, float* d_blob, int device, int sms, std::string& err);

// frame-major STFT / ISTFT on a stand-alone operator handle (api.cu)
int adn_stft_ld(const adn_stft* s);
int adn_stft_pad_frames(const adn_stft* s);
int adn_stft_padded_len(const adn_stft* s, int length);
// opt-in 3xTF32 tensor-core path of the frame-major transforms (and of adn_stft_inverse, which packs and then runs the
// frame-major inverse); returns the number of floats of finite slack the inverse needs behind its input buffer, 0 if the
// geometry does not qualify (the exact fp32 GEMM stays in use)
int adn_stft_enable_tc(adn_stft* s, int sms);
adn_status adn_stft_forward_fm(adn_stft* s, const float* d_xp, float* d_spec_fm, int rows, int n_frames, int Lp, cudaStream_t st);
adn_status adn_stft_inverse_fm(adn_stft* s, const float* d_fm_padded, float* d_y, int rows, int n_frames, cudaStream_t st);
