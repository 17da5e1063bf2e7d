// C ABI of libadn.so (include/adn.h).  Host-side C++ only: model construction from the
// flat weight blob, workspace management, stream-ordered launch sequence.
#include "adn.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "gtcrn.cuh"
#include "model_impl.h"

#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

std::string g_last_error;   // for failures without a handle (adn_create)
std::mutex g_err_mu;

void set_global_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = s;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Overlap-add weight for the row-gather GEMM: raw output block j (hop samples) is
//   sum_{q'=0}^{R-1} frame[j-R+1+q'] . Kinv[:, n' + (R-1-q')*hop]
// -> W[n'][q'*ld + r] (zero where the tap falls outside the frame).  This is exactly
// conv_transpose1d(inp, inverse_kernel, stride=hop) (GTCRN/STFT_Process.py:328).
std::vector<float> build_ola_weight(const float* inv_basis, int nfft, int hop, int ld, int R) {
  const int rows2f = nfft + 2;
  std::vector<float> w((size_t)hop * R * ld, 0.f);
  for (int n = 0; n < hop; ++n)
    for (int q = 0; q < R; ++q) {
      int src = n + (R - 1 - q) * hop;
      if (src >= nfft) continue;
      for (int r = 0; r < rows2f; ++r) w[(size_t)n * R * ld + (size_t)q * ld + r] = inv_basis[(size_t)r * nfft + src];
    }
  return w;
}

struct StftPlan {
  int nfft = 0, hop = 0, half = 0, center = 1, reflect = 1, norm_mul = 0;
  int rows2f = 0, ld = 0, R = 0;
  int n_frames(int L) const { return center ? L / hop + 1 : (L - nfft) / hop + 1; }
  int padded_len(int L) const { return round_up(L + (center ? 2 * half : 0), 4); }
  int out_len(int T) const { int raw = nfft + hop * (T - 1); return center ? raw - 2 * half : raw; }
  int pad_frames() const { return R - 1; }
  int blk_lo() const { return center ? half / hop : 0; }
  int blk_hi(int T) const {   // inclusive last raw block that intersects the kept range
    int raw = nfft + hop * (T - 1);
    int end = center ? raw - half : raw;
    return (end - 1) / hop;
  }
  void init(int nfft_, int hop_, int center_, int reflect_, int norm_mul_) {
    nfft = nfft_; hop = hop_; half = nfft_ / 2; center = center_; reflect = reflect_; norm_mul = norm_mul_;
    rows2f = nfft + 2;
    ld = round_up(rows2f, 8);
    R = (nfft + hop - 1) / hop;
  }
};

void fill_stft_gemm(GemmArgs& g, const StftPlan& p, const float* xp, int Lp, const float* fwd, int B, int T,
                    float* C, long long c_sB, long long c_sT, long long c_sN) {
  memset(&g, 0, sizeof(g));
  g.A = xp; g.a_sB = Lp; g.a_sT = p.hop; g.a_t0 = 0; g.TM = T;
  g.W = fwd; g.ldw = p.nfft;
  g.M = B * T; g.N = p.rows2f; g.K = p.nfft;
  g.C = C; g.c_sB = c_sB; g.c_sT = c_sT; g.c_sN = c_sN;
}

void fill_istft_gemm(GemmArgs& g, const StftPlan& p, const float* enh_padded, const float* ola_w,
                     const float* norm, int B, int T, void* out, int out_dtype) {
  memset(&g, 0, sizeof(g));
  const int lo = p.blk_lo(), hi = p.blk_hi(T);
  g.A = enh_padded; g.a_sB = (long long)(T + 2 * p.pad_frames()) * p.ld; g.a_sT = p.ld; g.a_t0 = lo;
  g.TM = hi - lo + 1;
  g.W = ola_w; g.ldw = p.R * p.ld;
  g.M = B * g.TM; g.N = p.hop; g.K = p.R * p.ld;
  g.norm = norm; g.norm_mul = p.norm_mul; g.hop = p.hop; g.shift = p.center ? p.half : 0;
  g.out_len = p.out_len(T); g.out_dtype = out_dtype; g.out = out;
}

int choose_bn(int N) {
  const int cands[3] = {256, 176, 128};
  int best = 128, best_pad = 1 << 30;
  for (int c : cands) {
    int pad = (N + c - 1) / c * c;
    if (pad < best_pad) { best_pad = pad; best = c; }
  }
  return best;
}

// (rows, cols) row-major -> zero-padded (n_pad, k_pad) hi|lo planes (3xTF32 operand split)
std::vector<float> split_pad_weight(const float* w, int rows, int cols, int n_pad, int k_pad) {
  std::vector<float> out((size_t)2 * n_pad * k_pad, 0.f);
  float* hi = out.data();
  float* lo = out.data() + (size_t)n_pad * k_pad;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      float v = w[(size_t)r * cols + c];
      uint32_t bits;
      memcpy(&bits, &v, 4);
      bits &= 0xFFFFE000u;
      float h;
      memcpy(&h, &bits, 4);
      hi[(size_t)r * k_pad + c] = h;
      lo[(size_t)r * k_pad + c] = v - h;
    }
  return out;
}

void tile_rows(int TM, int& bt, int& bb, int& tpc) {
  if (TM >= 128) { bt = 128; bb = 1; tpc = (TM + 127) / 128; }
  else { bt = TM; bb = 128 / TM; tpc = 1; }
}

}  // namespace

// last-error slot for handle-less entry points in other translation units (ends.cu)
void adn_internal_set_error(const std::string& s) { set_global_error(s); }

// ===================================================================================
struct adn_model {
  std::string err;
  ModelImpl* impl = nullptr;   // non-GTCRN families (GTCRN state is inline below)
  int device = 0;
  std::string family;
  std::map<std::string, std::string> meta;
  std::map<std::string, TensorRef> index;
  float* d_blob = nullptr;
  size_t nfloats = 0;

  int in_dtype = ADN_F32, out_dtype = ADN_F32;
  int L = 0, T = 0, Lp = 0, Lout = 0, chans = 1, out_chans = 1;   // (H-GTCRN: two microphones in, one channel out)  // GTCRN: L / Lout are the MODEL-rate window and its ISTFT length
  // GTCRN linear resampling (Export_GTCRN.py:619-632, :638-654, :671-688): the caller's window is io_L samples at
  // in_sample_rate, the output io_Lout samples at out_sample_rate; the inner run works on fp32 model-rate buffers
  bool rs_in = false, rs_out = false;
  int io_L = 0, io_Lout = 0, in_sr = 16000, out_sr = 16000;
  double in_scale_factor = 1.0, out_scale_factor = 1.0;
  int run_in_dtype = ADN_F32, run_out_dtype = ADN_F32, run_remove_dc = 1;
  float *rs_xc = nullptr, *rs_xm = nullptr, *rs_ym = nullptr, *rs_yr = nullptr;
  int rs_cap = 0;
  StftPlan stft;
  const float* d_fwd = nullptr;
  float* d_ola = nullptr;
  float* d_norm = nullptr;

  // tensor-core (tcgen05, 3xTF32) DFT GEMMs
  bool use_tc = false;
  int sms = 148;
  float* d_wf_hl = nullptr;   // forward basis  hi|lo planes, (n_pad, k_pad) each
  float* d_wo_hl = nullptr;   // overlap-add W  hi|lo planes
  int wf_npad = 0, wf_kpad = 0, wo_npad = 0, wo_kpad = 0;
  float* xp_hl = nullptr;     // A planes (hi|lo) of the padded waveform
  float* enh_hl = nullptr;    // A planes (hi|lo) of the enhanced spectrum
  size_t enh_plane = 0;
  tc::TcPlan stft_plan, istft_plan;
  tc::TcArgs stft_args{}, istft_args{};

  gtcrn::Weights w;
  gtcrn::Buffers buf{};
  int capacity = 0;
  size_t ws_bytes = 0;
  std::vector<void*> allocs;
  void* d_in = nullptr;     // device staging for adn_run_host
  void* d_out = nullptr;    // n_out consecutive (batch, chans, Lout) blocks
  int n_out = 1;
  int io_cap = 0;           // staging capacity when `impl` owns the workspace
  cudaStream_t own_stream = nullptr;
  cudaStream_t st_in = nullptr, st_out = nullptr;      // copy streams of adn_run_host
  cudaEvent_t ev_h2d[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_done[4] = {nullptr, nullptr, nullptr, nullptr};

  bool profiling = false;
  std::vector<cudaEvent_t> events;
  std::vector<const char*> ev_names;
  std::vector<float> ev_ms;
  size_t ev_used = 0;
  cudaStream_t ev_stream = nullptr;
  int last_launches = 0;
  int last_batch = 0;
  int stop_after = 0;       // diagnostics: stop the launch sequence after N kernels

  // CUDA graphs of whole runs (adn_run): one executable graph per (input, outputs, batch, stream); the first run of a batch
  // size is eager (workspace allocation, GEMM planning), the second is captured, later ones are one cudaGraphLaunch
  struct GraphEntry {
    const void* in; void* outs[4]; int batch; cudaStream_t st; cudaGraphExec_t exec; unsigned long long epoch;
  };
  std::vector<GraphEntry> graphs;
  std::map<int, int> eager_runs;
  bool use_graphs = true;
  unsigned long long graph_launches = 0;
};

namespace {

size_t dtype_size(int dt) { return dt == ADN_F32 ? 4 : 2; }

int parse_dtype(const std::string& s, int& out) {
  if (s == "F32") out = ADN_F32;
  else if (s == "INT16") out = ADN_I16;
  else if (s == "F16") out = ADN_F16;
  else return 0;
  return 1;
}

const float* dptr(adn_model* m, const std::string& name, size_t expect, bool& ok) {
  auto it = m->index.find(name);
  if (it == m->index.end()) {
    if (ok) m->err = "weight blob has no tensor '" + name + "'";
    ok = false;
    return nullptr;
  }
  if (expect && it->second.count != expect) {
    if (ok) m->err = "tensor '" + name + "' has " + std::to_string(it->second.count) + " floats, expected " +
                     std::to_string(expect);
    ok = false;
    return nullptr;
  }
  return m->d_blob + it->second.offset;
}

template <typename S>
void load_struct(adn_model* m, const float* hblob, const std::string& name, S& dst, bool& ok) {
  auto it = m->index.find(name);
  if (it == m->index.end() || it->second.count != sizeof(S) / sizeof(float)) {
    if (ok) m->err = "tensor '" + name + "' missing or not " + std::to_string(sizeof(S) / sizeof(float)) + " floats";
    ok = false;
    return;
  }
  memcpy(&dst, hblob + it->second.offset, sizeof(S));
}

gtcrn::GruPtrs gru_ptrs(adn_model* m, const std::string& p, int I, int H, bool& ok) {
  gtcrn::GruPtrs g;
  g.w_ih = dptr(m, p + ".w_ih", (size_t)3 * H * I, ok);
  g.w_hh = dptr(m, p + ".w_hh", (size_t)3 * H * H, ok);
  g.b_ih = dptr(m, p + ".b_ih", (size_t)3 * H, ok);
  g.b_hh = dptr(m, p + ".b_hh", (size_t)3 * H, ok);
  return g;
}

adn_status dev_alloc(adn_model* m, void** p, size_t bytes, bool zero) {
  ADN_CUDA_TRY(cudaMalloc(p, bytes), m->err);
  m->allocs.push_back(*p);
  m->ws_bytes += bytes;
  if (zero) ADN_CUDA_TRY(cudaMemset(*p, 0, bytes), m->err);
  return ADN_OK;
}

void free_workspace(adn_model* m) {
  adn_note_free();
  for (void* p : m->allocs) cudaFree(p);
  m->allocs.clear();
  m->ws_bytes = 0;
  m->capacity = 0;
}

size_t workspace_bytes_for(const adn_model* m, int B) {
  using namespace gtcrn;
  size_t T = m->T, f = 0;
  f += (size_t)B * m->Lp;                       // xp
  f += (size_t)B * T * SPEC_LD;                 // spec
  f += (size_t)B * T * FRAME_E0;                // e0
  f += (size_t)4 * B * T * FRAME16;             // e1..e4
  f += (size_t)B * T * 8 * E1_F + (size_t)B * T * (8 + 16 + 48);   // h1, zt, at, tgi
  f += (size_t)B * T * 3 * FRAME16;             // gi
  if (m->use_tc) f += (size_t)2 * B * m->Lp + (size_t)2 * (B * (T + 2 * m->stft.pad_frames()) * SPEC_LD + m->wo_kpad);
  f += (size_t)3 * B * T * FRAME16;             // xa, xb, inter
  f += (size_t)B * (T + 2 * m->stft.pad_frames()) * SPEC_LD;   // enh
  size_t bytes = f * sizeof(float);
  bytes += (size_t)B * m->io_L * dtype_size(m->in_dtype) + (size_t)B * m->io_Lout * dtype_size(m->out_dtype);
  return bytes;
}

adn_status ensure_capacity(adn_model* m, int B) {
  using namespace gtcrn;
  if (B <= m->capacity) return ADN_OK;
  ADN_CUDA_TRY(cudaDeviceSynchronize(), m->err);
  free_workspace(m);
  const size_t T = m->T;
  adn_status s;
#define A(ptr, nfl, zero) \
  if ((s = dev_alloc(m, (void**)&(ptr), (size_t)(nfl) * sizeof(float), zero)) != ADN_OK) return s;
  A(m->buf.xp, (size_t)B * m->Lp, false);
  A(m->buf.spec, (size_t)B * T * SPEC_LD, true);
  A(m->buf.e0, (size_t)B * T * FRAME_E0, false);
  m->buf.e[0] = nullptr;
  for (int i = 1; i <= 4; ++i) A(m->buf.e[i], (size_t)B * T * FRAME16, false);
  A(m->buf.h1, (size_t)B * T * 8 * E1_F, false);
  A(m->buf.zt, (size_t)B * T * 8, false);
  A(m->buf.at, (size_t)B * T * 8, false);
  A(m->buf.tgi, (size_t)B * T * 48, false);
  A(m->buf.thid, (size_t)B * T * 16, false);
  A(m->buf.gi, (size_t)B * T * 3 * FRAME16, false);
  m->buf.xp_hi = m->buf.xp_lo = m->buf.enh_hi = m->buf.enh_lo = nullptr;
  A(m->buf.xa, (size_t)B * T * FRAME16, false);
  A(m->buf.xb, (size_t)B * T * FRAME16, false);
  A(m->buf.inter, (size_t)B * T * FRAME16, false);
  A(m->buf.enh, (size_t)B * (T + 2 * m->stft.pad_frames()) * SPEC_LD, true);
  if (m->use_tc) {
    const StftPlan& p = m->stft;
    const size_t xp_plane = (size_t)B * m->Lp;
    m->enh_plane = (size_t)B * (T + 2 * p.pad_frames()) * SPEC_LD + m->wo_kpad;   // + slack for the k overrun
    A(m->xp_hl, 2 * xp_plane, false);
    A(m->enh_hl, 2 * m->enh_plane, true);
    m->buf.xp_hi = m->xp_hl; m->buf.xp_lo = m->xp_hl + xp_plane;
    m->buf.enh_hi = m->enh_hl; m->buf.enh_lo = m->enh_hl + m->enh_plane;
    int bt, bb, tpc;
    // forward: rows = frames, row stride = hop
    tile_rows((int)T, bt, bb, tpc);
    tc::TcArgs& a = m->stft_args;
    a = tc::TcArgs{};
    a.bb = bb; a.bt = bt; a.tiles_per_chunk = tpc; a.t0 = 0; a.TM = (int)T;
    a.N = p.rows2f; a.K = p.nfft;
    a.C = m->buf.spec; a.c_sB = (long long)T * SPEC_LD; a.c_sT = SPEC_LD;
    if (!tc::make_row_map(&m->stft_plan.map_a_hi, m->xp_hl, p.nfft, (int)T, p.hop, B, m->Lp, bt, bb, m->err) ||
        !tc::make_row_map(&m->stft_plan.map_a_lo, m->xp_hl + xp_plane, p.nfft, (int)T, p.hop, B, m->Lp, bt, bb, m->err))
      return ADN_ERR_CUDA;
    // inverse: rows = raw hop-blocks, each a run of R consecutive (zero-framed) spectrum frames
    const int lo = p.blk_lo(), hi = p.blk_hi((int)T), TM = hi - lo + 1;
    tile_rows(TM, bt, bb, tpc);
    tc::TcArgs& c = m->istft_args;
    c = tc::TcArgs{};
    c.bb = bb; c.bt = bt; c.tiles_per_chunk = tpc; c.t0 = lo; c.TM = TM;
    c.N = p.hop; c.K = p.R * p.ld;
    c.norm = m->d_norm; c.norm_mul = p.norm_mul; c.hop = p.hop; c.shift = p.center ? p.half : 0;
    c.out_len = m->Lout; c.out_dtype = m->run_out_dtype;
    const int rows = (int)T + 2 * p.pad_frames();
    if (!tc::make_row_map(&m->istft_plan.map_a_hi, m->enh_hl, m->wo_kpad, rows, p.ld, B, (long long)rows * p.ld, bt, bb, m->err) ||
        !tc::make_row_map(&m->istft_plan.map_a_lo, m->enh_hl + m->enh_plane, m->wo_kpad, rows, p.ld, B,
                          (long long)rows * p.ld, bt, bb, m->err))
      return ADN_ERR_CUDA;
  }
  if ((s = dev_alloc(m, &m->d_in, (size_t)B * m->io_L * dtype_size(m->in_dtype), false)) != ADN_OK) return s;
  if ((s = dev_alloc(m, &m->d_out, (size_t)B * m->io_Lout * dtype_size(m->out_dtype), false)) != ADN_OK) return s;
#undef A
  ADN_CUDA_TRY(cudaDeviceSynchronize(), m->err);   // the zero-fills above ran on the legacy default stream
  m->capacity = B;
  return ADN_OK;
}

void tick_cb(void* ctx, const char* name) {
  adn_model* m = (adn_model*)ctx;
  if (!m->profiling) return;
  if (m->ev_used >= m->events.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    m->events.push_back(e);
    m->ev_names.push_back(name);
  }
  m->ev_names[m->ev_used] = name;
  cudaEventRecord(m->events[m->ev_used], m->ev_stream);
  ++m->ev_used;
}

adn_status build_gtcrn(adn_model* m, const float* hblob) {
  bool ok = true;
  auto& w = m->w;
  load_struct(m, hblob, "enc_front", w.enc_front, ok);
  load_struct(m, hblob, "dec_tail", w.dec_tail, ok);
  for (int i = 0; i < 3; ++i) {
    load_struct(m, hblob, "enc_gt." + std::to_string(i), w.enc_gt[i], ok);
    load_struct(m, hblob, "dec_gt." + std::to_string(i), w.dec_gt[i], ok);
    std::string pe = "enc_tra." + std::to_string(i), pd = "dec_tra." + std::to_string(i);
    w.enc_tra[i].gru = gru_ptrs(m, pe, 8, 16, ok);
    w.enc_tra[i].fc_w = dptr(m, pe + ".fc_w", 128, ok);
    w.enc_tra[i].fc_b = dptr(m, pe + ".fc_b", 8, ok);
    w.dec_tra[i].gru = gru_ptrs(m, pd, 8, 16, ok);
    w.dec_tra[i].fc_w = dptr(m, pd + ".fc_w", 128, ok);
    w.dec_tra[i].fc_b = dptr(m, pd + ".fc_b", 8, ok);
  }
  for (int i = 0; i < 2; ++i) {
    std::string p = "dp." + std::to_string(i);
    for (int g = 0; g < 2; ++g) {
      for (int d = 0; d < 2; ++d)
        w.dp[i].intra[g][d] = gru_ptrs(m, p + ".intra." + std::to_string(g) + "." + std::to_string(d), 8, 4, ok);
      w.dp[i].inter[g] = gru_ptrs(m, p + ".inter." + std::to_string(g), 8, 8, ok);
    }
    w.dp[i].intra_fc_w = dptr(m, p + ".intra_fc_w", 256, ok);
    w.dp[i].intra_fc_b = dptr(m, p + ".intra_fc_b", 16, ok);
    w.dp[i].intra_ln_w = dptr(m, p + ".intra_ln_w", 528, ok);
    w.dp[i].intra_ln_b = dptr(m, p + ".intra_ln_b", 528, ok);
    w.dp[i].inter_fc_w = dptr(m, p + ".inter_fc_w", 256, ok);
    w.dp[i].inter_fc_b = dptr(m, p + ".inter_fc_b", 16, ok);
    w.dp[i].inter_ln_w = dptr(m, p + ".inter_ln_w", 528, ok);
    w.dp[i].inter_ln_b = dptr(m, p + ".inter_ln_b", 528, ok);
  }
  w.erb.bm = dptr(m, "erb.bm", 192 * 64, ok);
  w.erb.bm_lo = dptr(m, "erb.bm_lo", 64, ok);
  w.erb.bm_hi = dptr(m, "erb.bm_hi", 64, ok);
  w.erb.bs = dptr(m, "erb.bs", 64 * 192, ok);
  w.erb.bs_lo = dptr(m, "erb.bs_lo", 192, ok);
  w.erb.bs_hi = dptr(m, "erb.bs_hi", 192, ok);
  m->d_fwd = dptr(m, "stft.fwd", (size_t)514 * 512, ok);
  if (!ok) return ADN_ERR_INVALID;

  auto inv = m->index.find("istft.inv");
  auto nrm = m->index.find("istft.norm");
  if (inv == m->index.end() || inv->second.count != (size_t)514 * 512) {
    m->err = "tensor 'istft.inv' missing or wrong size";
    return ADN_ERR_INVALID;
  }
  if (nrm == m->index.end() || nrm->second.count != (size_t)m->Lout) {
    m->err = "tensor 'istft.norm' must have output_audio_length floats";
    return ADN_ERR_INVALID;
  }
  std::vector<float> ola = build_ola_weight(hblob + inv->second.offset, m->stft.nfft, m->stft.hop, m->stft.ld, m->stft.R);
  ADN_CUDA_TRY(cudaMalloc((void**)&m->d_ola, ola.size() * sizeof(float)), m->err);
  ADN_CUDA_TRY(cudaMemcpy(m->d_ola, ola.data(), ola.size() * sizeof(float), cudaMemcpyHostToDevice), m->err);
  m->d_norm = m->d_blob + nrm->second.offset;

  // tensor-core path: pre-split, zero-padded weight planes + their TMA maps
  const char* env = getenv("ADN_GEMM");
  m->use_tc = !(env && std::string(env) == "ffma");
  if (m->use_tc) {
    const StftPlan& p = m->stft;
    auto fwd = m->index.find("stft.fwd");
    m->stft_plan.bn = choose_bn(p.rows2f);
    m->wf_npad = (p.rows2f + m->stft_plan.bn - 1) / m->stft_plan.bn * m->stft_plan.bn;
    m->wf_kpad = round_up(p.nfft, 32);
    std::vector<float> wf = split_pad_weight(hblob + fwd->second.offset, p.rows2f, p.nfft, m->wf_npad, m->wf_kpad);
    m->istft_plan.bn = choose_bn(p.hop);
    m->wo_npad = (p.hop + m->istft_plan.bn - 1) / m->istft_plan.bn * m->istft_plan.bn;
    m->wo_kpad = round_up(p.R * p.ld, 32);
    std::vector<float> wo = split_pad_weight(ola.data(), p.hop, p.R * p.ld, m->wo_npad, m->wo_kpad);
    ADN_CUDA_TRY(cudaMalloc((void**)&m->d_wf_hl, wf.size() * sizeof(float)), m->err);
    ADN_CUDA_TRY(cudaMemcpy(m->d_wf_hl, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice), m->err);
    ADN_CUDA_TRY(cudaMalloc((void**)&m->d_wo_hl, wo.size() * sizeof(float)), m->err);
    ADN_CUDA_TRY(cudaMemcpy(m->d_wo_hl, wo.data(), wo.size() * sizeof(float), cudaMemcpyHostToDevice), m->err);
    if (!tc::make_weight_map(&m->stft_plan.map_w_hi, m->d_wf_hl, m->wf_kpad, m->wf_npad, m->stft_plan.bn, m->err) ||
        !tc::make_weight_map(&m->stft_plan.map_w_lo, m->d_wf_hl + (size_t)m->wf_npad * m->wf_kpad, m->wf_kpad,
                             m->wf_npad, m->stft_plan.bn, m->err) ||
        !tc::make_weight_map(&m->istft_plan.map_w_hi, m->d_wo_hl, m->wo_kpad, m->wo_npad, m->istft_plan.bn, m->err) ||
        !tc::make_weight_map(&m->istft_plan.map_w_lo, m->d_wo_hl + (size_t)m->wo_npad * m->wo_kpad, m->wo_kpad,
                             m->wo_npad, m->istft_plan.bn, m->err))
      return ADN_ERR_CUDA;
  }
  return ADN_OK;
}


// Two-range pipelining hooks for adn_run_host: the front stage (conditioning + STFT) of a range
// starts as soon as its H2D copy has landed, the D2H copy of a range starts as soon as its ISTFT
// is done; the backbone in between runs once over the whole batch.
struct PhaseSync {
  int mid;                          // ranges [0, mid) and [mid, batch); mid is even
  cudaEvent_t h2d_done[2];          // waited on before the front stage of each range
  cudaEvent_t out_ready[2];         // recorded after the tail stage of each range
};

adn_status gtcrn_run(adn_model* m, const void* d_in, void* d_out, int batch, cudaStream_t st, const PhaseSync* ps) {
  adn_status s = ensure_capacity(m, batch);
  if (s != ADN_OK) return s;
  m->ev_used = 0;
  m->ev_stream = st;
  int n = 0;
  tick_cb(m, "start");
  const int nr = ps ? 2 : 1;
  const int lo[2] = {0, ps ? ps->mid : 0}, hi[2] = {ps ? ps->mid : batch, batch};
  const size_t in_row = (size_t)m->L * dtype_size(m->run_in_dtype);
  GemmArgs g;

  for (int r = 0; r < nr; ++r) {
    const int b0 = lo[r], nb = hi[r] - lo[r];
    if (ps) ADN_CUDA_TRY(cudaStreamWaitEvent(st, ps->h2d_done[r], 0), m->err);
    gtcrn::launch_prep((const char*)d_in + b0 * in_row, m->run_in_dtype, m->buf.xp + (size_t)b0 * m->Lp,
                       m->buf.xp_hi ? m->buf.xp_hi + (size_t)b0 * m->Lp : nullptr,
                       m->buf.xp_lo ? m->buf.xp_lo + (size_t)b0 * m->Lp : nullptr, nb, m->L, m->Lp, m->stft.half,
                       /*remove_dc=*/m->run_remove_dc, m->stft.reflect, st);
    if (r == 0) { ++n; tick_cb(m, "prep"); }
    if (m->use_tc) {
      tc::TcArgs a = m->stft_args;
      a.b_off = b0;
      a.B = hi[r];
      a.m_tiles = a.bb > 1 ? (nb + a.bb - 1) / a.bb : nb * a.tiles_per_chunk;
      ADN_CUDA_TRY(tc::launch(m->stft_plan, a, EPI_STORE, m->sms, st), m->err);
      if (r == 0) { ++n; tick_cb(m, "stft_gemm_tc"); }
    } else {
      fill_stft_gemm(g, m->stft, m->buf.xp + (size_t)b0 * m->Lp, m->Lp, m->d_fwd, nb, m->T,
                     m->buf.spec + (size_t)b0 * m->T * gtcrn::SPEC_LD, (long long)m->T * gtcrn::SPEC_LD,
                     gtcrn::SPEC_LD, 1);
      launch_gemm_ffma(g, EPI_STORE, st);
      if (r == 0) { ++n; tick_cb(m, "stft_gemm"); }
    }
  }

  gtcrn::Dims d{batch, m->L, m->Lp, m->T};
  const int stop_bb = m->stop_after > 0 ? (m->stop_after > n ? m->stop_after - n : 1) : 0;
  n += gtcrn::launch_backbone(m->w, m->buf, d, m->stft.pad_frames(), st, tick_cb, m, stop_bb);
  m->last_launches = n;
  m->last_batch = batch;
  if (m->stop_after > 0 && n >= m->stop_after) {
    ADN_CUDA_TRY(cudaGetLastError(), m->err);
    return ADN_OK;
  }

  const size_t out_row = (size_t)m->Lout * dtype_size(m->run_out_dtype);
  for (int r = 0; r < nr; ++r) {
    const int b0 = lo[r], nb = hi[r] - lo[r];
    if (m->use_tc) {
      tc::TcArgs a = m->istft_args;
      a.b_off = b0;
      a.B = hi[r];
      a.m_tiles = a.bb > 1 ? (nb + a.bb - 1) / a.bb : nb * a.tiles_per_chunk;
      a.out = d_out;
      ADN_CUDA_TRY(tc::launch(m->istft_plan, a, EPI_ISTFT, m->sms, st), m->err);
      if (r == 0) { ++n; tick_cb(m, "istft_gemm_tc"); }
    } else {
      fill_istft_gemm(g, m->stft, m->buf.enh + (size_t)b0 * (m->T + 2 * m->stft.pad_frames()) * gtcrn::SPEC_LD,
                      m->d_ola, m->d_norm, nb, m->T, (char*)d_out + b0 * out_row, m->run_out_dtype);
      launch_gemm_ffma(g, EPI_ISTFT, st);
      if (r == 0) { ++n; tick_cb(m, "istft_gemm"); }
    }
    if (ps) ADN_CUDA_TRY(cudaEventRecord(ps->out_ready[r], st), m->err);
  }
  m->last_launches = n;
  m->last_batch = batch;
  ADN_CUDA_TRY(cudaGetLastError(), m->err);
  return ADN_OK;
}

__global__ void rs_scale_kernel(float* __restrict__ x, long long n, float s) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[i] *= s;
}

// clamp / truncate to int16, fp32 copy or fp16 cast (Export_GTCRN.py:689-693)
__global__ void rs_convert_kernel(const float* __restrict__ src, void* __restrict__ out, int out_dtype, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  if (out_dtype == ADN_I16) reinterpret_cast<int16_t*>(out)[i] = (int16_t)(int)fminf(fmaxf(v, -32768.0f), 32767.0f);
  else if (out_dtype == ADN_F32) reinterpret_cast<float*>(out)[i] = v;
  else reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
}

// GTCRN with in / out sample rates other than 16 kHz (Export_GTCRN.py:636-693).  Input: down-sampling resamples the raw
// samples first and then applies the PCM scale and the DC removal (inside the model's prep kernel); up-sampling scales and
// removes the mean at the INPUT rate, then resamples (:638-654).  Output: down-sampling resamples before the x32767 PCM
// scale, up-sampling after it (:671-688).  The 2^-15 input scale commutes exactly with the linear interpolation.
adn_status gtcrn_run_resampled(adn_model* m, const void* d_in, void* d_out, int batch, cudaStream_t st) {
  if (batch > m->rs_cap) {
    ADN_CUDA_TRY(cudaDeviceSynchronize(), m->err);
    adn_note_free();
    if (m->rs_cap) { cudaFree(m->rs_xc); cudaFree(m->rs_xm); cudaFree(m->rs_ym); cudaFree(m->rs_yr); }
    m->rs_cap = 0;
    ADN_CUDA_TRY(cudaMalloc((void**)&m->rs_xc, (size_t)batch * m->io_L * 4), m->err);
    ADN_CUDA_TRY(cudaMalloc((void**)&m->rs_xm, (size_t)batch * m->L * 4), m->err);
    ADN_CUDA_TRY(cudaMalloc((void**)&m->rs_ym, (size_t)batch * m->Lout * 4), m->err);
    ADN_CUDA_TRY(cudaMalloc((void**)&m->rs_yr, (size_t)batch * m->io_Lout * 4), m->err);
    m->rs_cap = batch;
  }
  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };
  const void* run_in = d_in;
  if (m->rs_in) {
    const long long nm = (long long)batch * m->L;
    if (m->in_sr > 16000) {
      if (adn_resample_linear(d_in, m->in_dtype, m->rs_xm, batch, m->io_L, m->L, m->in_scale_factor, st) != ADN_OK) return ADN_ERR_CUDA;
      if (m->in_dtype == ADN_I16) rs_scale_kernel<<<blocks(nm), 256, 0, st>>>(m->rs_xm, nm, 1.0f / 32768.0f);
    } else {
      // cast, PCM scale and DC removal at the input rate (the prep kernel without padding), then the resampler
      gtcrn::launch_prep(d_in, m->in_dtype, m->rs_xc, nullptr, nullptr, batch, m->io_L, m->io_L, 0, 1, 0, st);
      if (adn_resample_linear(m->rs_xc, ADN_F32, m->rs_xm, batch, m->io_L, m->L, m->in_scale_factor, st) != ADN_OK) return ADN_ERR_CUDA;
    }
    run_in = m->rs_xm;
  }
  adn_status s = gtcrn_run(m, run_in, m->rs_out ? (void*)m->rs_ym : d_out, batch, st, nullptr);
  if (s != ADN_OK || !m->rs_out || (m->stop_after > 0 && m->last_launches >= m->stop_after)) return s;
  const long long no = (long long)batch * m->io_Lout, nmo = (long long)batch * m->Lout;
  const bool pcm = m->out_dtype == ADN_I16;
  if (m->out_sr > 16000 && pcm) rs_scale_kernel<<<blocks(nmo), 256, 0, st>>>(m->rs_ym, nmo, 32767.0f);
  if (adn_resample_linear(m->rs_ym, ADN_F32, m->rs_yr, batch, m->Lout, m->io_Lout, m->out_scale_factor, st) != ADN_OK) return ADN_ERR_CUDA;
  if (m->out_sr < 16000 && pcm) rs_scale_kernel<<<blocks(no), 256, 0, st>>>(m->rs_yr, no, 32767.0f);
  rs_convert_kernel<<<blocks(no), 256, 0, st>>>(m->rs_yr, d_out, m->out_dtype, no);
  ADN_CUDA_TRY(cudaGetLastError(), m->err);
  return ADN_OK;
}

}  // namespace

static std::atomic<unsigned long long> g_alloc_epoch{1};
void adn_note_free() { g_alloc_epoch.fetch_add(1, std::memory_order_relaxed); }
unsigned long long adn_alloc_epoch() { return g_alloc_epoch.load(std::memory_order_relaxed); }

// Every entry point runs on the handle's device and puts the caller's current device back (a process that drives several GPUs,
// e.g. torch with one model per device, must not find its current device changed by a library call).
struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ===================================================================================
extern "C" {

const char* adn_version(void) { return "adn 0.1 sm_100a"; }

const char* adn_last_error(const adn_model* m) {
  if (m) return m->err.c_str();
  std::lock_guard<std::mutex> lk(g_err_mu);
  return g_last_error.c_str();
}

adn_status adn_create(adn_model** out, const adn_desc* desc, const float* weights, size_t nfloats,
                      int device_id) {
  if (!out) return ADN_ERR_INVALID;
  *out = nullptr;
  if (!desc || !weights) {
    set_global_error("adn_create: null descriptor or weights");
    return ADN_ERR_INVALID;
  }
  adn_model* m = new adn_model();
  auto fail = [&](adn_status s) {
    set_global_error(m->err);
    adn_destroy(m);
    return s;
  };
  m->device = device_id;
  for (int i = 0; i < desc->n_kv; ++i) m->meta[desc->keys[i]] = desc->values[i];
  for (int i = 0; i < desc->n_tensors; ++i) {
    const auto& t = desc->tensors[i];
    if (t.count > nfloats || t.offset > nfloats - t.count) {        // (no uint64 wrap-around)
      m->err = std::string("tensor '") + t.name + "' exceeds the blob";
      return fail(ADN_ERR_INVALID);
    }
    m->index[t.name] = TensorRef{t.offset, t.count};
  }
  auto need = [&](const char* k, std::string& v) {
    auto it = m->meta.find(k);
    if (it == m->meta.end() || it->second.empty()) {
      m->err = std::string("Required metadata key ") + k + " is missing.";
      return false;
    }
    v = it->second;
    return true;
  };
  std::string fam, sL, sin, sout, snfft, shop;
  if (!need("model_family", fam) || !need("input_audio_length", sL) || !need("input_audio_dtype", sin) ||
      !need("output_audio_dtype", sout))
    return fail(ADN_ERR_INVALID);
  const bool is_ss = fam == "mossformer2_ss";       // learned encoder / decoder: no STFT keys (feature_kind conv_encoder_decoder)
  if (!is_ss && (!need("nfft", snfft) || !need("hop_length", shop))) return fail(ADN_ERR_INVALID);
  if (is_ss) { snfft = "16"; shop = "8"; }          // Conv1d(k16, s8) framing, snip-edges
  m->family = fam;
  if (fam != "gtcrn" && fam != "mel_band_roformer" && fam != "mossformer2_se" && fam != "mossformergan_se" && fam != "dfsmn" && fam != "ulunas" && fam != "zipenhancer" && fam != "h_gtcrn" && !is_ss) {
    m->err = "unsupported model_family '" + fam + "'";
    return fail(ADN_ERR_UNSUPPORTED);
  }
  if (!parse_dtype(sin, m->in_dtype) || !parse_dtype(sout, m->out_dtype)) {
    m->err = "input/output_audio_dtype must be F32, F16 or INT16";
    return fail(ADN_ERR_INVALID);
  }
  m->L = atoi(sL.c_str());
  int nfft = atoi(snfft.c_str()), hop = atoi(shop.c_str());
  const bool is_gtcrn = fam == "gtcrn";
  if (is_gtcrn && (nfft != gtcrn::NFFT || hop != gtcrn::HOP)) {
    m->err = "gtcrn requires nfft=512, hop_length=256";
    return fail(ADN_ERR_INVALID);
  }
  m->io_L = m->L;
  m->run_in_dtype = m->in_dtype;
  m->run_out_dtype = m->out_dtype;
  if (is_gtcrn) {                                  // optional linear resampling (Export_GTCRN.py:619-632)
    auto opt = [&](const char* k, int& v) { auto it = m->meta.find(k); if (it != m->meta.end() && !it->second.empty()) v = atoi(it->second.c_str()); };
    int model_sr = 16000;
    opt("in_sample_rate", m->in_sr); opt("out_sample_rate", m->out_sr); opt("model_sample_rate", model_sr);
    if (model_sr != 16000 || m->in_sr <= 0 || m->out_sr <= 0) {
      m->err = "gtcrn runs at model_sample_rate 16000";
      return fail(ADN_ERR_INVALID);
    }
    m->rs_in = m->in_sr != 16000;
    m->rs_out = m->out_sr != 16000;
    m->in_scale_factor = 1.0 / ((double)m->in_sr / 16000.0);       // model_rate_scale (:626)
    m->out_scale_factor = (double)m->out_sr / 16000.0;             // out_sample_rate_scale (:625)
    if (m->rs_in) {
      m->L = (int)floor((double)m->io_L * m->in_scale_factor);     // F.interpolate(scale_factor=...) output size
      m->run_in_dtype = ADN_F32;
      m->run_remove_dc = m->in_sr > 16000 ? 1 : 0;                 // up-sampling: the mean is removed BEFORE the resampler (:627-628)
    }
    if (m->rs_out) m->run_out_dtype = ADN_F32;
  }
  if (is_gtcrn && m->L < nfft) {                   // (the other families validate their own model-rate window)
    m->err = "input_audio_length must be >= nfft";
    return fail(ADN_ERR_INVALID);
  }
  m->stft.init(nfft, hop, 1, 1, 0);
  m->T = m->stft.n_frames(m->L);
  m->Lp = m->stft.padded_len(m->L);
  m->Lout = m->stft.out_len(m->T);
  if (is_ss) {
    m->T = (m->L - nfft) / hop + 1;
    m->Lout = (m->T - 1) * hop + nfft;
    m->n_out = 2;
  }
  m->io_Lout = m->rs_out ? (int)floor((double)m->Lout * m->out_scale_factor) : m->Lout;
  m->chans = fam == "mel_band_roformer" ? 2 : 1;
  m->out_chans = m->chans;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device_id) {
    m->err = "no usable CUDA device " + std::to_string(device_id) + " (libadn has no CPU fallback)";
    return fail(ADN_ERR_CUDA);
  }
  cudaDeviceProp prop;
  DeviceGuard dg(device_id);
  if (!dg.ok || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) {
    m->err = "cudaSetDevice failed";
    return fail(ADN_ERR_CUDA);
  }
  if (prop.major != 10) {
    m->err = "libadn is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
    return fail(ADN_ERR_CUDA);
  }
  if (cudaMalloc((void**)&m->d_blob, nfloats * sizeof(float)) != cudaSuccess ||
      cudaMemcpy(m->d_blob, weights, nfloats * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    m->err = "failed to upload the weight blob";
    return fail(ADN_ERR_CUDA);
  }
  m->nfloats = nfloats;
  m->sms = prop.multiProcessorCount;
  adn_status s = ADN_OK;
  if (is_gtcrn) {
    s = build_gtcrn(m, weights);
  } else {
    m->impl = fam == "mossformer2_se"
                  ? mf2se_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                  : is_ss ? mf2ss_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                  : fam == "mossformergan_se" ? mfgan_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                  : fam == "dfsmn" ? dfsmn_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                  : fam == "ulunas" ? ulunas_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                  : fam == "zipenhancer" ? zipenh_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                  : fam == "h_gtcrn" ? hgtcrn_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err)
                          : mbr_create(m->meta, m->index, weights, m->d_blob, device_id, m->sms, m->err);
    if (!m->impl) s = ADN_ERR_INVALID;
    else {                                          // host staging follows the family's own I/O description
      adn_tensor_info tin, touts[4];
      m->impl->io_info(&tin, touts);
      m->L = tin.length; m->chans = tin.channels; m->out_chans = touts[0].channels; m->Lout = touts[0].length; m->n_out = m->impl->n_outputs();
      m->io_L = m->L; m->io_Lout = m->Lout;
    }
  }
  if (s != ADN_OK) return fail(s);
  if (cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    m->err = "cudaStreamCreate failed";
    return fail(ADN_ERR_CUDA);
  }
  {
    const char* ge = getenv("ADN_GRAPHS");             // ADN_GRAPHS=0: never replay runs as CUDA graphs
    m->use_graphs = !(ge && ge[0] == '0');
  }
  // weight uploads, operand splits and table memsets ran on the legacy default stream; runs use non-blocking streams
  if (cudaDeviceSynchronize() != cudaSuccess) {
    m->err = "device initialisation failed";
    return fail(ADN_ERR_CUDA);
  }
  *out = m;
  return ADN_OK;
}

void adn_destroy(adn_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  for (auto& ge : m->graphs) cudaGraphExecDestroy(ge.exec);
  delete m->impl;
  free_workspace(m);
  if (m->d_blob) cudaFree(m->d_blob);
  if (m->rs_cap) { cudaFree(m->rs_xc); cudaFree(m->rs_xm); cudaFree(m->rs_ym); cudaFree(m->rs_yr); }
  if (m->d_ola) cudaFree(m->d_ola);
  if (m->d_wf_hl) cudaFree(m->d_wf_hl);
  if (m->d_wo_hl) cudaFree(m->d_wo_hl);
  for (auto e : m->events) cudaEventDestroy(e);
  if (m->io_cap) { cudaFree(m->d_in); cudaFree(m->d_out); }
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m->st_in) cudaStreamDestroy(m->st_in);
  if (m->st_out) cudaStreamDestroy(m->st_out);
  for (int i = 0; i < 4; ++i) {
    if (m->ev_h2d[i]) cudaEventDestroy(m->ev_h2d[i]);
    if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
  }
  delete m;
}

adn_status adn_io_info(const adn_model* m, adn_tensor_info* in, adn_tensor_info* outs, int32_t* n_out) {
  if (!m || !in || !outs || !n_out) return ADN_ERR_INVALID;
  if (m->impl) {
    m->impl->io_info(in, outs);
    *n_out = m->impl->n_outputs();
    return ADN_OK;
  }
  memset(in, 0, sizeof(*in));
  memset(outs, 0, sizeof(*outs));
  strncpy(in->name, "noisy_audio", sizeof(in->name) - 1);       // Export_GTCRN.py:768
  in->dtype = m->in_dtype; in->channels = 1; in->length = m->io_L;
  strncpy(outs->name, "denoised_audio", sizeof(outs->name) - 1); // Export_GTCRN.py:769
  outs->dtype = m->out_dtype; outs->channels = 1; outs->length = m->io_Lout;
  *n_out = 1;
  return ADN_OK;
}

size_t adn_workspace_bytes(const adn_model* m, int32_t batch) {
  if (!m || batch <= 0) return 0;
  if (m->impl) return m->impl->workspace_bytes(batch);
  return workspace_bytes_for(m, batch);
}

int32_t adn_launches_per_run(const adn_model* m, int32_t batch) {
  (void)batch;
  // prep, stft, enc_front, 6x(gt_main, tra_gru, tra_apply), 2x(dp_intra, dp_inter), ln_res, dec_tail, istft
  if (m && m->impl) return m->impl->launches(batch);
  return m ? 28 : 0;
}

adn_status adn_run(adn_model* m, const void* d_in, void* const* d_outs, int32_t batch, void* stream) {
  if (!m) return ADN_ERR_INVALID;
  if (!d_in || !d_outs || !d_outs[0] || batch <= 0) {
    m->err = "adn_run: null buffer or non-positive batch";
    return ADN_ERR_INVALID;
  }
  DeviceGuard dg(m->device);
  if (!dg.ok) { m->err = "cudaSetDevice failed"; return ADN_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  auto eager = [&]() -> adn_status {
    if (m->impl) {
      m->ev_used = 0;
      m->ev_stream = st;
      m->impl->tick = tick_cb;
      m->impl->tick_ctx = m;
      tick_cb(m, "start");
      adn_status r = m->impl->run_multi(d_in, d_outs, batch, st);
      if (r != ADN_OK) m->err = m->impl->err;
      m->last_batch = batch;
      return r;
    }
    if (!m->rs_in && !m->rs_out) return gtcrn_run(m, d_in, d_outs[0], batch, st, nullptr);
    return gtcrn_run_resampled(m, d_in, d_outs[0], batch, st);
  };
  // Launch-bound regime (batch 1: 28 ... 514 dependent launches per run): replay the run as one CUDA graph.  Not on the legacy
  // default stream (it cannot be captured), not while profiling or dumping stages, not inside a caller's own capture.
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool graphable = m->use_graphs && !m->profiling && m->stop_after == 0 && st != nullptr && st != cudaStreamLegacy &&
                         st != cudaStreamPerThread && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone;
  if (!graphable) return eager();
  const int n_out = m->impl ? m->impl->n_outputs() : 1;
  const unsigned long long ep = adn_alloc_epoch();
  if (!m->graphs.empty() && m->graphs[0].epoch != ep) {        // some workspace was re-allocated since: addresses are stale
    for (auto& ge : m->graphs) cudaGraphExecDestroy(ge.exec);
    m->graphs.clear();
  }
  for (auto& ge : m->graphs) {
    bool same = ge.in == d_in && ge.batch == batch && ge.st == st;
    for (int o = 0; same && o < n_out; ++o) same = ge.outs[o] == d_outs[o];
    if (!same) continue;
    ADN_CUDA_TRY(cudaGraphLaunch(ge.exec, st), m->err);
    ++m->graph_launches;
    m->last_batch = batch;
    return ADN_OK;
  }
  int& seen = m->eager_runs[batch];
  if (seen < 1) { ++seen; return eager(); }
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    m->use_graphs = false;
    return eager();
  }
  adn_status r = eager();
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &graph);
  cudaGraphExec_t exec = nullptr;
  if (r == ADN_OK && ce == cudaSuccess && graph && adn_alloc_epoch() == ep) ce = cudaGraphInstantiate(&exec, graph, 0);
  else if (ce == cudaSuccess) ce = cudaErrorUnknown;
  if (graph) cudaGraphDestroy(graph);
  if (r != ADN_OK) return r;
  if (ce != cudaSuccess || !exec) {             // (a prohibited call inside the run, or an allocation: stay eager from now on)
    cudaGetLastError();
    m->use_graphs = false;
    return eager();
  }
  if (m->graphs.size() >= 32) {
    cudaGraphExecDestroy(m->graphs.front().exec);
    m->graphs.erase(m->graphs.begin());
  }
  adn_model::GraphEntry ge{};
  ge.in = d_in; ge.batch = batch; ge.st = st; ge.exec = exec; ge.epoch = ep;
  for (int o = 0; o < n_out && o < 4; ++o) ge.outs[o] = d_outs[o];
  m->graphs.push_back(ge);
  ADN_CUDA_TRY(cudaGraphLaunch(exec, st), m->err);
  ++m->graph_launches;
  return ADN_OK;
}

adn_status adn_run_host(adn_model* m, const void* h_in, void* const* h_outs, int32_t batch) {
  if (!m) return ADN_ERR_INVALID;
  if (!h_in || !h_outs || !h_outs[0] || batch <= 0) {
    m->err = "adn_run_host: null buffer or non-positive batch";
    return ADN_ERR_INVALID;
  }
  DeviceGuard dg(m->device);
  if (!dg.ok) { m->err = "cudaSetDevice failed"; return ADN_ERR_CUDA; }
  adn_status s = ADN_OK;
  if (m->impl) {
    if (batch > m->io_cap) {     // staging buffers for families that own their workspace
      ADN_CUDA_TRY(cudaDeviceSynchronize(), m->err);
      adn_note_free();
      if (m->io_cap) { cudaFree(m->d_in); cudaFree(m->d_out); }
      ADN_CUDA_TRY(cudaMalloc(&m->d_in, (size_t)batch * m->chans * m->io_L * dtype_size(m->in_dtype)), m->err);
      ADN_CUDA_TRY(cudaMalloc(&m->d_out, (size_t)m->n_out * batch * m->out_chans * m->io_Lout * dtype_size(m->out_dtype)), m->err);
      m->io_cap = batch;
    }
  } else {
    s = ensure_capacity(m, batch);
    if (s != ADN_OK) return s;
  }
  const size_t in_row = (size_t)m->chans * m->io_L * dtype_size(m->in_dtype);
  const size_t out_row = (size_t)m->out_chans * m->io_Lout * dtype_size(m->out_dtype);
  if (!m->ev_h2d[0]) {
    ADN_CUDA_TRY(cudaStreamCreateWithFlags(&m->st_in, cudaStreamNonBlocking), m->err);
    ADN_CUDA_TRY(cudaStreamCreateWithFlags(&m->st_out, cudaStreamNonBlocking), m->err);
    for (int i = 0; i < 4; ++i) {
      ADN_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming), m->err);
      ADN_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming), m->err);
    }
  }
  cudaStream_t st = m->own_stream;
  if (!m->impl && batch >= 64 && m->stop_after == 0 && !m->rs_in && !m->rs_out) {
    // Two ranges: the second half's H2D overlaps the first half's conditioning + STFT, the first
    // half's D2H overlaps the second half's ISTFT.  The backbone runs once over the whole batch
    // (its GRU kernels are latency-bound, so slicing it would cost more than the copies).
    PhaseSync ps;
    ps.mid = (batch / 2) & ~1;
    const int lo[2] = {0, ps.mid}, hi[2] = {ps.mid, batch};
    for (int r = 0; r < 2; ++r) {
      ADN_CUDA_TRY(cudaMemcpyAsync((char*)m->d_in + lo[r] * in_row, (const char*)h_in + lo[r] * in_row,
                                   (hi[r] - lo[r]) * in_row, cudaMemcpyHostToDevice, m->st_in), m->err);
      ADN_CUDA_TRY(cudaEventRecord(m->ev_h2d[r], m->st_in), m->err);
      ps.h2d_done[r] = m->ev_h2d[r];
      ps.out_ready[r] = m->ev_done[r];
    }
    s = gtcrn_run(m, m->d_in, m->d_out, batch, st, &ps);
    if (s != ADN_OK) return s;
    for (int r = 0; r < 2; ++r) {
      ADN_CUDA_TRY(cudaStreamWaitEvent(m->st_out, m->ev_done[r], 0), m->err);
      ADN_CUDA_TRY(cudaMemcpyAsync((char*)h_outs[0] + lo[r] * out_row, (char*)m->d_out + lo[r] * out_row,
                                   (hi[r] - lo[r]) * out_row, cudaMemcpyDeviceToHost, m->st_out), m->err);
    }
    ADN_CUDA_TRY(cudaStreamSynchronize(m->st_out), m->err);
    ADN_CUDA_TRY(cudaStreamSynchronize(st), m->err);
    return ADN_OK;
  }
  ADN_CUDA_TRY(cudaMemcpyAsync(m->d_in, h_in, batch * in_row, cudaMemcpyHostToDevice, st), m->err);
  void* outs[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int o = 0; o < m->n_out; ++o) {
    if (!h_outs[o]) { m->err = "adn_run_host: null output buffer"; return ADN_ERR_INVALID; }
    outs[o] = (char*)m->d_out + (size_t)o * batch * out_row;
  }
  s = adn_run(m, m->d_in, outs, batch, st);
  if (s != ADN_OK) return s;
  for (int o = 0; o < m->n_out; ++o)
    ADN_CUDA_TRY(cudaMemcpyAsync(h_outs[o], outs[o], batch * out_row, cudaMemcpyDeviceToHost, st), m->err);
  ADN_CUDA_TRY(cudaStreamSynchronize(st), m->err);
  return ADN_OK;
}

adn_status adn_debug_stop_after(adn_model* m, int32_t n_launches) {
  if (!m) return ADN_ERR_INVALID;
  m->stop_after = n_launches;
  if (m->impl) m->impl->set_stop_after(n_launches);
  return ADN_OK;
}

adn_status adn_set_profiling(adn_model* m, int32_t enabled) {
  if (!m) return ADN_ERR_INVALID;
  m->profiling = enabled != 0;
  return ADN_OK;
}

adn_status adn_last_kernel_times(adn_model* m, const char** names, float* ms, int32_t cap, int32_t* n) {
  if (!m || !n) return ADN_ERR_INVALID;
  *n = 0;
  if (m->ev_used < 2) return ADN_OK;
  ADN_CUDA_TRY(cudaEventSynchronize(m->events[m->ev_used - 1]), m->err);
  for (size_t i = 1; i < m->ev_used && (int)(i - 1) < cap; ++i) {
    float t = 0.f;
    ADN_CUDA_TRY(cudaEventElapsedTime(&t, m->events[i - 1], m->events[i]), m->err);
    names[i - 1] = m->ev_names[i];
    ms[i - 1] = t;
    *n = (int)i;
  }
  return ADN_OK;
}

adn_status adn_debug_read(adn_model* m, const char* name, float* h_dst, size_t count, size_t* actual) {
  if (!m || !name) return ADN_ERR_INVALID;
  if (!strcmp(name, "graph_launches")) {          // runs of this handle that were replayed as a CUDA graph
    if (actual) *actual = 1;
    if (h_dst && count) h_dst[0] = (float)m->graph_launches;
    return ADN_OK;
  }
  if (m->impl) {
    adn_status r = m->impl->debug_read(name, h_dst, count, actual);
    if (r != ADN_OK) m->err = m->impl->err;
    return r;
  }
  using namespace gtcrn;
  const size_t B = m->last_batch, T = m->T;
  if (B == 0) {
    m->err = "adn_debug_read: no run yet";
    return ADN_ERR_INVALID;
  }
  std::map<std::string, std::pair<const float*, size_t>> tbl = {
      {"xp", {m->buf.xp, B * m->Lp}},
      {"spec", {m->buf.spec, B * T * SPEC_LD}},
      {"e0", {m->buf.e0, B * T * FRAME_E0}},
      {"e1", {m->buf.e[1], B * T * FRAME16}},
      {"e2", {m->buf.e[2], B * T * FRAME16}},
      {"e3", {m->buf.e[3], B * T * FRAME16}},
      {"e4", {m->buf.e[4], B * T * FRAME16}},
      {"h1", {m->buf.h1, B * T * 8 * E1_F}},
      {"zt", {m->buf.zt, B * T * 8}},
      {"at", {m->buf.at, B * T * 8}},
      {"tgi", {m->buf.tgi, B * T * 48}},
      {"gi", {m->buf.gi, B * T * 3 * FRAME16}},
      {"xa", {m->buf.xa, B * T * FRAME16}},
      {"xb", {m->buf.xb, B * T * FRAME16}},
      {"inter", {m->buf.inter, B * T * FRAME16}},
      {"enh", {m->buf.enh, B * (T + 2 * m->stft.pad_frames()) * SPEC_LD}},
  };
  auto it = tbl.find(name);
  if (it == tbl.end()) {
    m->err = std::string("adn_debug_read: unknown tensor '") + name + "'";
    return ADN_ERR_INVALID;
  }
  if (actual) *actual = it->second.second;
  if (!h_dst) return ADN_OK;
  size_t nc = count < it->second.second ? count : it->second.second;
  ADN_CUDA_TRY(cudaSetDevice(m->device), m->err);
  ADN_CUDA_TRY(cudaDeviceSynchronize(), m->err);
  ADN_CUDA_TRY(cudaMemcpy(h_dst, it->second.first, nc * sizeof(float), cudaMemcpyDeviceToHost), m->err);
  return ADN_OK;
}

}  // extern "C"

// ===================================================================================
// stand-alone STFT / ISTFT operators
// ===================================================================================
struct adn_stft {
  int device = 0;
  StftPlan p;
  int T = 0;            // frames the norm table was built for
  float* d_fwd = nullptr;
  float* d_ola = nullptr;
  float* d_norm = nullptr;
  float* d_xp = nullptr;
  size_t xp_cap = 0;
  float* d_fm = nullptr;   // frame-major padded spectrum for the inverse
  size_t fm_cap = 0;
  // opt-in tensor-core path (adn_stft_enable_tc; model families only -- the public operators stay on the exact fp32 GEMM):
  // windowed-DFT / overlap-add weights as zero-padded tf32 hi | lo planes, plans keyed by (buffers, rows)
  bool tc_on = false;
  int sms = 148;
  float* d_wf_hl = nullptr; int wf_npad = 0, wf_kpad = 0, wf_bn = 0;
  float* d_wo_hl = nullptr; int wo_npad = 0, wo_kpad = 0, wo_bn = 0;
  struct TcSlot { const float* in = nullptr; float* out = nullptr; int rows = 0, frames = 0; tc::TcPlan plan; tc::TcArgs args{}; bool valid = false; };
  TcSlot fwd_slot, inv_slot;
  std::vector<float> h_ola;   // kept for the lazily built planes
  std::vector<float> h_fwd;
};

namespace {

__global__ void pack_to_frame_major_kernel(const float* __restrict__ spec, float* __restrict__ fm, int rows2f,
                                           int T, int ld, int pad) {
  // spec (B, 2F, T) -> fm (B, T+2*pad, ld); 32x32 smem transpose tiles
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const float* s = spec + (long long)b * rows2f * T;
  float* o = fm + (long long)b * (T + 2 * pad) * ld;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows2f && t < T) ? s[(long long)r * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, r = r0 + threadIdx.x;
    if (t < T && r < rows2f) o[(long long)(t + pad) * ld + r] = tile[threadIdx.x][i];
  }
}

}  // namespace

extern "C" {

adn_status adn_stft_create(adn_stft** out, const adn_stft_geom* g, const float* fwd_basis,
                           const float* inv_basis, const float* win_norm, int32_t n_frames, int device_id) {
  if (!out) return ADN_ERR_INVALID;
  *out = nullptr;
  if (!g || !fwd_basis || !inv_basis || !win_norm || g->nfft <= 0 || g->hop <= 0 || n_frames <= 0) {
    set_global_error("adn_stft_create: invalid argument");
    return ADN_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device_id || cudaSetDevice(device_id) != cudaSuccess) {
    set_global_error("no usable CUDA device (libadn has no CPU fallback)");
    return ADN_ERR_CUDA;
  }
  adn_stft* s = new adn_stft();
  s->device = device_id;
  s->p.init(g->nfft, g->hop, g->center, g->pad_reflect, g->norm_multiply);
  s->T = n_frames;
  const size_t nb = (size_t)s->p.rows2f * g->nfft;
  std::vector<float> ola = build_ola_weight(inv_basis, g->nfft, g->hop, s->p.ld, s->p.R);
  s->h_ola = ola;
  s->h_fwd.assign(fwd_basis, fwd_basis + nb);
  const int lout = s->p.out_len(n_frames);
  bool ok = cudaMalloc((void**)&s->d_fwd, nb * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->d_ola, ola.size() * 4) == cudaSuccess &&
            cudaMalloc((void**)&s->d_norm, (size_t)lout * 4) == cudaSuccess &&
            cudaMemcpy(s->d_fwd, fwd_basis, nb * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(s->d_ola, ola.data(), ola.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(s->d_norm, win_norm, (size_t)lout * 4, cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) {
    set_global_error(std::string("adn_stft_create: ") + cudaGetErrorString(cudaGetLastError()));
    adn_stft_destroy(s);
    return ADN_ERR_CUDA;
  }
  *out = s;
  return ADN_OK;
}

void adn_stft_destroy(adn_stft* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  cudaFree(s->d_fwd); cudaFree(s->d_ola); cudaFree(s->d_norm); cudaFree(s->d_xp); cudaFree(s->d_fm);
  cudaFree(s->d_wf_hl); cudaFree(s->d_wo_hl);
  delete s;
}

adn_status adn_stft_forward(adn_stft* s, const float* d_x, float* d_spec, int32_t batch, int32_t length,
                            void* stream) {
  if (!s || !d_x || !d_spec || batch <= 0 || length < s->p.nfft) {
    set_global_error("adn_stft_forward: invalid argument");
    return ADN_ERR_INVALID;
  }
  std::string err;
  ADN_CUDA_TRY(cudaSetDevice(s->device), err);
  cudaStream_t st = (cudaStream_t)stream;
  const int Lp = s->p.padded_len(length), T = s->p.n_frames(length);
  size_t need = (size_t)batch * Lp;
  if (need > s->xp_cap) {
    cudaDeviceSynchronize();
    adn_note_free();
    cudaFree(s->d_xp);
    if (cudaMalloc((void**)&s->d_xp, need * 4) != cudaSuccess) {
      set_global_error("adn_stft_forward: out of memory");
      s->d_xp = nullptr; s->xp_cap = 0;
      return ADN_ERR_CUDA;
    }
    s->xp_cap = need;
  }
  gtcrn::launch_prep(d_x, ADN_F32, s->d_xp, nullptr, nullptr, batch, length, Lp, s->p.center ? s->p.half : 0, 0,
                     s->p.reflect, st);
  GemmArgs g;
  fill_stft_gemm(g, s->p, s->d_xp, Lp, s->d_fwd, batch, T, d_spec, (long long)s->p.rows2f * T, 1, T);
  launch_gemm_ffma(g, EPI_STORE, st);
  if (cudaGetLastError() != cudaSuccess) {
    set_global_error("adn_stft_forward: launch failed");
    return ADN_ERR_CUDA;
  }
  return ADN_OK;
}

}  // extern "C"

// Frame-major variants (internal, model_impl.h) for model families that keep GTCRN's (rows, T, ld) spectrum layout: the
// forward takes an already padded waveform (rows, Lp) and writes (rows, T, ld) [Re | Im | pad]; the inverse takes the
// zero-framed (rows, T + 2 * pad_frames, ld) enhanced spectrum as the overlap-add GEMM's A operand directly.
int adn_stft_ld(const adn_stft* s) { return s->p.ld; }
int adn_stft_pad_frames(const adn_stft* s) { return s->p.pad_frames(); }
int adn_stft_padded_len(const adn_stft* s, int length) { return s->p.padded_len(length); }
namespace {
__global__ void istft_norm_kernel(float* __restrict__ y, const float* __restrict__ norm, int out_len, long long n, int mul) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float nv = __ldg(norm + (int)(i % out_len));
  y[i] = mul ? y[i] * nv : y[i] / nv;
}
}  // namespace

// Tensor-core variants of the frame-major transforms (3xTF32 on the fp32-A GEMM of gemm_tc.cu: the padded waveform / the
// zero-framed spectrum are the A operand as they lie in memory, overlapping rows through the TMA row stride).  The inverse needs
// the overlap-added block rows to tile the output exactly (no centre pad, or a centre pad of whole hops) and `slack` floats of
// finite memory behind the spectrum buffer (the K window of the last rows runs past it; those weights are zero).
int adn_stft_enable_tc(adn_stft* s, int sms) {
  const StftPlan& p = s->p;
  if (p.hop % 4 || p.nfft % 32 || (p.center && p.half % p.hop)) return 0;
  std::string err;
  s->sms = sms;
  s->wf_bn = choose_bn(p.rows2f);
  if (s->wf_bn == 176) s->wf_bn = 128;             // TMA-store epilogue: N tiles are whole 32-column boxes (gemm_tc.cu launch_bn)
  s->wf_npad = (p.rows2f + s->wf_bn - 1) / s->wf_bn * s->wf_bn;
  s->wf_kpad = round_up(p.nfft, 32);
  s->wo_bn = choose_bn(p.hop);
  s->wo_npad = (p.hop + s->wo_bn - 1) / s->wo_bn * s->wo_bn;
  s->wo_kpad = round_up(p.R * p.ld, 32);
  std::vector<float> wf = split_pad_weight(s->h_fwd.data(), p.rows2f, p.nfft, s->wf_npad, s->wf_kpad);
  std::vector<float> wo = split_pad_weight(s->h_ola.data(), p.hop, p.R * p.ld, s->wo_npad, s->wo_kpad);
  if (cudaMalloc((void**)&s->d_wf_hl, wf.size() * 4) != cudaSuccess || cudaMalloc((void**)&s->d_wo_hl, wo.size() * 4) != cudaSuccess ||
      cudaMemcpy(s->d_wf_hl, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(s->d_wo_hl, wo.data(), wo.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
    return 0;
  s->tc_on = true;
  return s->wo_kpad;                               // the slack (floats) the inverse needs behind its input buffer
}

adn_status adn_stft_forward_fm(adn_stft* s, const float* d_xp, float* d_spec_fm, int rows, int n_frames, int Lp, cudaStream_t st) {
  const StftPlan& p = s->p;
  if (s->tc_on && Lp % 4 == 0 && p.ld % 4 == 0) {
    adn_stft::TcSlot& e = s->fwd_slot;
    if (!(e.valid && e.in == d_xp && e.out == d_spec_fm && e.rows == rows && e.frames == n_frames)) {
      std::string err;
      e.valid = false;
      e.plan = tc::TcPlan{};
      e.plan.bn = s->wf_bn; e.plan.a_f32 = true;
      const int bt = n_frames >= 128 ? 128 : n_frames;
      const size_t plane = (size_t)s->wf_npad * s->wf_kpad;
      if (tc::make_row_map(&e.plan.map_a_hi, d_xp, p.nfft, n_frames, p.hop, rows, Lp, bt, 1, err) &&
          tc::make_weight_map(&e.plan.map_w_hi, s->d_wf_hl, s->wf_kpad, s->wf_npad, s->wf_bn, err) &&
          tc::make_weight_map(&e.plan.map_w_lo, s->d_wf_hl + plane, s->wf_kpad, s->wf_npad, s->wf_bn, err) &&
          tc::make_store_map(&e.plan.map_c, d_spec_fm, p.rows2f, n_frames, p.ld, rows, (long long)n_frames * p.ld, err)) {
        e.plan.map_a_lo = e.plan.map_a_hi; e.plan.map_w2_hi = e.plan.map_w_hi; e.plan.map_w2_lo = e.plan.map_w_lo;
        tc::TcArgs& a = e.args;
        a = tc::TcArgs{};
        a.bb = 1; a.bt = bt; a.tiles_per_chunk = (n_frames + 127) / 128; a.t0 = 0;
        a.B = rows; a.TM = n_frames; a.N = p.rows2f; a.K = p.nfft;
        a.m_tiles = rows * a.tiles_per_chunk;
        a.C = d_spec_fm; a.ldc = p.ld;
        e.in = d_xp; e.out = d_spec_fm; e.rows = rows; e.frames = n_frames; e.valid = true;
      }
    }
    if (e.valid && tc::launch(e.plan, e.args, EPI_LIN, s->sms, st) == cudaSuccess) return ADN_OK;
    cudaGetLastError();
  }
  GemmArgs g;
  fill_stft_gemm(g, s->p, d_xp, Lp, s->d_fwd, rows, n_frames, d_spec_fm, (long long)n_frames * s->p.ld, s->p.ld, 1);
  launch_gemm_ffma(g, EPI_STORE, st);
  return cudaGetLastError() == cudaSuccess ? ADN_OK : ADN_ERR_CUDA;
}
adn_status adn_stft_inverse_fm(adn_stft* s, const float* d_fm_padded, float* d_y, int rows, int n_frames, cudaStream_t st) {
  if (n_frames != s->T) return ADN_ERR_INVALID;
  const StftPlan& p = s->p;
  const int lo = p.blk_lo(), hi = p.blk_hi(n_frames), TM = hi - lo + 1, out_len = p.out_len(n_frames);
  if (s->tc_on && (long long)TM * p.hop == out_len && (p.center ? p.half == lo * p.hop : lo == 0) && out_len % 4 == 0) {
    adn_stft::TcSlot& e = s->inv_slot;
    if (!(e.valid && e.in == d_fm_padded && e.out == d_y && e.rows == rows && e.frames == n_frames)) {
      std::string err;
      e.valid = false;
      e.plan = tc::TcPlan{};
      e.plan.bn = s->wo_bn; e.plan.a_f32 = true;
      const int frows = n_frames + 2 * p.pad_frames();
      const int bt = TM >= 128 ? 128 : TM;
      const size_t plane = (size_t)s->wo_npad * s->wo_kpad;
      if (tc::make_row_map(&e.plan.map_a_hi, d_fm_padded, s->wo_kpad, frows, p.ld, rows, (long long)frows * p.ld, bt, 1, err) &&
          tc::make_weight_map(&e.plan.map_w_hi, s->d_wo_hl, s->wo_kpad, s->wo_npad, s->wo_bn, err) &&
          tc::make_weight_map(&e.plan.map_w_lo, s->d_wo_hl + plane, s->wo_kpad, s->wo_npad, s->wo_bn, err) &&
          tc::make_store_map(&e.plan.map_c, d_y, p.hop, TM, p.hop, rows, out_len, err)) {
        e.plan.map_a_lo = e.plan.map_a_hi; e.plan.map_w2_hi = e.plan.map_w_hi; e.plan.map_w2_lo = e.plan.map_w_lo;
        tc::TcArgs& a = e.args;
        a = tc::TcArgs{};
        a.bb = 1; a.bt = bt; a.tiles_per_chunk = (TM + 127) / 128; a.t0 = lo;
        a.B = rows; a.TM = TM; a.N = p.hop; a.K = p.R * p.ld;
        a.m_tiles = rows * a.tiles_per_chunk;
        a.C = d_y; a.ldc = p.hop;
        e.in = d_fm_padded; e.out = d_y; e.rows = rows; e.frames = n_frames; e.valid = true;
      }
    }
    if (e.valid && tc::launch(e.plan, e.args, EPI_LIN, s->sms, st) == cudaSuccess) {
      const long long n = (long long)rows * out_len;
      istft_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_y, s->d_norm, out_len, n, p.norm_mul);
      return cudaGetLastError() == cudaSuccess ? ADN_OK : ADN_ERR_CUDA;
    }
    cudaGetLastError();
  }
  GemmArgs g;
  fill_istft_gemm(g, s->p, d_fm_padded, s->d_ola, s->d_norm, rows, n_frames, d_y, ADN_F32);
  launch_gemm_ffma(g, EPI_ISTFT, st);
  return cudaGetLastError() == cudaSuccess ? ADN_OK : ADN_ERR_CUDA;
}

extern "C" {

adn_status adn_stft_inverse(adn_stft* s, const float* d_spec, float* d_y, int32_t batch, int32_t n_frames,
                            void* stream) {
  if (!s || !d_spec || !d_y || batch <= 0 || n_frames != s->T) {
    set_global_error("adn_stft_inverse: invalid argument (n_frames must equal the plan's)");
    return ADN_ERR_INVALID;
  }
  std::string err;
  ADN_CUDA_TRY(cudaSetDevice(s->device), err);
  cudaStream_t st = (cudaStream_t)stream;
  const int T = n_frames, pad = s->p.pad_frames(), ld = s->p.ld;
  size_t need = (size_t)batch * (T + 2 * pad) * ld + (s->tc_on ? (size_t)s->wo_kpad : 0);   // + the K-window slack of the tensor-core path
  if (need > s->fm_cap) {
    cudaDeviceSynchronize();
    adn_note_free();
    cudaFree(s->d_fm);
    if (cudaMalloc((void**)&s->d_fm, need * 4) != cudaSuccess) {
      set_global_error("adn_stft_inverse: out of memory");
      s->d_fm = nullptr; s->fm_cap = 0;
      return ADN_ERR_CUDA;
    }
    s->fm_cap = need;
  }
  cudaMemsetAsync(s->d_fm, 0, need * 4, st);
  dim3 grid((T + 31) / 32, (s->p.rows2f + 31) / 32, batch), block(32, 8);
  pack_to_frame_major_kernel<<<grid, block, 0, st>>>(d_spec, s->d_fm, s->p.rows2f, T, ld, pad);
  if (adn_stft_inverse_fm(s, s->d_fm, d_y, batch, T, st) != ADN_OK) {
    set_global_error("adn_stft_inverse: launch failed");
    return ADN_ERR_CUDA;
  }
  return ADN_OK;
}

}  // extern "C"
