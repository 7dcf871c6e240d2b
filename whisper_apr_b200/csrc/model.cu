// Host side of libwhisper_b200.so: model handle, `.apr` upload, workspace, the encoder launch sequence and the C ABI.
//
// Mirrors, for the mel + encoder path only:
//   WhisperApr::{load_from_apr, compute_mel, encode}      src/lib.rs:673-754, 407-449
//   load_encoder_weights / load_*_weights                  src/lib.rs:757-841, 931-993
//   Encoder::{forward, forward_mel, forward_batch{,_padded}} src/model/encoder.rs:450-478, 566-660
//   EncoderBlock::forward                                  src/model/encoder.rs:346-361
//   transcribe_batch_optimized steps 1-2                   src/lib.rs:1162-1170
//   split_into_chunks / to_padded_tensor                   src/audio/batch.rs:219-240, 107-127
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "apr.h"
#include "wb_internal.h"

namespace wb {

std::atomic<long long> g_launch_count{0};
static thread_local std::string g_err;
int set_error(int status, const std::string& msg) {
  g_err = msg;
  return status;
}
const char* last_error() { return g_err.c_str(); }

namespace {

constexpr int N_SAMPLES_30S = 480000;   // lib.rs:408
constexpr int N_FRAMES_30S = 3000;      // lib.rs:409
constexpr int N_FFT = 400, HOP = 160, N_FREQ = 201;

typedef __nv_bfloat16 bf16;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }                 // every early return (WB_CUDA_OK) gives the memory back
  int ensure(size_t count) {
    if (count <= n) return WB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    WB_CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
    n = count;
    return WB_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct EventPair {                         // timing events that cannot leak on an early return
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  EventPair() {
    if (cudaEventCreate(&e0) != cudaSuccess) e0 = nullptr;
    if (cudaEventCreate(&e1) != cudaSuccess) e1 = nullptr;
  }
  EventPair(const EventPair&) = delete;
  EventPair& operator=(const EventPair&) = delete;
  ~EventPair() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

struct LayerW {
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float* sqkv = nullptr;                 // [3d] per-column scale (quantised files) or nullptr
  float so = 1.f, s1 = 1.f, s2 = 1.f;    // per-tensor scales
  // Int8 / Int4 `.apr` payloads stay PACKED in HBM (the file's own bytes: i8, or two's-complement nibbles, low nibble first);
  // the bf16 pointers above then alias the model's per-kind expansion buffers, refilled for every layer (see expand_layer).
  uint8_t *pqkv = nullptr, *po = nullptr, *p1 = nullptr, *p2 = nullptr;
};

struct Workspace {
  int cap = 0;
  DevBuf<float> audio, logmel, mel_f32, x, out_f32;
  DevBuf<int> n_valid, max_key;
  DevBuf<bf16> mel_bf16, c1, xn, qkv, att, hid, out_bf16;
};

}  // namespace
}  // namespace wb

struct wb_model {
  int device = 0;
  wb_config cfg{};
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int max_batch = 32;
  std::mutex mu;
  std::vector<void*> allocs;            // weight allocations (freed in wb_model_free)
  // mel
  wb::MelTables mel{};
  std::map<int, wb::MelTables> htk_tables;   // BatchPreprocessor filterbanks by n_mels (built on first use)
  // conv stem
  wb::bf16 *conv1_w = nullptr, *conv2_w = nullptr;
  float *conv1_b = nullptr, *conv2_b = nullptr;
  float conv1_s = 1.f, conv2_s = 1.f;
  float* pe = nullptr;
  std::vector<wb::LayerW> layers;
  int quant = 0;                        // 0: f32 payloads (bf16 weights resident); 2 / 3: int8 / int4 payloads resident, expanded per layer
  wb::bf16 *xp_qkv = nullptr, *xp_o = nullptr, *xp_1 = nullptr, *xp_2 = nullptr;   // expansion buffers (12 d^2 bf16: L2-sized)
  float *lnp_g = nullptr, *lnp_b = nullptr;
  wb::Workspace ws;
  // copy/compute overlap of the host-buffer entry point: two staging slots, copy-in and copy-out streams
  struct Slot {
    wb::DevBuf<float> audio;
    wb::DevBuf<int> n_valid;
    wb::DevBuf<uint8_t> out;
    int* h_n_valid = nullptr;          // pinned
    int h_cap = 0;
    cudaEvent_t in_done = nullptr, compute_done = nullptr, out_done = nullptr;
    bool busy = false;
  } slot[2];
  int next_slot = 0;
  cudaStream_t in_stream = nullptr, out_stream = nullptr;
  // CUDA graphs of the fused mel + encoder step, keyed by its device pointers and batch size: the step is 234 launches and as
  // many host-side tensor-map encodes; a replay is one cudaGraphLaunch.  First sighting of a key runs eagerly, the second is captured.
  struct StepGraph {
    const void* in = nullptr; const void* n_valid = nullptr; void* out = nullptr; int B = 0; int dtype = 0;
    int seen = 0; cudaGraphExec_t exec = nullptr; long long launches = 0;
  };
  std::vector<StepGraph> graphs;
  bool use_graphs = true;
  // per-kernel timing (wb_profile_*): CUDA events recorded on the launching stream around every launch
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;      // start/stop pairs
  std::vector<int> prof_cat;
};

namespace wb {
namespace {

struct DeviceGuard {
  int prev = 0;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (ok) cudaSetDevice(prev);
  }
};

enum ProfCat { PC_MEL_STFT = 0, PC_MEL_FINALIZE = 1, PC_GEMM = 2, PC_ATTENTION = 3, PC_LAYERNORM = 4, PC_OTHER = 5, PC_COUNT = 6 };
struct ProfScope {
  wb_model* m;
  cudaEvent_t stop = nullptr;
  ProfScope(wb_model* model, int cat) : m(model) {
    if (!m->prof_on) return;
    cudaEvent_t start;
    if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) { stop = nullptr; return; }
    cudaEventRecord(start, m->stream);
    m->prof_ev.push_back(start);
    m->prof_ev.push_back(stop);
    m->prof_cat.push_back(cat);
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, m->stream);
  }
};
#define WB_PROF(cat, expr)      \
  do {                          \
    ProfScope _ps(m, cat);      \
    rc = (expr);                \
  } while (0);                  \
  if (rc != WB_OK) return rc

template <typename T>
int dev_alloc(wb_model* m, size_t count, T** out) {
  void* p = nullptr;
  WB_CUDA_OK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
  m->allocs.push_back(p);
  *out = static_cast<T*>(p);
  return WB_OK;
}

// ---- `.apr` -> device ------------------------------------------------------------------------
struct Uploader {
  wb_model* m;
  const AprFile* f;
  DevBuf<uint8_t> staging;

  // Raw payload of `name` on the device (staging), clamped to `max_elems` elements.  *n_out = 0 if the tensor is
  // absent or runs past the end of the file (the reference keeps its default in both cases).
  int stage(const std::string& name, size_t max_elems, size_t* n_out, float* scale_out) {
    *n_out = 0;
    *scale_out = 1.f;
    const AprTensor* t = f->find(name);
    if (!t) return WB_OK;
    size_t nbytes = 0;
    const uint8_t* src = f->payload(*t, &nbytes);
    if (!src) return WB_OK;
    size_t n = std::min<size_t>(static_cast<size_t>(t->n_elements), max_elems);   // lib.rs:772-774 min-clamp
    if (n == 0) return WB_OK;
    size_t copy_bytes = f->cfg.quantization == 2 ? n : (f->cfg.quantization == 3 ? (n + 1) / 2 : n * 4);
    int rc = staging.ensure(copy_bytes + 16);
    if (rc != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpyAsync(staging.p, src, copy_bytes, cudaMemcpyHostToDevice, m->stream));
    *n_out = n;
    *scale_out = t->scale;
    return WB_OK;
  }
  // f32 parameter (bias / LayerNorm / positional embedding): dst pre-filled with the default.
  int load_f32(const std::string& name, float* dst, size_t count) {
    size_t n;
    float s;
    int rc = stage(name, count, &n, &s);
    if (rc != WB_OK || n == 0) return rc;
    switch (f->cfg.quantization) {
      case 2: rc = launch_i8_to_f32(reinterpret_cast<const int8_t*>(staging.p), s, dst, n, m->stream); break;
      case 3: rc = launch_i4_to_f32(staging.p, s, dst, n, m->stream); break;
      default: WB_CUDA_OK(cudaMemcpyAsync(dst, staging.p, n * 4, cudaMemcpyDeviceToDevice, m->stream)); break;
    }
    if (rc != WB_OK) return rc;
    WB_CUDA_OK(cudaStreamSynchronize(m->stream));      // staging is reused by the next tensor
    return WB_OK;
  }
  // Quantised GEMM weight: the payload bytes go to the device unchanged (dst pre-zeroed; absent / short tensors keep zeros).
  int load_packed(const std::string& name, uint8_t* dst, size_t count, float* scale_out) {
    *scale_out = 1.f;
    const AprTensor* t = f->find(name);
    if (!t) return WB_OK;
    size_t nbytes = 0;
    const uint8_t* src = f->payload(*t, &nbytes);
    if (!src) return WB_OK;
    const size_t n = std::min<size_t>(static_cast<size_t>(t->n_elements), count);
    if (n == 0) return WB_OK;
    const size_t copy_bytes = f->cfg.quantization == 2 ? n : n / 2;     // a trailing odd nibble (never for these shapes) is dropped
    WB_CUDA_OK(cudaMemcpyAsync(dst, src, copy_bytes, cudaMemcpyHostToDevice, m->stream));
    *scale_out = t->scale;
    return WB_OK;
  }
  // GEMM weight -> bf16 (quantised payloads keep their integer value; *scale_out carries the per-tensor scale).
  int load_bf16(const std::string& name, bf16* dst, size_t count, float* scale_out) {
    size_t n;
    float s;
    *scale_out = 1.f;
    int rc = stage(name, count, &n, &s);
    if (rc != WB_OK || n == 0) return rc;
    switch (f->cfg.quantization) {
      case 2: rc = launch_i8_to_bf16(reinterpret_cast<const int8_t*>(staging.p), dst, n, m->stream); *scale_out = s; break;
      case 3: rc = launch_i4_to_bf16(staging.p, dst, n, m->stream); *scale_out = s; break;
      default: rc = launch_f32_to_bf16(reinterpret_cast<const float*>(staging.p), dst, n, m->stream); break;
    }
    if (rc != WB_OK) return rc;
    WB_CUDA_OK(cudaStreamSynchronize(m->stream));
    return WB_OK;
  }
};

int fill_f32(wb_model* m, float* dst, size_t count, float value) {
  std::vector<float> h(count, value);
  WB_CUDA_OK(cudaMemcpy(dst, h.data(), count * 4, cudaMemcpyHostToDevice));
  return WB_OK;
}

int new_f32_param(wb_model* m, Uploader& up, const std::string& name, size_t count, float dflt, float** out) {
  int rc = dev_alloc(m, count, out);
  if (rc != WB_OK) return rc;
  rc = fill_f32(m, *out, count, dflt);
  if (rc != WB_OK) return rc;
  return up.load_f32(name, *out, count);
}

int new_weight(wb_model* m, Uploader& up, const std::string& name, size_t count, bf16** out, float* scale) {
  int rc = dev_alloc(m, count, out);
  if (rc != WB_OK) return rc;
  WB_CUDA_OK(cudaMemset(*out, 0, count * sizeof(bf16)));
  return up.load_bf16(name, *out, count, scale);
}

// filters [n_mels][201] -> device tables (dense rows + the first/last non-zero bin of every row + the periodic Hann window)
int upload_mel_tables(wb_model* m, const std::vector<float>& filt, int n_mels, MelTables* out) {
  std::vector<int> lo(n_mels, 0), len(n_mels, 0);
  for (int j = 0; j < n_mels; ++j) {
    int first = -1, last = -1;
    for (int k = 0; k < N_FREQ; ++k)
      if (filt[static_cast<size_t>(j) * N_FREQ + k] != 0.0f) {
        if (first < 0) first = k;
        last = k;
      }
    if (first >= 0) { lo[j] = first; len[j] = last - first + 1; }
  }
  const std::vector<float> win = hann_window_periodic(N_FFT);
  float *d_win, *d_filt;
  int *d_lo, *d_len;
  int rc;
  if ((rc = dev_alloc(m, win.size(), &d_win)) != WB_OK) return rc;
  if ((rc = dev_alloc(m, filt.size(), &d_filt)) != WB_OK) return rc;
  if ((rc = dev_alloc(m, lo.size(), &d_lo)) != WB_OK) return rc;
  if ((rc = dev_alloc(m, len.size(), &d_len)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpy(d_win, win.data(), win.size() * 4, cudaMemcpyHostToDevice));
  WB_CUDA_OK(cudaMemcpy(d_filt, filt.data(), filt.size() * 4, cudaMemcpyHostToDevice));
  WB_CUDA_OK(cudaMemcpy(d_lo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice));
  WB_CUDA_OK(cudaMemcpy(d_len, len.data(), len.size() * 4, cudaMemcpyHostToDevice));
  out->window = d_win;
  out->filters = d_filt;
  out->span_lo = d_lo;
  out->span_len = d_len;
  out->n_mels = n_mels;
  out->packed = nullptr;
  out->span_off = nullptr;
  out->nnz = 0;
  std::vector<int> off(n_mels, 0);
  std::vector<float> packed;
  for (int j = 0; j < n_mels; ++j) {                      // every span padded with zero weights to a multiple of 4 (16 B loads)
    off[j] = static_cast<int>(packed.size());
    for (int k = 0; k < len[j]; ++k) packed.push_back(filt[static_cast<size_t>(j) * N_FREQ + lo[j] + k]);
    while (packed.size() % 4 != 0) packed.push_back(0.0f);
  }
  if (n_mels <= 256 && packed.size() <= 2048 && !packed.empty()) {
    float* d_packed;
    int* d_off;
    if ((rc = dev_alloc(m, packed.size(), &d_packed)) != WB_OK) return rc;
    if ((rc = dev_alloc(m, off.size(), &d_off)) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpy(d_packed, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    WB_CUDA_OK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    out->packed = d_packed;
    out->span_off = d_off;
    out->nnz = static_cast<int>(packed.size());
  }
  return WB_OK;
}

int build_mel_tables(wb_model* m, const AprFile& f) {
  std::vector<float> filt;
  int n_mels = static_cast<int>(m->cfg.n_mels);
  if (f.has_filterbank && f.fb_freqs == N_FREQ && f.fb_mels > 0) {
    // lib.rs:738-741: the embedded (slaney) filterbank defines n_mels of the mel stage
    n_mels = static_cast<int>(f.fb_mels);
    filt.resize(static_cast<size_t>(n_mels) * N_FREQ);
    memcpy(filt.data(), f.fb_data, filt.size() * 4);
  } else {
    if (n_mels <= 0) return set_error(WB_ERR_FORMAT, "model has no mel filterbank and n_mels == 0");
    filt = htk_filterbank(n_mels, N_FFT, 16000);          // lib.rs:297-298 -> MelFilterbank::new
  }
  return upload_mel_tables(m, filt, n_mels, &m->mel);
}

int load_weights(wb_model* m, const AprFile& f) {
  const size_t d = m->cfg.n_audio_state, nm = m->cfg.n_mels, L = m->cfg.n_audio_layer, ctx = m->cfg.n_audio_ctx;
  Uploader up;
  up.m = m;
  up.f = &f;
  int rc = WB_OK;
  const bool quant = f.cfg.quantization == 2 || f.cfg.quantization == 3;
  auto done = [&](int r) { up.staging.release(); return r; };

  // conv stem: [out][in][3] -> bf16 -> [out][3][in]
  {
    bf16* tmp = nullptr;
    DevBuf<bf16> scratch;
    if ((rc = scratch.ensure(std::max(d * nm * 3, d * d * 3))) != WB_OK) return done(rc);
    tmp = scratch.p;
    WB_CUDA_OK(cudaMemset(tmp, 0, d * nm * 3 * sizeof(bf16)));
    if ((rc = up.load_bf16("encoder.conv1.weight", tmp, d * nm * 3, &m->conv1_s)) != WB_OK) return done(rc);
    if ((rc = dev_alloc(m, d * nm * 3, &m->conv1_w)) != WB_OK) return done(rc);
    if ((rc = launch_conv_repack(tmp, m->conv1_w, static_cast<int>(d), static_cast<int>(nm), m->stream)) != WB_OK) return done(rc);
    WB_CUDA_OK(cudaStreamSynchronize(m->stream));
    WB_CUDA_OK(cudaMemset(tmp, 0, d * d * 3 * sizeof(bf16)));
    if ((rc = up.load_bf16("encoder.conv2.weight", tmp, d * d * 3, &m->conv2_s)) != WB_OK) return done(rc);
    if ((rc = dev_alloc(m, d * d * 3, &m->conv2_w)) != WB_OK) return done(rc);
    if ((rc = launch_conv_repack(tmp, m->conv2_w, static_cast<int>(d), static_cast<int>(d), m->stream)) != WB_OK) return done(rc);
    WB_CUDA_OK(cudaStreamSynchronize(m->stream));
    scratch.release();
  }
  if ((rc = new_f32_param(m, up, "encoder.conv1.bias", d, 0.f, &m->conv1_b)) != WB_OK) return done(rc);
  if ((rc = new_f32_param(m, up, "encoder.conv2.bias", d, 0.f, &m->conv2_b)) != WB_OK) return done(rc);

  // positional embedding: embed_positions.weight, else positional_embedding, else the default table (lib.rs:793-800)
  {
    if ((rc = dev_alloc(m, ctx * d, &m->pe)) != WB_OK) return done(rc);
    const std::vector<float> pe = default_positional_embedding(static_cast<int>(ctx), static_cast<int>(d));
    WB_CUDA_OK(cudaMemcpy(m->pe, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice));
    const AprTensor* t = f.find("encoder.embed_positions.weight");
    size_t nb = 0;
    const char* name = (t && f.payload(*t, &nb)) ? "encoder.embed_positions.weight" : "encoder.positional_embedding";
    if ((rc = up.load_f32(name, m->pe, ctx * d)) != WB_OK) return done(rc);
  }

  m->quant = quant ? static_cast<int>(f.cfg.quantization) : 0;
  if (quant) {
    if ((rc = dev_alloc(m, 3 * d * d, &m->xp_qkv)) != WB_OK || (rc = dev_alloc(m, d * d, &m->xp_o)) != WB_OK ||
        (rc = dev_alloc(m, 4 * d * d, &m->xp_1)) != WB_OK || (rc = dev_alloc(m, 4 * d * d, &m->xp_2)) != WB_OK)
      return done(rc);
  }
  m->layers.resize(L);
  for (size_t i = 0; i < L; ++i) {
    LayerW& w = m->layers[i];
    const std::string p = "encoder.layers." + std::to_string(i);
    if ((rc = new_f32_param(m, up, p + ".self_attn_layer_norm.weight", d, 1.f, &w.ln1_g)) != WB_OK) return done(rc);
    if ((rc = new_f32_param(m, up, p + ".self_attn_layer_norm.bias", d, 0.f, &w.ln1_b)) != WB_OK) return done(rc);
    if ((rc = new_f32_param(m, up, p + ".final_layer_norm.weight", d, 1.f, &w.ln2_g)) != WB_OK) return done(rc);
    if ((rc = new_f32_param(m, up, p + ".final_layer_norm.bias", d, 0.f, &w.ln2_b)) != WB_OK) return done(rc);
    // fused QKV: rows [0,d) = q_proj, [d,2d) = k_proj, [2d,3d) = v_proj (three separate GEMMs in attention.rs:912-914)
    if (quant) {
      // dequant-in-prologue: upload the packed bytes as they are; every layer shares one set of bf16 expansion buffers
      const size_t qb = f.cfg.quantization == 2 ? d * d : d * d / 2;        // bytes of one d x d tensor (d is even)
      if ((rc = dev_alloc(m, 3 * qb, &w.pqkv)) != WB_OK || (rc = dev_alloc(m, qb, &w.po)) != WB_OK ||
          (rc = dev_alloc(m, 4 * qb, &w.p1)) != WB_OK || (rc = dev_alloc(m, 4 * qb, &w.p2)) != WB_OK)
        return done(rc);
      WB_CUDA_OK(cudaMemsetAsync(w.pqkv, 0, 3 * qb, m->stream));
      WB_CUDA_OK(cudaMemsetAsync(w.po, 0, qb, m->stream));
      WB_CUDA_OK(cudaMemsetAsync(w.p1, 0, 4 * qb, m->stream));
      WB_CUDA_OK(cudaMemsetAsync(w.p2, 0, 4 * qb, m->stream));
      w.wqkv = m->xp_qkv; w.wo = m->xp_o; w.w1 = m->xp_1; w.w2 = m->xp_2;
      if ((rc = dev_alloc(m, 3 * d, &w.bqkv)) != WB_OK) return done(rc);
      if ((rc = fill_f32(m, w.bqkv, 3 * d, 0.f)) != WB_OK) return done(rc);
      float sc[3] = {1.f, 1.f, 1.f};
      const char* proj[3] = {".self_attn.q_proj", ".self_attn.k_proj", ".self_attn.v_proj"};
      for (int k = 0; k < 3; ++k) {
        if ((rc = up.load_packed(p + proj[k] + ".weight", w.pqkv + k * qb, d * d, &sc[k])) != WB_OK) return done(rc);
        if ((rc = up.load_f32(p + proj[k] + ".bias", w.bqkv + k * d, d)) != WB_OK) return done(rc);
      }
      if ((rc = dev_alloc(m, 3 * d, &w.sqkv)) != WB_OK) return done(rc);
      std::vector<float> h(3 * d);
      for (int k = 0; k < 3; ++k) std::fill(h.begin() + k * d, h.begin() + (k + 1) * d, sc[k]);
      WB_CUDA_OK(cudaMemcpy(w.sqkv, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
      if ((rc = up.load_packed(p + ".self_attn.out_proj.weight", w.po, d * d, &w.so)) != WB_OK) return done(rc);
      if ((rc = new_f32_param(m, up, p + ".self_attn.out_proj.bias", d, 0.f, &w.bo)) != WB_OK) return done(rc);
      if ((rc = up.load_packed(p + ".fc1.weight", w.p1, 4 * d * d, &w.s1)) != WB_OK) return done(rc);
      if ((rc = new_f32_param(m, up, p + ".fc1.bias", 4 * d, 0.f, &w.b1)) != WB_OK) return done(rc);
      if ((rc = up.load_packed(p + ".fc2.weight", w.p2, 4 * d * d, &w.s2)) != WB_OK) return done(rc);
      if ((rc = new_f32_param(m, up, p + ".fc2.bias", d, 0.f, &w.b2)) != WB_OK) return done(rc);
      continue;
    }
    if ((rc = dev_alloc(m, 3 * d * d, &w.wqkv)) != WB_OK) return done(rc);
    WB_CUDA_OK(cudaMemset(w.wqkv, 0, 3 * d * d * sizeof(bf16)));
    if ((rc = dev_alloc(m, 3 * d, &w.bqkv)) != WB_OK) return done(rc);
    if ((rc = fill_f32(m, w.bqkv, 3 * d, 0.f)) != WB_OK) return done(rc);
    float sc[3] = {1.f, 1.f, 1.f};
    const char* proj[3] = {".self_attn.q_proj", ".self_attn.k_proj", ".self_attn.v_proj"};
    for (int k = 0; k < 3; ++k) {
      if ((rc = up.load_bf16(p + proj[k] + ".weight", w.wqkv + k * d * d, d * d, &sc[k])) != WB_OK) return done(rc);
      if ((rc = up.load_f32(p + proj[k] + ".bias", w.bqkv + k * d, d)) != WB_OK) return done(rc);
    }
    if (quant) {
      if ((rc = dev_alloc(m, 3 * d, &w.sqkv)) != WB_OK) return done(rc);
      std::vector<float> h(3 * d);
      for (int k = 0; k < 3; ++k) std::fill(h.begin() + k * d, h.begin() + (k + 1) * d, sc[k]);
      WB_CUDA_OK(cudaMemcpy(w.sqkv, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    if ((rc = new_weight(m, up, p + ".self_attn.out_proj.weight", d * d, &w.wo, &w.so)) != WB_OK) return done(rc);
    if ((rc = new_f32_param(m, up, p + ".self_attn.out_proj.bias", d, 0.f, &w.bo)) != WB_OK) return done(rc);
    if ((rc = new_weight(m, up, p + ".fc1.weight", 4 * d * d, &w.w1, &w.s1)) != WB_OK) return done(rc);
    if ((rc = new_f32_param(m, up, p + ".fc1.bias", 4 * d, 0.f, &w.b1)) != WB_OK) return done(rc);
    if ((rc = new_weight(m, up, p + ".fc2.weight", 4 * d * d, &w.w2, &w.s2)) != WB_OK) return done(rc);
    if ((rc = new_f32_param(m, up, p + ".fc2.bias", d, 0.f, &w.b2)) != WB_OK) return done(rc);
  }
  if ((rc = new_f32_param(m, up, "encoder.layer_norm.weight", d, 1.f, &m->lnp_g)) != WB_OK) return done(rc);
  if ((rc = new_f32_param(m, up, "encoder.layer_norm.bias", d, 0.f, &m->lnp_b)) != WB_OK) return done(rc);
  return done(WB_OK);
}

// ---- workspace ---------------------------------------------------------------------------------
int ensure_workspace(wb_model* m, int B) {
  Workspace& w = m->ws;
  if (B <= w.cap) return WB_OK;
  // a mel of T frames gives S = (T - 1) / 2 + 1 <= n_audio_ctx positions (validate_mel_len), so T <= 2 * ctx; the fused paths
  // always run T = 3000 / S = 1500 and are refused up front when the header's n_audio_ctx is smaller (check_fused_dims)
  const size_t d = m->cfg.n_audio_state, nm = std::max<size_t>(m->cfg.n_mels, m->mel.n_mels);
  const size_t S = std::max<size_t>(m->cfg.n_audio_ctx, (N_FRAMES_30S - 1) / 2 + 1);
  const size_t T = std::max<size_t>(N_FRAMES_30S, 2 * static_cast<size_t>(m->cfg.n_audio_ctx)), b = static_cast<size_t>(B);
  int rc;
  if ((rc = w.audio.ensure(b * N_SAMPLES_30S)) != WB_OK) return rc;
  if ((rc = w.n_valid.ensure(b)) != WB_OK) return rc;
  if ((rc = w.max_key.ensure(b)) != WB_OK) return rc;
  if ((rc = w.logmel.ensure(b * T * nm)) != WB_OK) return rc;
  if ((rc = w.mel_f32.ensure(b * T * nm)) != WB_OK) return rc;
  if ((rc = w.mel_bf16.ensure(b * (T + 2) * nm)) != WB_OK) return rc;
  if ((rc = w.c1.ensure(b * (T + 2) * d)) != WB_OK) return rc;
  if ((rc = w.x.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.xn.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.qkv.ensure(b * S * 3 * d)) != WB_OK) return rc;
  if ((rc = w.att.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.hid.ensure(b * S * 4 * d)) != WB_OK) return rc;
  if ((rc = w.out_f32.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.out_bf16.ensure(b * S * d)) != WB_OK) return rc;
  w.cap = B;
  for (auto& g : m->graphs)                 // workspace pointers are baked into captured launches
    if (g.exec) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  return WB_OK;
}

// ---- encoder launch sequence ------------------------------------------------------------------------
// Precondition: ws.mel_bf16 holds [B][T+2][n_mels] bf16 with rows 1..T = mel frames and zero guard rows.
// d_out: [B][S][d] f32 or bf16 (device).  n_layers < 0 -> all layers; ln_post as Encoder::forward (encoder.rs:477).
int encode_device(wb_model* m, int B, int T, void* d_out, wb_dtype out_dtype, int n_layers, bool ln_post) {
  const int d = static_cast<int>(m->cfg.n_audio_state), nm = static_cast<int>(m->cfg.n_mels);
  const int H = static_cast<int>(m->cfg.n_audio_head);
  const int S = (T - 1) / 2 + 1;               // conv2: (T + 2 - 3) / 2 + 1 (encoder.rs:79)
  Workspace& w = m->ws;
  cudaStream_t st = m->stream;
  int rc;
  const int L = n_layers < 0 ? static_cast<int>(m->layers.size()) : std::min<int>(n_layers, static_cast<int>(m->layers.size()));

  // c1 guard rows (0 and T+1) are zero: conv2's padding
  WB_PROF(PC_OTHER, launch_fill_bf16_rows(w.c1.p, static_cast<long long>(T + 2) * d, B, d, st));
  WB_PROF(PC_OTHER, launch_fill_bf16_rows(w.c1.p + static_cast<long long>(T + 1) * d, static_cast<long long>(T + 2) * d, B, d, st));

  GemmDesc g{};
  // conv1 + GELU: row t of the operand = padded frames t, t+1, t+2 (3*nm contiguous values)
  g.A = w.mel_bf16.p; g.a_row_stride = nm; g.a_batch_stride = static_cast<long long>(T + 2) * nm;
  g.rows_per_batch = T; g.n_batch = B;
  g.W = m->conv1_w; g.N = d; g.K = 3 * nm;
  g.epilogue = EPI_GELU_BF16; g.alpha = m->conv1_s; g.col_scale = nullptr; g.bias = m->conv1_b;
  g.out = w.c1.p; g.ldc = d; g.out_rows_per_batch = T + 2; g.out_row_off = 1; g.pe = nullptr;
  WB_PROF(PC_GEMM, launch_gemm(g, st));
  // conv2 (stride 2) + GELU + positional embedding -> fp32 residual stream
  g.A = w.c1.p; g.a_row_stride = 2LL * d; g.a_batch_stride = static_cast<long long>(T + 2) * d;
  g.rows_per_batch = S; g.n_batch = B;
  g.W = m->conv2_w; g.N = d; g.K = 3 * d;
  g.epilogue = EPI_GELU_PE_F32; g.alpha = m->conv2_s; g.bias = m->conv2_b;
  g.out = w.x.p; g.ldc = d; g.out_rows_per_batch = S; g.out_row_off = 0; g.pe = m->pe;
  WB_PROF(PC_GEMM, launch_gemm(g, st));

  const int M = B * S;
  auto flat = [&](const bf16* A, int K, const bf16* W, int N, int epi, float alpha, const float* cs, const float* bias, void* out) {
    GemmDesc q{};
    q.A = A; q.a_row_stride = K; q.a_batch_stride = static_cast<long long>(M) * K; q.rows_per_batch = M; q.n_batch = 1;
    q.W = W; q.N = N; q.K = K; q.epilogue = epi; q.alpha = alpha; q.col_scale = cs; q.bias = bias;
    q.out = out; q.ldc = N; q.out_rows_per_batch = M; q.out_row_off = 0; q.pe = nullptr;
    return launch_gemm(q, st);
  };
  // Quantised models: a layer's packed weights are expanded to bf16 (exact integer values; the scale stays in the GEMM epilogue)
  // right before the GEMM that consumes them, into buffers every layer reuses.  With M = B x 1500 rows per launch each weight tile
  // is consumed by ~190 row tiles, so expanding once per launch costs 1/190th of converting inside every CTA, and the 12 d^2
  // bf16 (39 MB for d = 1280) stay L2-resident for the GEMM that follows; HBM only ever holds the packed bytes.
  const size_t dd = static_cast<size_t>(d) * d;
  auto expand = [&](const uint8_t* packed, bf16* dst, size_t n) {
    return m->quant == 2 ? launch_i8_to_bf16(reinterpret_cast<const int8_t*>(packed), dst, n, st) : launch_i4_to_bf16(packed, dst, n, st);
  };
  for (int i = 0; i < L; ++i) {
    const LayerW& lw = m->layers[i];
    WB_PROF(PC_LAYERNORM, launch_layernorm(w.x.p, lw.ln1_g, lw.ln1_b, M, d, w.xn.p, nullptr, st));
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.pqkv, lw.wqkv, 3 * dd)); }
    WB_PROF(PC_GEMM, flat(w.xn.p, d, lw.wqkv, 3 * d, EPI_BF16, 1.f, lw.sqkv, lw.bqkv, w.qkv.p));
    WB_PROF(PC_ATTENTION, launch_attention(w.qkv.p, w.att.p, B, S, d, H, st));
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.po, lw.wo, dd)); }
    WB_PROF(PC_GEMM, flat(w.att.p, d, lw.wo, d, EPI_RESID_F32, lw.so, nullptr, lw.bo, w.x.p));
    WB_PROF(PC_LAYERNORM, launch_layernorm(w.x.p, lw.ln2_g, lw.ln2_b, M, d, w.xn.p, nullptr, st));
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.p1, lw.w1, 4 * dd)); }
    WB_PROF(PC_GEMM, flat(w.xn.p, d, lw.w1, 4 * d, EPI_GELU_BF16, lw.s1, nullptr, lw.b1, w.hid.p));
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.p2, lw.w2, 4 * dd)); }
    WB_PROF(PC_GEMM, flat(w.hid.p, 4 * d, lw.w2, d, EPI_RESID_F32, lw.s2, nullptr, lw.b2, w.x.p));
  }
  if (ln_post) {
    WB_PROF(PC_LAYERNORM, launch_layernorm(w.x.p, m->lnp_g, m->lnp_b, M, d, out_dtype == WB_BF16 ? static_cast<bf16*>(d_out) : nullptr,
                                           out_dtype == WB_BF16 ? nullptr : static_cast<float*>(d_out), st));
  } else {
    if (out_dtype == WB_BF16) rc = launch_f32_to_bf16(w.x.p, static_cast<bf16*>(d_out), static_cast<size_t>(M) * d, st);
    else {
      WB_CUDA_OK(cudaMemcpyAsync(d_out, w.x.p, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

// mel of B chunks already in ws.audio ([B][480000], n_valid per chunk in ws.n_valid) -> optional f32 [B][3000][m] and/or
// the bf16 padded operand in ws.mel_bf16.
int mel_device(wb_model* m, const float* d_audio, const int* d_n_valid, int B, float* d_mel_f32, bool want_bf16) {
  Workspace& w = m->ws;
  const int nm = m->mel.n_mels;
  const int n_frames = (N_SAMPLES_30S - N_FFT) / HOP + 1;     // 2998 (mel.rs:245-249)
  int rc;
  WB_PROF(PC_MEL_STFT, launch_mel_stft(d_audio, N_SAMPLES_30S, d_n_valid, N_SAMPLES_30S, HOP, n_frames, B, m->mel, w.logmel.p,
                                       w.max_key.p, m->stream));
  if (want_bf16) {
    const long long bs = static_cast<long long>(N_FRAMES_30S + 2) * nm;
    WB_PROF(PC_OTHER, launch_fill_bf16_rows(w.mel_bf16.p, bs, B, nm, m->stream));
    WB_PROF(PC_OTHER, launch_fill_bf16_rows(w.mel_bf16.p + static_cast<long long>(N_FRAMES_30S + 1) * nm, bs, B, nm, m->stream));
  }
  WB_PROF(PC_MEL_FINALIZE, launch_mel_finalize(w.logmel.p, w.max_key.p, n_frames, N_FRAMES_30S, nm, B, d_mel_f32,
                                               want_bf16 ? w.mel_bf16.p : nullptr, m->stream));
  return WB_OK;
}

int check_encoder_dims(const wb_model* m) {
  const wb_config& c = m->cfg;
  if (c.n_audio_state == 0 || c.n_audio_state % 128 != 0 || c.n_audio_head == 0 || c.n_audio_state != c.n_audio_head * 64)
    return set_error(WB_ERR_MODEL, "encoder kernels need n_audio_state % 128 == 0 and d_head == 64");
  if (c.n_mels == 0 || c.n_mels % 8 != 0) return set_error(WB_ERR_MODEL, "encoder kernels need n_mels % 8 == 0");
  if (static_cast<int>(c.n_mels) != m->mel.n_mels)
    return set_error(WB_ERR_MODEL, "filterbank n_mels differs from the model's n_mels");
  return WB_OK;
}

// The fused 30 s paths always produce 1500 positions: a header with a smaller n_audio_ctx gets the reference's own error
// (Encoder::forward, encoder.rs:456-461) instead of a workspace / positional-embedding overrun.
int check_fused_dims(const wb_model* m) {
  int rc = check_encoder_dims(m);
  if (rc != WB_OK) return rc;
  const size_t S = (N_FRAMES_30S - 1) / 2 + 1;
  if (S > m->cfg.n_audio_ctx)
    return set_error(WB_ERR_MODEL, "sequence length " + std::to_string(S) + " exceeds max " + std::to_string(m->cfg.n_audio_ctx));
  return WB_OK;
}

}  // namespace
}  // namespace wb

using namespace wb;

// =====================================================================================================
extern "C" {

const char* wb_version(void) { return "whisper_b200 0.1.0 (sm_100a)"; }
const char* wb_last_error(void) { return wb::last_error(); }

int wb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int wb_model_from_apr(const uint8_t* bytes, size_t n_bytes, int device, wb_model** out) {
  if (!out) return set_error(WB_ERR_MODEL, "null output handle");
  *out = nullptr;
  AprFile f;
  int rc = parse_apr(bytes, n_bytes, &f);
  if (rc != WB_OK) return rc;
  if (f.cfg.quantization == 1) return set_error(WB_ERR_FORMAT, "F16 .apr payloads have no reader (as in the reference)");
  // an untrusted header sizes every device allocation: refuse dimensions no Whisper variant comes near
  if (f.cfg.n_audio_state > 16384 || f.cfg.n_audio_layer > 512 || f.cfg.n_audio_ctx > 65536 || f.cfg.n_mels > 1024 ||
      f.cfg.n_audio_head > 256)
    return set_error(WB_ERR_FORMAT, "unreasonable model dimensions in the .apr header");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(WB_ERR_CUDA, "no CUDA device: libwhisper_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return set_error(WB_ERR_CUDA, "invalid CUDA device ordinal");
  int cc_major = 0;
  WB_CUDA_OK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  if (cc_major != 10) return set_error(WB_ERR_CUDA, "kernels are built for sm_100a only; device has compute capability " + std::to_string(cc_major) + ".x");
  DeviceGuard guard(device);
  wb_model* m = new wb_model();
  m->device = device;
  m->cfg = f.cfg;
  m->use_graphs = getenv("WB_NO_GRAPH") == nullptr;       // A/B switch: plain launches instead of graph replay
  auto fail = [&](int r) {
    std::string keep = wb::last_error();
    wb_model_free(m);
    return set_error(r, keep);
  };
  if (cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(set_error(WB_ERR_CUDA, "stream"));
  m->stream = m->own_stream;
  if ((rc = gemm_init()) != WB_OK) return fail(rc);
  if ((rc = attention_init()) != WB_OK) return fail(rc);
  if ((rc = mel_init()) != WB_OK) return fail(rc);
  if ((rc = build_mel_tables(m, f)) != WB_OK) return fail(rc);
  if (m->cfg.n_mels == 0) m->cfg.n_mels = m->mel.n_mels;
  if ((rc = load_weights(m, f)) != WB_OK) return fail(rc);
  if (cudaStreamSynchronize(m->stream) != cudaSuccess)
    return fail(set_error(WB_ERR_CUDA, std::string("model upload failed: ") + cudaGetErrorString(cudaGetLastError())));
  *out = m;
  return WB_OK;
}

int wb_model_config(const wb_model* m, wb_config* out) {
  if (!m || !out) return set_error(WB_ERR_MODEL, "null argument");
  *out = m->cfg;
  return WB_OK;
}

void wb_model_free(wb_model* m) {
  if (!m) return;
  DeviceGuard guard(m->device);
  cudaDeviceSynchronize();
  for (void* p : m->allocs) cudaFree(p);
  Workspace& w = m->ws;
  w.audio.release(); w.logmel.release(); w.mel_f32.release(); w.x.release(); w.out_f32.release();
  w.n_valid.release(); w.max_key.release();
  w.mel_bf16.release(); w.c1.release(); w.xn.release(); w.qkv.release(); w.att.release(); w.hid.release(); w.out_bf16.release();
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (auto& sl : m->slot) {
    sl.audio.release(); sl.n_valid.release(); sl.out.release();
    if (sl.h_n_valid) cudaFreeHost(sl.h_n_valid);
    if (sl.in_done) cudaEventDestroy(sl.in_done);
    if (sl.compute_done) cudaEventDestroy(sl.compute_done);
    if (sl.out_done) cudaEventDestroy(sl.out_done);
  }
  if (m->in_stream) cudaStreamDestroy(m->in_stream);
  if (m->out_stream) cudaStreamDestroy(m->out_stream);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
}

int wb_model_set_stream(wb_model* m, void* cuda_stream) {
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  std::lock_guard<std::mutex> lk(m->mu);
  m->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : m->own_stream;
  return WB_OK;
}

int wb_model_set_max_batch(wb_model* m, int max_chunks) {
  if (!m || max_chunks < 1) return set_error(WB_ERR_MODEL, "max batch must be >= 1");
  std::lock_guard<std::mutex> lk(m->mu);
  m->max_batch = max_chunks;
  return WB_OK;
}

int wb_sync(const wb_model* cm) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  for (auto& sl : m->slot) {
    if (sl.busy) {
      WB_CUDA_OK(cudaEventSynchronize(sl.out_done));
      sl.busy = false;
    }
  }
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));
  return WB_OK;
}

// ---------------------------------------------------------------------------------------------- mel
// MelFilterbank::compute for one segment with the given filter tables (host in, host out).  Caller holds the model lock.
static int mel_compute_one(wb_model* m, const MelTables& tab, const float* audio, size_t n, size_t hop, float* out, size_t out_capacity,
                           size_t* n_frames_out) {
  if (n_frames_out) *n_frames_out = 0;
  if (n == 0) return WB_OK;                                                        // mel.rs:236-238
  if (hop == 0) return set_error(WB_ERR_AUDIO, "hop_length must be positive");     // mel.rs:240-242
  const size_t n_frames = n >= N_FFT ? (n - N_FFT) / hop + 1 : 0;                  // mel.rs:245-249
  if (n_frames == 0) return WB_OK;
  const int nm = tab.n_mels;
  if (n > 0x7fff0000ull || hop > 0x7fff0000ull) return set_error(WB_ERR_AUDIO, "audio too long for one call");
  if (!audio || !out || out_capacity < n_frames * nm) return set_error(WB_ERR_AUDIO, "output buffer too small");
  DevBuf<float> d_audio, d_log, d_out;
  DevBuf<int> d_key;
  int rc;
  auto cleanup = [&](int r) { d_audio.release(); d_log.release(); d_out.release(); d_key.release(); return r; };
  if ((rc = d_audio.ensure((n + 3) & ~static_cast<size_t>(3))) != WB_OK) return cleanup(rc);
  if ((rc = d_log.ensure(n_frames * nm)) != WB_OK) return cleanup(rc);
  if ((rc = d_out.ensure(n_frames * nm)) != WB_OK) return cleanup(rc);
  if ((rc = d_key.ensure(1)) != WB_OK) return cleanup(rc);
  if (cudaMemcpyAsync(d_audio.p, audio, n * 4, cudaMemcpyHostToDevice, m->stream) != cudaSuccess)
    return cleanup(set_error(WB_ERR_CUDA, "H2D audio copy failed"));
  rc = launch_mel_stft(d_audio.p, static_cast<long long>(d_audio.n), nullptr, static_cast<int>(n), static_cast<int>(hop),
                       static_cast<int>(n_frames), 1, tab, d_log.p, d_key.p, m->stream);
  if (rc != WB_OK) return cleanup(rc);
  rc = launch_mel_finalize(d_log.p, d_key.p, static_cast<int>(n_frames), static_cast<int>(n_frames), nm, 1, d_out.p, nullptr, m->stream);
  if (rc != WB_OK) return cleanup(rc);
  if (cudaMemcpyAsync(out, d_out.p, n_frames * nm * 4, cudaMemcpyDeviceToHost, m->stream) != cudaSuccess ||
      cudaStreamSynchronize(m->stream) != cudaSuccess)
    return cleanup(set_error(WB_ERR_CUDA, std::string("mel kernels failed: ") + cudaGetErrorString(cudaGetLastError())));
  if (n_frames_out) *n_frames_out = n_frames;
  return cleanup(WB_OK);
}

int wb_mel_compute(const wb_model* cm, const float* audio, size_t n, size_t hop, float* out, size_t out_capacity,
                   size_t* n_frames_out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  return mel_compute_one(m, m->mel, audio, n, hop, out, out_capacity, n_frames_out);
}

// BatchPreprocessor::process_batch (src/audio/batch.rs:157-176): every segment goes through MelFilterbank::compute with the
// preprocessor's OWN filterbank -- MelFilterbank::new(n_mels, n_fft, sample_rate), the HTK triangles (batch.rs:143), not the model's --
// and is neither padded nor truncated.  Tables for a given n_mels are built once per model and kept.
int wb_batch_preprocess(const wb_model* cm, const float* const* audio, const size_t* n_samples, int B, size_t n_mels, size_t hop,
                        float* const* mels_out, const size_t* out_capacity, size_t* frame_counts, size_t* max_frames_out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  if (B < 0 || (B > 0 && (!audio || !n_samples || !mels_out || !out_capacity || !frame_counts)))
    return set_error(WB_ERR_AUDIO, "null argument");
  if (n_mels == 0 || n_mels > 1024) return set_error(WB_ERR_AUDIO, "n_mels out of range");
  if (max_frames_out) *max_frames_out = 0;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  auto it = m->htk_tables.find(static_cast<int>(n_mels));
  if (it == m->htk_tables.end()) {
    MelTables t{};
    int rc = upload_mel_tables(m, htk_filterbank(static_cast<int>(n_mels), N_FFT, 16000), static_cast<int>(n_mels), &t);
    if (rc != WB_OK) return rc;
    it = m->htk_tables.emplace(static_cast<int>(n_mels), t).first;
  }
  size_t max_frames = 0;
  for (int i = 0; i < B; ++i) {
    size_t nf = 0;
    int rc = mel_compute_one(m, it->second, audio[i], n_samples[i], hop, mels_out[i], out_capacity[i], &nf);
    if (rc != WB_OK) return rc;
    frame_counts[i] = nf;
    max_frames = std::max(max_frames, nf);
  }
  if (max_frames_out) *max_frames_out = max_frames;
  return WB_OK;
}

static int compute_mel_host(wb_model* m, const float* const* audio, const size_t* n_samples, const float* contiguous, int B,
                            float* out) {
  const int nm = m->mel.n_mels;
  const size_t per_out = static_cast<size_t>(N_FRAMES_30S) * nm;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    std::vector<int> nv(nb);
    for (int i = 0; i < nb; ++i) {
      const size_t n = contiguous ? N_SAMPLES_30S : std::min<size_t>(n_samples[b0 + i], N_SAMPLES_30S);   // lib.rs:413-425
      nv[i] = static_cast<int>(n);
      const float* src = contiguous ? contiguous + static_cast<size_t>(b0 + i) * N_SAMPLES_30S : audio[b0 + i];
      if (n) WB_CUDA_OK(cudaMemcpyAsync(m->ws.audio.p + static_cast<size_t>(i) * N_SAMPLES_30S, src, n * 4, cudaMemcpyHostToDevice, m->stream));
    }
    WB_CUDA_OK(cudaMemcpyAsync(m->ws.n_valid.p, nv.data(), nb * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    if ((rc = mel_device(m, m->ws.audio.p, m->ws.n_valid.p, nb, m->ws.mel_f32.p, false)) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpyAsync(out + static_cast<size_t>(b0) * per_out, m->ws.mel_f32.p, nb * per_out * 4, cudaMemcpyDeviceToHost, m->stream));
    WB_CUDA_OK(cudaStreamSynchronize(m->stream));     // nv goes out of scope; out is host-visible
  }
  return WB_OK;
}

int wb_compute_mel(const wb_model* cm, const float* audio, size_t n, float* out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || !out || (n && !audio)) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const float* ptrs[1] = {audio};
  const size_t lens[1] = {n};
  return compute_mel_host(m, ptrs, lens, nullptr, 1, out);
}

int wb_compute_mel_batch(const wb_model* cm, const float* audio, int B, float* out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || !out || !audio || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  return compute_mel_host(m, nullptr, nullptr, audio, B, out);
}

int wb_compute_mel_batch_dev(const wb_model* cm, const float* d_audio, int B, float* d_mel_out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || !d_audio || !d_mel_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const size_t per_out = static_cast<size_t>(N_FRAMES_30S) * m->mel.n_mels;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    rc = mel_device(m, d_audio + static_cast<size_t>(b0) * N_SAMPLES_30S, nullptr, nb, d_mel_out + static_cast<size_t>(b0) * per_out, false);
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

// ------------------------------------------------------------------------------------------ encoder
static int encode_same_len(wb_model* m, const float* const* mels, const float* d_mel, int B, int T, void* out_host, void* out_dev,
                           size_t out_stride_elems, wb_dtype dt, int n_layers, bool ln_post) {
  // B mels of T frames each (host pointers `mels` or one device array `d_mel` [B][T][nm]) -> [B][S][d]
  const int nm = static_cast<int>(m->cfg.n_mels), d = static_cast<int>(m->cfg.n_audio_state);
  const int S = (T - 1) / 2 + 1;
  const size_t esz = dt == WB_BF16 ? 2 : 4;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    const float* src_dev;
    if (d_mel) {
      src_dev = d_mel + static_cast<size_t>(b0) * T * nm;
    } else {
      for (int i = 0; i < nb; ++i)
        WB_CUDA_OK(cudaMemcpyAsync(m->ws.mel_f32.p + static_cast<size_t>(i) * T * nm, mels[b0 + i], static_cast<size_t>(T) * nm * 4,
                                   cudaMemcpyHostToDevice, m->stream));
      src_dev = m->ws.mel_f32.p;
    }
    if ((rc = launch_mel_pad_bf16(src_dev, m->ws.mel_bf16.p, nb, T, nm, m->stream)) != WB_OK) return rc;
    void* dst_dev;
    if (out_dev) dst_dev = static_cast<uint8_t*>(out_dev) + static_cast<size_t>(b0) * out_stride_elems * esz;
    else dst_dev = dt == WB_BF16 ? static_cast<void*>(m->ws.out_bf16.p) : static_cast<void*>(m->ws.out_f32.p);
    if ((rc = encode_device(m, nb, T, dst_dev, dt, n_layers, ln_post)) != WB_OK) return rc;
    if (out_host) {
      const size_t row_bytes = static_cast<size_t>(S) * d * esz;
      WB_CUDA_OK(cudaMemcpy2DAsync(static_cast<uint8_t*>(out_host) + static_cast<size_t>(b0) * out_stride_elems * esz,
                                   out_stride_elems * esz, dst_dev, row_bytes, row_bytes, nb, cudaMemcpyDeviceToHost, m->stream));
      WB_CUDA_OK(cudaStreamSynchronize(m->stream));
    }
  }
  return WB_OK;
}

static int validate_mel_len(const wb_model* m, size_t mel_len, int* T_out) {
  const size_t nm = m->cfg.n_mels;
  if (mel_len % nm != 0)                                                             // encoder.rs:568-574
    return set_error(WB_ERR_MODEL, "mel size " + std::to_string(mel_len) + " not divisible by n_mels " + std::to_string(nm));
  const size_t T = mel_len / nm;
  const size_t S = T == 0 ? 0 : (T - 1) / 2 + 1;
  if (S > m->cfg.n_audio_ctx)                                                        // encoder.rs:456-461
    return set_error(WB_ERR_MODEL, "sequence length " + std::to_string(S) + " exceeds max " + std::to_string(m->cfg.n_audio_ctx));
  *T_out = static_cast<int>(T);
  return WB_OK;
}

int wb_encode(const wb_model* cm, const float* mel, size_t mel_len, float* out, size_t out_capacity, size_t* seq_len_out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  int rc = check_encoder_dims(m);
  if (rc != WB_OK) return rc;
  int T = 0;
  if ((rc = validate_mel_len(m, mel_len, &T)) != WB_OK) return rc;
  const size_t S = T == 0 ? 0 : (T - 1) / 2 + 1;
  if (seq_len_out) *seq_len_out = S;
  if (T == 0) return WB_OK;
  if (!mel || !out || out_capacity < S * m->cfg.n_audio_state) return set_error(WB_ERR_MODEL, "output buffer too small");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const float* ptrs[1] = {mel};
  return encode_same_len(m, ptrs, nullptr, 1, T, out, nullptr, S * m->cfg.n_audio_state, WB_F32, -1, true);
}

int wb_encode_batch(const wb_model* cm, const float* const* mels, const size_t* mel_lens, int B, float* out, size_t out_capacity,
                    size_t* seq_lens, size_t* max_seq_out) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || B < 0) return set_error(WB_ERR_MODEL, "null model");
  int rc = check_encoder_dims(m);
  if (rc != WB_OK) return rc;
  const size_t d = m->cfg.n_audio_state;
  std::vector<int> Ts(B);
  size_t max_seq = 0;
  for (int i = 0; i < B; ++i) {
    if ((rc = validate_mel_len(m, mel_lens[i], &Ts[i])) != WB_OK) return rc;
    const size_t S = Ts[i] == 0 ? 0 : (Ts[i] - 1) / 2 + 1;
    if (seq_lens) seq_lens[i] = S;
    max_seq = std::max(max_seq, S);
  }
  if (max_seq_out) *max_seq_out = max_seq;
  if (B == 0 || max_seq == 0) return WB_OK;
  if (!out || out_capacity < static_cast<size_t>(B) * max_seq * d) return set_error(WB_ERR_MODEL, "output buffer too small");
  memset(out, 0, static_cast<size_t>(B) * max_seq * d * sizeof(float));            // zero padding (encoder.rs:641-643)
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  // runs of equal length share one device pass
  int i = 0;
  while (i < B) {
    int j = i + 1;
    while (j < B && Ts[j] == Ts[i]) ++j;
    if (Ts[i] > 0) {
      rc = encode_same_len(m, mels + i, nullptr, j - i, Ts[i], out + static_cast<size_t>(i) * max_seq * d, nullptr, max_seq * d, WB_F32, -1, true);
      if (rc != WB_OK) return rc;
    }
    i = j;
  }
  return WB_OK;
}

int wb_encode_batch_dev(const wb_model* cm, const float* d_mel, int B, void* d_out, wb_dtype out_dtype) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || !d_mel || !d_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const size_t S = (N_FRAMES_30S - 1) / 2 + 1;
  return encode_same_len(m, nullptr, d_mel, B, N_FRAMES_30S, nullptr, d_out, S * m->cfg.n_audio_state, out_dtype, -1, true);
}

// ------------------------------------------------------------------------------------- fused hot path
// One fused mel + encoder step on the model's stream, replayed from a CUDA graph once its (pointers, batch) key has been seen twice.
static int mel_encode_step(wb_model* m, const float* d_audio, const int* d_n_valid, int nb, void* d_out, wb_dtype out_dtype) {
  auto eager = [&]() -> int {
    int rc = mel_device(m, d_audio, d_n_valid, nb, nullptr, true);
    if (rc != WB_OK) return rc;
    return encode_device(m, nb, N_FRAMES_30S, d_out, out_dtype, -1, true);
  };
  if (!m->use_graphs || m->prof_on) return eager();
  wb_model::StepGraph* g = nullptr;
  for (auto& e : m->graphs)
    if (e.in == d_audio && e.n_valid == d_n_valid && e.out == d_out && e.B == nb && e.dtype == static_cast<int>(out_dtype)) g = &e;
  if (!g) {
    if (m->graphs.size() >= 16) {                         // callers that never repeat their pointers do not accumulate graphs
      for (auto& e : m->graphs)
        if (e.exec) cudaGraphExecDestroy(e.exec);
      m->graphs.clear();
    }
    wb_model::StepGraph e;
    e.in = d_audio; e.n_valid = d_n_valid; e.out = d_out; e.B = nb; e.dtype = static_cast<int>(out_dtype); e.seen = 1;
    m->graphs.push_back(e);
    return eager();
  }
  if (g->exec) {
    WB_CUDA_OK(cudaGraphLaunch(g->exec, m->stream));
    count_launch(static_cast<int>(g->launches));
    return WB_OK;
  }
  // second sighting: capture the launch sequence (thread-local mode: other threads' CUDA calls are unaffected)
  const long long before = g_launch_count.load();
  WB_CUDA_OK(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
  int rc = eager();
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
  if (rc != WB_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    m->use_graphs = false;                                // fall back to plain launches for this model
    if (rc != WB_OK) return rc;
    return eager();
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    m->use_graphs = false;
    return eager();
  }
  g->exec = exec;
  g->launches = g_launch_count.load() - before;
  WB_CUDA_OK(cudaGraphLaunch(exec, m->stream));
  return WB_OK;
}

int wb_mel_encode_batch_dev(const wb_model* cm, const float* d_audio, int B, void* d_out, wb_dtype out_dtype) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || !d_audio || !d_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const size_t d = m->cfg.n_audio_state, S = (N_FRAMES_30S - 1) / 2 + 1, esz = out_dtype == WB_BF16 ? 2 : 4;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
    if ((rc = mel_encode_step(m, d_audio + static_cast<size_t>(b0) * N_SAMPLES_30S, nullptr, nb,
                              static_cast<uint8_t*>(d_out) + static_cast<size_t>(b0) * S * d * esz, out_dtype)) != WB_OK)
      return rc;
  }
  return WB_OK;
}

static int prepare_slot(wb_model* m, wb_model::Slot& sl, int nb, size_t out_bytes) {
  int rc;
  if (!m->in_stream) {
    WB_CUDA_OK(cudaStreamCreateWithFlags(&m->in_stream, cudaStreamNonBlocking));
    WB_CUDA_OK(cudaStreamCreateWithFlags(&m->out_stream, cudaStreamNonBlocking));
  }
  if (!sl.in_done) {
    WB_CUDA_OK(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
    WB_CUDA_OK(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
    WB_CUDA_OK(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
  }
  if (sl.busy) {                                         // the batch that used this slot two calls ago must have left it
    WB_CUDA_OK(cudaEventSynchronize(sl.out_done));
    sl.busy = false;
  }
  if ((rc = sl.audio.ensure(static_cast<size_t>(nb) * N_SAMPLES_30S)) != WB_OK) return rc;
  if ((rc = sl.n_valid.ensure(nb)) != WB_OK) return rc;
  if ((rc = sl.out.ensure(out_bytes)) != WB_OK) return rc;
  if (sl.h_cap < nb) {
    if (sl.h_n_valid) cudaFreeHost(sl.h_n_valid);
    sl.h_n_valid = nullptr;
    WB_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&sl.h_n_valid), static_cast<size_t>(nb) * sizeof(int), cudaHostAllocDefault));
    sl.h_cap = nb;
  }
  return WB_OK;
}

// transcribe_batch_optimized steps 1-2 from host buffers.  Per micro-batch: copy-in stream (audio H2D) -> the model's stream (mel +
// encoder) -> copy-out stream (states D2H), chained by events over two staging slots.
static int mel_encode_batch_enqueue(wb_model* m, const float* const* audio, const size_t* n_samples, int B, void* out, wb_dtype out_dtype) {
  int rc;
  const size_t d = m->cfg.n_audio_state, S = (N_FRAMES_30S - 1) / 2 + 1, esz = out_dtype == WB_BF16 ? 2 : 4;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    const size_t out_bytes = static_cast<size_t>(nb) * S * d * esz;
    wb_model::Slot& sl = m->slot[m->next_slot];
    m->next_slot ^= 1;
    if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
    if ((rc = prepare_slot(m, sl, nb, out_bytes)) != WB_OK) return rc;
    for (int i = 0; i < nb; ++i) {
      const size_t n = std::min<size_t>(n_samples[b0 + i], N_SAMPLES_30S);
      sl.h_n_valid[i] = static_cast<int>(n);
      if (n) WB_CUDA_OK(cudaMemcpyAsync(sl.audio.p + static_cast<size_t>(i) * N_SAMPLES_30S, audio[b0 + i], n * 4, cudaMemcpyHostToDevice, m->in_stream));
    }
    WB_CUDA_OK(cudaMemcpyAsync(sl.n_valid.p, sl.h_n_valid, nb * sizeof(int), cudaMemcpyHostToDevice, m->in_stream));
    WB_CUDA_OK(cudaEventRecord(sl.in_done, m->in_stream));
    WB_CUDA_OK(cudaStreamWaitEvent(m->stream, sl.in_done, 0));
    if ((rc = mel_encode_step(m, sl.audio.p, sl.n_valid.p, nb, sl.out.p, out_dtype)) != WB_OK) return rc;
    WB_CUDA_OK(cudaEventRecord(sl.compute_done, m->stream));
    WB_CUDA_OK(cudaStreamWaitEvent(m->out_stream, sl.compute_done, 0));
    WB_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(out) + static_cast<size_t>(b0) * S * d * esz, sl.out.p, out_bytes, cudaMemcpyDeviceToHost, m->out_stream));
    WB_CUDA_OK(cudaEventRecord(sl.out_done, m->out_stream));
    sl.busy = true;
  }
  return WB_OK;
}

int wb_mel_encode_batch_async(const wb_model* cm, const float* const* audio, const size_t* n_samples, int B, void* out, wb_dtype out_dtype) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m || !audio || !n_samples || !out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  return mel_encode_batch_enqueue(m, audio, n_samples, B, out, out_dtype);
}

int wb_mel_encode_batch(const wb_model* cm, const float* const* audio, const size_t* n_samples, int B, void* out, wb_dtype out_dtype) {
  int rc = wb_mel_encode_batch_async(cm, audio, n_samples, B, out, out_dtype);
  if (rc != WB_OK) return rc;
  return wb_sync(cm);
}

// ------------------------------------------------------------------------------------------ chunking
size_t wb_split_into_chunks(size_t n_samples, size_t chunk_size, size_t overlap, size_t* starts, size_t* lens, size_t capacity) {
  if (n_samples == 0 || chunk_size == 0) return 0;                       // batch.rs:220-222
  size_t step = chunk_size > overlap ? chunk_size - overlap : 0;         // saturating_sub
  if (step < 1) step = 1;
  size_t count = 0, start = 0;
  while (start < n_samples) {
    const size_t end = std::min(start + chunk_size, n_samples);
    if (count < capacity) {
      if (starts) starts[count] = start;
      if (lens) lens[count] = end - start;
    }
    ++count;
    start += step;
    if (end >= n_samples) break;
  }
  return count;
}

int wb_to_padded_tensor(const float* const* mels, const size_t* frame_counts, int B, size_t n_mels, size_t max_frames, float* out) {
  if (B < 0 || !out) return set_error(WB_ERR_AUDIO, "null argument");
  memset(out, 0, static_cast<size_t>(B) * n_mels * max_frames * sizeof(float));
  for (int b = 0; b < B; ++b)
    for (size_t f = 0; f < frame_counts[b] && f < max_frames; ++f)
      for (size_t j = 0; j < n_mels; ++j) out[(static_cast<size_t>(b) * n_mels + j) * max_frames + f] = mels[b][f * n_mels + j];
  return WB_OK;
}

// ---------------------------------------------------------------------------------------- test hooks
static int debug_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return set_error(WB_ERR_CUDA, "no CUDA device: libwhisper_b200 has no CPU fallback");
  }
  WB_CUDA_OK(cudaSetDevice(device));
  return WB_OK;
}

int wb_debug_gemm(int device, const float* A, const float* W, const float* bias, const float* resid_or_pe, int M, int N, int K,
                  int epilogue, float alpha, float* out) {
  int rc = debug_device(device);
  if (rc != WB_OK) return rc;
  DevBuf<float> fa, fw, fb, fo, fpe;
  DevBuf<bf16> ba, bw, bo;
  auto cleanup = [&](int r) { fa.release(); fw.release(); fb.release(); fo.release(); fpe.release(); ba.release(); bw.release(); bo.release(); return r; };
  const size_t na = static_cast<size_t>(M) * K, nw = static_cast<size_t>(N) * K, no = static_cast<size_t>(M) * N;
  if ((rc = fa.ensure(na)) || (rc = fw.ensure(nw)) || (rc = fo.ensure(no)) || (rc = ba.ensure(na)) || (rc = bw.ensure(nw)) ||
      (rc = bo.ensure(no)) || (rc = fb.ensure(N)) || (rc = fpe.ensure(no)))
    return cleanup(rc);
  cudaMemcpy(fa.p, A, na * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(fw.p, W, nw * 4, cudaMemcpyHostToDevice);
  if (bias) cudaMemcpy(fb.p, bias, static_cast<size_t>(N) * 4, cudaMemcpyHostToDevice);
  launch_f32_to_bf16(fa.p, ba.p, na, nullptr);
  launch_f32_to_bf16(fw.p, bw.p, nw, nullptr);
  GemmDesc g{};
  g.A = ba.p; g.a_row_stride = K; g.a_batch_stride = static_cast<long long>(M) * K; g.rows_per_batch = M; g.n_batch = 1;
  g.W = bw.p; g.N = N; g.K = K; g.epilogue = epilogue; g.alpha = alpha; g.col_scale = nullptr; g.bias = bias ? fb.p : nullptr;
  g.ldc = N; g.out_rows_per_batch = M; g.out_row_off = 0; g.pe = nullptr;
  const bool bf_out = epilogue == EPI_BF16 || epilogue == EPI_GELU_BF16;
  if (epilogue == EPI_RESID_F32) {
    if (!resid_or_pe) return cleanup(set_error(WB_ERR_MODEL, "residual input required"));
    cudaMemcpy(fo.p, resid_or_pe, no * 4, cudaMemcpyHostToDevice);
  } else if (epilogue == EPI_GELU_PE_F32) {
    if (!resid_or_pe) return cleanup(set_error(WB_ERR_MODEL, "pe input required"));
    cudaMemcpy(fpe.p, resid_or_pe, no * 4, cudaMemcpyHostToDevice);
    g.pe = fpe.p;
  }
  g.out = bf_out ? static_cast<void*>(bo.p) : static_cast<void*>(fo.p);
  if ((rc = launch_gemm(g, nullptr)) != WB_OK) return cleanup(rc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(set_error(WB_ERR_CUDA, std::string("gemm kernel failed: ") + cudaGetErrorString(e)));
  if (bf_out) {
    std::vector<bf16> h(no);
    cudaMemcpy(h.data(), bo.p, no * 2, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < no; ++i) out[i] = __bfloat162float(h[i]);
  } else {
    cudaMemcpy(out, fo.p, no * 4, cudaMemcpyDeviceToHost);
  }
  return cleanup(WB_OK);
}

int wb_debug_attention(int device, const float* qkv, int B, int S, int d, int n_heads, float* out) {
  int rc = debug_device(device);
  if (rc != WB_OK) return rc;
  DevBuf<float> f;
  DevBuf<bf16> bq, bo;
  auto cleanup = [&](int r) { f.release(); bq.release(); bo.release(); return r; };
  const size_t nq = static_cast<size_t>(B) * S * 3 * d, no = static_cast<size_t>(B) * S * d;
  if ((rc = f.ensure(nq)) || (rc = bq.ensure(nq)) || (rc = bo.ensure(no))) return cleanup(rc);
  cudaMemcpy(f.p, qkv, nq * 4, cudaMemcpyHostToDevice);
  launch_f32_to_bf16(f.p, bq.p, nq, nullptr);
  if ((rc = launch_attention(bq.p, bo.p, B, S, d, n_heads, nullptr)) != WB_OK) return cleanup(rc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(set_error(WB_ERR_CUDA, std::string("attention kernel failed: ") + cudaGetErrorString(e)));
  std::vector<bf16> h(no);
  cudaMemcpy(h.data(), bo.p, no * 2, cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < no; ++i) out[i] = __bfloat162float(h[i]);
  return cleanup(WB_OK);
}

int wb_debug_gemm_bench(int device, int n_batch, int rows, int N, int K, int epilogue, int iters, float* ms_per_launch) {
  int rc = debug_device(device);
  if (rc != WB_OK) return rc;
  if (!ms_per_launch || iters <= 0) return set_error(WB_ERR_MODEL, "bad argument");
  DevBuf<float> f, fb, fo;
  DevBuf<bf16> ba, bw, bo;
  auto cleanup = [&](int r) { f.release(); fb.release(); fo.release(); ba.release(); bw.release(); bo.release(); return r; };
  const size_t M = static_cast<size_t>(n_batch) * rows, na = M * K, nw = static_cast<size_t>(N) * K, no = M * N;
  const size_t nf = std::max(static_cast<size_t>(rows) * K, nw);
  const bool bf_out = epilogue == EPI_BF16 || epilogue == EPI_GELU_BF16;
  if ((rc = f.ensure(nf)) || (rc = fb.ensure(N)) || (rc = ba.ensure(na)) || (rc = bw.ensure(nw)) ||
      (rc = bf_out ? bo.ensure(no) : fo.ensure(no)))
    return cleanup(rc);
  std::vector<float> h(nf);
  uint32_t x = 777u;
  for (size_t i = 0; i < nf; ++i) {
    x = x * 1664525u + 1013904223u;
    h[i] = (static_cast<float>(x >> 8) / 8388608.0f - 1.0f) * 0.05f;
  }
  cudaMemcpy(f.p, h.data(), nf * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(fb.p, h.data(), static_cast<size_t>(N) * 4, cudaMemcpyHostToDevice);
  launch_f32_to_bf16(f.p, bw.p, nw, nullptr);
  launch_f32_to_bf16(f.p, ba.p, static_cast<size_t>(rows) * K, nullptr);
  for (int b = 1; b < n_batch; ++b)
    cudaMemcpyAsync(ba.p + static_cast<size_t>(b) * rows * K, ba.p, static_cast<size_t>(rows) * K * 2, cudaMemcpyDeviceToDevice, nullptr);
  if (!bf_out) cudaMemsetAsync(fo.p, 0, no * 4, nullptr);
  GemmDesc g{};
  g.A = ba.p; g.a_row_stride = K; g.a_batch_stride = static_cast<long long>(rows) * K; g.rows_per_batch = rows; g.n_batch = n_batch;
  g.W = bw.p; g.N = N; g.K = K; g.epilogue = epilogue; g.alpha = 1.0f; g.col_scale = nullptr; g.bias = fb.p;
  g.ldc = N; g.out_rows_per_batch = rows; g.out_row_off = 0; g.pe = nullptr;
  g.out = bf_out ? static_cast<void*>(bo.p) : static_cast<void*>(fo.p);
  if (epilogue == EPI_GELU_PE_F32) return cleanup(set_error(WB_ERR_MODEL, "pe epilogue not benchmarked"));
  EventPair ev;
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  if (!e0 || !e1) return cleanup(set_error(WB_ERR_CUDA, "cudaEventCreate failed"));
  for (int i = 0; i < 2; ++i)
    if ((rc = launch_gemm(g, nullptr)) != WB_OK) return cleanup(rc);
  cudaEventRecord(e0, nullptr);
  for (int i = 0; i < iters; ++i)
    if ((rc = launch_gemm(g, nullptr)) != WB_OK) return cleanup(rc);
  cudaEventRecord(e1, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(set_error(WB_ERR_CUDA, std::string("gemm kernel failed: ") + cudaGetErrorString(e)));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_launch = ms / iters;
  return cleanup(WB_OK);
}

int wb_debug_attention_bench(int device, int B, int S, int d, int n_heads, int iters, float* ms_per_launch) {
  int rc = debug_device(device);
  if (rc != WB_OK) return rc;
  if (!ms_per_launch || iters <= 0) return set_error(WB_ERR_MODEL, "bad argument");
  DevBuf<float> f;
  DevBuf<bf16> bq, bo;
  auto cleanup = [&](int r) { f.release(); bq.release(); bo.release(); return r; };
  const size_t per = static_cast<size_t>(S) * 3 * d, nq = per * B, no = static_cast<size_t>(B) * S * d;
  if ((rc = f.ensure(per)) || (rc = bq.ensure(nq)) || (rc = bo.ensure(no))) return cleanup(rc);
  std::vector<float> h(per);
  uint32_t x = 12345u;
  for (size_t i = 0; i < per; ++i) {                      // LCG noise in [-2, 2): scores with a realistic spread
    x = x * 1664525u + 1013904223u;
    h[i] = (static_cast<float>(x >> 8) / 8388608.0f - 1.0f) * 2.0f;
  }
  cudaMemcpy(f.p, h.data(), per * 4, cudaMemcpyHostToDevice);
  launch_f32_to_bf16(f.p, bq.p, per, nullptr);
  for (int b = 1; b < B; ++b) cudaMemcpyAsync(bq.p + b * per, bq.p, per * 2, cudaMemcpyDeviceToDevice, nullptr);
  EventPair ev;
  cudaEvent_t e0 = ev.e0, e1 = ev.e1;
  if (!e0 || !e1) return cleanup(set_error(WB_ERR_CUDA, "cudaEventCreate failed"));
  for (int i = 0; i < 2; ++i)
    if ((rc = launch_attention(bq.p, bo.p, B, S, d, n_heads, nullptr)) != WB_OK) return cleanup(rc);
  cudaEventRecord(e0, nullptr);
  for (int i = 0; i < iters; ++i)
    if ((rc = launch_attention(bq.p, bo.p, B, S, d, n_heads, nullptr)) != WB_OK) return cleanup(rc);
  cudaEventRecord(e1, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(set_error(WB_ERR_CUDA, std::string("attention kernel failed: ") + cudaGetErrorString(e)));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_launch = ms / iters;
  return cleanup(WB_OK);
}

long long wb_launch_count(void) { return wb::g_launch_count.load(); }

long long wb_debug_apr_tensor_bytes(const uint8_t* bytes, size_t n_bytes, const char* name) {
  AprFile f;
  if (!name || parse_apr(bytes, n_bytes, &f) != WB_OK) return -2;
  const AprTensor* t = f.find(name);
  if (!t) return -1;
  size_t nb = 0;
  return f.payload(*t, &nb) ? static_cast<long long>(nb) : -1;      // -1: "tensor data out of bounds" (format/mod.rs:610-628)
}

int wb_profile_enable(wb_model* m, int on) {
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  std::lock_guard<std::mutex> lk(m->mu);
  m->prof_on = on != 0;
  return WB_OK;
}

int wb_profile_read(wb_model* m, float* ms_by_cat, int* launches_by_cat, int n_cat) {
  if (!m || !ms_by_cat || !launches_by_cat) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));
  for (int i = 0; i < n_cat; ++i) { ms_by_cat[i] = 0.f; launches_by_cat[i] = 0; }
  for (size_t i = 0; i < m->prof_cat.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]);
    const int c = m->prof_cat[i];
    if (c < n_cat) { ms_by_cat[c] += ms; launches_by_cat[c] += 1; }
  }
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  m->prof_ev.clear();
  m->prof_cat.clear();
  return WB_OK;
}

int wb_debug_encode(const wb_model* cm, const float* mel, size_t mel_len, int n_layers, int ln_post, float* out, size_t out_capacity) {
  wb_model* m = const_cast<wb_model*>(cm);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  int rc = check_encoder_dims(m);
  if (rc != WB_OK) return rc;
  int T = 0;
  if ((rc = validate_mel_len(m, mel_len, &T)) != WB_OK) return rc;
  if (T == 0) return WB_OK;
  const size_t S = (T - 1) / 2 + 1;
  if (!mel || !out || out_capacity < S * m->cfg.n_audio_state) return set_error(WB_ERR_MODEL, "output buffer too small");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const float* ptrs[1] = {mel};
  return encode_same_len(m, ptrs, nullptr, 1, T, out, nullptr, S * m->cfg.n_audio_state, WB_F32, n_layers, ln_post != 0);
}

int wb_debug_layernorm(int device, const float* x, const float* gamma, const float* beta, int rows, int d, float* out) {
  int rc = debug_device(device);
  if (rc != WB_OK) return rc;
  DevBuf<float> fx, fg, fb, fo;
  auto cleanup = [&](int r) { fx.release(); fg.release(); fb.release(); fo.release(); return r; };
  const size_t n = static_cast<size_t>(rows) * d;
  if ((rc = fx.ensure(n)) || (rc = fg.ensure(d)) || (rc = fb.ensure(d)) || (rc = fo.ensure(n))) return cleanup(rc);
  cudaMemcpy(fx.p, x, n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(fg.p, gamma, static_cast<size_t>(d) * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(fb.p, beta, static_cast<size_t>(d) * 4, cudaMemcpyHostToDevice);
  if ((rc = launch_layernorm(fx.p, fg.p, fb.p, rows, d, nullptr, fo.p, nullptr)) != WB_OK) return cleanup(rc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(set_error(WB_ERR_CUDA, std::string("layernorm kernel failed: ") + cudaGetErrorString(e)));
  cudaMemcpy(out, fo.p, n * 4, cudaMemcpyDeviceToHost);
  return cleanup(WB_OK);
}

}  // extern "C"
