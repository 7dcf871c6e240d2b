// Host-side data model of libwhisper_b200.so, shared by loader.cu, pipeline.cu, decoder.cu, audio_pre.cu, api.cu and debug.cu.
//
//   wb_model   the handle the C ABI hands out: a SET of devices (parallel::configure_thread_pool's successor, src/parallel.rs:34-60)
//   Replica    one device's copy of the weights (replicated) with its own streams, workspace, staging slots and graphs
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "apr.h"
#include "wb_internal.h"

namespace wb {

constexpr int N_SAMPLES_30S = 480000;   // lib.rs:408
constexpr int N_FRAMES_30S = 3000;      // lib.rs:409
constexpr int N_POS_30S = 1500;         // conv2, stride 2 (encoder.rs:79)
constexpr int N_FFT = 400, HOP = 160, N_FREQ = 201;

// internal output format of the encoder: 16-bit states in the OPERAND format (they feed the decoder's cross-attention K/V GEMM)
constexpr wb_dtype WB_OP16 = static_cast<wb_dtype>(2);
inline size_t dtype_size(wb_dtype dt) { return dt == WB_F32 ? 4 : 2; }

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

// One encoder block's device tensors.
struct LayerW {
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  op16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float* sqkv = nullptr;                 // [3d] per-column scale of the fused QKV GEMM (quantised files) or nullptr
  float so = 1.f, s1 = 1.f, s2 = 1.f;    // per-tensor scales (`.apr` scale table) of the other three GEMMs
  // per-channel int8 (model/quantized.rs:1769-1813): one scale per output row of W = per output column of the GEMM
  float *cso = nullptr, *cs1 = nullptr, *cs2 = nullptr;
  // Int8 / Int4 payloads stay PACKED in HBM (the file's own bytes: i8, or two's-complement nibbles, low nibble first);
  // the op16 pointers above then alias the replica's per-kind expansion buffers, refilled for every layer.
  uint8_t *pqkv = nullptr, *po = nullptr, *p1 = nullptr, *p2 = nullptr;
};

// One decoder block (src/model/decoder.rs DecoderBlock: ln1 / self_attn / ln2 / cross_attn / ln3 / ffn), f32 for the per-token path,
// op16 [2d][d] (k_proj rows then v_proj rows) for the cross-attention K/V precompute GEMM.
struct DecLayerW {
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr, *ln3_g = nullptr, *ln3_b = nullptr;
  float *sa_wqkv = nullptr, *sa_bqkv = nullptr;     // [3d][d], [3d]  (q | k | v rows)
  float *sa_wo = nullptr, *sa_bo = nullptr;
  float *ca_wq = nullptr, *ca_bq = nullptr, *ca_wo = nullptr, *ca_bo = nullptr;
  op16* ca_wkv = nullptr;                            // [2d][d]
  float* ca_bkv = nullptr;                           // [2d]
  float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
};

struct DecoderW {
  bool loaded = false;
  int d = 0, n_heads = 0, n_layers = 0, n_vocab = 0, n_ctx = 0;
  float* tok_emb = nullptr;       // [n_vocab][d]
  float* pos_emb = nullptr;       // [n_ctx][d]
  float *ln_g = nullptr, *ln_b = nullptr;
  std::vector<DecLayerW> layers;
  uint8_t* suppress[2] = {nullptr, nullptr};   // [n_vocab] 1 = suppressed (WhisperTokenSuppressor); [0] timestamps suppressed, [1] not
};

struct Workspace {
  int cap = 0;
  DevBuf<float> audio, logmel, mel_f32, x, out_f32;
  DevBuf<int> n_valid, max_key;
  DevBuf<unsigned int> mel_done;        // [B] zeroed: mel_finalize's last-block protocol
  DevBuf<unsigned int> ln_ready;        // [rows / 32] zeroed: residual GEMM -> layernorm_follow hand-over counters (re-armed by the follower)
  DevBuf<op16> mel_bf16, c1, xn, qkv, att, hid, out_bf16;
  int guard_T = -1;                     // T for which c1's zero guard rows (conv2's padding) are in place
};

enum ProfCat { PC_MEL_STFT = 0, PC_MEL_FINALIZE = 1, PC_GEMM = 2, PC_ATTENTION = 3, PC_LAYERNORM = 4, PC_OTHER = 5, PC_COUNT = 6 };

struct Replica {
  int device = 0;
  wb_config cfg{};
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // the LayerNorm that follows a residual GEMM runs beside it on this stream (fork / join through the two events)
  cudaStream_t ln_stream = nullptr;
  cudaEvent_t ln_fork = nullptr, ln_join = nullptr;
  bool attn_bf16 = true;                // q, k, v and P in bf16 inside an fp16-operand build (WB_ATTN_FP16=1: operand format)
  bool zigzag = true;                   // WB_NO_ZIGZAG=1: every kernel walks its rows ascending (A/B switch, pipeline.cu)
  bool ln_follow = false;               // WB_LN_FOLLOW=1 / wb_debug_set_ln_follow: experiment, measured slower under the power cap
  int max_batch = 32;
  std::mutex mu;                        // calls on one replica serialise ("concurrent calls on one handle serialise per device stream")
  std::vector<void*> allocs;            // weight allocations (freed in free_replica)
  // mel
  MelTables mel{};
  std::map<int, MelTables> htk_tables;  // BatchPreprocessor filterbanks by n_mels (built on first use)
  // conv stem
  op16 *conv1_w = nullptr, *conv2_w = nullptr;
  float *conv1_b = nullptr, *conv2_b = nullptr;
  float conv1_s = 1.f, conv2_s = 1.f;
  float* pe = nullptr;
  std::vector<LayerW> layers;
  int quant = 0;                        // 0: op16 weights resident; 2 / 3: int8 / int4 payloads resident, expanded per layer
  op16 *xp_qkv = nullptr, *xp_o = nullptr, *xp_1 = nullptr, *xp_2 = nullptr;   // expansion buffers (12 d^2 op16: L2-sized)
  float *lnp_g = nullptr, *lnp_b = nullptr;
  DecoderW dec;
  Workspace ws;
  // ragged mel batches (BatchPreprocessor::process_batch, MelFilterbank::compute): arena, rows and host staging, reused
  struct Ragged {
    DevBuf<float> audio, logmel;
    DevBuf<int> keys;
    DevBuf<uint8_t> tables;
    std::vector<uint8_t> h_tables;
    std::vector<float> h_audio, h_out;
  } rag;
  // copy/compute overlap of the host-buffer entry point: two staging slots, copy-in and copy-out streams
  struct Slot {
    DevBuf<float> audio;
    DevBuf<int> n_valid;
    DevBuf<uint8_t> out;
    int* h_n_valid = nullptr;          // pinned
    int h_cap = 0;
    cudaEvent_t in_done = nullptr, compute_done = nullptr, out_done = nullptr;
    bool busy = false;
  } slot[2];
  int next_slot = 0;
  cudaStream_t in_stream = nullptr, out_stream = nullptr;
  // CUDA graphs of the fused mel + encoder step, keyed by its device pointers and batch size: the step is ~230 launches and as
  // many host-side tensor-map encodes; a replay is one cudaGraphLaunch.  First sighting of a key runs eagerly, the second is captured.
  struct StepGraph {
    const void* in = nullptr; const void* n_valid = nullptr; const void* seg_off = nullptr; long long stride = 0;
    void* out = nullptr; int B = 0; int dtype = 0;
    int seen = 0; cudaGraphExec_t exec = nullptr; long long launches = 0;
  };
  std::vector<StepGraph> graphs;
  bool use_graphs = true;
  // per-kernel timing (wb_profile_*): CUDA events recorded on the launching stream around every launch
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;      // start/stop pairs
  std::vector<int> prof_cat;
  // decoder state (decoder.cu): cross-attention K/V of the current batch, self-attention cache, per-step scratch
  struct DecodeState* dstate = nullptr;
  // peers this device may store to (cudaDeviceEnablePeerAccess done)
  std::vector<int> peers_enabled;
  cudaEvent_t done_event = nullptr;      // "everything enqueued on this replica so far has finished" (cross-device waits)
};

struct ProfScope {
  Replica* m;
  cudaEvent_t stop = nullptr;
  ProfScope(Replica* model, int cat) : m(model) {
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
    ::wb::ProfScope _ps(m, cat); \
    rc = (expr);                \
  } while (0);                  \
  if (rc != WB_OK) return rc

// NVTX ranges named after the reference's renacer spans (src/audio/mel.rs:234 `step_f_mel`, .renacer.toml:11-32 `step_g_encode`):
// resolved from libnvToolsExt at run time when a profiler has it loaded; a no-op otherwise (nvtx.cpp).
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtx_push(name); }
  ~NvtxRange() { nvtx_pop(); }
};

template <typename T>
int dev_alloc(Replica* m, size_t count, T** out) {
  void* p = nullptr;
  WB_CUDA_OK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
  m->allocs.push_back(p);
  *out = static_cast<T*>(p);
  return WB_OK;
}

// ---- loader.cu
int load_replica(Replica* m, const AprFile& f, const uint8_t* pinned_base);     // streams, kernel init, tables, weights (async; caller syncs)
void free_replica(Replica* m);
int upload_mel_tables(Replica* m, const std::vector<float>& filt, int n_mels, MelTables* out);
int requantize_int8_per_channel(Replica* m);

// ---- pipeline.cu
int ensure_workspace(Replica* m, int B);
int check_encoder_dims(const Replica* m);
int check_fused_dims(const Replica* m);
int validate_mel_len(const Replica* m, size_t mel_len, int* T_out);
int encode_device(Replica* m, int B, int T, void* d_out, wb_dtype out_dtype, int n_layers, bool ln_post);
int mel_device(Replica* m, const float* d_audio, long long audio_stride, const long long* d_seg_off, const int* d_n_valid, int B,
               float* d_mel_f32, bool want_bf16);
int encode_same_len(Replica* m, const float* const* mels, const float* d_mel, int B, int T, void* out_host, void* out_dev,
                    size_t out_stride_elems, wb_dtype dt, int n_layers, bool ln_post);
int mel_encode_step(Replica* m, const float* d_audio, long long audio_stride, const long long* d_seg_off, const int* d_n_valid, int nb,
                    void* d_out, wb_dtype out_dtype);
// micro-batch `mb` of this replica's share: audio H2D on the copy-in stream, fused step, then either the D2H of the states on the
// copy-out stream (out_host) or nothing more (states were written to d_out_final, possibly a peer device's memory, by ln_post)
int enqueue_microbatch(Replica* m, const float* const* audio, const size_t* n_samples, int nb, void* out_host, void* d_out_final,
                       wb_dtype out_dtype);
int sync_replica(Replica* m);
int compute_mel_host(Replica* m, const float* const* audio, const size_t* n_samples, const float* contiguous, int B, float* out);
int mel_compute_ragged(Replica* m, const MelTables& tab, const float* const* audio, const size_t* n_samples, int B, size_t hop,
                       float* const* mels_out, const size_t* out_capacity, size_t* frame_counts, size_t* max_frames_out);

// ---- decoder.cu
int load_decoder(Replica* m, const AprFile& f, struct Uploader& up);
void free_decode_state(Replica* m);
int decoder_cross_kv(Replica* m, const op16* d_states, int B);
int decoder_greedy(Replica* m, const op16* d_states, int B, const int* initial_tokens, int n_init, int max_tokens, int suppress_timestamps,
                   int* tokens_out, int* lens_out);
int decoder_greedy_s(Replica* m, const op16* d_states, int B, int S, const int* initial_tokens, int n_init, int max_tokens,
                     int suppress_timestamps, int* tokens_out, int* lens_out, float* logits_last_host);
int decoder_debug_cross_kv(Replica* m, const op16* d_states, int S, int layer, float* k_out, float* v_out);

}  // namespace wb

struct wb_model {
  wb_config cfg{};
  std::vector<wb::Replica*> reps;
  std::mutex mu;
  const uint8_t* registered = nullptr;   // caller bytes pinned with cudaHostRegister during the upload
  // gather buffer of wb_mel_encode_gather: [B][1500][d] on reps[gather_rank]'s device
  void* gather_buf = nullptr;
  size_t gather_bytes = 0;
  int gather_dev_index = -1;
};
