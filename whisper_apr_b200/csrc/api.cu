// The C ABI of libwhisper_b200.so (include/whisper_b200.h): handle management, the reference-shaped entry points and the
// multi-device layer (chunk sharding + the gather of encoder states), the successor of src/parallel.rs (ordered parallel_map over
// independent items, :82-118) for a box of B200s.
//
// Sharding rule: chunk i of a B-chunk call runs on device floor(i * G / B) -- contiguous blocks, order preserved,
// weights replicated, no collective on the data path.  Every replica runs its own three-stream pipeline (pipeline.cu); the host
// thread feeds the replicas round-robin, one micro-batch at a time, so all devices start working at once.
#include <thread>

#include "model.h"

namespace wb {

std::atomic<long long> g_launch_count{0};
static thread_local std::string g_err;
int set_error(int status, const std::string& msg) {
  g_err = msg;
  return status;
}
const char* last_error() { return g_err.c_str(); }

namespace {

inline Replica* rep0(const wb_model* h) { return (h && !h->reps.empty()) ? h->reps[0] : nullptr; }

// [start, end) of the chunks device g of G owns: chunk i -> device floor(i * G / B)
inline void shard_range(int B, int G, int g, int* start, int* end) {
  // smallest i with floor(i * G / B) >= g  is ceil(g * B / G)
  *start = static_cast<int>((static_cast<long long>(g) * B + G - 1) / G);
  *end = static_cast<int>((static_cast<long long>(g + 1) * B + G - 1) / G);
  if (*end > B) *end = B;
}

// Device `from` may store into device `to`'s memory (cudaDeviceEnablePeerAccess, once per ordered pair).
int enable_peer(Replica* from, int to_device) {
  if (from->device == to_device) return WB_OK;
  if (std::find(from->peers_enabled.begin(), from->peers_enabled.end(), to_device) != from->peers_enabled.end()) return WB_OK;
  DeviceGuard guard(from->device);
  int can = 0;
  WB_CUDA_OK(cudaDeviceCanAccessPeer(&can, from->device, to_device));
  if (!can) return set_error(WB_ERR_CUDA, "devices " + std::to_string(from->device) + " and " + std::to_string(to_device) + " have no peer access (NVLink / PCIe P2P)");
  cudaError_t e = cudaDeviceEnablePeerAccess(to_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_error(WB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
  cudaGetLastError();
  from->peers_enabled.push_back(to_device);
  return WB_OK;
}

// Feed B chunks to the handle's replicas.  out_host: [B][1500][d] host buffer (states copied back per micro-batch), or nullptr;
// d_out: device buffer on replica `out_rep` (peer stores from the others), or nullptr.
int sharded_enqueue(wb_model* h, const float* const* audio, const size_t* n_samples, int B, void* out_host, void* d_out, wb_dtype dt) {
  const int G = static_cast<int>(h->reps.size());
  const size_t d = h->cfg.n_audio_state, per = static_cast<size_t>(N_POS_30S) * d * dtype_size(dt);
  std::vector<int> pos(G), end(G);
  for (int g = 0; g < G; ++g) shard_range(B, G, g, &pos[g], &end[g]);
  bool more = true;
  while (more) {                                    // round-robin: micro-batch k of every device before micro-batch k + 1 of any
    more = false;
    for (int g = 0; g < G; ++g) {
      if (pos[g] >= end[g]) continue;
      Replica* m = h->reps[g];
      std::lock_guard<std::mutex> lk(m->mu);
      DeviceGuard guard(m->device);
      const int nb = std::min(m->max_batch, end[g] - pos[g]);
      void* oh = out_host ? static_cast<uint8_t*>(out_host) + static_cast<size_t>(pos[g]) * per : nullptr;
      void* od = d_out ? static_cast<uint8_t*>(d_out) + static_cast<size_t>(pos[g]) * per : nullptr;
      int rc = enqueue_microbatch(m, audio + pos[g], n_samples + pos[g], nb, oh, od, dt);
      if (rc != WB_OK) return rc;
      pos[g] += nb;
      if (pos[g] < end[g]) more = true;
    }
  }
  return WB_OK;
}

}  // namespace
}  // namespace wb

using namespace wb;

// =====================================================================================================
extern "C" {

const char* wb_version(void) { return kOp16IsFp16 ? "whisper_b200 0.2.0 (sm_100a, fp16 operands)" : "whisper_b200 0.2.0 (sm_100a, bf16 operands)"; }
const char* wb_operand_format(void) { return kOp16IsFp16 ? "fp16" : "bf16"; }
const char* wb_last_error(void) { return wb::last_error(); }

int wb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int wb_model_from_apr_devices(const uint8_t* bytes, size_t n_bytes, const int* devices, int n_devices, wb_model** out) {
  if (!out) return set_error(WB_ERR_MODEL, "null output handle");
  *out = nullptr;
  AprFile f;
  int rc = parse_apr(bytes, n_bytes, &f);
  if (rc != WB_OK) return rc;
  if (f.cfg.quantization == 1) return set_error(WB_ERR_FORMAT, "F16 .apr payloads have no reader (as in the reference)");
  // an untrusted header sizes every device allocation: refuse dimensions no Whisper variant comes near
  if (f.cfg.n_audio_state > 16384 || f.cfg.n_audio_layer > 512 || f.cfg.n_audio_ctx > 65536 || f.cfg.n_mels > 1024 ||
      f.cfg.n_audio_head > 256 || f.cfg.n_text_state > 16384 || f.cfg.n_text_layer > 512 || f.cfg.n_text_ctx > 65536 ||
      f.cfg.n_vocab > (1u << 22))
    return set_error(WB_ERR_FORMAT, "unreasonable model dimensions in the .apr header");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_error(WB_ERR_CUDA, "no CUDA device: libwhisper_b200 has no CPU fallback");
  }
  if (!devices || n_devices < 1 || n_devices > WB_MAX_DEVICES) return set_error(WB_ERR_CUDA, "empty or oversized device list");
  for (int i = 0; i < n_devices; ++i) {
    if (devices[i] < 0 || devices[i] >= ndev) return set_error(WB_ERR_CUDA, "invalid CUDA device ordinal");
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) return set_error(WB_ERR_CUDA, "device listed twice");
    int cc_major = 0;
    WB_CUDA_OK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, devices[i]));
    if (cc_major != 10)
      return set_error(WB_ERR_CUDA, "kernels are built for sm_100a only; device " + std::to_string(devices[i]) + " has compute capability " +
                                        std::to_string(cc_major) + ".x");
  }
  std::unique_ptr<wb_model> h(new wb_model());
  h->cfg = f.cfg;
  // page-lock the caller's bytes for the duration of the upload: the H2D pieces become true DMA (and run concurrently on every
  // device's own PCIe link).  Failure is not an error -- the copies then go through the driver's staging buffer.
  bool registered = false;
  if (getenv("WB_NO_HOST_REGISTER") == nullptr && n_bytes >= (8u << 20)) {
    if (cudaHostRegister(const_cast<uint8_t*>(bytes), n_bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly) == cudaSuccess) registered = true;
    else {
      cudaGetLastError();
      if (cudaHostRegister(const_cast<uint8_t*>(bytes), n_bytes, cudaHostRegisterPortable) == cudaSuccess) registered = true;
      else cudaGetLastError();
    }
  }
  auto fail = [&](int r) {
    const std::string keep = wb::last_error();
    for (Replica* m : h->reps) free_replica(m);
    h->reps.clear();
    if (registered) cudaHostUnregister(const_cast<uint8_t*>(bytes));
    return set_error(r, keep);
  };
  for (int i = 0; i < n_devices; ++i) {
    Replica* m = new Replica();
    m->device = devices[i];
    h->reps.push_back(m);
  }
  if (n_devices == 1) {
    if ((rc = load_replica(h->reps[0], f, registered ? bytes : nullptr)) != WB_OK) return fail(rc);
  } else {
    // every device pulls the same pinned bytes over its own PCIe link at the same time: one host thread per replica
    std::vector<int> rcs(n_devices, WB_OK);
    std::vector<std::string> errs(n_devices);
    std::vector<std::thread> th;
    for (int i = 0; i < n_devices; ++i)
      th.emplace_back([&, i]() {
        rcs[i] = load_replica(h->reps[i], f, registered ? bytes : nullptr);
        if (rcs[i] != WB_OK) errs[i] = wb::last_error();
      });
    for (auto& t : th) t.join();
    for (int i = 0; i < n_devices; ++i)
      if (rcs[i] != WB_OK) return fail(set_error(rcs[i], errs[i]));
  }
  if (registered) cudaHostUnregister(const_cast<uint8_t*>(bytes));
  h->cfg = h->reps[0]->cfg;
  *out = h.release();
  return WB_OK;
}

int wb_model_from_apr(const uint8_t* bytes, size_t n_bytes, int device, wb_model** out) {
  return wb_model_from_apr_devices(bytes, n_bytes, &device, 1, out);
}

int wb_model_config(const wb_model* h, wb_config* out) {
  if (!h || !out) return set_error(WB_ERR_MODEL, "null argument");
  *out = h->cfg;
  return WB_OK;
}

int wb_model_n_devices(const wb_model* h) { return h ? static_cast<int>(h->reps.size()) : 0; }

int wb_model_device(const wb_model* h, int index) {
  if (!h || index < 0 || index >= static_cast<int>(h->reps.size())) return -1;
  return h->reps[index]->device;
}

void wb_model_free(wb_model* h) {
  if (!h) return;
  if (h->gather_buf && h->gather_dev_index >= 0) {
    DeviceGuard guard(h->reps[h->gather_dev_index]->device);
    cudaDeviceSynchronize();
    cudaFree(h->gather_buf);
  }
  for (Replica* m : h->reps) free_replica(m);
  delete h;
}

int wb_model_set_stream(wb_model* h, void* cuda_stream) {
  Replica* m = rep0(h);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  if (h->reps.size() != 1) return set_error(WB_ERR_MODEL, "wb_model_set_stream applies to single-device handles (a stream belongs to one device)");
  std::lock_guard<std::mutex> lk(m->mu);
  m->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : m->own_stream;
  return WB_OK;
}

int wb_model_set_max_batch(wb_model* h, int max_chunks) {
  if (!h || max_chunks < 1) return set_error(WB_ERR_MODEL, "max batch must be >= 1");
  for (Replica* m : h->reps) {
    std::lock_guard<std::mutex> lk(m->mu);
    m->max_batch = max_chunks;
  }
  return WB_OK;
}

int wb_model_requantize(wb_model* h, int mode) {
  if (!h) return set_error(WB_ERR_MODEL, "null model");
  if (mode != WB_QUANT_INT8_PER_CHANNEL) return set_error(WB_ERR_MODEL, "unknown requantisation mode");
  for (Replica* m : h->reps) {
    std::lock_guard<std::mutex> lk(m->mu);
    int rc = requantize_int8_per_channel(m);
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

int wb_sync(const wb_model* ch) {
  wb_model* h = const_cast<wb_model*>(ch);
  if (!h) return set_error(WB_ERR_MODEL, "null model");
  for (Replica* m : h->reps) {
    std::lock_guard<std::mutex> lk(m->mu);
    int rc = sync_replica(m);
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

// ---------------------------------------------------------------------------------------------- mel
int wb_mel_compute(const wb_model* h, const float* audio, size_t n, size_t hop, float* out, size_t out_capacity, size_t* n_frames_out) {
  Replica* m = rep0(h);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  if (n_frames_out) *n_frames_out = 0;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const float* ptrs[1] = {audio};
  const size_t lens[1] = {n};
  float* outs[1] = {out};
  const size_t caps[1] = {out_capacity};
  size_t cnt[1] = {0};
  int rc = mel_compute_ragged(m, m->mel, ptrs, lens, 1, hop, outs, caps, cnt, nullptr);
  if (rc == WB_OK && n_frames_out) *n_frames_out = cnt[0];
  return rc;
}

// BatchPreprocessor::process_batch (src/audio/batch.rs:157-176): every segment goes through MelFilterbank::compute with the
// preprocessor's OWN filterbank -- MelFilterbank::new(n_mels, n_fft, sample_rate), the HTK triangles (batch.rs:143), not the model's --
// and is neither padded nor truncated.  Tables for a given n_mels are built once per model and kept.
int wb_batch_preprocess(const wb_model* h, const float* const* audio, const size_t* n_samples, int B, size_t n_mels, size_t hop,
                        float* const* mels_out, const size_t* out_capacity, size_t* frame_counts, size_t* max_frames_out) {
  Replica* m = rep0(h);
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
  return mel_compute_ragged(m, it->second, audio, n_samples, B, hop, mels_out, out_capacity, frame_counts, max_frames_out);
}

int wb_compute_mel(const wb_model* h, const float* audio, size_t n, float* out) {
  Replica* m = rep0(h);
  if (!m || !out || (n && !audio)) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const float* ptrs[1] = {audio};
  const size_t lens[1] = {n};
  return compute_mel_host(m, ptrs, lens, nullptr, 1, out);
}

int wb_compute_mel_batch(const wb_model* h, const float* audio, int B, float* out) {
  Replica* m = rep0(h);
  if (!m || !out || !audio || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  return compute_mel_host(m, nullptr, nullptr, audio, B, out);
}

int wb_compute_mel_batch_dev(const wb_model* h, const float* d_audio, int B, float* d_mel_out) {
  Replica* m = rep0(h);
  if (!m || !d_audio || !d_mel_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const size_t per_out = static_cast<size_t>(N_FRAMES_30S) * m->mel.n_mels;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    rc = mel_device(m, d_audio + static_cast<size_t>(b0) * N_SAMPLES_30S, N_SAMPLES_30S, nullptr, nullptr, nb, d_mel_out + static_cast<size_t>(b0) * per_out, false);
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

// ------------------------------------------------------------------------------------------ encoder
int wb_encode(const wb_model* h, const float* mel, size_t mel_len, float* out, size_t out_capacity, size_t* seq_len_out) {
  Replica* m = rep0(h);
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

int wb_encode_batch(const wb_model* h, const float* const* mels, const size_t* mel_lens, int B, float* out, size_t out_capacity,
                    size_t* seq_lens, size_t* max_seq_out) {
  Replica* m = rep0(h);
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

int wb_encode_batch_dev(const wb_model* h, const float* d_mel, int B, void* d_out, wb_dtype out_dtype) {
  Replica* m = rep0(h);
  if (!m || !d_mel || !d_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  return encode_same_len(m, nullptr, d_mel, B, N_FRAMES_30S, nullptr, d_out, static_cast<size_t>(N_POS_30S) * m->cfg.n_audio_state, out_dtype, -1, true);
}

// ------------------------------------------------------------------------------------- fused hot path
int wb_mel_encode_batch_dev(const wb_model* h, const float* d_audio, int B, void* d_out, wb_dtype out_dtype) {
  Replica* m = rep0(h);
  if (!m || !d_audio || !d_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const size_t d = m->cfg.n_audio_state, S = N_POS_30S, esz = dtype_size(out_dtype);
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
    if ((rc = mel_encode_step(m, d_audio + static_cast<size_t>(b0) * N_SAMPLES_30S, N_SAMPLES_30S, nullptr, nullptr, nb,
                              static_cast<uint8_t*>(d_out) + static_cast<size_t>(b0) * S * d * esz, out_dtype)) != WB_OK)
      return rc;
  }
  return WB_OK;
}

int wb_mel_encode_batch_async(const wb_model* ch, const float* const* audio, const size_t* n_samples, int B, void* out, wb_dtype out_dtype) {
  wb_model* h = const_cast<wb_model*>(ch);
  if (!rep0(h) || !audio || !n_samples || !out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(rep0(h));
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(h->mu);
  return sharded_enqueue(h, audio, n_samples, B, out, nullptr, out_dtype);
}

int wb_mel_encode_batch(const wb_model* h, const float* const* audio, const size_t* n_samples, int B, void* out, wb_dtype out_dtype) {
  int rc = wb_mel_encode_batch_async(h, audio, n_samples, B, out, out_dtype);
  if (rc != WB_OK) return rc;
  return wb_sync(h);
}

// The same sharded call with the encoder states left ON A DEVICE: device `gather_index` (an index into the handle's device list)
// receives all B chunks' states, [B][1500][d] in chunk order, in a library-owned buffer.  The other devices' final LayerNorm kernels
// store their rows straight into that buffer over NVLink (peer stores issued by the compute kernel itself, per micro-batch -- the
// gather overlaps the next micro-batch and costs no extra pass, no staging copy and no NCCL call).  *d_states_out is valid until
// the next gather call on this handle; wb_sync completes it.
int wb_mel_encode_gather(const wb_model* ch, const float* const* audio, const size_t* n_samples, int B, wb_dtype out_dtype,
                         int gather_index, void** d_states_out) {
  wb_model* h = const_cast<wb_model*>(ch);
  if (!rep0(h) || !audio || !n_samples || !d_states_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  const int G = static_cast<int>(h->reps.size());
  if (gather_index < 0 || gather_index >= G) return set_error(WB_ERR_MODEL, "gather index outside the handle's device list");
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(rep0(h));
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(h->mu);
  const size_t bytes = static_cast<size_t>(B) * N_POS_30S * h->cfg.n_audio_state * dtype_size(out_dtype);
  Replica* root = h->reps[gather_index];
  if (h->gather_dev_index != gather_index || h->gather_bytes < bytes) {
    if ((rc = wb_sync(h)) != WB_OK) return rc;                 // nobody may still be writing the old buffer
    if (h->gather_buf) {
      DeviceGuard guard(h->reps[h->gather_dev_index]->device);
      cudaFree(h->gather_buf);
      h->gather_buf = nullptr;
      h->gather_bytes = 0;
    }
    DeviceGuard guard(root->device);
    WB_CUDA_OK(cudaMalloc(&h->gather_buf, std::max<size_t>(bytes, 256)));
    h->gather_bytes = bytes;
    h->gather_dev_index = gather_index;
    for (Replica* m : h->reps) {
      // captured step graphs hold the old destination
      for (auto& g : m->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
      m->graphs.clear();
    }
  }
  for (Replica* m : h->reps)
    if ((rc = enable_peer(m, root->device)) != WB_OK) return rc;
  if ((rc = sharded_enqueue(h, audio, n_samples, B, nullptr, h->gather_buf, out_dtype)) != WB_OK) return rc;
  // the root's stream waits for every other replica's last step: work enqueued on the root after this call sees all states
  for (Replica* m : h->reps) {
    if (m == root) continue;
    DeviceGuard guard(m->device);
    WB_CUDA_OK(cudaEventRecord(m->done_event, m->stream));
    WB_CUDA_OK(cudaStreamWaitEvent(root->stream, m->done_event, 0));
  }
  *d_states_out = h->gather_buf;
  return WB_OK;
}

// ---- peer-visible device buffers across PROCESSES (one process per GPU, torchrun): the gather rank exports its states buffer, the
// other ranks open it and pass `peer + offset` as d_out of wb_mel_encode_batch_dev, so their final LayerNorm stores over NVLink too.
int wb_ipc_alloc(int device, size_t bytes, void** d_ptr, uint8_t* handle64) {
  if (!d_ptr || !handle64) return set_error(WB_ERR_MODEL, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard guard(device);
  WB_CUDA_OK(cudaMalloc(d_ptr, std::max<size_t>(bytes, 256)));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, *d_ptr);
  if (e != cudaSuccess) {
    cudaFree(*d_ptr);
    *d_ptr = nullptr;
    return set_error(WB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &hd, 64);
  return WB_OK;
}
int wb_ipc_open(int device, const uint8_t* handle64, void** d_ptr) {
  if (!d_ptr || !handle64) return set_error(WB_ERR_MODEL, "null argument");
  DeviceGuard guard(device);
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_error(WB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  return WB_OK;
}
int wb_ipc_close(int device, void* d_ptr) {
  DeviceGuard guard(device);
  WB_CUDA_OK(cudaIpcCloseMemHandle(d_ptr));
  return WB_OK;
}
int wb_ipc_free(int device, void* d_ptr) {
  DeviceGuard guard(device);
  WB_CUDA_OK(cudaFree(d_ptr));
  return WB_OK;
}
// Copy a device buffer (any device this process can address, e.g. the gather buffer) to the host: test / consumer convenience.
int wb_read_device(int device, const void* d_src, void* host_dst, size_t bytes) {
  DeviceGuard guard(device);
  WB_CUDA_OK(cudaMemcpy(host_dst, d_src, bytes, cudaMemcpyDeviceToHost));
  return WB_OK;
}

// ------------------------------------------------------------------------------------------ decoder
namespace {
// host f32 states [B][S][d] -> op16 on the replica's device (DecodeState::states is private to decoder.cu: use the workspace)
int states_to_device_bf16(Replica* m, const float* states, size_t rows, DevBuf<float>& f32, DevBuf<op16>& b16) {
  const size_t n = rows * m->cfg.n_audio_state;
  int rc;
  if ((rc = f32.ensure(n)) != WB_OK || (rc = b16.ensure(n)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(f32.p, states, n * 4, cudaMemcpyHostToDevice, m->stream));
  return launch_f32_to_op16(f32.p, b16.p, n, m->stream);
}

// transcribe_batch_optimized steps 1-3 for one replica's shard: mel + encoder (op16 states stay in HBM) -> cross K/V + greedy loop
int transcribe_shard(Replica* m, const float* const* audio, const size_t* n_samples, int cnt, const int32_t* init, int n_init, int max_tokens,
                     int suppress_ts, int32_t* tokens_out, int32_t* lens_out) {
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const int step = std::min(m->max_batch, 32);
  for (int b0 = 0; b0 < cnt; b0 += step) {
    const int nb = std::min(step, cnt - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    std::vector<int> nv(nb);
    for (int i = 0; i < nb; ++i) {
      const size_t n = std::min<size_t>(n_samples[b0 + i], N_SAMPLES_30S);
      nv[i] = static_cast<int>(n);
      if (n) WB_CUDA_OK(cudaMemcpyAsync(m->ws.audio.p + static_cast<size_t>(i) * N_SAMPLES_30S, audio[b0 + i], n * 4, cudaMemcpyHostToDevice, m->stream));
    }
    WB_CUDA_OK(cudaMemcpy(m->ws.n_valid.p, nv.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
    if ((rc = mel_encode_step(m, m->ws.audio.p, N_SAMPLES_30S, nullptr, m->ws.n_valid.p, nb, m->ws.out_bf16.p, WB_OP16)) != WB_OK) return rc;
    if ((rc = decoder_greedy(m, m->ws.out_bf16.p, nb, init, n_init, max_tokens, suppress_ts, tokens_out + static_cast<size_t>(b0) * max_tokens,
                             lens_out + b0)) != WB_OK)
      return rc;
  }
  return WB_OK;
}
}  // namespace

int wb_decoder_available(const wb_model* h) {
  Replica* m = rep0(h);
  return (m && m->dec.loaded) ? 1 : 0;
}

int wb_decode_greedy(const wb_model* h, const float* states, size_t seq_len, int B, const int32_t* initial_tokens, int n_init, int max_tokens,
                     int suppress_timestamps, int32_t* tokens_out, int32_t* lens_out) {
  Replica* m = rep0(h);
  if (!m || !states || !tokens_out || !lens_out || B < 1) return set_error(WB_ERR_MODEL, "null argument");
  if (max_tokens < 1) return set_error(WB_ERR_MODEL, "max_tokens must be positive");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  DevBuf<float> f32;
  DevBuf<op16> b16;
  const int step = 32;
  for (int b0 = 0; b0 < B; b0 += step) {
    const int nb = std::min(step, B - b0);
    int rc = states_to_device_bf16(m, states + static_cast<size_t>(b0) * seq_len * m->cfg.n_audio_state, static_cast<size_t>(nb) * seq_len, f32, b16);
    if (rc != WB_OK) return rc;
    if ((rc = decoder_greedy_s(m, b16.p, nb, static_cast<int>(seq_len), initial_tokens, n_init, max_tokens, suppress_timestamps,
                               tokens_out + static_cast<size_t>(b0) * max_tokens, lens_out + b0, nullptr)) != WB_OK)
      return rc;
  }
  return WB_OK;
}

int wb_transcribe_tokens_batch(const wb_model* ch, const float* const* audio, const size_t* n_samples, int B, const int32_t* initial_tokens,
                               int n_init, int max_tokens, int suppress_timestamps, int32_t* tokens_out, int32_t* lens_out) {
  wb_model* h = const_cast<wb_model*>(ch);
  if (!rep0(h) || !audio || !n_samples || !tokens_out || !lens_out || B < 0) return set_error(WB_ERR_MODEL, "null argument");
  if (max_tokens < 1) return set_error(WB_ERR_MODEL, "max_tokens must be positive");
  int rc = check_fused_dims(rep0(h));
  if (rc != WB_OK) return rc;
  if (!rep0(h)->dec.loaded) return set_error(WB_ERR_MODEL, "the .apr file carries no decoder tensors");
  std::lock_guard<std::mutex> lk(h->mu);
  const int G = static_cast<int>(h->reps.size());
  if (G == 1) return transcribe_shard(h->reps[0], audio, n_samples, B, initial_tokens, n_init, max_tokens, suppress_timestamps, tokens_out, lens_out);
  // the decoder is sharded exactly like the encoder (chunk i -> device floor(i*G/B)): the states never leave their device and no
  // gather is needed; one host thread per device drives its encode -> decode sequence
  std::vector<int> rcs(G, WB_OK);
  std::vector<std::string> errs(G);
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g) {
    int s0, s1;
    shard_range(B, G, g, &s0, &s1);
    if (s1 <= s0) continue;
    th.emplace_back([&, g, s0, s1]() {
      rcs[g] = transcribe_shard(h->reps[g], audio + s0, n_samples + s0, s1 - s0, initial_tokens, n_init, max_tokens, suppress_timestamps,
                                tokens_out + static_cast<size_t>(s0) * max_tokens, lens_out + s0);
      if (rcs[g] != WB_OK) errs[g] = wb::last_error();
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < G; ++g)
    if (rcs[g] != WB_OK) return set_error(rcs[g], errs[g]);
  return WB_OK;
}

// test hooks of the decoder: logits of forward_one after feeding `tokens`; the cross-attention K/V of one layer
int wb_debug_decoder_logits(const wb_model* h, const float* states, size_t seq_len, const int32_t* tokens, int n_tokens, float* logits_out) {
  Replica* m = rep0(h);
  if (!m || !states || !tokens || !logits_out || n_tokens < 1) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  DevBuf<float> f32;
  DevBuf<op16> b16;
  int rc = states_to_device_bf16(m, states, seq_len, f32, b16);
  if (rc != WB_OK) return rc;
  std::vector<int32_t> toks(n_tokens + 1);
  int len = 0;
  return decoder_greedy_s(m, b16.p, 1, static_cast<int>(seq_len), tokens, n_tokens, n_tokens + 1, 1, toks.data(), &len, logits_out);
}

int wb_debug_cross_kv(const wb_model* h, const float* states, size_t seq_len, int layer, float* k_out, float* v_out) {
  Replica* m = rep0(h);
  if (!m || !states || !k_out || !v_out) return set_error(WB_ERR_MODEL, "null argument");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  DevBuf<float> f32;
  DevBuf<op16> b16;
  int rc = states_to_device_bf16(m, states, seq_len, f32, b16);
  if (rc != WB_OK) return rc;
  return decoder_debug_cross_kv(m, b16.p, static_cast<int>(seq_len), layer, k_out, v_out);
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

}  // extern "C"
