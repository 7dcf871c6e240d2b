// C ABI of the stages in front of the hot path: WAV ingest, sinc resampling, voice-activity detection,
// split_into_chunks as zero-copy chunk VIEWS into uploaded streams, and the streaming processor's overlap-carry chunk assembly for
// many streams at once.  Kernels and host parsers: audio_pre.cu.
#include "model.h"

namespace wb {
int wav_parse(const uint8_t* data, size_t n, wb_wav_info* out);
int launch_pcm_to_mono(const uint8_t* d_raw, int kind, int channels, size_t n_frames, float* d_out, cudaStream_t st);
size_t resample_out_len(size_t n_in, uint32_t source_rate, uint32_t target_rate);
int launch_resample(const float* d_in, size_t n_in, uint32_t source_rate, uint32_t target_rate, int half_len, double beta, float* d_out,
                    size_t n_out, cudaStream_t st);
int launch_vad(const float* d_arena, const long long* d_off, const long long* d_len, const long long* d_frame_off, int n_streams,
               long long max_frames, const wb_vad_config& c, float* d_energy, float* d_zcr, uint8_t* d_events, float* d_seg, int seg_cap,
               int* d_nseg, cudaStream_t st);
int launch_assemble_chunks(const int2* d_ready, int n_ready, float* d_acc, long long acc_stride, int chunk_samples, int overlap_samples,
                           float* d_chunks, long long chunk_stride, int* d_n_valid, cudaStream_t st);
}  // namespace wb

using namespace wb;

namespace {
inline Replica* rep0(const wb_model* h) { return (h && !h->reps.empty()) ? h->reps[0] : nullptr; }
constexpr int DEFAULT_KERNEL_HALF_LEN = 16;       // resampler.rs:22
constexpr double DEFAULT_KAISER_BETA = 6.0;       // resampler.rs:25

inline long long vad_frames(size_t n, size_t frame_size) {
  const size_t full = n / frame_size, rem = n - full * frame_size;
  return static_cast<long long>(full + ((rem > 0 && rem >= frame_size / 2) ? 1 : 0));      // vad.rs:562-565
}
}  // namespace

// StreamingProcessor's chunk assembly for N streams (audio_pre.cu: assemble_chunks_kernel); the host keeps the lengths.
struct wb_stream_set {
  wb_model* model = nullptr;
  int n_streams = 0;
  int chunk_samples = 0, overlap_samples = 0;
  long long acc_stride = 0;                  // capacity of one stream's accumulator (floats)
  long long chunk_stride = 0;
  DevBuf<float> acc;                         // [n_streams][acc_stride]
  std::vector<int> len;                      // samples held per stream (carried overlap included)
  std::vector<int> fresh;                    // samples pushed since the stream's last chunk
  DevBuf<float> chunks;                      // [cap][chunk_stride]
  DevBuf<int> n_valid;
  DevBuf<int2> ready;
  DevBuf<uint8_t> out;
};

extern "C" {

// ------------------------------------------------------------------------------------------------ WAV
int wb_wav_parse(const uint8_t* bytes, size_t n_bytes, wb_wav_info* out) {
  if (!out) return set_error(WB_ERR_AUDIO, "null argument");
  memset(out, 0, sizeof *out);
  return wav_parse(bytes, n_bytes, out);
}

int wb_wav_decode(const wb_model* h, const uint8_t* bytes, size_t n_bytes, float* out, size_t out_capacity, wb_wav_info* info) {
  Replica* m = rep0(h);
  if (!m || !info) return set_error(WB_ERR_MODEL, "null argument");
  int rc = wb_wav_parse(bytes, n_bytes, info);
  if (rc != WB_OK) return rc;
  if (info->n_frames == 0) return WB_OK;
  if (!out || out_capacity < info->n_frames) return set_error(WB_ERR_AUDIO, "output buffer too small");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  DevBuf<uint8_t> raw;
  DevBuf<float> mono;
  if ((rc = raw.ensure(info->data_bytes + 16)) != WB_OK || (rc = mono.ensure(info->n_frames)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(raw.p, bytes + info->data_offset, info->data_bytes, cudaMemcpyHostToDevice, m->stream));
  if ((rc = launch_pcm_to_mono(raw.p, info->sample_kind, info->channels, info->n_frames, mono.p, m->stream)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(out, mono.p, info->n_frames * 4, cudaMemcpyDeviceToHost, m->stream));
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));
  return WB_OK;
}

// -------------------------------------------------------------------------------------------- resample
size_t wb_resample_len(size_t n, uint32_t source_rate, uint32_t target_rate) {
  if (source_rate == 0 || target_rate == 0) return 0;
  return resample_out_len(n, source_rate, target_rate);
}

int wb_resample_with_params(const wb_model* h, const float* audio, size_t n, uint32_t source_rate, uint32_t target_rate, int kernel_half_len,
                            double kaiser_beta, float* out, size_t out_capacity, size_t* n_out) {
  Replica* m = rep0(h);
  if (!m) return set_error(WB_ERR_MODEL, "null model");
  if (n_out) *n_out = 0;
  if (source_rate == 0 || target_rate == 0) return set_error(WB_ERR_AUDIO, "sample rate must be non-zero");         // resampler.rs:88-90
  if (kernel_half_len <= 0) return set_error(WB_ERR_AUDIO, "kernel half-length must be non-zero");                    // resampler.rs:91-95
  if (n == 0 || !audio) return set_error(WB_ERR_AUDIO, "cannot resample empty audio");                               // resampler.rs:137-139
  const size_t no = resample_out_len(n, source_rate, target_rate);
  if (no == 0) return set_error(WB_ERR_AUDIO, "output length would be zero");
  if (!out || out_capacity < no) return set_error(WB_ERR_AUDIO, "output buffer too small");
  if (n_out) *n_out = no;
  if (source_rate == target_rate) {                                                                                   // resampler.rs:142-144
    memcpy(out, audio, n * 4);
    return WB_OK;
  }
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  DevBuf<float> in, o;
  int rc;
  if ((rc = in.ensure(n)) != WB_OK || (rc = o.ensure(no)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(in.p, audio, n * 4, cudaMemcpyHostToDevice, m->stream));
  if ((rc = launch_resample(in.p, n, source_rate, target_rate, kernel_half_len, kaiser_beta, o.p, no, m->stream)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(out, o.p, no * 4, cudaMemcpyDeviceToHost, m->stream));
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));
  return WB_OK;
}

int wb_resample(const wb_model* h, const float* audio, size_t n, uint32_t source_rate, uint32_t target_rate, float* out, size_t out_capacity,
                size_t* n_out) {
  return wb_resample_with_params(h, audio, n, source_rate, target_rate, DEFAULT_KERNEL_HALF_LEN, DEFAULT_KAISER_BETA, out, out_capacity, n_out);
}

// WAV bytes -> mono f32 -> 16 kHz in one go: one upload of the raw payload, conversion and resampling on the device, one download.
int wb_ingest_wav_16k(const wb_model* h, const uint8_t* bytes, size_t n_bytes, float* out, size_t out_capacity, size_t* n_out, wb_wav_info* info) {
  Replica* m = rep0(h);
  if (!m || !info) return set_error(WB_ERR_MODEL, "null argument");
  if (n_out) *n_out = 0;
  int rc = wb_wav_parse(bytes, n_bytes, info);
  if (rc != WB_OK) return rc;
  if (info->sample_rate == 0) return set_error(WB_ERR_AUDIO, "sample rate must be non-zero");
  if (info->n_frames == 0) return set_error(WB_ERR_AUDIO, "cannot resample empty audio");
  const size_t no = resample_out_len(info->n_frames, info->sample_rate, 16000);
  if (no == 0) return set_error(WB_ERR_AUDIO, "output length would be zero");
  if (!out || out_capacity < no) return set_error(WB_ERR_AUDIO, "output buffer too small");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  DevBuf<uint8_t> raw;
  DevBuf<float> mono, o;
  if ((rc = raw.ensure(info->data_bytes + 16)) != WB_OK || (rc = mono.ensure(info->n_frames)) != WB_OK || (rc = o.ensure(no)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(raw.p, bytes + info->data_offset, info->data_bytes, cudaMemcpyHostToDevice, m->stream));
  if ((rc = launch_pcm_to_mono(raw.p, info->sample_kind, info->channels, info->n_frames, mono.p, m->stream)) != WB_OK) return rc;
  const float* src = mono.p;
  if (info->sample_rate != 16000) {
    if ((rc = launch_resample(mono.p, info->n_frames, info->sample_rate, 16000, DEFAULT_KERNEL_HALF_LEN, DEFAULT_KAISER_BETA, o.p, no, m->stream)) != WB_OK) return rc;
    src = o.p;
  }
  WB_CUDA_OK(cudaMemcpyAsync(out, src, no * 4, cudaMemcpyDeviceToHost, m->stream));
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));
  if (n_out) *n_out = no;
  return WB_OK;
}

// ------------------------------------------------------------------------------------------------- VAD
void wb_vad_config_default(wb_vad_config* c) {
  if (!c) return;
  c->sample_rate = 16000; c->frame_size = 480; c->energy_threshold = 2.0f; c->zcr_threshold = 0.3f;      // vad.rs:54-66
  c->min_speech_frames = 3; c->min_silence_frames = 10; c->smoothing = 0.95f;
}

int wb_vad_detect_batch(const wb_model* h, const float* const* audio, const size_t* n_samples, int B, const wb_vad_config* cfg,
                        float* segments, int seg_capacity, int* n_segments, uint8_t* const* frame_events, size_t* n_frames_out) {
  Replica* m = rep0(h);
  if (!m || B < 0 || (B > 0 && (!audio || !n_samples || !n_segments))) return set_error(WB_ERR_MODEL, "null argument");
  wb_vad_config c;
  if (cfg) c = *cfg; else wb_vad_config_default(&c);
  if (c.frame_size == 0 || c.sample_rate == 0) return set_error(WB_ERR_AUDIO, "frame size and sample rate must be non-zero");
  if (seg_capacity < 0 || (seg_capacity > 0 && !segments)) return set_error(WB_ERR_AUDIO, "null argument");
  if (B == 0) return WB_OK;
  std::vector<long long> off(B), len(B), foff(B);
  long long arena = 0, frames = 0, max_frames = 0;
  for (int i = 0; i < B; ++i) {
    off[i] = arena;
    len[i] = static_cast<long long>(n_samples[i]);
    foff[i] = frames;
    const long long nf = vad_frames(n_samples[i], c.frame_size);
    if (n_frames_out) n_frames_out[i] = static_cast<size_t>(nf);
    arena += static_cast<long long>((n_samples[i] + 3) & ~static_cast<size_t>(3));
    frames += nf;
    max_frames = std::max(max_frames, nf);
  }
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  cudaStream_t st = m->stream;
  DevBuf<float> d_audio, d_energy, d_zcr, d_seg;
  DevBuf<long long> d_tab;
  DevBuf<uint8_t> d_ev;
  DevBuf<int> d_nseg;
  int rc;
  if ((rc = d_audio.ensure(static_cast<size_t>(arena) + 4)) || (rc = d_energy.ensure(static_cast<size_t>(frames) + 1)) || (rc = d_zcr.ensure(static_cast<size_t>(frames) + 1)) ||
      (rc = d_ev.ensure(static_cast<size_t>(frames) + 1)) || (rc = d_seg.ensure(static_cast<size_t>(B) * std::max(seg_capacity, 1) * 3)) ||
      (rc = d_nseg.ensure(B)) || (rc = d_tab.ensure(3 * static_cast<size_t>(B))))
    return rc;
  std::vector<float> h_audio(static_cast<size_t>(arena));
  for (int i = 0; i < B; ++i)
    if (n_samples[i]) memcpy(h_audio.data() + off[i], audio[i], n_samples[i] * 4);
  std::vector<long long> tab(3 * static_cast<size_t>(B));
  memcpy(tab.data(), off.data(), 8 * static_cast<size_t>(B));
  memcpy(tab.data() + B, len.data(), 8 * static_cast<size_t>(B));
  memcpy(tab.data() + 2 * B, foff.data(), 8 * static_cast<size_t>(B));
  if (arena) WB_CUDA_OK(cudaMemcpyAsync(d_audio.p, h_audio.data(), static_cast<size_t>(arena) * 4, cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemcpyAsync(d_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st));
  if ((rc = launch_vad(d_audio.p, d_tab.p, d_tab.p + B, d_tab.p + 2 * B, B, max_frames, c, d_energy.p, d_zcr.p, d_ev.p, d_seg.p, seg_capacity,
                       d_nseg.p, st)) != WB_OK)
    return rc;
  std::vector<uint8_t> h_ev(static_cast<size_t>(frames));
  if (frame_events && frames) WB_CUDA_OK(cudaMemcpyAsync(h_ev.data(), d_ev.p, static_cast<size_t>(frames), cudaMemcpyDeviceToHost, st));
  if (seg_capacity) WB_CUDA_OK(cudaMemcpyAsync(segments, d_seg.p, static_cast<size_t>(B) * seg_capacity * 3 * 4, cudaMemcpyDeviceToHost, st));
  WB_CUDA_OK(cudaMemcpyAsync(n_segments, d_nseg.p, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, st));
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_error(WB_ERR_CUDA, std::string("VAD kernels failed: ") + cudaGetErrorString(e));
  if (frame_events)
    for (int i = 0; i < B; ++i) {
      const long long nf = (i + 1 < B ? foff[i + 1] : frames) - foff[i];
      if (frame_events[i] && nf) memcpy(frame_events[i], h_ev.data() + foff[i], static_cast<size_t>(nf));
    }
  return WB_OK;
}

// ----------------------------------------------------------------------------------- chunk views of streams
// audio::split_into_chunks (src/audio/batch.rs:219-240) + compute_mel + Encoder::forward_batch for every chunk of every stream,
// WITHOUT materialising the chunks: each stream is uploaded once, a chunk is an (offset, length) view into it that the mel kernel
// reads in place (samples past the view read as 0 = compute_mel's zero padding to 30 s).  Compared with cutting on the host and
// sending padded 30 s chunks, a 5 s / 0.5 s-overlap workload moves 6.6x fewer bytes over PCIe.  out: [total chunks][1500][d],
// stream-major, chunk order inside a stream as split_into_chunks produces it.  Streams are sharded over the handle's devices
// (stream s -> device floor(s * G / n_streams)).
int wb_stream_encode_views(const wb_model* ch, const float* const* streams, const size_t* stream_lens, int n_streams, size_t chunk_size,
                           size_t overlap, void* out, wb_dtype out_dtype, size_t out_capacity_chunks, size_t* chunk_counts, size_t* total_chunks) {
  wb_model* h = const_cast<wb_model*>(ch);
  Replica* m0 = rep0(h);
  if (!m0 || n_streams < 0 || (n_streams > 0 && (!streams || !stream_lens))) return set_error(WB_ERR_MODEL, "null argument");
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(m0);
  if (rc != WB_OK) return rc;
  if (chunk_size > static_cast<size_t>(N_SAMPLES_30S)) return set_error(WB_ERR_AUDIO, "chunk size above 30 s (compute_mel would truncate it)");
  std::vector<size_t> counts(n_streams), first(n_streams + 1, 0);
  for (int s = 0; s < n_streams; ++s) {
    counts[s] = wb_split_into_chunks(stream_lens[s], chunk_size, overlap, nullptr, nullptr, 0);
    if (chunk_counts) chunk_counts[s] = counts[s];
    first[s + 1] = first[s] + counts[s];
  }
  const size_t total = first[n_streams];
  if (total_chunks) *total_chunks = total;
  if (total == 0) return WB_OK;
  if (!out || out_capacity_chunks < total) return set_error(WB_ERR_MODEL, "output buffer too small");
  std::lock_guard<std::mutex> hl(h->mu);
  const int G = static_cast<int>(h->reps.size());
  const size_t d = h->cfg.n_audio_state, per = static_cast<size_t>(N_POS_30S) * d * dtype_size(out_dtype);
  for (int g = 0; g < G; ++g) {
    const int s0 = static_cast<int>((static_cast<long long>(g) * n_streams + G - 1) / G);
    const int s1 = std::min<int>(n_streams, static_cast<int>((static_cast<long long>(g + 1) * n_streams + G - 1) / G));
    if (s1 <= s0) continue;
    Replica* m = h->reps[g];
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard guard(m->device);
    // this device's streams -> arena; chunk table
    std::vector<long long> seg_off;
    std::vector<int> n_valid;
    long long arena = 0;
    Replica::Ragged& r = m->rag;
    std::vector<long long> soff(s1 - s0);
    for (int s = s0; s < s1; ++s) {
      soff[s - s0] = arena;
      arena += static_cast<long long>((stream_lens[s] + 3) & ~static_cast<size_t>(3));
    }
    r.h_audio.resize(static_cast<size_t>(arena));
    for (int s = s0; s < s1; ++s) {
      if (stream_lens[s]) memcpy(r.h_audio.data() + soff[s - s0], streams[s], stream_lens[s] * 4);
      std::vector<size_t> st(counts[s]), ln(counts[s]);
      wb_split_into_chunks(stream_lens[s], chunk_size, overlap, st.data(), ln.data(), counts[s]);
      for (size_t c = 0; c < counts[s]; ++c) {
        seg_off.push_back(soff[s - s0] + static_cast<long long>(st[c]));
        n_valid.push_back(static_cast<int>(ln[c]));
      }
    }
    const size_t n_chunks = seg_off.size();
    if (n_chunks == 0) continue;
    if ((rc = r.audio.ensure(static_cast<size_t>(arena) + 4)) != WB_OK) return rc;
    const size_t t_bytes = n_chunks * 12 + 16;
    if ((rc = r.tables.ensure(t_bytes)) != WB_OK) return rc;
    r.h_tables.resize(t_bytes);
    memcpy(r.h_tables.data(), seg_off.data(), n_chunks * 8);
    memcpy(r.h_tables.data() + n_chunks * 8, n_valid.data(), n_chunks * 4);
    cudaStream_t st = m->stream;
    WB_CUDA_OK(cudaMemcpyAsync(r.audio.p, r.h_audio.data(), static_cast<size_t>(arena) * 4, cudaMemcpyHostToDevice, st));
    WB_CUDA_OK(cudaMemcpyAsync(r.tables.p, r.h_tables.data(), t_bytes, cudaMemcpyHostToDevice, st));
    const long long* d_off = reinterpret_cast<const long long*>(r.tables.p);
    const int* d_nv = reinterpret_cast<const int*>(r.tables.p + n_chunks * 8);
    uint8_t* dst = static_cast<uint8_t*>(out) + first[s0] * per;
    for (size_t c0 = 0; c0 < n_chunks; c0 += m->max_batch) {
      const int nb = static_cast<int>(std::min<size_t>(m->max_batch, n_chunks - c0));
      if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
      void* d_o = out_dtype != WB_F32 ? static_cast<void*>(m->ws.out_bf16.p) : static_cast<void*>(m->ws.out_f32.p);
      if ((rc = mel_encode_step(m, r.audio.p, 0, d_off + c0, d_nv + c0, nb, d_o, out_dtype)) != WB_OK) return rc;
      WB_CUDA_OK(cudaMemcpyAsync(dst + c0 * per, d_o, static_cast<size_t>(nb) * per, cudaMemcpyDeviceToHost, st));
    }
  }
  for (Replica* m : h->reps) {
    DeviceGuard guard(m->device);
    cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) return set_error(WB_ERR_CUDA, std::string("stream view encode failed: ") + cudaGetErrorString(e));
  }
  return WB_OK;
}

// The device-resident form: the streams already sit in HBM (d_arena), d_seg_off / d_n_valid describe n_chunks views into them.
// Enqueues on the model's stream, does not synchronise.  d_out: [n_chunks][1500][d].
int wb_mel_encode_views_dev(const wb_model* h, const float* d_arena, const long long* d_seg_off, const int* d_n_valid, int n_chunks,
                            void* d_out, wb_dtype out_dtype) {
  Replica* m = rep0(h);
  if (!m || !d_arena || !d_seg_off || !d_n_valid || !d_out || n_chunks < 0) return set_error(WB_ERR_MODEL, "null argument");
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  const size_t per = static_cast<size_t>(N_POS_30S) * m->cfg.n_audio_state * dtype_size(out_dtype);
  for (int c0 = 0; c0 < n_chunks; c0 += m->max_batch) {
    const int nb = std::min(m->max_batch, n_chunks - c0);
    if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
    if ((rc = mel_encode_step(m, d_arena, 0, d_seg_off + c0, d_n_valid + c0, nb, static_cast<uint8_t*>(d_out) + static_cast<size_t>(c0) * per,
                              out_dtype)) != WB_OK)
      return rc;
  }
  return WB_OK;
}

// ------------------------------------------------------------------------------- streaming chunk assembly
int wb_stream_set_new(const wb_model* ch, int n_streams, size_t chunk_samples, size_t overlap_samples, wb_stream_set** out) {
  wb_model* h = const_cast<wb_model*>(ch);
  Replica* m = rep0(h);
  if (!m || !out) return set_error(WB_ERR_MODEL, "null argument");
  *out = nullptr;
  if (n_streams < 1) return set_error(WB_ERR_AUDIO, "need at least one stream");
  if (chunk_samples == 0 || chunk_samples > static_cast<size_t>(N_SAMPLES_30S)) return set_error(WB_ERR_AUDIO, "chunk size must be 1..480000 samples");
  if (overlap_samples >= chunk_samples) return set_error(WB_ERR_AUDIO, "overlap must be smaller than the chunk");
  std::unique_ptr<wb_stream_set> s(new wb_stream_set());
  s->model = h;
  s->n_streams = n_streams;
  s->chunk_samples = static_cast<int>(chunk_samples);
  s->overlap_samples = static_cast<int>(overlap_samples);
  s->acc_stride = static_cast<long long>((2 * chunk_samples + overlap_samples + 3) & ~static_cast<size_t>(3));
  s->chunk_stride = static_cast<long long>((chunk_samples + 3) & ~static_cast<size_t>(3));
  s->len.assign(n_streams, 0);
  s->fresh.assign(n_streams, 0);
  DeviceGuard guard(m->device);
  int rc = s->acc.ensure(static_cast<size_t>(n_streams) * s->acc_stride);
  if (rc != WB_OK) return rc;
  *out = s.release();
  return WB_OK;
}

void wb_stream_set_free(wb_stream_set* s) {
  if (!s) return;
  Replica* m = rep0(s->model);
  if (m) {
    DeviceGuard guard(m->device);
    cudaStreamSynchronize(m->stream);
    delete s;
  } else {
    delete s;
  }
}

// StreamingProcessor::push_audio for `count` (stream, samples) pairs.  A stream holds at most two chunks of audio: take chunks first.
int wb_stream_set_push(wb_stream_set* s, const int* stream_ids, const float* const* samples, const size_t* n_samples, int count) {
  if (!s || count < 0 || (count > 0 && (!stream_ids || !samples || !n_samples))) return set_error(WB_ERR_MODEL, "null argument");
  Replica* m = rep0(s->model);
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  for (int i = 0; i < count; ++i) {
    const int id = stream_ids[i];
    if (id < 0 || id >= s->n_streams) return set_error(WB_ERR_AUDIO, "stream id out of range");
    if (static_cast<long long>(s->len[id]) + static_cast<long long>(n_samples[i]) > s->acc_stride)
      return set_error(WB_ERR_AUDIO, "stream " + std::to_string(id) + " holds more than two chunks of audio: take chunks first");
  }
  for (int i = 0; i < count; ++i) {
    const int id = stream_ids[i];
    if (n_samples[i] == 0) continue;
    WB_CUDA_OK(cudaMemcpyAsync(s->acc.p + static_cast<long long>(id) * s->acc_stride + s->len[id], samples[i], n_samples[i] * 4, cudaMemcpyHostToDevice, m->stream));
    s->len[id] += static_cast<int>(n_samples[i]);
    s->fresh[id] += static_cast<int>(n_samples[i]);
  }
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));            // the caller's sample buffers may be reused on return
  return WB_OK;
}

// has_chunk for every stream: ids of the streams that hold a full chunk (StreamingProcessor state ChunkReady, streaming.rs:770-777)
int wb_stream_set_ready(const wb_stream_set* s, int* ids_out, int capacity) {
  if (!s) return 0;
  int n = 0;
  for (int i = 0; i < s->n_streams; ++i)
    if (s->len[i] >= s->chunk_samples) {
      if (ids_out && n < capacity) ids_out[n] = i;
      ++n;
    }
  return n;
}

// get_chunk (streaming.rs:843-870) for every ready stream -- with flush != 0, flush() (:872-905) for every stream holding fresh audio --
// assembled ON THE DEVICE ([carried overlap | samples] zero padded to the chunk size, the chunk's tail kept as the next overlap),
// then compute_mel + encoder over the assembled batch.  out: [n][1500][d] host; ids_out / n_valid_out [n]: which stream each chunk
// belongs to and how many of its samples are real; *n_chunks_out = n (<= capacity; the rest stay queued).
int wb_stream_set_encode(wb_stream_set* s, int flush, void* out, wb_dtype out_dtype, int* ids_out, size_t* n_valid_out, int capacity,
                         int* n_chunks_out) {
  if (!s || !n_chunks_out) return set_error(WB_ERR_MODEL, "null argument");
  *n_chunks_out = 0;
  Replica* m = rep0(s->model);
  if (out_dtype != WB_F32 && out_dtype != WB_BF16) return set_error(WB_ERR_MODEL, "output dtype must be WB_F32 or WB_BF16");
  int rc = check_fused_dims(m);
  if (rc != WB_OK) return rc;
  std::vector<int2> ready;
  for (int i = 0; i < s->n_streams && static_cast<int>(ready.size()) < capacity; ++i)
    if (s->len[i] >= s->chunk_samples || (flush && s->fresh[i] > 0)) ready.push_back(make_int2(i, s->len[i]));
  const int n = static_cast<int>(ready.size());
  if (n == 0) return WB_OK;
  if (!out) return set_error(WB_ERR_MODEL, "null output");
  std::lock_guard<std::mutex> lk(m->mu);
  DeviceGuard guard(m->device);
  cudaStream_t st = m->stream;
  if ((rc = s->chunks.ensure(static_cast<size_t>(n) * s->chunk_stride)) || (rc = s->n_valid.ensure(n)) || (rc = s->ready.ensure(n))) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(s->ready.p, ready.data(), n * sizeof(int2), cudaMemcpyHostToDevice, st));
  if ((rc = launch_assemble_chunks(s->ready.p, n, s->acc.p, s->acc_stride, s->chunk_samples, s->overlap_samples, s->chunks.p, s->chunk_stride,
                                   s->n_valid.p, st)) != WB_OK)
    return rc;
  const size_t d = m->cfg.n_audio_state, per = static_cast<size_t>(N_POS_30S) * d * dtype_size(out_dtype);
  for (int c0 = 0; c0 < n; c0 += m->max_batch) {
    const int nb = std::min(m->max_batch, n - c0);
    if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
    void* d_o = out_dtype != WB_F32 ? static_cast<void*>(m->ws.out_bf16.p) : static_cast<void*>(m->ws.out_f32.p);
    if ((rc = mel_encode_step(m, s->chunks.p + static_cast<long long>(c0) * s->chunk_stride, s->chunk_stride, nullptr, s->n_valid.p + c0, nb, d_o,
                              out_dtype)) != WB_OK)
      return rc;
    WB_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(out) + static_cast<size_t>(c0) * per, d_o, static_cast<size_t>(nb) * per, cudaMemcpyDeviceToHost, st));
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_error(WB_ERR_CUDA, std::string("stream chunk encode failed: ") + cudaGetErrorString(e));
  // mirror what the kernel did to the accumulators
  for (int i = 0; i < n; ++i) {
    const int id = ready[i].x, len = ready[i].y;
    const int take = std::min(len, s->chunk_samples);
    const int keep = take > s->overlap_samples ? s->overlap_samples : 0;
    s->len[id] = keep + (len - take);
    s->fresh[id] = len - take;
    if (ids_out) ids_out[i] = id;
    if (n_valid_out) n_valid_out[i] = static_cast<size_t>(take);
  }
  *n_chunks_out = n;
  return WB_OK;
}

// test hook: the assembled chunk batch of the last wb_stream_set_encode call ([n][chunk_samples] f32)
int wb_debug_stream_set_chunks(const wb_stream_set* s, int n, float* out) {
  if (!s || !out || n < 0) return set_error(WB_ERR_MODEL, "null argument");
  Replica* m = rep0(s->model);
  DeviceGuard guard(m->device);
  for (int i = 0; i < n; ++i)
    WB_CUDA_OK(cudaMemcpy(out + static_cast<size_t>(i) * s->chunk_samples, s->chunks.p + static_cast<long long>(i) * s->chunk_stride,
                          static_cast<size_t>(s->chunk_samples) * 4, cudaMemcpyDeviceToHost));
  return WB_OK;
}

}  // extern "C"
