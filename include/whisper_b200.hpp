// whisper_b200.hpp -- C++17 host-side mirror of the reference's Rust interface for the mel + encoder path, header-only, over the C ABI
// of libwhisper_b200.so (whisper_b200.h).  The reference is compiled code (Rust) whose toolchain is not in this image, so this is the
// compiled-language binding a caller links against; it keeps the reference's names, argument meaning and error behaviour:
//
//   wb::WhisperApr::load_from_apr(bytes)            WhisperApr::load_from_apr        src/lib.rs:673-754
//   .config()                                       WhisperApr::config               src/lib.rs:330-333
//   .compute_mel(audio)                             WhisperApr::compute_mel          src/lib.rs:407-443
//   .mel_filterbank_compute(audio, hop)             MelFilterbank::compute           src/audio/mel.rs:233-310
//   .encode(mel)                                    WhisperApr::encode               src/lib.rs:446-449 (Encoder::forward_mel)
//   .encode_batch_padded(mels)                      Encoder::forward_batch_padded    src/model/encoder.rs:625-660 (BatchEncoderOutput)
//   .mel_encode_batch(chunks)                       transcribe_batch_optimized 1-2   src/lib.rs:1162-1170
//   .transcribe_tokens_batch(chunks, prompt, n)     transcribe_batch_optimized 1-3   src/lib.rs:1162-1201 (token ids)
//   .parse_wav(bytes) / .ingest_wav_16k(bytes)      audio::wav::parse_wav            src/audio/wav.rs:99-293 (+ SincResampler to 16 kHz)
//   .resample(audio, from, to)                      SincResampler::resample          src/audio/resampler.rs:136-250
//   .vad_detect(streams)                            VoiceActivityDetector::detect    src/vad.rs:554-700 (VadConfig::default())
//   .stream_encode_views(streams, size, overlap)    split_into_chunks + compute_mel + forward_batch, chunks read in place
//   wb::split_into_chunks(n, size, overlap)         audio::split_into_chunks         src/audio/batch.rs:219-240
//   wb::WhisperError { kind, what() }               WhisperError::{Audio,Model,Format}  src/error.rs:6-44   (Result<T, WhisperError> -> throw)
//   wb::device_count()                              parallel::thread_count           src/parallel.rs:155-170
//
// Buffers are std::vector<float> like the reference's Vec<f32>; `&[f32]` parameters are (pointer, length) spans.  No CPU fallback: every
// compute call throws WhisperError{Cuda} without a CUDA device.  tests/cpp/host_mirror.cpp exercises it (tests/test_cpp_host.py).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "whisper_b200.h"

namespace wb {

enum class ErrorKind { Audio = WB_ERR_AUDIO, Model = WB_ERR_MODEL, Format = WB_ERR_FORMAT, Cuda = WB_ERR_CUDA };

class WhisperError : public std::runtime_error {
 public:
  WhisperError(int status, const std::string& msg) : std::runtime_error(kind_name(status) + " error: " + msg), status_(status) {}
  ErrorKind kind() const { return static_cast<ErrorKind>(status_); }
  int status() const { return status_; }
  static std::string kind_name(int status) {
    switch (status) {
      case WB_ERR_AUDIO: return "Audio";
      case WB_ERR_MODEL: return "Model";
      case WB_ERR_FORMAT: return "Format";
      case WB_ERR_CUDA: return "Cuda";
      default: return "status " + std::to_string(status);
    }
  }

 private:
  int status_;
};

inline void check(int status) {
  if (status != WB_OK) {
    const char* m = wb_last_error();
    throw WhisperError(status, m ? m : "");
  }
}

inline int device_count() { return wb_device_count(); }
inline std::string version() { return wb_version(); }

// audio::split_into_chunks: (start, len) of every chunk
inline std::vector<std::pair<size_t, size_t>> split_into_chunks(size_t n_samples, size_t chunk_size, size_t overlap) {
  const size_t n = wb_split_into_chunks(n_samples, chunk_size, overlap, nullptr, nullptr, 0);
  std::vector<size_t> st(n), ln(n);
  wb_split_into_chunks(n_samples, chunk_size, overlap, st.data(), ln.data(), n);
  std::vector<std::pair<size_t, size_t>> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = {st[i], ln[i]};
  return out;
}

// BatchEncoderOutput (src/model/encoder.rs:662-720): [batch][max_seq][d_model] zero padded + the true lengths
struct BatchEncoderOutput {
  std::vector<float> data;
  std::vector<size_t> seq_lengths;
  size_t max_seq_len = 0, d_model = 0;
  size_t batch_size() const { return seq_lengths.size(); }
  const float* get(size_t b) const { return data.data() + b * max_seq_len * d_model; }
};

class WhisperApr {
 public:
  static constexpr size_t N_SAMPLES_30S = 480000, N_FRAMES_30S = 3000, N_POS_30S = 1500;

  // devices: CUDA ordinals the handle spans (weights replicated; batch calls shard over them).  Empty = {0}.
  static WhisperApr load_from_apr(const uint8_t* bytes, size_t n_bytes, const std::vector<int>& devices = {}) {
    wb_model* h = nullptr;
    if (devices.empty()) check(wb_model_from_apr(bytes, n_bytes, 0, &h));
    else check(wb_model_from_apr_devices(bytes, n_bytes, devices.data(), static_cast<int>(devices.size()), &h));
    return WhisperApr(h);
  }
  static WhisperApr load_from_apr(const std::vector<uint8_t>& bytes, const std::vector<int>& devices = {}) {
    return load_from_apr(bytes.data(), bytes.size(), devices);
  }

  WhisperApr(WhisperApr&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  WhisperApr& operator=(WhisperApr&& o) noexcept {
    if (this != &o) { wb_model_free(h_); h_ = o.h_; o.h_ = nullptr; }
    return *this;
  }
  WhisperApr(const WhisperApr&) = delete;
  WhisperApr& operator=(const WhisperApr&) = delete;
  ~WhisperApr() { wb_model_free(h_); }

  wb_config config() const {
    wb_config c{};
    check(wb_model_config(h_, &c));
    return c;
  }
  int n_devices() const { return wb_model_n_devices(h_); }
  bool has_decoder() const { return wb_decoder_available(h_) != 0; }
  wb_model* handle() const { return h_; }

  // MelFilterbank::compute: [n_frames][n_mels], frame-major; empty for inputs shorter than one window; hop 0 -> Audio error
  std::vector<float> mel_filterbank_compute(const float* audio, size_t n, size_t hop = 160) const {
    const size_t m = config().n_mels;
    const size_t cap = (n >= 400 && hop > 0) ? ((n - 400) / hop + 1) * m : 0;
    std::vector<float> out(cap);
    size_t frames = 0;
    check(wb_mel_compute(h_, audio, n, hop, out.data(), out.size(), &frames));
    out.resize(frames * m);
    return out;
  }
  // WhisperApr::compute_mel: pad / truncate to 30 s, frames padded to 3000 with -1.0 -> [3000][n_mels]
  std::vector<float> compute_mel(const float* audio, size_t n) const {
    std::vector<float> out(N_FRAMES_30S * config().n_mels);
    check(wb_compute_mel(h_, audio, n, out.data()));
    return out;
  }
  std::vector<float> compute_mel(const std::vector<float>& audio) const { return compute_mel(audio.data(), audio.size()); }

  // WhisperApr::encode: mel [n_frames][n_mels] -> [S][d], S = (n_frames - 1) / 2 + 1
  std::vector<float> encode(const float* mel, size_t mel_len) const {
    const wb_config c = config();
    const size_t frames = c.n_mels ? mel_len / c.n_mels : 0;
    std::vector<float> out((frames ? (frames - 1) / 2 + 1 : 0) * c.n_audio_state);
    size_t S = 0;
    check(wb_encode(h_, mel, mel_len, out.data(), out.size(), &S));
    out.resize(S * c.n_audio_state);
    return out;
  }
  std::vector<float> encode(const std::vector<float>& mel) const { return encode(mel.data(), mel.size()); }

  // Encoder::forward_batch_padded
  BatchEncoderOutput encode_batch_padded(const std::vector<std::vector<float>>& mels) const {
    const wb_config c = config();
    BatchEncoderOutput r;
    r.d_model = c.n_audio_state;
    const int B = static_cast<int>(mels.size());
    if (B == 0) return r;
    std::vector<const float*> p(B);
    std::vector<size_t> ln(B);
    size_t max_s = 0;
    for (int i = 0; i < B; ++i) {
      p[i] = mels[i].data();
      ln[i] = mels[i].size();
      const size_t frames = c.n_mels ? ln[i] / c.n_mels : 0;
      max_s = std::max(max_s, frames ? (frames - 1) / 2 + 1 : size_t{0});
    }
    r.data.assign(static_cast<size_t>(B) * max_s * r.d_model, 0.f);
    r.seq_lengths.resize(B);
    check(wb_encode_batch(h_, p.data(), ln.data(), B, r.data.data(), r.data.size(), r.seq_lengths.data(), &r.max_seq_len));
    r.data.resize(static_cast<size_t>(B) * r.max_seq_len * r.d_model);
    return r;
  }

  // transcribe_batch_optimized steps 1-2: chunks (each padded / truncated to 30 s) -> [B][1500][d] f32
  std::vector<float> mel_encode_batch(const std::vector<std::vector<float>>& chunks) const {
    const int B = static_cast<int>(chunks.size());
    std::vector<float> out(static_cast<size_t>(B) * N_POS_30S * config().n_audio_state);
    if (B == 0) return out;
    std::vector<const float*> p(B);
    std::vector<size_t> ln(B);
    for (int i = 0; i < B; ++i) { p[i] = chunks[i].data(); ln[i] = chunks[i].size(); }
    check(wb_mel_encode_batch(h_, p.data(), ln.data(), B, out.data(), WB_F32));
    return out;
  }

  // ---- ingest in front of the path and the streaming front end (8f-3)
  // WavData (wav.rs:40-60): mono f32 samples + what the header said
  struct WavData {
    std::vector<float> samples;
    wb_wav_info info{};
  };
  WavData parse_wav(const uint8_t* bytes, size_t n_bytes) const {
    WavData w;
    check(wb_wav_parse(bytes, n_bytes, &w.info));
    w.samples.resize(static_cast<size_t>(w.info.n_frames));
    check(wb_wav_decode(h_, bytes, n_bytes, w.samples.data(), w.samples.size(), &w.info));
    return w;
  }
  // WAV bytes -> mono -> 16 kHz (what compute_mel is fed from a file)
  WavData ingest_wav_16k(const uint8_t* bytes, size_t n_bytes) const {
    WavData w;
    check(wb_wav_parse(bytes, n_bytes, &w.info));
    w.samples.resize(wb_resample_len(static_cast<size_t>(w.info.n_frames), w.info.sample_rate, 16000) + 1);
    size_t n = 0;
    check(wb_ingest_wav_16k(h_, bytes, n_bytes, w.samples.data(), w.samples.size(), &n, &w.info));
    w.samples.resize(n);
    return w;
  }
  std::vector<float> resample(const float* audio, size_t n, uint32_t source_rate, uint32_t target_rate) const {
    std::vector<float> out(wb_resample_len(n, source_rate, target_rate) + 1);
    size_t n_out = 0;
    check(wb_resample(h_, audio, n, source_rate, target_rate, out.data(), out.size(), &n_out));
    out.resize(n_out);
    return out;
  }
  // SpeechSegment (vad.rs:80-100): start / end in seconds, mean energy
  struct SpeechSegment { float start, end, energy; };
  std::vector<std::vector<SpeechSegment>> vad_detect(const std::vector<std::vector<float>>& streams, const wb_vad_config* cfg = nullptr,
                                                     int seg_capacity = 64) const {
    const int B = static_cast<int>(streams.size());
    std::vector<std::vector<SpeechSegment>> out(B);
    if (B == 0) return out;
    std::vector<const float*> p(B);
    std::vector<size_t> ln(B);
    for (int i = 0; i < B; ++i) { p[i] = streams[i].data(); ln[i] = streams[i].size(); }
    std::vector<float> seg(static_cast<size_t>(B) * seg_capacity * 3);
    std::vector<int> cnt(B);
    check(wb_vad_detect_batch(h_, p.data(), ln.data(), B, cfg, seg.data(), seg_capacity, cnt.data(), nullptr, nullptr));
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < std::min(cnt[b], seg_capacity); ++i) {
        const float* q = seg.data() + (static_cast<size_t>(b) * seg_capacity + i) * 3;
        out[b].push_back({q[0], q[1], q[2]});
      }
    return out;
  }
  // every chunk of every stream (views into the uploaded streams) -> [total chunks][1500][d]; counts[i] = chunks of stream i
  std::vector<float> stream_encode_views(const std::vector<std::vector<float>>& streams, size_t chunk_size, size_t overlap,
                                         std::vector<size_t>* counts = nullptr) const {
    const int n = static_cast<int>(streams.size());
    std::vector<const float*> p(n);
    std::vector<size_t> ln(n), cnt(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
      p[i] = streams[i].data();
      ln[i] = streams[i].size();
      total += wb_split_into_chunks(ln[i], chunk_size, overlap, nullptr, nullptr, 0);
    }
    std::vector<float> out(total * N_POS_30S * config().n_audio_state);
    size_t got = 0;
    if (n) check(wb_stream_encode_views(h_, p.data(), ln.data(), n, chunk_size, overlap, out.data(), WB_F32, total, cnt.data(), &got));
    out.resize(got * N_POS_30S * config().n_audio_state);
    if (counts) *counts = cnt;
    return out;
  }

  // transcribe_batch_optimized steps 1-3 up to token ids (greedy): one token list per chunk, initial tokens included
  std::vector<std::vector<int32_t>> transcribe_tokens_batch(const std::vector<std::vector<float>>& chunks, const std::vector<int32_t>& initial_tokens,
                                                            int max_tokens, bool suppress_timestamps = true) const {
    const int B = static_cast<int>(chunks.size());
    std::vector<std::vector<int32_t>> out(B);
    if (B == 0) return out;
    std::vector<const float*> p(B);
    std::vector<size_t> ln(B);
    for (int i = 0; i < B; ++i) { p[i] = chunks[i].data(); ln[i] = chunks[i].size(); }
    std::vector<int32_t> toks(static_cast<size_t>(B) * max_tokens), lens(B);
    check(wb_transcribe_tokens_batch(h_, p.data(), ln.data(), B, initial_tokens.data(), static_cast<int>(initial_tokens.size()), max_tokens,
                                     suppress_timestamps ? 1 : 0, toks.data(), lens.data()));
    for (int b = 0; b < B; ++b) out[b].assign(toks.begin() + static_cast<size_t>(b) * max_tokens, toks.begin() + static_cast<size_t>(b) * max_tokens + lens[b]);
    return out;
  }

 private:
  explicit WhisperApr(wb_model* h) : h_(h) {}
  wb_model* h_ = nullptr;
};

}  // namespace wb
