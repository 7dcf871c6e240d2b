// Audio ingest in front of the mel, on the device (SURVEY 8f-4 and the chunk-assembly half of 8f-3): WAV payload -> mono f32,
// Kaiser-windowed sinc resampling to 16 kHz, energy / zero-crossing voice-activity detection for thousands of streams at once, and
// the overlap-carry chunk assembly of the streaming processor.
//
// Replaces, with the same arithmetic (f32 where the reference is f32, f64 where it is f64, same operation order):
//   parse_wav, convert_{8,16,24,32}bit_pcm, convert_32bit_float, convert_to_mono     src/audio/wav.rs:99-293
//   SincResampler::{resample, windowed_sinc, kaiser_window}, bessel_i0                src/audio/resampler.rs:136-250, 260-276
//   VoiceActivityDetector::{detect, process_frame, is_speech_frame, frame_energy,
//                           zero_crossing_rate}                                       src/vad.rs:554-700
//   StreamingProcessor::get_chunk / flush (overlap carry, zero pad)                   src/audio/streaming.rs:843-905
#include "model.h"

namespace wb {
namespace {

constexpr uint16_t WAVE_FORMAT_PCM = 1, WAVE_FORMAT_IEEE_FLOAT = 3, WAVE_FORMAT_EXTENSIBLE = 0xFFFE;

inline uint16_t rd16(const uint8_t* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) {
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24);
}

// One thread per output (mono) sample.  kind: 0 u8, 1 i16, 2 i24, 3 i32, 4 f32.
__global__ void __launch_bounds__(256) pcm_to_mono_kernel(const uint8_t* __restrict__ raw, int kind, int channels, size_t n_frames,
                                                          float* __restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_frames) return;
  float v[2] = {0.f, 0.f};
  for (int c = 0; c < channels; ++c) {
    const size_t s = i * channels + c;
    float x;
    switch (kind) {
      case 0: x = (static_cast<float>(raw[s]) - 128.0f) / 128.0f; break;                                           // wav.rs:238-240
      case 1: {
        const int16_t t = static_cast<int16_t>(raw[2 * s] | (raw[2 * s + 1] << 8));
        x = static_cast<float>(t) / 32768.0f;                                                                       // wav.rs:228-235
        break;
      }
      case 2: {
        int32_t t = raw[3 * s] | (raw[3 * s + 1] << 8) | (raw[3 * s + 2] << 16);
        if (t & 0x800000) t |= 0xFF000000;                                                                          // wav.rs:244-252
        x = static_cast<float>(t) / 8388608.0f;
        break;
      }
      case 3: {
        const int32_t t = static_cast<int32_t>(raw[4 * s] | (raw[4 * s + 1] << 8) | (raw[4 * s + 2] << 16) | (static_cast<uint32_t>(raw[4 * s + 3]) << 24));
        x = static_cast<float>(t) / 2147483648.0f;                                                                  // wav.rs:256-263
        break;
      }
      default: {
        const uint32_t u = raw[4 * s] | (raw[4 * s + 1] << 8) | (raw[4 * s + 2] << 16) | (static_cast<uint32_t>(raw[4 * s + 3]) << 24);
        x = __uint_as_float(u);                                                                                     // wav.rs:267-271
      }
    }
    v[c] = x;
  }
  out[i] = channels == 2 ? __fdiv_rn(__fadd_rn(v[0], v[1]), 2.0f) : v[0];                                           // wav.rs:275-285
}

// bessel_i0 (resampler.rs:260-276): the series with the reference's stopping rule
__device__ __forceinline__ double bessel_i0_dev(double x) {
  double sum = 1.0, term = 1.0;
  const double q = (x * x) / 4.0;
  for (int k = 1; k < 50; ++k) {
    term *= q / static_cast<double>(k * k);
    sum += term;
    if (fabs(term) < 1e-15 * fabs(sum)) break;
  }
  return sum;
}

// SincResampler::resample (resampler.rs:136-206): one thread per output sample, f64 like the reference.
__global__ void __launch_bounds__(128) sinc_resample_kernel(const float* __restrict__ in, long long n_in, double ratio, int half_len, double beta,
                                                            double i0_beta, float* __restrict__ out, long long n_out) {
  const long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const double PI = 3.14159265358979323846;
  const double cutoff = ratio < 1.0 ? ratio : 1.0;
  const double in_pos = static_cast<double>(o) / ratio;
  const double fl = floor(in_pos);
  const long long center = static_cast<long long>(fl);
  const double frac = in_pos - fl;
  double sum = 0.0, wsum = 0.0;
  for (int k = -half_len; k <= half_len; ++k) {
    const long long idx = center + k;
    if (idx < 0 || idx >= n_in) continue;
    const double x = static_cast<double>(k) - frac;
    const double sarg = cutoff * x;
    const double sinc = fabs(sarg) < 1e-10 ? 1.0 : sin(PI * sarg) / (PI * sarg);
    const double warg = x / static_cast<double>(half_len);
    double win = 0.0;
    if (!(fabs(warg) > 1.0)) win = bessel_i0_dev(beta * sqrt(fmax(fma(warg, -warg, 1.0), 0.0))) / i0_beta;      // x.mul_add(-x, 1.0)
    const double v = sinc * win;
    sum += static_cast<double>(in[idx]) * v;
    wsum += v;
  }
  out[o] = fabs(wsum) > 1e-10 ? static_cast<float>(sum / wsum) : 0.0f;
}

// VAD features: one thread per frame, the reference's sequential f32 sums (no FMA contraction: Rust does not fuse).
__global__ void __launch_bounds__(128) vad_features_kernel(const float* __restrict__ arena, const long long* __restrict__ stream_off,
                                                           const long long* __restrict__ stream_len, const long long* __restrict__ frame_off,
                                                           int n_streams, int frame_size, float* __restrict__ energy, float* __restrict__ zcr) {
  const int s = blockIdx.y;
  const long long n = stream_len[s];
  const long long n_full = n / frame_size, rem = n - n_full * frame_size;
  const long long n_frames = n_full + ((rem > 0 && rem >= frame_size / 2) ? 1 : 0);                                // vad.rs:562-565
  const long long f = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const float* x = arena + stream_off[s] + f * frame_size;
  const int len = static_cast<int>(f < n_full ? frame_size : rem);
  float sum = 0.f;
  int crossings = 0;
  float prev = x[0];
  sum = __fadd_rn(sum, __fmul_rn(prev, prev));
  for (int i = 1; i < len; ++i) {
    const float v = x[i];
    sum = __fadd_rn(sum, __fmul_rn(v, v));                                                                         // vad.rs:671-674
    crossings += ((prev >= 0.f) != (v >= 0.f)) ? 1 : 0;                                                            // vad.rs:682-685
    prev = v;
  }
  energy[frame_off[s] + f] = __fsqrt_rn(__fdiv_rn(sum, static_cast<float>(len)));
  zcr[frame_off[s] + f] = len < 2 ? 0.f : __fdiv_rn(static_cast<float>(crossings), static_cast<float>(len - 1));
}

struct VadParams {
  int frame_size, min_speech_frames, min_silence_frames, sample_rate;
  float energy_threshold, zcr_threshold, smoothing;
};

// The sequential half of VoiceActivityDetector::detect (vad.rs:554-607 over process_frame :609-660): one thread per STREAM walks
// its frames -- thousands of streams advance side by side.  events[frame]: 0 Continue, 1 SpeechStart, 2 SpeechEnd.
__global__ void __launch_bounds__(128) vad_scan_kernel(const float* __restrict__ energy, const float* __restrict__ zcr,
                                                       const long long* __restrict__ stream_len, const long long* __restrict__ frame_off,
                                                       int n_streams, VadParams p, uint8_t* __restrict__ events, float* __restrict__ seg,
                                                       int seg_cap, int* __restrict__ n_seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const long long n = stream_len[s];
  const long long n_full = n / p.frame_size, rem = n - n_full * p.frame_size;
  const long long n_frames = n_full + ((rem > 0 && rem >= p.frame_size / 2) ? 1 : 0);
  float noise_floor = 0.001f;
  int state = 0, speech_frames = 0, silence_frames = 0;       // state 0 Silence, 1 Speech
  long long current_sample = 0;
  bool in_seg = false;
  float seg_start = 0.f, energy_sum = 0.f;
  int frame_count = 0, count = 0;
  float* my_seg = seg + static_cast<size_t>(s) * seg_cap * 3;
  const float sr = static_cast<float>(p.sample_rate);
  for (long long f = 0; f < n_frames; ++f) {
    const float e = energy[frame_off[s] + f], z = zcr[frame_off[s] + f];
    if (state == 0) noise_floor = fmaf(p.smoothing, noise_floor, __fmul_rn(__fsub_rn(1.0f, p.smoothing), e));      // vad.rs:614-619
    const bool is_speech = (e > __fmul_rn(noise_floor, p.energy_threshold)) && (z > 0.05f && z < p.zcr_threshold);  // vad.rs:663-672
    int ev = 0;
    if (state == 0) {
      if (is_speech) {
        ++speech_frames;
        silence_frames = 0;
        if (speech_frames >= p.min_speech_frames) { state = 1; ev = 1; }
      } else {
        speech_frames = 0;
      }
    } else {
      if (is_speech) {
        silence_frames = 0;
        ++speech_frames;
      } else {
        ++silence_frames;
        speech_frames = 0;
        if (silence_frames >= p.min_silence_frames) { state = 0; ev = 2; }
      }
    }
    events[frame_off[s] + f] = static_cast<uint8_t>(ev);
    const float time = __fdiv_rn(static_cast<float>(current_sample), sr);
    if (ev == 1) {
      in_seg = true; seg_start = time; energy_sum = e; frame_count = 1;
    } else if (ev == 2) {
      if (in_seg) {
        if (count < seg_cap) { my_seg[3 * count] = seg_start; my_seg[3 * count + 1] = time; my_seg[3 * count + 2] = __fdiv_rn(energy_sum, static_cast<float>(max(frame_count, 1))); }
        ++count;
        in_seg = false;
      }
    } else if (in_seg) {
      energy_sum = __fadd_rn(energy_sum, e);
      ++frame_count;
    }
    current_sample += (f < n_full ? p.frame_size : rem);
  }
  if (in_seg) {                                                // unterminated speech segment (vad.rs:596-604)
    const float time = __fdiv_rn(static_cast<float>(current_sample), sr);
    if (count < seg_cap) { my_seg[3 * count] = seg_start; my_seg[3 * count + 1] = time; my_seg[3 * count + 2] = __fdiv_rn(energy_sum, static_cast<float>(max(frame_count, 1))); }
    ++count;
  }
  n_seg[s] = count;
}

// StreamingProcessor::get_chunk for many streams (streaming.rs:843-870): chunk = [carried overlap | fresh samples] zero padded to
// chunk_samples; the last overlap_samples of what was taken become the next chunk's prefix.  One block per ready stream.
__global__ void __launch_bounds__(256) assemble_chunks_kernel(const int2* __restrict__ ready, float* __restrict__ acc, long long acc_stride,
                                                              int chunk_samples, int overlap_samples, float* __restrict__ chunks,
                                                              long long chunk_stride, int* __restrict__ n_valid) {
  const int s = ready[blockIdx.x].x;                   // {stream id, samples held}: the host keeps the lengths
  float* a = acc + static_cast<long long>(s) * acc_stride;
  const int len = ready[blockIdx.x].y;
  const int take = min(len, chunk_samples);
  float* c = chunks + static_cast<long long>(blockIdx.x) * chunk_stride;
  for (int i = threadIdx.x; i < chunk_stride; i += blockDim.x) c[i] = i < take ? a[i] : 0.f;
  __syncthreads();                                              // every read of a[0 .. take) is done before the carry overwrites the front
  const int keep = take > overlap_samples ? overlap_samples : 0;    // streaming.rs:850-854: only when the chunk is longer than the overlap
  const int rest = len - take;
  // new accumulator: [last `keep` samples of the chunk | the `rest` samples beyond it]; both sources lie at or after their
  // destination, so staging through registers in two passes is enough
  float tmp[8];
  const int total = keep + rest;
  for (int base = 0; base < total; base += blockDim.x * 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * blockDim.x + threadIdx.x;
      tmp[k] = i < total ? (i < keep ? a[take - keep + i] : a[take + (i - keep)]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * blockDim.x + threadIdx.x;
      if (i < total) a[i] = tmp[k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_valid[blockIdx.x] = take;
}

}  // namespace

// -------------------------------------------------------------------------------------------------------------- host
// parse_wav's chunk walk (wav.rs:99-224); error strings are WavError's Display texts.
int wav_parse(const uint8_t* data, size_t n, wb_wav_info* out) {
  if (!data || n < 44) return set_error(WB_ERR_AUDIO, "WAV file too small");
  if (memcmp(data, "RIFF", 4) != 0) return set_error(WB_ERR_AUDIO, "missing RIFF header");
  if (memcmp(data + 8, "WAVE", 4) != 0) return set_error(WB_ERR_AUDIO, "missing WAVE format");
  size_t pos = 12;
  uint32_t sample_rate = 0;
  uint16_t channels = 0, bits = 0, audio_format = 0, sub_format = 0;
  while (pos + 8 <= n) {
    const uint8_t* id = data + pos;
    const size_t size = rd32(data + pos + 4);
    if (memcmp(id, "fmt ", 4) == 0) {
      if (pos + 8 + size > n || size < 16) return set_error(WB_ERR_AUDIO, "fmt chunk truncated");
      audio_format = rd16(data + pos + 8);
      channels = rd16(data + pos + 10);
      sample_rate = rd32(data + pos + 12);
      bits = rd16(data + pos + 22);
      if (audio_format == WAVE_FORMAT_EXTENSIBLE && size >= 40) {
        const size_t off = pos + 8 + 24;
        if (off + 2 <= n) sub_format = rd16(data + off);
      }
      pos += 8 + size;
    } else if (memcmp(id, "data", 4) == 0) {
      const size_t start = pos + 8;
      const size_t end = std::min(start + size, n);
      const uint16_t eff = audio_format == WAVE_FORMAT_EXTENSIBLE ? sub_format : audio_format;
      const std::string unsup = "unsupported format " + std::to_string(audio_format) + " with " + std::to_string(bits) + " bits";
      if (eff != WAVE_FORMAT_PCM && eff != WAVE_FORMAT_IEEE_FLOAT) return set_error(WB_ERR_AUDIO, unsup);
      int kind;
      if (eff == WAVE_FORMAT_PCM && bits == 16) kind = 1;
      else if (eff == WAVE_FORMAT_PCM && bits == 8) kind = 0;
      else if (eff == WAVE_FORMAT_PCM && bits == 24) kind = 2;
      else if (eff == WAVE_FORMAT_PCM && bits == 32) kind = 3;
      else if (eff == WAVE_FORMAT_IEEE_FLOAT && bits == 32) kind = 4;
      else return set_error(WB_ERR_AUDIO, unsup);
      if (channels != 1 && channels != 2) return set_error(WB_ERR_AUDIO, "unsupported channel count " + std::to_string(channels));
      const size_t bps = kind == 0 ? 1 : (kind == 1 ? 2 : (kind == 2 ? 3 : 4));
      const size_t n_samples = (end - start) / bps;                 // chunks_exact drops a trailing partial sample
      out->sample_rate = sample_rate;
      out->channels = channels;
      out->bits_per_sample = bits;
      out->sample_kind = static_cast<uint16_t>(kind);
      out->data_offset = start;
      out->data_bytes = end - start;
      out->n_frames = n_samples / channels;                         // chunks_exact(2) drops an unpaired trailing sample
      return WB_OK;
    } else {
      pos += 8 + size;
      if (size % 2 != 0) pos += 1;
    }
  }
  return set_error(WB_ERR_AUDIO, "no data chunk");
}

int launch_pcm_to_mono(const uint8_t* d_raw, int kind, int channels, size_t n_frames, float* d_out, cudaStream_t st) {
  if (n_frames == 0) return WB_OK;
  pcm_to_mono_kernel<<<static_cast<unsigned>((n_frames + 255) / 256), 256, 0, st>>>(d_raw, kind, channels, n_frames, d_out);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

static double bessel_i0_host(double x) {
  double sum = 1.0, term = 1.0;
  const double q = (x * x) / 4.0;
  for (int k = 1; k < 50; ++k) {
    term *= q / static_cast<double>(k * k);
    sum += term;
    if (fabs(term) < 1e-15 * fabs(sum)) break;
  }
  return sum;
}

size_t resample_out_len(size_t n_in, uint32_t source_rate, uint32_t target_rate) {
  if (source_rate == target_rate) return n_in;
  const double ratio = static_cast<double>(target_rate) / static_cast<double>(source_rate);
  return static_cast<size_t>(ceil(static_cast<double>(n_in) * ratio));
}

int launch_resample(const float* d_in, size_t n_in, uint32_t source_rate, uint32_t target_rate, int half_len, double beta, float* d_out,
                    size_t n_out, cudaStream_t st) {
  if (n_out == 0) return WB_OK;
  const double ratio = static_cast<double>(target_rate) / static_cast<double>(source_rate);
  sinc_resample_kernel<<<static_cast<unsigned>((n_out + 127) / 128), 128, 0, st>>>(d_in, static_cast<long long>(n_in), ratio, half_len, beta,
                                                                                    bessel_i0_host(beta), d_out, static_cast<long long>(n_out));
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_vad(const float* d_arena, const long long* d_off, const long long* d_len, const long long* d_frame_off, int n_streams,
               long long max_frames, const wb_vad_config& c, float* d_energy, float* d_zcr, uint8_t* d_events, float* d_seg, int seg_cap,
               int* d_nseg, cudaStream_t st) {
  if (n_streams <= 0) return WB_OK;
  if (max_frames > 0) {
    dim3 grid(static_cast<unsigned>((max_frames + 127) / 128), n_streams);
    vad_features_kernel<<<grid, 128, 0, st>>>(d_arena, d_off, d_len, d_frame_off, n_streams, static_cast<int>(c.frame_size), d_energy, d_zcr);
    count_launch();
  }
  VadParams p;
  p.frame_size = static_cast<int>(c.frame_size); p.min_speech_frames = static_cast<int>(c.min_speech_frames);
  p.min_silence_frames = static_cast<int>(c.min_silence_frames); p.sample_rate = static_cast<int>(c.sample_rate);
  p.energy_threshold = c.energy_threshold; p.zcr_threshold = c.zcr_threshold; p.smoothing = c.smoothing;
  vad_scan_kernel<<<(n_streams + 127) / 128, 128, 0, st>>>(d_energy, d_zcr, d_len, d_frame_off, n_streams, p, d_events, d_seg, seg_cap, d_nseg);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_assemble_chunks(const int2* d_ready, int n_ready, float* d_acc, long long acc_stride, int chunk_samples, int overlap_samples,
                           float* d_chunks, long long chunk_stride, int* d_n_valid, cudaStream_t st) {
  if (n_ready <= 0) return WB_OK;
  assemble_chunks_kernel<<<n_ready, 256, 0, st>>>(d_ready, d_acc, acc_stride, chunk_samples, overlap_samples, d_chunks, chunk_stride, d_n_valid);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

}  // namespace wb
