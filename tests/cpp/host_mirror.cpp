// Compiled-language caller of the C ABI through the C++ mirror of the reference's interface (include/whisper_b200.hpp).
//   host_mirror                               CPU-only checks: error kinds / messages of the reference, chunking known answers
//   host_mirror <model.apr> <audio.f32> <out_prefix>
//                                             load, compute_mel, encode, mel_encode_batch (and greedy tokens when the file has a decoder)
//                                             on cuda:0; writes <out_prefix>.mel / .states / .tokens for the Python test to compare
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>

#include "whisper_b200.hpp"

template <class T>
static std::vector<T> read_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<T> out(raw.size() / sizeof(T));
  std::memcpy(out.data(), raw.data(), out.size() * sizeof(T));
  return out;
}
template <class T>
static void write_file(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}
#define EXPECT(cond)                                                         \
  do {                                                                       \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

static int cpu_checks() {
  // AprReader::new on garbage: Format error with the reference's texts (src/format/mod.rs:484-522)
  const uint8_t junk[64] = {'N', 'O', 'P', 'E'};
  try {
    auto m = wb::WhisperApr::load_from_apr(junk, sizeof junk);
    EXPECT(false && "garbage accepted");
  } catch (const wb::WhisperError& e) {
    EXPECT(e.kind() == wb::ErrorKind::Format || e.kind() == wb::ErrorKind::Cuda);     // no device: Cuda comes first
    EXPECT(std::string(e.what()).find("error: ") != std::string::npos);
  }
  // audio::split_into_chunks (src/audio/batch.rs:219-240) known answers: 10 s at 16 kHz, 5 s chunks, 0.5 s overlap
  auto ch = wb::split_into_chunks(160000, 80000, 8000);
  EXPECT(ch.size() == 3);
  EXPECT(ch[0].first == 0 && ch[0].second == 80000);
  EXPECT(ch[1].first == 72000 && ch[1].second == 80000);
  EXPECT(ch[2].first == 144000 && ch[2].second == 16000);
  EXPECT(wb::split_into_chunks(0, 80000, 8000).empty());
  // WAV walk errors (src/audio/wav.rs:99-224)
  wb_wav_info info{};
  EXPECT(wb_wav_parse(junk, 8, &info) == WB_ERR_AUDIO);
  EXPECT(std::string(wb_last_error()).find("WAV file too small") != std::string::npos);
  EXPECT(!wb::version().empty());
  std::printf("cpu checks ok (devices visible: %d)\n", wb::device_count());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 4) return cpu_checks();
  try {
    const auto bytes = read_file<uint8_t>(argv[1]);
    const auto audio = read_file<float>(argv[2]);
    const std::string prefix = argv[3];
    auto model = wb::WhisperApr::load_from_apr(bytes);
    const wb_config cfg = model.config();
    std::printf("loaded: d %u, layers %u, mels %u, decoder %d\n", cfg.n_audio_state, cfg.n_audio_layer, cfg.n_mels, model.has_decoder() ? 1 : 0);
    const auto mel = model.compute_mel(audio);
    write_file(prefix + ".mel", mel);
    const auto states_two_step = model.encode(mel);
    const auto states = model.mel_encode_batch({audio, std::vector<float>(audio.begin(), audio.begin() + audio.size() / 2)});
    write_file(prefix + ".states", states);
    // the fused call and the two-step call are the same computation on the same kernels: bit-identical for chunk 0
    EXPECT(states_two_step.size() == 1500u * cfg.n_audio_state);
    EXPECT(std::memcmp(states_two_step.data(), states.data(), states_two_step.size() * sizeof(float)) == 0);
    // error behaviour of Encoder::forward (encoder.rs:450-461): a mel whose length is not a multiple of n_mels
    try {
      model.encode(mel.data(), mel.size() - 1);
      EXPECT(false && "ragged mel accepted");
    } catch (const wb::WhisperError& e) {
      EXPECT(e.kind() == wb::ErrorKind::Model);
    }
    // MelFilterbank::compute: hop 0 is an Audio error, a short input is empty
    try {
      model.mel_filterbank_compute(audio.data(), audio.size(), 0);
      EXPECT(false && "hop 0 accepted");
    } catch (const wb::WhisperError& e) {
      EXPECT(e.kind() == wb::ErrorKind::Audio);
    }
    EXPECT(model.mel_filterbank_compute(audio.data(), 100).empty());
    EXPECT(model.mel_filterbank_compute(audio.data(), 16000).size() == 98u * cfg.n_mels);       // mel.rs:660-668
    // ingest + streaming rows through the same mirror: resample 8 kHz -> 16 kHz (length law of SincResampler: ceil(n * to / from)), VAD on
    // silence (no segments), two 5 s chunks with 0.5 s overlap of the first 9.5 s read in place as views
    {
      const std::vector<float> lo(audio.begin(), audio.begin() + 8000);
      const auto up = model.resample(lo.data(), lo.size(), 8000, 16000);
      EXPECT(up.size() == 16000u);
      write_file(prefix + ".resampled", up);
      const auto segs = model.vad_detect({std::vector<float>(16000, 0.f)});
      EXPECT(segs.size() == 1 && segs[0].empty());
      std::vector<size_t> counts;
      const std::vector<float> stream(audio.begin(), audio.begin() + 152000);
      const auto views = model.stream_encode_views({stream}, 80000, 8000, &counts);
      EXPECT(counts.size() == 1 && counts[0] == 2);
      EXPECT(views.size() == 2u * 1500u * cfg.n_audio_state);
      write_file(prefix + ".views", views);
    }
    if (model.has_decoder()) {
      const auto toks = model.transcribe_tokens_batch({audio}, {50258, 50259, 50359, 50363}, 12);
      write_file(prefix + ".tokens", toks[0]);
    }
    std::printf("gpu run ok\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("FAILED: %s\n", e.what());
    return 1;
  }
}
