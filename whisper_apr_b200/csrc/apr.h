// `.apr` v1 container parser (host side, C++).  Mirrors AprReader::new / load_tensor / read_mel_filterbank of the
// reference (src/format/mod.rs:484-522, 610-672, 736-780); no copy of the file is made (the reference copies it
// whole, format/mod.rs:484) -- tensors are addressed in place and uploaded straight from the caller's buffer.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/whisper_b200.h"

namespace wb {

struct AprTensor {
  std::string name;
  uint64_t offset = 0;       // relative to the data section
  uint64_t size = 0;         // bytes
  uint64_t n_elements = 0;
  uint32_t shape[4] = {0, 0, 0, 0};
  uint8_t n_dims = 0;
  float scale = 1.0f;        // Int8 / Int4 per-tensor scale (format/mod.rs:495-501)
};

struct AprFile {
  wb_config cfg{};
  std::vector<AprTensor> tensors;
  const uint8_t* bytes = nullptr;
  size_t n_bytes = 0;
  size_t data_offset = 0;
  bool has_filterbank = false;
  uint32_t fb_mels = 0, fb_freqs = 0;
  const uint8_t* fb_data = nullptr;      // little-endian f32 [fb_mels][fb_freqs]

  const AprTensor* find(const std::string& name) const;
  // Bytes of a tensor payload, or nullptr when it runs past the end of the file (the reference's load_tensor
  // returns Err in that case and the loader silently keeps the default, src/lib.rs:769-800).
  const uint8_t* payload(const AprTensor& t, size_t* n_bytes_out) const;
};

// Returns WB_OK or WB_ERR_FORMAT (message via wb_last_error), same failure cases as AprReader::new.
int parse_apr(const uint8_t* bytes, size_t n, AprFile* out);

// MelFilterbank::new fallback filterbank (src/audio/mel.rs:144-197), [n_mels][n_fft/2+1] f32.
std::vector<float> htk_filterbank(int n_mels, int n_fft, int sample_rate);
// MelFilterbank::hann_window (src/audio/mel.rs:215-219), periodic, evaluated in f32.
std::vector<float> hann_window_periodic(int n);
// Encoder::create_positional_embedding (src/model/encoder.rs:429-441).
std::vector<float> default_positional_embedding(int max_len, int d_model);

}  // namespace wb
