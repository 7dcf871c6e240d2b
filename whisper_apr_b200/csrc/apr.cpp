#include "apr.h"

#include <math.h>
#include <string.h>

#include "wb_internal.h"

namespace wb {

namespace {
constexpr size_t HEADER_SIZE = 48;
constexpr size_t DESC_SIZE = 96;

inline uint16_t rd16(const uint8_t* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) {
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
         (static_cast<uint32_t>(p[3]) << 24);
}
inline uint64_t rd64(const uint8_t* p) { return static_cast<uint64_t>(rd32(p)) | (static_cast<uint64_t>(rd32(p + 4)) << 32); }
inline float rdf32(const uint8_t* p) {
  uint32_t u = rd32(p);
  float f;
  memcpy(&f, &u, 4);
  return f;
}
}  // namespace

const AprTensor* AprFile::find(const std::string& name) const {
  for (const auto& t : tensors)
    if (t.name == name) return &t;
  return nullptr;
}

const uint8_t* AprFile::payload(const AprTensor& t, size_t* n_out) const {
  // every bound is checked by subtraction / division so that a hostile descriptor (n_elements >= 2^62, offset near 2^64) cannot
  // wrap: the reference answers "tensor data out of bounds" for all of them (format/mod.rs:610-628)
  if (data_offset > n_bytes || t.offset > n_bytes - data_offset) return nullptr;
  const uint64_t start = static_cast<uint64_t>(data_offset) + t.offset;
  const uint64_t avail = n_bytes - start;
  uint64_t need;
  switch (cfg.quantization) {
    case 2:                                             // int8: one byte per element (format/mod.rs:653-655)
      if (t.n_elements > avail) return nullptr;
      need = t.n_elements;
      break;
    case 3:                                             // int4 (defined extension): two per byte
      if (t.n_elements / 2 > avail || (t.n_elements + 1) / 2 > avail) return nullptr;
      need = (t.n_elements + 1) / 2;
      break;
    default:                                            // everything else is read as f32 (format/mod.rs:619-628)
      if (t.n_elements > avail / 4) return nullptr;
      need = t.n_elements * 4;
      break;
  }
  *n_out = static_cast<size_t>(need);
  return bytes + start;
}

int parse_apr(const uint8_t* bytes, size_t n, AprFile* f) {
  if (bytes == nullptr || n < 4 || memcmp(bytes, "APR1", 4) != 0) return set_error(WB_ERR_FORMAT, "invalid magic");
  if (n < 4 + HEADER_SIZE) return set_error(WB_ERR_FORMAT, "header too short");
  const uint8_t* h = bytes + 4;
  const uint16_t version = rd16(h);
  if (version > 1) return set_error(WB_ERR_FORMAT, "unsupported format version: " + std::to_string(version));
  const uint8_t quant = h[3];
  if (quant > 3) return set_error(WB_ERR_FORMAT, "invalid quantization type: " + std::to_string(quant));
  wb_config& c = f->cfg;
  c.model_type = h[2];
  c.quantization = quant;
  c.n_tensors = rd16(h + 5);
  const uint8_t flags = h[7];
  const bool has_vocab = flags & 1;
  c.has_filterbank = (flags & 2) ? 1 : 0;
  c.n_vocab = rd32(h + 8);
  c.n_audio_ctx = rd32(h + 12);
  c.n_audio_state = rd32(h + 16);
  c.n_audio_head = rd32(h + 20);
  c.n_audio_layer = rd32(h + 24);
  c.n_text_ctx = rd32(h + 28);
  c.n_text_state = rd32(h + 32);
  c.n_text_head = rd32(h + 36);
  c.n_text_layer = rd32(h + 40);
  c.n_mels = rd32(h + 44);
  f->bytes = bytes;
  f->n_bytes = n;

  const size_t nt = c.n_tensors;
  const size_t index_start = 4 + HEADER_SIZE;
  if (nt > 0 && n < index_start + nt * DESC_SIZE) return set_error(WB_ERR_FORMAT, "file too short for tensor index");
  const size_t scale_table = index_start + nt * DESC_SIZE;
  const bool quantised = (quant == 2 || quant == 3);
  f->data_offset = scale_table + (quantised ? nt * 4 : 0);
  f->tensors.resize(nt);
  uint64_t total = 0;
  for (size_t i = 0; i < nt; ++i) {
    const uint8_t* d = bytes + index_start + i * DESC_SIZE;
    AprTensor& t = f->tensors[i];
    size_t len = 0;
    while (len < 48 && d[len] != 0) ++len;
    t.name.assign(reinterpret_cast<const char*>(d), len);
    t.offset = rd64(d + 48);
    t.size = rd64(d + 56);
    t.n_elements = rd64(d + 64);
    for (int k = 0; k < 4; ++k) t.shape[k] = rd32(d + 72 + 4 * k);
    t.n_dims = d[88];
    if (quantised && scale_table + 4 * i + 4 <= n) t.scale = rdf32(bytes + scale_table + 4 * i);
    total += t.size;
  }
  // trailing sections: [vocab], [filterbank], crc32 (written, not verified by the reference: format/mod.rs:484-522)
  f->has_filterbank = false;
  if (c.has_filterbank) {
    uint64_t pos = static_cast<uint64_t>(f->data_offset) + total;
    bool ok = true;
    if (has_vocab) {
      if (pos + 4 > n) ok = false;
      else pos += 4ull + rd32(bytes + pos);
    }
    if (ok && pos + 4 <= n) {
      const uint32_t fsz = rd32(bytes + pos);
      const uint64_t body = pos + 4;
      if (body + fsz <= n && fsz >= 8) {
        const uint32_t nm = rd32(bytes + body), nf = rd32(bytes + body + 4);
        if (static_cast<uint64_t>(nm) * nf * 4 + 8 <= fsz) {
          f->has_filterbank = true;
          f->fb_mels = nm;
          f->fb_freqs = nf;
          f->fb_data = bytes + body + 8;
        }
      }
    }
  }
  return WB_OK;
}

std::vector<float> hann_window_periodic(int n) {
  std::vector<float> w(n);
  const float pi = 3.14159265358979323846f;
  for (int i = 0; i < n; ++i) w[i] = 0.5f * (1.0f - cosf(2.0f * pi * static_cast<float>(i) / static_cast<float>(n)));
  return w;
}

std::vector<float> htk_filterbank(int n_mels, int n_fft, int sample_rate) {
  const int n_freqs = n_fft / 2 + 1;
  std::vector<float> filt(static_cast<size_t>(n_mels) * n_freqs, 0.f);
  auto hz_to_mel = [](float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); };
  auto mel_to_hz = [](float mel) { return 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); };
  const float mel_min = hz_to_mel(0.0f), mel_max = hz_to_mel(static_cast<float>(sample_rate) / 2.0f);
  std::vector<long> bins(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) {
    const float mel = mel_min + (mel_max - mel_min) * static_cast<float>(i) / static_cast<float>(n_mels + 1);
    const float f = mel_to_hz(mel);
    bins[i] = static_cast<long>(floorf((static_cast<float>(n_fft) + 1.0f) * f / static_cast<float>(sample_rate)));
  }
  for (int m = 0; m < n_mels; ++m) {
    const long lo = bins[m], ce = bins[m + 1], hi = bins[m + 2];
    for (long k = lo; k < ce; ++k)
      if (k < n_freqs && ce > lo) filt[static_cast<size_t>(m) * n_freqs + k] = static_cast<float>(k - lo) / static_cast<float>(ce - lo);
    for (long k = ce; k < hi; ++k)
      if (k < n_freqs && hi > ce) filt[static_cast<size_t>(m) * n_freqs + k] = static_cast<float>(hi - k) / static_cast<float>(hi - ce);
  }
  return filt;
}

std::vector<float> default_positional_embedding(int max_len, int d_model) {
  std::vector<float> pe(static_cast<size_t>(max_len) * d_model, 0.f);
  for (int pos = 0; pos < max_len; ++pos)
    for (int i = 0; i < d_model / 2; ++i) {
      const float angle = static_cast<float>(pos) / powf(10000.0f, 2.0f * static_cast<float>(i) / static_cast<float>(d_model));
      pe[static_cast<size_t>(pos) * d_model + 2 * i] = sinf(angle);
      pe[static_cast<size_t>(pos) * d_model + 2 * i + 1] = cosf(angle);
    }
  return pe;
}

}  // namespace wb
