// The `.apr` uploader shared by loader.cu (encoder tensors) and decoder.cu (decoder tensors): see loader.cu's header comment.
#pragma once
#include "model.h"

namespace wb {

constexpr size_t PIECE = 64ull << 20;     // H2D granularity of the file image

int launch_fill_f32(float* dst, size_t n, float v, cudaStream_t s);
int launch_copy_f32_bytes(const uint8_t* src, float* dst, size_t n, cudaStream_t s);          // little-endian f32 bytes, any alignment
int launch_f32_bytes_to_op16(const uint8_t* src, op16* dst, size_t n, cudaStream_t s);

// ------------------------------------------------------------------------------------------------------------------
struct Uploader {
  Replica* m = nullptr;
  const AprFile* f = nullptr;
  DevBuf<uint8_t> image;              // device copy of file bytes [img_lo, img_hi)
  size_t img_lo = 0, img_hi = 0;
  std::vector<cudaEvent_t> piece_done;
  size_t waited_pieces = 0;           // the compute stream already waits for pieces [0, waited_pieces)

  ~Uploader() {
    for (cudaEvent_t e : piece_done)
      if (e) cudaEventDestroy(e);
  }

  // Enqueue the H2D of the tensor-data section.  `pinned` says the caller's buffer is page-locked (true DMA, non-blocking).
  int start() {
    uint64_t hi = f->data_offset;
    for (const auto& t : f->tensors) {
      size_t nb = 0;
      const uint8_t* p = f->payload(t, &nb);
      if (p) hi = std::max<uint64_t>(hi, static_cast<uint64_t>(p - f->bytes) + nb);
    }
    img_lo = f->data_offset & ~static_cast<size_t>(255);     // device alignment of a tensor == its file offset's alignment (mod 256)
    img_hi = static_cast<size_t>(hi);
    if (img_hi <= img_lo) return WB_OK;
    int rc = image.ensure(img_hi - img_lo + 16);
    if (rc != WB_OK) return rc;
    if (!m->in_stream) {
      WB_CUDA_OK(cudaStreamCreateWithFlags(&m->in_stream, cudaStreamNonBlocking));
      WB_CUDA_OK(cudaStreamCreateWithFlags(&m->out_stream, cudaStreamNonBlocking));
    }
    for (size_t off = img_lo; off < img_hi; off += PIECE) {
      const size_t nb = std::min(PIECE, img_hi - off);
      WB_CUDA_OK(cudaMemcpyAsync(image.p + (off - img_lo), f->bytes + off, nb, cudaMemcpyHostToDevice, m->in_stream));
      cudaEvent_t e = nullptr;
      WB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      piece_done.push_back(e);
      WB_CUDA_OK(cudaEventRecord(e, m->in_stream));
    }
    return WB_OK;
  }

  // Device pointer of the raw payload of `name`, clamped to `max_elems` elements; the compute stream is made to wait for the last
  // piece the payload touches.  *n_out = 0 if the tensor is absent or runs past the end of the file (the reference keeps its
  // default in both cases).
  int payload(const std::string& name, size_t max_elems, const uint8_t** dev, size_t* n_out, float* scale_out) {
    *n_out = 0;
    *scale_out = 1.f;
    *dev = nullptr;
    const AprTensor* t = f->find(name);
    if (!t) return WB_OK;
    size_t nbytes = 0;
    const uint8_t* src = f->payload(*t, &nbytes);
    if (!src) return WB_OK;
    const size_t n = std::min<size_t>(static_cast<size_t>(t->n_elements), max_elems);   // lib.rs:772-774 min-clamp
    if (n == 0) return WB_OK;
    const size_t off = static_cast<size_t>(src - f->bytes);
    const size_t used = f->cfg.quantization == 2 ? n : (f->cfg.quantization == 3 ? (n + 1) / 2 : n * 4);
    const size_t last_piece = (off + used - 1 - img_lo) / PIECE;
    while (waited_pieces <= last_piece && waited_pieces < piece_done.size()) {
      WB_CUDA_OK(cudaStreamWaitEvent(m->stream, piece_done[waited_pieces], 0));
      ++waited_pieces;
    }
    *dev = image.p + (off - img_lo);
    *n_out = n;
    *scale_out = t->scale;
    return WB_OK;
  }
  bool present(const std::string& name) const {
    const AprTensor* t = f->find(name);
    size_t nb = 0;
    return t && f->payload(*t, &nb);
  }

  // f32 parameter (bias / LayerNorm / embedding): dst pre-filled with the default.
  int load_f32(const std::string& name, float* dst, size_t count) {
    const uint8_t* src;
    size_t n;
    float s;
    int rc = payload(name, count, &src, &n, &s);
    if (rc != WB_OK || n == 0) return rc;
    switch (f->cfg.quantization) {
      case 2: return launch_i8_to_f32(reinterpret_cast<const int8_t*>(src), s, dst, n, m->stream);
      case 3: return launch_i4_to_f32(src, s, dst, n, m->stream);
      default: return launch_copy_f32_bytes(src, dst, n, m->stream);
    }
  }
  // Quantised GEMM weight: the payload bytes stay as they are (dst pre-zeroed; absent / short tensors keep zeros).
  int load_packed(const std::string& name, uint8_t* dst, size_t count, float* scale_out) {
    const uint8_t* src;
    size_t n;
    int rc = payload(name, count, &src, &n, scale_out);
    if (rc != WB_OK || n == 0) return rc;
    const size_t copy_bytes = f->cfg.quantization == 2 ? n : n / 2;     // a trailing odd nibble (never for these shapes) is dropped
    WB_CUDA_OK(cudaMemcpyAsync(dst, src, copy_bytes, cudaMemcpyDeviceToDevice, m->stream));
    return WB_OK;
  }
  // GEMM weight -> op16 (quantised payloads keep their integer value; *scale_out carries the per-tensor scale).
  int load_bf16(const std::string& name, op16* dst, size_t count, float* scale_out) {
    const uint8_t* src;
    size_t n;
    float s;
    *scale_out = 1.f;
    int rc = payload(name, count, &src, &n, &s);
    if (rc != WB_OK || n == 0) return rc;
    switch (f->cfg.quantization) {
      case 2: *scale_out = s; return launch_i8_to_op16(reinterpret_cast<const int8_t*>(src), dst, n, m->stream);
      case 3: *scale_out = s; return launch_i4_to_op16(src, dst, n, m->stream);
      default: return launch_f32_bytes_to_op16(src, dst, n, m->stream);
    }
  }
};


}  // namespace wb
