// `.apr` -> device.  Mirrors WhisperApr::load_from_apr's tensor walk (src/lib.rs:673-754), load_encoder_weights /
// load_{layer_norm,attention,ffn}_weights (src/lib.rs:757-841, 931-993) and load_decoder_weights (src/lib.rs:843-929): every
// tensor is looked up by name, clamped to the destination length, and silently keeps its default when absent or out of bounds.
//
// Upload path ("zero-copy streaming upload"): the caller's bytes are never copied on the host (the reference clones
// the whole file, format/mod.rs:484).  The tensor-data section goes to the device ONCE as an image, in 64 MB pieces on the copy
// stream (the caller's buffer is page-locked by wb_model_from_apr_devices for the duration, so the pieces are true DMA and the
// call does not block); every tensor is then cut out of the image by a conversion kernel on the compute stream that waits only for
// the piece holding its last byte.  No per-tensor synchronisation, no host-side staging vectors; the image is freed after the load.
#include "loader.h"

namespace wb {

namespace {

__global__ void fill_f32_kernel(float* __restrict__ dst, size_t n, float v) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}
__global__ void copy_f32_unaligned_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n) {
  // f32 payloads sit at 4-byte aligned file offsets in every writer of the reference; a foreign writer may not honour that
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const bool aligned = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  for (; i < n; i += stride) {
    if (aligned) {
      dst[i] = reinterpret_cast<const float*>(src)[i];
    } else {
      const uint8_t* p = src + 4 * i;
      const uint32_t u = p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24);
      dst[i] = __uint_as_float(u);
    }
  }
}
__global__ void f32_unaligned_to_bf16_kernel(const uint8_t* __restrict__ src, op16* __restrict__ dst, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    const uint8_t* p = src + 4 * i;
    const uint32_t u = p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24);
    dst[i] = float_to_op16(__uint_as_float(u));
  }
}
// per-channel symmetric int8 of a op16 weight matrix [rows][cols] (quantize_f32_to_i8_per_channel, model/quantized.rs:1769-1794 over
// quantize_f32_to_i8 :1732-1756): scale = absmax / 127 (1.0 when absmax < 1e-10), q = clamp(round(x / scale), -128, 127).
// One warp per row.
__global__ void __launch_bounds__(256) quant_i8_rows_kernel(const op16* __restrict__ w, int rows, int cols, int8_t* __restrict__ q,
                                                            float* __restrict__ scales) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const op16* wr = w + static_cast<size_t>(row) * cols;
  float mx = 0.f;
  for (int i = lane; i < cols; i += 32) mx = fmaxf(mx, fabsf(op16_to_float(wr[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float scale = mx < 1e-10f ? 1.0f : mx / 127.0f;
  for (int i = lane; i < cols; i += 32) {
    const float v = roundf(op16_to_float(wr[i]) / scale);          // f32::round: half away from zero, as roundf
    q[static_cast<size_t>(row) * cols + i] = static_cast<int8_t>(fminf(fmaxf(v, -128.f), 127.f));
  }
  if (lane == 0) scales[row] = scale;
}

inline unsigned grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  return static_cast<unsigned>(g > 148 * 16 ? 148 * 16 : (g == 0 ? 1 : g));
}

}  // namespace

int launch_copy_f32_bytes(const uint8_t* src, float* dst, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  copy_f32_unaligned_kernel<<<grid_for(n), 256, 0, s>>>(src, dst, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_f32_bytes_to_op16(const uint8_t* src, op16* dst, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) return launch_f32_to_op16(reinterpret_cast<const float*>(src), dst, n, s);
  f32_unaligned_to_bf16_kernel<<<grid_for(n), 256, 0, s>>>(src, dst, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_fill_f32(float* dst, size_t n, float v, cudaStream_t s) {
  if (n == 0) return WB_OK;
  fill_f32_kernel<<<grid_for(n), 256, 0, s>>>(dst, n, v);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

namespace {

int new_f32_param(Replica* m, Uploader& up, const std::string& name, size_t count, float dflt, float** out) {
  int rc = dev_alloc(m, count, out);
  if (rc != WB_OK) return rc;
  if ((rc = launch_fill_f32(*out, count, dflt, m->stream)) != WB_OK) return rc;
  return up.load_f32(name, *out, count);
}

int new_weight(Replica* m, Uploader& up, const std::string& name, size_t count, op16** out, float* scale) {
  int rc = dev_alloc(m, count, out);
  if (rc != WB_OK) return rc;
  WB_CUDA_OK(cudaMemsetAsync(*out, 0, count * sizeof(op16), m->stream));
  return up.load_bf16(name, *out, count, scale);
}

int build_mel_tables(Replica* m, const AprFile& f) {
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

int load_encoder(Replica* m, const AprFile& f, Uploader& up) {
  const size_t d = m->cfg.n_audio_state, nm = m->cfg.n_mels, L = m->cfg.n_audio_layer, ctx = m->cfg.n_audio_ctx;
  int rc = WB_OK;
  const bool quant = f.cfg.quantization == 2 || f.cfg.quantization == 3;
  cudaStream_t st = m->stream;

  // conv stem: [out][in][3] -> op16 -> [out][3][in]
  {
    DevBuf<op16> t1, t2;
    if ((rc = t1.ensure(d * nm * 3)) != WB_OK || (rc = t2.ensure(d * d * 3)) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemsetAsync(t1.p, 0, d * nm * 3 * sizeof(op16), st));
    WB_CUDA_OK(cudaMemsetAsync(t2.p, 0, d * d * 3 * sizeof(op16), st));
    if ((rc = up.load_bf16("encoder.conv1.weight", t1.p, d * nm * 3, &m->conv1_s)) != WB_OK) return rc;
    if ((rc = dev_alloc(m, d * nm * 3, &m->conv1_w)) != WB_OK) return rc;
    if ((rc = launch_conv_repack(t1.p, m->conv1_w, static_cast<int>(d), static_cast<int>(nm), st)) != WB_OK) return rc;
    if ((rc = up.load_bf16("encoder.conv2.weight", t2.p, d * d * 3, &m->conv2_s)) != WB_OK) return rc;
    if ((rc = dev_alloc(m, d * d * 3, &m->conv2_w)) != WB_OK) return rc;
    if ((rc = launch_conv_repack(t2.p, m->conv2_w, static_cast<int>(d), static_cast<int>(d), st)) != WB_OK) return rc;
    WB_CUDA_OK(cudaStreamSynchronize(st));        // t1 / t2 go out of scope (3 MB + 10 MB; the only wait of the load besides the last)
  }
  if ((rc = new_f32_param(m, up, "encoder.conv1.bias", d, 0.f, &m->conv1_b)) != WB_OK) return rc;
  if ((rc = new_f32_param(m, up, "encoder.conv2.bias", d, 0.f, &m->conv2_b)) != WB_OK) return rc;

  // positional embedding: embed_positions.weight, else positional_embedding, else the default table (lib.rs:793-800)
  {
    if ((rc = dev_alloc(m, ctx * d, &m->pe)) != WB_OK) return rc;
    const char* name = up.present("encoder.embed_positions.weight") ? "encoder.embed_positions.weight" : "encoder.positional_embedding";
    const AprTensor* t = f.find(name);
    if (!(up.present(name) && t && t->n_elements >= ctx * d)) {    // the table is needed (in part): built on the host like the reference's
      const std::vector<float> pe = default_positional_embedding(static_cast<int>(ctx), static_cast<int>(d));
      WB_CUDA_OK(cudaMemcpyAsync(m->pe, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice, st));
      WB_CUDA_OK(cudaStreamSynchronize(st));
    }
    if ((rc = up.load_f32(name, m->pe, ctx * d)) != WB_OK) return rc;
  }

  m->quant = quant ? static_cast<int>(f.cfg.quantization) : 0;
  if (quant) {
    if ((rc = dev_alloc(m, 3 * d * d, &m->xp_qkv)) != WB_OK || (rc = dev_alloc(m, d * d, &m->xp_o)) != WB_OK ||
        (rc = dev_alloc(m, 4 * d * d, &m->xp_1)) != WB_OK || (rc = dev_alloc(m, 4 * d * d, &m->xp_2)) != WB_OK)
      return rc;
  }
  m->layers.resize(L);
  const char* proj[3] = {".self_attn.q_proj", ".self_attn.k_proj", ".self_attn.v_proj"};
  for (size_t i = 0; i < L; ++i) {
    LayerW& w = m->layers[i];
    const std::string p = "encoder.layers." + std::to_string(i);
    if ((rc = new_f32_param(m, up, p + ".self_attn_layer_norm.weight", d, 1.f, &w.ln1_g)) != WB_OK) return rc;
    if ((rc = new_f32_param(m, up, p + ".self_attn_layer_norm.bias", d, 0.f, &w.ln1_b)) != WB_OK) return rc;
    if ((rc = new_f32_param(m, up, p + ".final_layer_norm.weight", d, 1.f, &w.ln2_g)) != WB_OK) return rc;
    if ((rc = new_f32_param(m, up, p + ".final_layer_norm.bias", d, 0.f, &w.ln2_b)) != WB_OK) return rc;
    // fused QKV: rows [0,d) = q_proj, [d,2d) = k_proj, [2d,3d) = v_proj (three separate GEMMs in attention.rs:912-914)
    if ((rc = dev_alloc(m, 3 * d, &w.bqkv)) != WB_OK) return rc;
    if ((rc = launch_fill_f32(w.bqkv, 3 * d, 0.f, st)) != WB_OK) return rc;
    float sc[3] = {1.f, 1.f, 1.f};
    if (quant) {
      // the packed bytes stay as they are in HBM; every layer shares one set of op16 expansion buffers
      const size_t qb = f.cfg.quantization == 2 ? d * d : d * d / 2;        // bytes of one d x d tensor (d is even)
      if ((rc = dev_alloc(m, 3 * qb, &w.pqkv)) != WB_OK || (rc = dev_alloc(m, qb, &w.po)) != WB_OK ||
          (rc = dev_alloc(m, 4 * qb, &w.p1)) != WB_OK || (rc = dev_alloc(m, 4 * qb, &w.p2)) != WB_OK)
        return rc;
      WB_CUDA_OK(cudaMemsetAsync(w.pqkv, 0, 3 * qb, st));
      WB_CUDA_OK(cudaMemsetAsync(w.po, 0, qb, st));
      WB_CUDA_OK(cudaMemsetAsync(w.p1, 0, 4 * qb, st));
      WB_CUDA_OK(cudaMemsetAsync(w.p2, 0, 4 * qb, st));
      w.wqkv = m->xp_qkv; w.wo = m->xp_o; w.w1 = m->xp_1; w.w2 = m->xp_2;
      for (int k = 0; k < 3; ++k) {
        if ((rc = up.load_packed(p + proj[k] + ".weight", w.pqkv + k * qb, d * d, &sc[k])) != WB_OK) return rc;
        if ((rc = up.load_f32(p + proj[k] + ".bias", w.bqkv + k * d, d)) != WB_OK) return rc;
      }
      if ((rc = up.load_packed(p + ".self_attn.out_proj.weight", w.po, d * d, &w.so)) != WB_OK) return rc;
      if ((rc = up.load_packed(p + ".fc1.weight", w.p1, 4 * d * d, &w.s1)) != WB_OK) return rc;
      if ((rc = up.load_packed(p + ".fc2.weight", w.p2, 4 * d * d, &w.s2)) != WB_OK) return rc;
    } else {
      if ((rc = dev_alloc(m, 3 * d * d, &w.wqkv)) != WB_OK) return rc;
      WB_CUDA_OK(cudaMemsetAsync(w.wqkv, 0, 3 * d * d * sizeof(op16), st));
      for (int k = 0; k < 3; ++k) {
        if ((rc = up.load_bf16(p + proj[k] + ".weight", w.wqkv + k * d * d, d * d, &sc[k])) != WB_OK) return rc;
        if ((rc = up.load_f32(p + proj[k] + ".bias", w.bqkv + k * d, d)) != WB_OK) return rc;
      }
      if ((rc = new_weight(m, up, p + ".self_attn.out_proj.weight", d * d, &w.wo, &w.so)) != WB_OK) return rc;
      if ((rc = new_weight(m, up, p + ".fc1.weight", 4 * d * d, &w.w1, &w.s1)) != WB_OK) return rc;
      if ((rc = new_weight(m, up, p + ".fc2.weight", 4 * d * d, &w.w2, &w.s2)) != WB_OK) return rc;
    }
    if (quant) {                                   // per-column scale vector of the fused QKV GEMM: three per-tensor scales
      if ((rc = dev_alloc(m, 3 * d, &w.sqkv)) != WB_OK) return rc;
      for (int k = 0; k < 3; ++k)
        if ((rc = launch_fill_f32(w.sqkv + k * d, d, sc[k], st)) != WB_OK) return rc;
    }
    if ((rc = new_f32_param(m, up, p + ".self_attn.out_proj.bias", d, 0.f, &w.bo)) != WB_OK) return rc;
    if ((rc = new_f32_param(m, up, p + ".fc1.bias", 4 * d, 0.f, &w.b1)) != WB_OK) return rc;
    if ((rc = new_f32_param(m, up, p + ".fc2.bias", d, 0.f, &w.b2)) != WB_OK) return rc;
  }
  if ((rc = new_f32_param(m, up, "encoder.layer_norm.weight", d, 1.f, &m->lnp_g)) != WB_OK) return rc;
  if ((rc = new_f32_param(m, up, "encoder.layer_norm.bias", d, 0.f, &m->lnp_b)) != WB_OK) return rc;
  return WB_OK;
}

}  // namespace

// filters [n_mels][201] -> device tables (dense rows + the first/last non-zero bin of every row + the periodic Hann window)
int upload_mel_tables(Replica* m, const std::vector<float>& filt, int n_mels, MelTables* out) {
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
  out->packed_lo = nullptr;
  out->packed_len = nullptr;
  out->nnz = 0;
  std::vector<int> off(n_mels, 0);
  std::vector<float> packed;
  std::vector<int> plo(n_mels, 0), plen(n_mels, 0);
  for (int j = 0; j < n_mels; ++j) {                      // every span padded with zero weights to a multiple of 4 (16 B loads)
    off[j] = static_cast<int>(packed.size());             // a multiple of 4: every span starts 16-byte aligned
    if (len[j] == 0) continue;
    plo[j] = lo[j];
    plen[j] = (len[j] + 3) & ~3;                          // lo + plen <= 204: the kernel keeps bins 201..203 of every frame at zero
    for (int k = 0; k < plen[j]; ++k) packed.push_back(k < len[j] ? filt[static_cast<size_t>(j) * N_FREQ + lo[j] + k] : 0.0f);
  }
  if (n_mels <= 256 && packed.size() <= 2048 && !packed.empty()) {
    float* d_packed;
    int* d_off;
    if ((rc = dev_alloc(m, packed.size(), &d_packed)) != WB_OK) return rc;
    if ((rc = dev_alloc(m, off.size(), &d_off)) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpy(d_packed, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
    WB_CUDA_OK(cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    int *d_plo, *d_plen;
    if ((rc = dev_alloc(m, plo.size(), &d_plo)) != WB_OK || (rc = dev_alloc(m, plen.size(), &d_plen)) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpy(d_plo, plo.data(), plo.size() * 4, cudaMemcpyHostToDevice));
    WB_CUDA_OK(cudaMemcpy(d_plen, plen.data(), plen.size() * 4, cudaMemcpyHostToDevice));
    out->packed = d_packed;
    out->span_off = d_off;
    out->packed_lo = d_plo;
    out->packed_len = d_plen;
    out->nnz = static_cast<int>(packed.size());
  }
  return WB_OK;
}

int load_replica(Replica* m, const AprFile& f, const uint8_t* /*pinned_base*/) {
  DeviceGuard guard(m->device);
  m->cfg = f.cfg;
  m->use_graphs = getenv("WB_NO_GRAPH") == nullptr;       // A/B switch: plain launches instead of graph replay
  m->attn_bf16 = getenv("WB_ATTN_FP16") == nullptr;      // A/B switch: attention operands in the build's operand format instead of bf16
  m->zigzag = getenv("WB_NO_ZIGZAG") == nullptr;         // A/B switch: alternate the row direction from kernel to kernel (L2 reuse)
  m->ln_follow = getenv("WB_LN_FOLLOW") != nullptr;       // experiment switch (default off): LayerNorm as a concurrent follower of the residual GEMMs
  if (cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess) return set_error(WB_ERR_CUDA, "cudaStreamCreate failed");
  m->stream = m->own_stream;
  WB_CUDA_OK(cudaEventCreateWithFlags(&m->done_event, cudaEventDisableTiming));
  int rc;
  if ((rc = gemm_init()) != WB_OK) return rc;
  if ((rc = attention_init()) != WB_OK) return rc;
  if ((rc = mel_init()) != WB_OK) return rc;
  Uploader up;
  up.m = m;
  up.f = &f;
  if ((rc = up.start()) != WB_OK) return rc;               // the H2D of the whole tensor section is in flight from here on
  if ((rc = build_mel_tables(m, f)) != WB_OK) return rc;
  if (m->cfg.n_mels == 0) m->cfg.n_mels = m->mel.n_mels;
  if ((rc = load_encoder(m, f, up)) != WB_OK) return rc;
  if ((rc = load_decoder(m, f, up)) != WB_OK) return rc;
  // the image is released when `up` goes out of scope: everything that reads it must have run
  if (cudaStreamSynchronize(m->stream) != cudaSuccess || cudaStreamSynchronize(m->in_stream) != cudaSuccess)
    return set_error(WB_ERR_CUDA, std::string("model upload failed: ") + cudaGetErrorString(cudaGetLastError()));
  return WB_OK;
}

// On-device requantisation of resident op16 weights to per-channel int8 (model/quantized.rs:1769-1813): one scale per output row
// of every linear weight; the packed int8 rows replace the op16 matrices in HBM and are expanded per layer like `.apr` int8 payloads,
// the per-row scale is applied per output column in the GEMM epilogue.
int requantize_int8_per_channel(Replica* m) {
  if (m->quant != 0) return set_error(WB_ERR_MODEL, "per-channel requantisation needs a model loaded from f32 payloads");
  DeviceGuard guard(m->device);
  const size_t d = m->cfg.n_audio_state;
  int rc;
  cudaStream_t st = m->stream;
  if ((rc = dev_alloc(m, 3 * d * d, &m->xp_qkv)) != WB_OK || (rc = dev_alloc(m, d * d, &m->xp_o)) != WB_OK ||
      (rc = dev_alloc(m, 4 * d * d, &m->xp_1)) != WB_OK || (rc = dev_alloc(m, 4 * d * d, &m->xp_2)) != WB_OK)
    return rc;
  std::vector<void*> old;
  auto quant_rows = [&](const op16* w, int rows, int cols, uint8_t** q, float** scales) -> int {
    int r;
    if ((r = dev_alloc(m, static_cast<size_t>(rows) * cols, q)) != WB_OK || (r = dev_alloc(m, rows, scales)) != WB_OK) return r;
    quant_i8_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(w, rows, cols, reinterpret_cast<int8_t*>(*q), *scales);
    count_launch();
    WB_CUDA_OK(cudaGetLastError());
    return WB_OK;
  };
  for (LayerW& w : m->layers) {
    const int di = static_cast<int>(d);
    if ((rc = quant_rows(w.wqkv, 3 * di, di, &w.pqkv, &w.sqkv)) != WB_OK) return rc;
    if ((rc = quant_rows(w.wo, di, di, &w.po, &w.cso)) != WB_OK) return rc;
    if ((rc = quant_rows(w.w1, 4 * di, di, &w.p1, &w.cs1)) != WB_OK) return rc;
    if ((rc = quant_rows(w.w2, di, 4 * di, &w.p2, &w.cs2)) != WB_OK) return rc;
    old.push_back(w.wqkv); old.push_back(w.wo); old.push_back(w.w1); old.push_back(w.w2);
    w.wqkv = m->xp_qkv; w.wo = m->xp_o; w.w1 = m->xp_1; w.w2 = m->xp_2;
    w.so = w.s1 = w.s2 = 1.f;
  }
  WB_CUDA_OK(cudaStreamSynchronize(st));
  for (void* p : old) {                                  // the op16 matrices are gone from HBM: 2 B -> 1 B per weight
    auto it = std::find(m->allocs.begin(), m->allocs.end(), p);
    if (it != m->allocs.end()) m->allocs.erase(it);
    cudaFree(p);
  }
  m->quant = 2;
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  return WB_OK;
}

void free_replica(Replica* m) {
  if (!m) return;
  DeviceGuard guard(m->device);
  cudaDeviceSynchronize();
  free_decode_state(m);
  for (void* p : m->allocs) cudaFree(p);
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
  for (auto& sl : m->slot) {
    if (sl.h_n_valid) cudaFreeHost(sl.h_n_valid);
    if (sl.in_done) cudaEventDestroy(sl.in_done);
    if (sl.compute_done) cudaEventDestroy(sl.compute_done);
    if (sl.out_done) cudaEventDestroy(sl.out_done);
  }
  if (m->done_event) cudaEventDestroy(m->done_event);
  if (m->ln_fork) cudaEventDestroy(m->ln_fork);
  if (m->ln_join) cudaEventDestroy(m->ln_join);
  if (m->ln_stream) cudaStreamDestroy(m->ln_stream);
  if (m->in_stream) cudaStreamDestroy(m->in_stream);
  if (m->out_stream) cudaStreamDestroy(m->out_stream);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;                              // DevBuf members (workspace, slots) release their memory
}

}  // namespace wb
