// The device pipeline of one replica: workspace, the mel and encoder launch sequences, the fused step (CUDA-graph replay) and the
// copy/compute overlapped micro-batch queue behind the host-buffer entry points.
//
// Mirrors, for the mel + encoder path only:
//   WhisperApr::{compute_mel, encode}                        src/lib.rs:407-449
//   Encoder::{forward, forward_mel, forward_batch{,_padded}} src/model/encoder.rs:450-478, 566-660
//   EncoderBlock::forward                                    src/model/encoder.rs:346-361
//   transcribe_batch_optimized steps 1-2                     src/lib.rs:1162-1170
//   BatchPreprocessor::process_batch                         src/audio/batch.rs:157-176
#include "model.h"

namespace wb {

// ---- workspace ---------------------------------------------------------------------------------
int ensure_workspace(Replica* m, int B) {
  Workspace& w = m->ws;
  if (B <= w.cap) return WB_OK;
  // a mel of T frames gives S = (T - 1) / 2 + 1 <= n_audio_ctx positions (validate_mel_len), so T <= 2 * ctx; the fused paths
  // always run T = 3000 / S = 1500 and are refused up front when the header's n_audio_ctx is smaller (check_fused_dims)
  const size_t d = m->cfg.n_audio_state, nm = std::max<size_t>(m->cfg.n_mels, m->mel.n_mels);
  const size_t S = std::max<size_t>(m->cfg.n_audio_ctx, (N_FRAMES_30S - 1) / 2 + 1);
  const size_t T = std::max<size_t>(N_FRAMES_30S, 2 * static_cast<size_t>(m->cfg.n_audio_ctx)), b = static_cast<size_t>(B);
  int rc;
  if ((rc = w.audio.ensure(b * N_SAMPLES_30S)) != WB_OK) return rc;
  if ((rc = w.n_valid.ensure(b)) != WB_OK) return rc;
  if ((rc = w.max_key.ensure(b)) != WB_OK) return rc;
  if ((rc = w.mel_done.ensure(b)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemsetAsync(w.mel_done.p, 0, b * sizeof(unsigned int), m->stream));
  if ((rc = launch_mel_init_keys(w.max_key.p, B, m->stream)) != WB_OK) return rc;    // armed once; mel_finalize re-arms after every use
  if ((rc = w.logmel.ensure(b * T * nm)) != WB_OK) return rc;
  if ((rc = w.mel_f32.ensure(b * T * nm)) != WB_OK) return rc;
  if ((rc = w.mel_bf16.ensure(b * (T + 2) * nm)) != WB_OK) return rc;
  if ((rc = w.c1.ensure(b * (T + 2) * d)) != WB_OK) return rc;
  if ((rc = w.x.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.xn.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.ln_ready.ensure(b * S / 32 + 2)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemsetAsync(w.ln_ready.p, 0, (b * S / 32 + 2) * sizeof(unsigned int), m->stream));
  if ((rc = w.qkv.ensure(b * S * 3 * d)) != WB_OK) return rc;
  if ((rc = w.att.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.hid.ensure(b * S * 4 * d)) != WB_OK) return rc;
  if ((rc = w.out_f32.ensure(b * S * d)) != WB_OK) return rc;
  if ((rc = w.out_bf16.ensure(b * S * d)) != WB_OK) return rc;
  w.cap = B;
  w.guard_T = -1;
  for (auto& g : m->graphs)                 // workspace pointers are baked into captured launches
    if (g.exec) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
  return WB_OK;
}

// ---- encoder launch sequence ------------------------------------------------------------------------
// Precondition: ws.mel_bf16 holds [B][T+2][n_mels] op16 with rows 1..T = mel frames and zero guard rows.
// d_out: [B][S][d] f32 or op16 (device).  n_layers < 0 -> all layers; ln_post as Encoder::forward (encoder.rs:477).
int encode_device(Replica* m, int B, int T, void* d_out, wb_dtype out_dtype, int n_layers, bool ln_post) {
  const int d = static_cast<int>(m->cfg.n_audio_state), nm = static_cast<int>(m->cfg.n_mels);
  const int H = static_cast<int>(m->cfg.n_audio_head);
  const int S = (T - 1) / 2 + 1;               // conv2: (T + 2 - 3) / 2 + 1 (encoder.rs:79)
  Workspace& w = m->ws;
  cudaStream_t st = m->stream;
  int rc;
  NvtxRange nvtx("step_g_encode");            // the reference's renacer span of the encoder stage (.renacer.toml:11-32)
  const int L = n_layers < 0 ? static_cast<int>(m->layers.size()) : std::min<int>(n_layers, static_cast<int>(m->layers.size()));

  // c1 guard rows (0 and T+1 of every batch entry) are zero: conv2's padding.  conv1 only ever writes rows 1..T, so they stay zero
  // from one step to the next; they are re-laid only when T (the row pitch of the batch entries) changes.
  if (w.guard_T != T) {
    WB_CUDA_OK(cudaMemsetAsync(w.c1.p, 0, w.c1.n * sizeof(op16), st));
    w.guard_T = T;
  }

  GemmDesc g{};
  // conv1 + GELU: row t of the operand = padded frames t, t+1, t+2 (3*nm contiguous values)
  g.A = w.mel_bf16.p; g.a_row_stride = nm; g.a_batch_stride = static_cast<long long>(T + 2) * nm;
  g.rows_per_batch = T; g.n_batch = B;
  g.W = m->conv1_w; g.N = d; g.K = 3 * nm;
  g.epilogue = EPI_GELU_BF16; g.alpha = m->conv1_s; g.col_scale = nullptr; g.bias = m->conv1_b;
  g.out = w.c1.p; g.ldc = d; g.out_rows_per_batch = T + 2; g.out_row_off = 1; g.pe = nullptr;
  WB_PROF(PC_GEMM, launch_gemm(g, st));
  // conv2 (stride 2) + GELU + positional embedding -> fp32 residual stream
  g.A = w.c1.p; g.a_row_stride = 2LL * d; g.a_batch_stride = static_cast<long long>(T + 2) * d;
  g.rows_per_batch = S; g.n_batch = B;
  g.W = m->conv2_w; g.N = d; g.K = 3 * d;
  g.epilogue = EPI_GELU_PE_F32; g.alpha = m->conv2_s; g.bias = m->conv2_b;
  g.out = w.x.p; g.ldc = d; g.out_rows_per_batch = S; g.out_row_off = 0; g.pe = m->pe;
  WB_PROF(PC_GEMM, launch_gemm(g, st));

  const int M = B * S;
  // Direction of travel over the rows, flipped after every kernel of a layer (`zig()`): a kernel that walks its rows in the opposite
  // direction of its producer starts on what the producer wrote last -- the part of a 123-491 MB activation that is still in the
  // 126 MB L2 -- instead of on what was evicted first.  WB_NO_ZIGZAG=1: every kernel ascending (A/B switch).
  bool dir = false;
  auto zig = [&]() -> bool {
    const bool d0 = dir;
    if (m->zigzag) dir = !dir;
    return d0;
  };
  bool qkv_out_bf16 = false;
  auto flat = [&](const op16* A, int K, const op16* W, int N, int epi, float alpha, const float* cs, const float* bias, void* out) {
    GemmDesc q{};
    q.reverse = zig();
    q.out_bf16 = qkv_out_bf16;
    q.A = A; q.a_row_stride = K; q.a_batch_stride = static_cast<long long>(M) * K; q.rows_per_batch = M; q.n_batch = 1;
    q.W = W; q.N = N; q.K = K; q.epilogue = epi; q.alpha = alpha; q.col_scale = cs; q.bias = bias;
    q.out = out; q.ldc = N; q.out_rows_per_batch = M; q.out_row_off = 0; q.pe = nullptr;
    return launch_gemm(q, st);
  };
  // Quantised models: a layer's packed weights are expanded to op16 (exact integer values; the scale stays in the GEMM epilogue)
  // right before the GEMM that consumes them, into buffers every layer reuses.  With M = B x 1500 rows per launch each weight tile
  // is consumed by ~190 row tiles, so expanding once per launch costs 1/190th of converting inside every CTA, and the 12 d^2
  // op16 (39 MB for d = 1280) stay L2-resident for the GEMM that follows; HBM only ever holds the packed bytes.
  const size_t dd = static_cast<size_t>(d) * d;
  auto expand = [&](const uint8_t* packed, op16* dst, size_t n) {
    return m->quant == 2 ? launch_i8_to_op16(reinterpret_cast<const int8_t*>(packed), dst, n, st) : launch_i4_to_op16(packed, dst, n, st);
  };
  // EXPERIMENT (off by default, WB_LN_FOLLOW=1): a residual GEMM (x += ...) and the LayerNorm of its result run SIDE BY SIDE: the
  // LayerNorm kernel is put on a second stream before the GEMM is launched, polls the GEMM's per-32-row completion counters and
  // normalises every row group out of L2 as soon as its last column tile has been reduced -- no HBM re-read of the residual stream, no
  // LayerNorm time on the critical path beyond the last group.  It works (bit-identical states, 64 fewer serial launches) and it is
  // 1-2 % SLOWER on the large-v3 step (profiles/r02w_ln_follow_ab.txt): the step is pinned at the 1 kW power cap, so hiding a low-power
  // HBM-bound kernel under a tensor-bound one frees no energy, and the follower's polling and L2 traffic cost the GEMM a little.
  // Off while the per-kernel profile is on and for widths the follower has no instantiation for.
  const bool follow = m->ln_follow && !m->prof_on && d % 128 == 0 && d <= 1280 && M >= 32;
  if (follow && !m->ln_stream) {
    WB_CUDA_OK(cudaStreamCreateWithFlags(&m->ln_stream, cudaStreamNonBlocking));
    WB_CUDA_OK(cudaEventCreateWithFlags(&m->ln_fork, cudaEventDisableTiming));
    WB_CUDA_OK(cudaEventCreateWithFlags(&m->ln_join, cudaEventDisableTiming));
  }
  // x += A . W^T (+ bias), then xn = LayerNorm(x; g, b) -- fused in time when `follow`
  auto resid_then_ln = [&](const op16* A, int K, const op16* W, float alpha, const float* cs, const float* bias, const float* g, const float* bta) -> int {
    GemmDesc q{};
    q.A = A; q.a_row_stride = K; q.a_batch_stride = static_cast<long long>(M) * K; q.rows_per_batch = M; q.n_batch = 1;
    q.W = W; q.N = d; q.K = K; q.epilogue = EPI_RESID_F32; q.alpha = alpha; q.col_scale = cs; q.bias = bias;
    q.out = w.x.p; q.ldc = d; q.out_rows_per_batch = M; q.out_row_off = 0; q.pe = nullptr;
    q.reverse = follow ? false : zig();
    if (g == nullptr) {                                                             // no LayerNorm behind this one
      WB_PROF(PC_GEMM, launch_gemm(q, st));
      return WB_OK;
    }
    if (!follow) {
      WB_PROF(PC_GEMM, launch_gemm(q, st));
      WB_PROF(PC_LAYERNORM, launch_layernorm(w.x.p, g, bta, M, d, w.xn.p, false, nullptr, st, zig()));
      return WB_OK;
    }
    q.ready = w.ln_ready.p;
    WB_CUDA_OK(cudaEventRecord(m->ln_fork, st));
    WB_CUDA_OK(cudaStreamWaitEvent(m->ln_stream, m->ln_fork, 0));
    const unsigned int need = static_cast<unsigned int>(gemm_tiles_n(d));
    int r = launch_layernorm_follow(w.x.p, g, bta, M, d, w.xn.p, w.ln_ready.p, need, true, m->ln_stream);
    if (r != WB_OK) return r;
    WB_CUDA_OK(cudaEventRecord(m->ln_join, m->ln_stream));
    r = launch_gemm(q, st);
    WB_CUDA_OK(cudaStreamWaitEvent(st, m->ln_join, 0));
    if (r != WB_OK) return r;
    return launch_layernorm_follow(w.x.p, g, bta, M, d, w.xn.p, w.ln_ready.p, need, false, st);      // sweep: normally finds nothing
  };
  for (int i = 0; i < L; ++i) {
    const LayerW& lw = m->layers[i];
    if (i == 0) { WB_PROF(PC_LAYERNORM, launch_layernorm(w.x.p, lw.ln1_g, lw.ln1_b, M, d, w.xn.p, false, nullptr, st, zig())); }
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.pqkv, lw.wqkv, 3 * dd)); }
    qkv_out_bf16 = m->attn_bf16;                            // q, k, v leave the QKV product as bf16: the attention's two products involve no weights
    WB_PROF(PC_GEMM, flat(w.xn.p, d, lw.wqkv, 3 * d, EPI_BF16, 1.f, lw.sqkv, lw.bqkv, w.qkv.p));
    qkv_out_bf16 = false;
    WB_PROF(PC_ATTENTION, launch_attention(w.qkv.p, w.att.p, B, S, d, H, st, zig(), m->attn_bf16));
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.po, lw.wo, dd)); }
    if ((rc = resid_then_ln(w.att.p, d, lw.wo, lw.so, lw.cso, lw.bo, lw.ln2_g, lw.ln2_b)) != WB_OK) return rc;
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.p1, lw.w1, 4 * dd)); }
    WB_PROF(PC_GEMM, flat(w.xn.p, d, lw.w1, 4 * d, EPI_GELU_BF16, lw.s1, lw.cs1, lw.b1, w.hid.p));
    if (m->quant) { WB_PROF(PC_OTHER, expand(lw.p2, lw.w2, 4 * dd)); }
    // fc2's result is normalised for the NEXT layer's attention (its ln1); the last layer is followed by ln_post below
    const bool next = i + 1 < L;
    if ((rc = resid_then_ln(w.hid.p, 4 * d, lw.w2, lw.s2, lw.cs2, lw.b2, next ? m->layers[i + 1].ln1_g : nullptr, next ? m->layers[i + 1].ln1_b : nullptr)) != WB_OK)
      return rc;
  }
  if (ln_post) {
    // 16-bit states: bf16 when the caller asked for WB_BF16 (a user-facing format), the operand format when they feed the decoder's
    // cross-attention K/V GEMM (WB_OP16, internal)
    WB_PROF(PC_LAYERNORM, launch_layernorm(w.x.p, m->lnp_g, m->lnp_b, M, d, out_dtype != WB_F32 ? d_out : nullptr, out_dtype == WB_BF16,
                                           out_dtype == WB_F32 ? static_cast<float*>(d_out) : nullptr, st, zig()));
  } else {
    if (out_dtype == WB_BF16) return set_error(WB_ERR_MODEL, "truncated encodes are f32 only");
    if (out_dtype == WB_OP16) rc = launch_f32_to_op16(w.x.p, static_cast<op16*>(d_out), static_cast<size_t>(M) * d, st);
    else {
      WB_CUDA_OK(cudaMemcpyAsync(d_out, w.x.p, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

// mel of B chunks already in ws.audio ([B][480000], n_valid per chunk in ws.n_valid) -> optional f32 [B][3000][m] and/or
// the op16 padded operand in ws.mel_bf16.
int mel_device(Replica* m, const float* d_audio, long long audio_stride, const long long* d_seg_off, const int* d_n_valid, int B,
               float* d_mel_f32, bool want_bf16) {
  Workspace& w = m->ws;
  const int nm = m->mel.n_mels;
  const int n_frames = (N_SAMPLES_30S - N_FFT) / HOP + 1;     // 2998 (mel.rs:245-249)
  int rc;
  NvtxRange nvtx("step_f_mel");               // the reference's span name for MelFilterbank::compute (src/audio/mel.rs:234)
  MelBatch job;
  job.audio = d_audio; job.audio_stride = audio_stride; job.seg_off = d_seg_off; job.n_valid = d_n_valid; job.n_valid_all = N_SAMPLES_30S;
  job.hop = HOP; job.B = B; job.n_frames = n_frames;
  WB_PROF(PC_MEL_STFT, launch_mel_stft(job, m->mel, w.logmel.p, w.max_key.p, m->stream));
  WB_PROF(PC_MEL_FINALIZE, launch_mel_finalize(w.logmel.p, w.max_key.p, w.mel_done.p, n_frames, N_FRAMES_30S, nm, B, d_mel_f32,
                                               want_bf16 ? w.mel_bf16.p : nullptr, m->stream));
  return WB_OK;
}

int check_encoder_dims(const Replica* m) {
  const wb_config& c = m->cfg;
  if (c.n_audio_state == 0 || c.n_audio_state % 128 != 0 || c.n_audio_head == 0 || c.n_audio_state != c.n_audio_head * 64)
    return set_error(WB_ERR_MODEL, "encoder kernels need n_audio_state % 128 == 0 and d_head == 64");
  if (c.n_mels == 0 || c.n_mels % 8 != 0) return set_error(WB_ERR_MODEL, "encoder kernels need n_mels % 8 == 0");
  if (static_cast<int>(c.n_mels) != m->mel.n_mels)
    return set_error(WB_ERR_MODEL, "filterbank n_mels differs from the model's n_mels");
  return WB_OK;
}

// The fused 30 s paths always produce 1500 positions: a header with a smaller n_audio_ctx gets the reference's own error
// (Encoder::forward, encoder.rs:456-461) instead of a workspace / positional-embedding overrun.
int check_fused_dims(const Replica* m) {
  int rc = check_encoder_dims(m);
  if (rc != WB_OK) return rc;
  const size_t S = (N_FRAMES_30S - 1) / 2 + 1;
  if (S > m->cfg.n_audio_ctx)
    return set_error(WB_ERR_MODEL, "sequence length " + std::to_string(S) + " exceeds max " + std::to_string(m->cfg.n_audio_ctx));
  return WB_OK;
}


int compute_mel_host(Replica* m, const float* const* audio, const size_t* n_samples, const float* contiguous, int B,
                            float* out) {
  const int nm = m->mel.n_mels;
  const size_t per_out = static_cast<size_t>(N_FRAMES_30S) * nm;
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    std::vector<int> nv(nb);
    for (int i = 0; i < nb; ++i) {
      const size_t n = contiguous ? N_SAMPLES_30S : std::min<size_t>(n_samples[b0 + i], N_SAMPLES_30S);   // lib.rs:413-425
      nv[i] = static_cast<int>(n);
      const float* src = contiguous ? contiguous + static_cast<size_t>(b0 + i) * N_SAMPLES_30S : audio[b0 + i];
      if (n) WB_CUDA_OK(cudaMemcpyAsync(m->ws.audio.p + static_cast<size_t>(i) * N_SAMPLES_30S, src, n * 4, cudaMemcpyHostToDevice, m->stream));
    }
    WB_CUDA_OK(cudaMemcpyAsync(m->ws.n_valid.p, nv.data(), nb * sizeof(int), cudaMemcpyHostToDevice, m->stream));
    if ((rc = mel_device(m, m->ws.audio.p, N_SAMPLES_30S, nullptr, m->ws.n_valid.p, nb, m->ws.mel_f32.p, false)) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpyAsync(out + static_cast<size_t>(b0) * per_out, m->ws.mel_f32.p, nb * per_out * 4, cudaMemcpyDeviceToHost, m->stream));
    WB_CUDA_OK(cudaStreamSynchronize(m->stream));     // nv goes out of scope; out is host-visible
  }
  return WB_OK;
}


int encode_same_len(Replica* m, const float* const* mels, const float* d_mel, int B, int T, void* out_host, void* out_dev,
                           size_t out_stride_elems, wb_dtype dt, int n_layers, bool ln_post) {
  // B mels of T frames each (host pointers `mels` or one device array `d_mel` [B][T][nm]) -> [B][S][d]
  const int nm = static_cast<int>(m->cfg.n_mels), d = static_cast<int>(m->cfg.n_audio_state);
  const int S = (T - 1) / 2 + 1;
  const size_t esz = dtype_size(dt);
  for (int b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = std::min(m->max_batch, B - b0);
    int rc = ensure_workspace(m, nb);
    if (rc != WB_OK) return rc;
    const float* src_dev;
    if (d_mel) {
      src_dev = d_mel + static_cast<size_t>(b0) * T * nm;
    } else {
      for (int i = 0; i < nb; ++i)
        WB_CUDA_OK(cudaMemcpyAsync(m->ws.mel_f32.p + static_cast<size_t>(i) * T * nm, mels[b0 + i], static_cast<size_t>(T) * nm * 4,
                                   cudaMemcpyHostToDevice, m->stream));
      src_dev = m->ws.mel_f32.p;
    }
    if ((rc = launch_mel_pad_op16(src_dev, m->ws.mel_bf16.p, nb, T, nm, m->stream)) != WB_OK) return rc;
    void* dst_dev;
    if (out_dev) dst_dev = static_cast<uint8_t*>(out_dev) + static_cast<size_t>(b0) * out_stride_elems * esz;
    else dst_dev = dt != WB_F32 ? static_cast<void*>(m->ws.out_bf16.p) : static_cast<void*>(m->ws.out_f32.p);
    if ((rc = encode_device(m, nb, T, dst_dev, dt, n_layers, ln_post)) != WB_OK) return rc;
    if (out_host) {
      const size_t row_bytes = static_cast<size_t>(S) * d * esz;
      WB_CUDA_OK(cudaMemcpy2DAsync(static_cast<uint8_t*>(out_host) + static_cast<size_t>(b0) * out_stride_elems * esz,
                                   out_stride_elems * esz, dst_dev, row_bytes, row_bytes, nb, cudaMemcpyDeviceToHost, m->stream));
      WB_CUDA_OK(cudaStreamSynchronize(m->stream));
    }
  }
  return WB_OK;
}

int validate_mel_len(const Replica* m, size_t mel_len, int* T_out) {
  const size_t nm = m->cfg.n_mels;
  if (mel_len % nm != 0)                                                             // encoder.rs:568-574
    return set_error(WB_ERR_MODEL, "mel size " + std::to_string(mel_len) + " not divisible by n_mels " + std::to_string(nm));
  const size_t T = mel_len / nm;
  const size_t S = T == 0 ? 0 : (T - 1) / 2 + 1;
  if (S > m->cfg.n_audio_ctx)                                                        // encoder.rs:456-461
    return set_error(WB_ERR_MODEL, "sequence length " + std::to_string(S) + " exceeds max " + std::to_string(m->cfg.n_audio_ctx));
  *T_out = static_cast<int>(T);
  return WB_OK;
}


// One fused mel + encoder step on the model's stream, replayed from a CUDA graph once its (pointers, batch) key has been seen twice.
int mel_encode_step(Replica* m, const float* d_audio, long long audio_stride, const long long* d_seg_off, const int* d_n_valid, int nb,
                    void* d_out, wb_dtype out_dtype) {
  auto eager = [&]() -> int {
    int rc = mel_device(m, d_audio, audio_stride, d_seg_off, d_n_valid, nb, nullptr, true);
    if (rc != WB_OK) return rc;
    return encode_device(m, nb, N_FRAMES_30S, d_out, out_dtype, -1, true);
  };
  if (!m->use_graphs || m->prof_on) return eager();
  Replica::StepGraph* g = nullptr;
  for (auto& e : m->graphs)
    if (e.in == d_audio && e.n_valid == d_n_valid && e.seg_off == d_seg_off && e.stride == audio_stride && e.out == d_out && e.B == nb &&
        e.dtype == static_cast<int>(out_dtype))
      g = &e;
  if (!g) {
    if (m->graphs.size() >= 16) {                         // callers that never repeat their pointers do not accumulate graphs
      for (auto& e : m->graphs)
        if (e.exec) cudaGraphExecDestroy(e.exec);
      m->graphs.clear();
    }
    Replica::StepGraph e;
    e.in = d_audio; e.n_valid = d_n_valid; e.seg_off = d_seg_off; e.stride = audio_stride; e.out = d_out; e.B = nb; e.dtype = static_cast<int>(out_dtype); e.seen = 1;
    m->graphs.push_back(e);
    return eager();
  }
  if (g->exec) {
    WB_CUDA_OK(cudaGraphLaunch(g->exec, m->stream));
    count_launch(static_cast<int>(g->launches));
    return WB_OK;
  }
  // second sighting: capture the launch sequence (thread-local mode: other threads' CUDA calls are unaffected)
  const long long before = g_launch_count.load();
  WB_CUDA_OK(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
  int rc = eager();
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
  if (rc != WB_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    m->use_graphs = false;                                // fall back to plain launches for this model
    if (rc != WB_OK) return rc;
    return eager();
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) {
    cudaGetLastError();
    m->use_graphs = false;
    return eager();
  }
  g->exec = exec;
  g->launches = g_launch_count.load() - before;
  WB_CUDA_OK(cudaGraphLaunch(exec, m->stream));
  return WB_OK;
}


int prepare_slot(Replica* m, Replica::Slot& sl, int nb, size_t out_bytes) {
  int rc;
  if (!m->in_stream) {
    WB_CUDA_OK(cudaStreamCreateWithFlags(&m->in_stream, cudaStreamNonBlocking));
    WB_CUDA_OK(cudaStreamCreateWithFlags(&m->out_stream, cudaStreamNonBlocking));
  }
  if (!sl.in_done) {
    WB_CUDA_OK(cudaEventCreateWithFlags(&sl.in_done, cudaEventDisableTiming));
    WB_CUDA_OK(cudaEventCreateWithFlags(&sl.compute_done, cudaEventDisableTiming));
    WB_CUDA_OK(cudaEventCreateWithFlags(&sl.out_done, cudaEventDisableTiming));
  }
  if (sl.busy) {                                         // the batch that used this slot two calls ago must have left it
    WB_CUDA_OK(cudaEventSynchronize(sl.out_done));
    sl.busy = false;
  }
  if ((rc = sl.audio.ensure(static_cast<size_t>(nb) * N_SAMPLES_30S)) != WB_OK) return rc;
  if ((rc = sl.n_valid.ensure(nb)) != WB_OK) return rc;
  if (out_bytes && (rc = sl.out.ensure(out_bytes)) != WB_OK) return rc;
  if (sl.h_cap < nb) {
    if (sl.h_n_valid) cudaFreeHost(sl.h_n_valid);
    sl.h_n_valid = nullptr;
    WB_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&sl.h_n_valid), static_cast<size_t>(nb) * sizeof(int), cudaHostAllocDefault));
    sl.h_cap = nb;
  }
  return WB_OK;
}

// transcribe_batch_optimized steps 1-2 from host buffers, one micro-batch: copy-in stream (audio H2D) -> the replica's stream
// (mel + encoder) -> copy-out stream (states D2H), chained by events over two staging slots.  With d_out_final the final LayerNorm
// writes the states straight to that address (this device's or, over NVLink, a peer's memory) and nothing is copied out.
int enqueue_microbatch(Replica* m, const float* const* audio, const size_t* n_samples, int nb, void* out_host, void* d_out_final,
                       wb_dtype out_dtype) {
  int rc;
  const size_t d = m->cfg.n_audio_state, S = N_POS_30S, esz = dtype_size(out_dtype);
  const size_t out_bytes = static_cast<size_t>(nb) * S * d * esz;
  Replica::Slot& sl = m->slot[m->next_slot];
  m->next_slot ^= 1;
  if ((rc = ensure_workspace(m, nb)) != WB_OK) return rc;
  if ((rc = prepare_slot(m, sl, nb, d_out_final ? 0 : out_bytes)) != WB_OK) return rc;
  for (int i = 0; i < nb; ++i) {
    const size_t n = std::min<size_t>(n_samples[i], N_SAMPLES_30S);
    sl.h_n_valid[i] = static_cast<int>(n);
    if (n) WB_CUDA_OK(cudaMemcpyAsync(sl.audio.p + static_cast<size_t>(i) * N_SAMPLES_30S, audio[i], n * 4, cudaMemcpyHostToDevice, m->in_stream));
  }
  WB_CUDA_OK(cudaMemcpyAsync(sl.n_valid.p, sl.h_n_valid, nb * sizeof(int), cudaMemcpyHostToDevice, m->in_stream));
  WB_CUDA_OK(cudaEventRecord(sl.in_done, m->in_stream));
  WB_CUDA_OK(cudaStreamWaitEvent(m->stream, sl.in_done, 0));
  void* d_out = d_out_final ? d_out_final : static_cast<void*>(sl.out.p);
  if ((rc = mel_encode_step(m, sl.audio.p, N_SAMPLES_30S, nullptr, sl.n_valid.p, nb, d_out, out_dtype)) != WB_OK) return rc;
  WB_CUDA_OK(cudaEventRecord(sl.compute_done, m->stream));
  if (out_host) {
    WB_CUDA_OK(cudaStreamWaitEvent(m->out_stream, sl.compute_done, 0));
    WB_CUDA_OK(cudaMemcpyAsync(out_host, sl.out.p, out_bytes, cudaMemcpyDeviceToHost, m->out_stream));
    WB_CUDA_OK(cudaEventRecord(sl.out_done, m->out_stream));
  } else {
    WB_CUDA_OK(cudaEventRecord(sl.out_done, m->stream));       // the slot's audio buffer is free once the step has run
  }
  sl.busy = true;
  return WB_OK;
}

int sync_replica(Replica* m) {
  DeviceGuard guard(m->device);
  for (auto& sl : m->slot) {
    if (sl.busy) {
      WB_CUDA_OK(cudaEventSynchronize(sl.out_done));
      sl.busy = false;
    }
  }
  WB_CUDA_OK(cudaStreamSynchronize(m->stream));
  return WB_OK;
}

// BatchPreprocessor::process_batch (src/audio/batch.rs:157-176) on the device, and MelFilterbank::compute for one segment as its
// B = 1 case: every segment through the mel with the given tables, neither padded nor truncated.  The whole ragged batch is ONE
// upload (segments packed into an arena), ONE stft launch over a tile table, one in-place finalize and ONE download; the arena, the
// log-mel rows and the host staging vectors belong to the replica and are reused from call to call.
int mel_compute_ragged(Replica* m, const MelTables& tab, const float* const* audio, const size_t* n_samples, int B, size_t hop,
                       float* const* mels_out, const size_t* out_capacity, size_t* frame_counts, size_t* max_frames_out) {
  if (max_frames_out) *max_frames_out = 0;
  const int nm = tab.n_mels;
  std::vector<long long> seg_off(B), row_off(B);
  std::vector<int> n_valid(B), n_fr(B);
  std::vector<int2> tiles;
  long long arena = 0, rows = 0, max_rows = 0;
  for (int i = 0; i < B; ++i) {
    const size_t n = n_samples[i];
    size_t nf = 0;
    if (n != 0) {                                                                      // mel.rs:236-238: empty audio -> empty result
      if (hop == 0) return set_error(WB_ERR_AUDIO, "hop_length must be positive");     // mel.rs:240-242
      nf = n >= N_FFT ? (n - N_FFT) / hop + 1 : 0;                                     // mel.rs:245-249
      if (n > 0x7fff0000ull || hop > 0x7fff0000ull || nf > 0x7fff0000ull) return set_error(WB_ERR_AUDIO, "audio too long for one call");
      if (nf && (!audio[i] || !mels_out[i] || out_capacity[i] < nf * nm)) return set_error(WB_ERR_AUDIO, "output buffer too small");
    }
    if (frame_counts) frame_counts[i] = nf;
    n_fr[i] = static_cast<int>(nf);
    n_valid[i] = static_cast<int>(n);
    seg_off[i] = arena;
    row_off[i] = rows;
    if (nf) {
      arena += static_cast<long long>((n + 3) & ~static_cast<size_t>(3));              // every segment starts 16 B aligned
      rows += static_cast<long long>(nf);
      max_rows = std::max<long long>(max_rows, static_cast<long long>(nf));
      for (size_t f0 = 0; f0 < nf; f0 += 32) tiles.push_back(make_int2(i, static_cast<int>(f0)));
    }
  }
  if (max_frames_out) *max_frames_out = static_cast<size_t>(max_rows);
  if (rows == 0) return WB_OK;
  int rc;
  Replica::Ragged& r = m->rag;
  if ((rc = r.audio.ensure(static_cast<size_t>(arena) + 4)) != WB_OK || (rc = r.logmel.ensure(static_cast<size_t>(rows) * nm)) != WB_OK ||
      (rc = r.keys.ensure(B)) != WB_OK)
    return rc;
  // one table blob: seg_off | row_off | n_valid | n_frames | tiles
  const size_t t_bytes = static_cast<size_t>(B) * (8 + 8 + 4 + 4) + tiles.size() * sizeof(int2) + 64;
  if ((rc = r.tables.ensure(t_bytes)) != WB_OK) return rc;
  r.h_tables.resize(t_bytes);
  uint8_t* hb = r.h_tables.data();
  size_t o_seg = 0, o_row = o_seg + 8 * static_cast<size_t>(B), o_nv = o_row + 8 * static_cast<size_t>(B), o_nf = o_nv + 4 * static_cast<size_t>(B);
  size_t o_tiles = (o_nf + 4 * static_cast<size_t>(B) + 7) & ~static_cast<size_t>(7);
  memcpy(hb + o_seg, seg_off.data(), 8 * static_cast<size_t>(B));
  memcpy(hb + o_row, row_off.data(), 8 * static_cast<size_t>(B));
  memcpy(hb + o_nv, n_valid.data(), 4 * static_cast<size_t>(B));
  memcpy(hb + o_nf, n_fr.data(), 4 * static_cast<size_t>(B));
  memcpy(hb + o_tiles, tiles.data(), tiles.size() * sizeof(int2));
  r.h_audio.resize(static_cast<size_t>(arena));
  for (int i = 0; i < B; ++i)
    if (n_fr[i]) memcpy(r.h_audio.data() + seg_off[i], audio[i], static_cast<size_t>(n_valid[i]) * 4);
  cudaStream_t st = m->stream;
  WB_CUDA_OK(cudaMemcpyAsync(r.audio.p, r.h_audio.data(), static_cast<size_t>(arena) * 4, cudaMemcpyHostToDevice, st));
  WB_CUDA_OK(cudaMemcpyAsync(r.tables.p, hb, t_bytes, cudaMemcpyHostToDevice, st));
  if ((rc = launch_mel_init_keys(r.keys.p, B, st)) != WB_OK) return rc;
  MelBatch job;
  job.audio = r.audio.p;
  job.seg_off = reinterpret_cast<const long long*>(r.tables.p + o_seg);
  job.row_off = reinterpret_cast<const long long*>(r.tables.p + o_row);
  job.n_valid = reinterpret_cast<const int*>(r.tables.p + o_nv);
  job.n_frames_arr = reinterpret_cast<const int*>(r.tables.p + o_nf);
  job.tiles = reinterpret_cast<const int2*>(r.tables.p + o_tiles);
  job.n_tiles = static_cast<int>(tiles.size());
  job.hop = static_cast<int>(hop);
  job.B = B;
  {
    NvtxRange nvtx("step_f_mel");
    WB_PROF(PC_MEL_STFT, launch_mel_stft(job, tab, r.logmel.p, r.keys.p, st));
    WB_PROF(PC_MEL_FINALIZE, launch_mel_finalize_ragged(r.logmel.p, r.keys.p, job.n_frames_arr, job.row_off, nm, B, max_rows, r.logmel.p, st));
  }
  r.h_out.resize(static_cast<size_t>(rows) * nm);
  if (cudaMemcpyAsync(r.h_out.data(), r.logmel.p, r.h_out.size() * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return set_error(WB_ERR_CUDA, std::string("mel kernels failed: ") + cudaGetErrorString(cudaGetLastError()));
  for (int i = 0; i < B; ++i)
    if (n_fr[i]) memcpy(mels_out[i], r.h_out.data() + row_off[i] * nm, static_cast<size_t>(n_fr[i]) * nm * 4);
  return WB_OK;
}

}  // namespace wb
