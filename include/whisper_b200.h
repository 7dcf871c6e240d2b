/* whisper_b200.h -- C ABI of libwhisper_b200.so
 *
 * B200-native (sm_100a) drop-in for ONE path of paiml/whisper.apr: batched log-mel
 * extraction over 30 s chunks followed by the Whisper encoder forward pass.
 *
 * The reference has no FFI boundary for this path (its only extern surface is
 * wasm-bindgen); the boundary it offers is its Rust public API.  Every entry point
 * below is shaped 1:1 after the Rust signature it replaces (cited as file:line,
 * relative to the reference checkout) so that a thin `extern "C"` crate can bind it
 * (INTEGRATION.md shows the binding).  Plain pointers and sizes only.
 *
 * Conventions
 *   - Status: 0 OK; 1 Audio, 2 Model, 3 Format mirror WhisperError::{Audio,Model,Format}
 *     (src/error.rs:6-44); 4 = CUDA failure.  wb_last_error() returns the message
 *     (thread-local), like WhisperError's Display.
 *   - Caller owns every host buffer.  The library owns device memory behind wb_model.
 *   - All calls are synchronous with respect to their host outputs.  The *_dev entry
 *     points take DEVICE pointers, enqueue on the model's stream and do not synchronise.
 *   - There is no CPU fallback: without a CUDA device every compute call fails with
 *     WB_ERR_CUDA.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_OK 0
#define WB_ERR_AUDIO 1
#define WB_ERR_MODEL 2
#define WB_ERR_FORMAT 3
#define WB_ERR_CUDA 4

typedef struct wb_model wb_model;

/* ModelConfig (src/model/mod.rs:35-62) as read from the .apr header
 * (AprHeader::to_model_config, src/format/mod.rs:249-279), plus the file's quantization byte. */
typedef struct wb_config {
  uint32_t model_type;
  uint32_t n_vocab;
  uint32_t n_audio_ctx;
  uint32_t n_audio_state;
  uint32_t n_audio_head;
  uint32_t n_audio_layer;
  uint32_t n_text_ctx;
  uint32_t n_text_state;
  uint32_t n_text_head;
  uint32_t n_text_layer;
  uint32_t n_mels;
  uint32_t quantization; /* 0 F32, 1 F16 (unsupported, as in the reference), 2 Int8, 3 Int4 */
  uint32_t has_filterbank;
  uint32_t n_tensors;
} wb_config;

/* Output element type of the fused batch entry point (a storage format for the caller; independent of the operand format). */
typedef enum wb_dtype { WB_F32 = 0, WB_BF16 = 1 } wb_dtype;

/* On-device requantisation modes of wb_model_requantize. */
#define WB_QUANT_INT8_PER_CHANNEL 1

/* ---- library ---------------------------------------------------------------------- */
const char* wb_version(void);
/* "fp16" or "bf16": the 16-bit format of every tensor-core operand in this build (IEEE fp16 by default). */
const char* wb_operand_format(void);
/* WhisperError Display (src/error.rs).  Thread-local; valid until the next failing call. */
const char* wb_last_error(void);
/* parallel::thread_count (src/parallel.rs:155-170) -> number of visible CUDA devices. */
int wb_device_count(void);

/* ---- model ------------------------------------------------------------------------ */
/* WhisperApr::load_from_apr (src/lib.rs:673-754) restricted to what this path needs:
 * AprReader::new (src/format/mod.rs:484-522), header -> config, load_encoder_weights
 * (src/lib.rs:757-841, 931-993), read_mel_filterbank -> MelFilterbank::from_apr_data
 * (src/lib.rs:738-741; falls back to MelFilterbank::new's HTK filters when the file has
 * none, src/lib.rs:297-298).  Missing tensors keep their defaults, lengths are clamped,
 * exactly as the reference loader does.  Int8 / Int4 weights stay packed on the host side of
 * the upload and are expanded on the device.  `device` is the CUDA ordinal. */
int wb_model_from_apr(const uint8_t* bytes, size_t n_bytes, int device, wb_model** out);
/* The same load onto a SET of devices of this process: weights and filterbank are replicated (one host thread per device pulls the
 * caller's page-locked bytes over that device's own link), every device gets its own streams and workspace.  This is what
 * parallel::configure_thread_pool (src/parallel.rs:34-60) becomes on a box of B200s: the handle's device list is the pool.
 * wb_mel_encode_batch{,_async}, wb_mel_encode_gather and wb_transcribe_tokens_batch shard their chunks over the list
 * (chunk i of B -> device floor(i * n_devices / B): contiguous blocks, order preserved like parallel_map, src/parallel.rs:82-118);
 * every other entry point runs on devices[0]. */
int wb_model_from_apr_devices(const uint8_t* bytes, size_t n_bytes, const int* devices, int n_devices, wb_model** out);
/* parallel::thread_count for a handle: the number of devices it spans; and the CUDA ordinal of its index-th device (-1 if out of range). */
int wb_model_n_devices(const wb_model* m);
int wb_model_device(const wb_model* m, int index);
/* Requantise the resident bf16 linear weights on the device.  WB_QUANT_INT8_PER_CHANNEL: quantize_f32_to_i8_per_channel
 * (src/model/quantized.rs:1769-1794) -- one scale = absmax / 127 per output channel, round half away from zero, clamp [-128, 127];
 * the int8 rows replace the bf16 matrices in HBM (1 B per weight) and the per-channel scale is applied per output column in the GEMM
 * epilogue, which is dequantize_i8_to_f32_per_channel (:1797-1813) folded into the product.  Needs a model loaded from f32 payloads. */
int wb_model_requantize(wb_model* m, int mode);
/* WhisperApr::config (src/lib.rs:330-333). */
int wb_model_config(const wb_model* m, wb_config* out);
void wb_model_free(wb_model* m);
/* parallel::configure_thread_pool (src/parallel.rs:34-60) analogue: the stream the model's
 * kernels are enqueued on (a cudaStream_t; NULL restores the model's own stream). */
int wb_model_set_stream(wb_model* m, void* cuda_stream);
/* Upper bound on chunks processed per device pass (workspace is sized for it). */
int wb_model_set_max_batch(wb_model* m, int max_chunks);

/* ---- mel -------------------------------------------------------------------------- */
/* MelFilterbank::compute (src/audio/mel.rs:233-310).  out: [n_frames][n_mels] f32, frame-major,
 * capacity out_capacity floats; *n_frames_out = (n - 400) / hop + 1, or 0 for n < 400 / n == 0
 * (then nothing is written).  hop == 0 -> WB_ERR_AUDIO ("hop_length must be positive"). */
int wb_mel_compute(const wb_model* m, const float* audio, size_t n, size_t hop, float* out, size_t out_capacity,
                   size_t* n_frames_out);
/* WhisperApr::compute_mel (src/lib.rs:407-443): pad/truncate to 480000 samples, mel, pad the
 * frame axis to 3000 with -1.0.  out: [3000][n_mels] f32.  (n_mels comes from the model, not the
 * reference's hard-coded 80 -- see DESIGN.md.) */
int wb_compute_mel(const wb_model* m, const float* audio, size_t n, float* out);
/* Batched compute_mel: `audio` is B chunks of exactly 480000 samples, contiguous. */
int wb_compute_mel_batch(const wb_model* m, const float* audio, int B, float* out);

/* ---- encoder ---------------------------------------------------------------------- */
/* WhisperApr::encode == Encoder::forward_mel (src/lib.rs:446-449, src/model/encoder.rs:566-581).
 * mel: [n_frames][n_mels] f32 (mel_len floats).  out: [S][d] f32 with
 * S = (n_frames - 1) / 2 + 1; *seq_len_out = S.  mel_len % n_mels != 0 -> WB_ERR_MODEL
 * ("mel size .. not divisible by n_mels"); S > n_audio_ctx -> WB_ERR_MODEL
 * ("sequence length .. exceeds max .."), as encoder.rs:450-461. */
int wb_encode(const wb_model* m, const float* mel, size_t mel_len, float* out, size_t out_capacity, size_t* seq_len_out);
/* Encoder::forward_batch_padded (src/model/encoder.rs:625-660): B mels of possibly different
 * lengths -> out [B][max_seq][d] zero padded, seq_lens[B], *max_seq_out.  out_capacity in floats.
 * Encoder::forward_batch (encoder.rs:599-608) is the same call read back per item. */
int wb_encode_batch(const wb_model* m, const float* const* mels, const size_t* mel_lens, int B, float* out,
                    size_t out_capacity, size_t* seq_lens, size_t* max_seq_out);

/* ---- fused hot path --------------------------------------------------------------- */
/* transcribe_batch_optimized steps 1-2 (src/lib.rs:1162-1170): compute_mel per item then
 * Encoder::forward_batch.  audio[i] has n_samples[i] samples (padded/truncated to 30 s as
 * compute_mel does).  out: [B][1500][d] in out_dtype (WB_F32 as the reference, or WB_BF16).
 * This is the measured entry point (bench.py `e2e`): host buffers in, host buffer out. */
int wb_mel_encode_batch(const wb_model* m, const float* const* audio, const size_t* n_samples, int B, void* out,
                        wb_dtype out_dtype);
/* The same call without the final wait: host->device copies, compute and the device->host copy of the result are enqueued
 * on three streams (copy-in, the model's stream, copy-out) over two staging slots, so the copies of one batch overlap the
 * compute of its neighbours; a third call blocks until the oldest batch has left its slot.  `audio[i]` and `out` must
 * stay valid (and should be pinned) until wb_sync returns; results are complete after wb_sync. */
int wb_mel_encode_batch_async(const wb_model* m, const float* const* audio, const size_t* n_samples, int B, void* out,
                              wb_dtype out_dtype);
/* Same work with inputs already resident in HBM: d_audio [B][480000] f32 (device), d_out
 * [B][1500][d] (device).  Enqueues on the model's stream; does not synchronise. */
int wb_mel_encode_batch_dev(const wb_model* m, const float* d_audio, int B, void* d_out, wb_dtype out_dtype);
/* Device-resident pieces of the same path, for per-kernel measurement. */
int wb_compute_mel_batch_dev(const wb_model* m, const float* d_audio, int B, float* d_mel_out);
int wb_encode_batch_dev(const wb_model* m, const float* d_mel, int B, void* d_out, wb_dtype out_dtype);
/* Block until every device of the handle is idle. */
int wb_sync(const wb_model* m);
/* The sharded call with the encoder states left ON A DEVICE ("a final NVLink gather of encoder states"): the device at
 * index `gather_index` of the handle's list receives all B chunks' states, [B][1500][d] in chunk order, in a library-owned buffer
 * returned in *d_states_out (valid until the next gather call on the handle; complete after wb_sync, and ordered before any work
 * enqueued afterwards on that device's stream).  The other devices' final LayerNorm kernels store their rows straight into that
 * buffer over NVLink -- peer stores issued by the compute kernel itself, micro-batch by micro-batch, so the transfer overlaps the
 * next micro-batch: no staging copy, no extra pass, no NCCL call. */
int wb_mel_encode_gather(const wb_model* m, const float* const* audio, const size_t* n_samples, int B, wb_dtype out_dtype,
                         int gather_index, void** d_states_out);
/* Peer-visible device buffers across PROCESSES (one process per GPU under torchrun): the gather rank allocates and exports its states
 * buffer (64-byte cudaIpcMemHandle), the other ranks open it and pass `peer + offset` as d_out of wb_mel_encode_batch_dev -- their
 * final LayerNorm then stores over NVLink exactly as in the single-process gather. */
int wb_ipc_alloc(int device, size_t bytes, void** d_ptr, uint8_t* handle64);
int wb_ipc_open(int device, const uint8_t* handle64, void** d_ptr);
int wb_ipc_close(int device, void* d_ptr);
int wb_ipc_free(int device, void* d_ptr);
/* cudaMemcpy of a device buffer this process can address (e.g. the gather buffer) to the host. */
int wb_read_device(int device, const void* d_src, void* host_dst, size_t bytes);

/* ---- decoder front half ------------------------------------------------ */
/* 1 when the .apr file carried decoder tensors (load_decoder_weights, src/lib.rs:843-929) and the decode entry points work. */
int wb_decoder_available(const wb_model* m);
/* WhisperApr::decode with DecodingStrategy::Greedy (src/lib.rs:529-598; GreedyDecoder::decode, src/inference/greedy.rs:118-146) for B
 * chunks side by side: states [B][seq_len][d] f32 (host) are the encoder outputs; the cross-attention K/V are computed once per
 * chunk (two tensor-core GEMMs per decoder layer; decoder.rs:2276-2296 does it on the first token, :2017-2040 per token), then
 * Decoder::forward_one (decoder.rs:2125-2172) runs per position with a KV cache, WhisperTokenSuppressor::apply
 * (src/inference/processors.rs:126-147; timestamps suppressed unless suppress_timestamps == 0) and the first-maximum argmax
 * (greedy.rs:83-96), all on the device.  A chunk stops after EOT; the call stops when every chunk has, or at max_tokens tokens in
 * total (initial ones included).  tokens_out [B][max_tokens] (padded with EOT), lens_out [B]. */
int wb_decode_greedy(const wb_model* m, const float* states, size_t seq_len, int B, const int32_t* initial_tokens, int n_init,
                     int max_tokens, int suppress_timestamps, int32_t* tokens_out, int32_t* lens_out);
/* transcribe_batch_optimized steps 1-3 up to token ids (src/lib.rs:1162-1201): mel + encoder + greedy decode with the encoder
 * states never leaving HBM (bf16), sharded over the handle's devices like wb_mel_encode_batch -- the decoder is sharded the same way,
 * so no gather is needed.  Only token ids come back. */
int wb_transcribe_tokens_batch(const wb_model* m, const float* const* audio, const size_t* n_samples, int B,
                               const int32_t* initial_tokens, int n_init, int max_tokens, int suppress_timestamps, int32_t* tokens_out,
                               int32_t* lens_out);

/* ---- chunking --------------------------------------------------------------------- */
/* audio::split_into_chunks (src/audio/batch.rs:219-240).  Writes up to `capacity` (start,len)
 * pairs; returns the number of chunks the reference would produce (may exceed capacity). */
size_t wb_split_into_chunks(size_t n_samples, size_t chunk_size, size_t overlap, size_t* starts, size_t* lens,
                            size_t capacity);
/* BatchPreprocessor::process_batch (src/audio/batch.rs:157-176): segment i (n_samples[i] samples, any length) ->
 * mels_out[i] = MelFilterbank::compute(segment, hop) with the preprocessor's own HTK filterbank MelFilterbank::new(n_mels, 400,
 * 16000) (batch.rs:143) -- not the model's -- no 30 s padding; frame_counts[i] frames of n_mels values, *max_frames_out as
 * BatchMelResult::max_frames.  out_capacity[i] in floats.  AudioConfig::default(): n_mels 80, hop 160. */
int wb_batch_preprocess(const wb_model* m, const float* const* audio, const size_t* n_samples, int B, size_t n_mels, size_t hop,
                        float* const* mels_out, const size_t* out_capacity, size_t* frame_counts, size_t* max_frames_out);
/* BatchMelResult::to_padded_tensor (src/audio/batch.rs:107-127): [B][n_mels][max_frames], zero padded. */
int wb_to_padded_tensor(const float* const* mels, const size_t* frame_counts, int B, size_t n_mels, size_t max_frames,
                        float* out);

/* ---- ingest in front of the hot path: WAV, resampling, VAD ---------------------------- */
/* What parse_wav (src/audio/wav.rs:99-224) finds in a RIFF/WAVE buffer.  sample_kind: 0 u8, 1 i16, 2 i24, 3 i32 PCM, 4 f32. */
typedef struct wb_wav_info {
  uint32_t sample_rate;
  uint16_t channels, bits_per_sample, sample_kind, reserved;
  uint64_t data_offset, data_bytes;
  uint64_t n_frames; /* samples per channel == length of the mono result */
} wb_wav_info;
/* The chunk walk of parse_wav (fmt / WAVE_FORMAT_EXTENSIBLE / data; unknown chunks skipped with even alignment); host only, no GPU.
 * Errors are WB_ERR_AUDIO with WavError's texts ("WAV file too small", "missing RIFF header", "missing WAVE format",
 * "fmt chunk truncated", "unsupported format ..", "unsupported channel count ..", "no data chunk"). */
int wb_wav_parse(const uint8_t* bytes, size_t n_bytes, wb_wav_info* out);
/* parse_wav's payload half on the device: convert_{8,16,24,32}bit_pcm / convert_32bit_float + convert_to_mono (wav.rs:226-286):
 * the raw payload is uploaded as it is, converted and down-mixed by one kernel.  out: info->n_frames mono f32 samples. */
int wb_wav_decode(const wb_model* m, const uint8_t* bytes, size_t n_bytes, float* out, size_t out_capacity, wb_wav_info* info);
/* SincResampler::resample (src/audio/resampler.rs:136-206; new / with_params :66-110): Kaiser-windowed sinc interpolation, f64
 * accumulation, weight-normalised, output length ceil(n * target / source); same rates -> copy.  Errors as the reference:
 * "sample rate must be non-zero", "kernel half-length must be non-zero", "cannot resample empty audio". */
size_t wb_resample_len(size_t n, uint32_t source_rate, uint32_t target_rate);
int wb_resample(const wb_model* m, const float* audio, size_t n, uint32_t source_rate, uint32_t target_rate, float* out,
                size_t out_capacity, size_t* n_out);
int wb_resample_with_params(const wb_model* m, const float* audio, size_t n, uint32_t source_rate, uint32_t target_rate,
                            int kernel_half_len, double kaiser_beta, float* out, size_t out_capacity, size_t* n_out);
/* WAV bytes -> mono -> 16 kHz with one upload of the raw payload and one download (what WhisperApr feeds compute_mel from a file). */
int wb_ingest_wav_16k(const wb_model* m, const uint8_t* bytes, size_t n_bytes, float* out, size_t out_capacity, size_t* n_out,
                      wb_wav_info* info);
/* VadConfig (src/vad.rs:36-66). */
typedef struct wb_vad_config {
  uint32_t sample_rate, frame_size;
  float energy_threshold, zcr_threshold;
  uint32_t min_speech_frames, min_silence_frames;
  float smoothing;
} wb_vad_config;
void wb_vad_config_default(wb_vad_config* c);
/* VoiceActivityDetector::detect (src/vad.rs:554-607) for B streams at once: per-frame RMS energy and zero-crossing rate (one thread
 * per frame, the reference's sequential f32 sums), then process_frame's state machine with its adaptive noise floor (one thread per
 * stream, :609-672).  segments: [B][seg_capacity][3] = (start s, end s, mean energy); n_segments[B] (may exceed the capacity: the
 * reference's count); frame_events (optional): per stream, one byte per frame, 0 Continue / 1 SpeechStart / 2 SpeechEnd;
 * n_frames_out (optional) [B].  cfg == NULL: VadConfig::default(). */
int wb_vad_detect_batch(const wb_model* m, const float* const* audio, const size_t* n_samples, int B, const wb_vad_config* cfg,
                        float* segments, int seg_capacity, int* n_segments, uint8_t* const* frame_events, size_t* n_frames_out);

/* ---- streaming front end on the device ------------------------------------------------- */
/* audio::split_into_chunks (src/audio/batch.rs:219-240) + compute_mel + Encoder::forward_batch for every chunk of every stream without
 * materialising the chunks: each stream is uploaded once and a chunk is an (offset, length) VIEW the mel kernel reads in place
 * (samples past the view read as 0 = compute_mel's zero padding to 30 s, src/lib.rs:413-420).  out: [total chunks][1500][d],
 * stream-major; chunk_counts [n_streams] and *total_chunks as split_into_chunks gives them.  Streams shard over the handle's devices. */
int wb_stream_encode_views(const wb_model* m, const float* const* streams, const size_t* stream_lens, int n_streams, size_t chunk_size,
                           size_t overlap, void* out, wb_dtype out_dtype, size_t out_capacity_chunks, size_t* chunk_counts,
                           size_t* total_chunks);
/* Device-resident form of the same work (per-kernel measurement, and callers whose streams already live in HBM): d_arena holds the
 * streams, d_seg_off[i] / d_n_valid[i] are chunk i's first sample (element offset into d_arena) and length.  Asynchronous on the
 * model's stream. */
int wb_mel_encode_views_dev(const wb_model* m, const float* d_arena, const long long* d_seg_off, const int* d_n_valid, int n_chunks,
                            void* d_out, wb_dtype out_dtype);
/* The chunk-assembly half of StreamingProcessor (src/audio/streaming.rs) for n_streams streams whose accumulators live in HBM:
 * push = push_audio (:672-675); ready = has_chunk; encode = get_chunk (:843-870: [carried overlap | samples] zero padded to the chunk
 * size, the chunk's last overlap_samples kept as the next chunk's prefix) -- or flush (:872-905) for every stream holding fresh audio --
 * assembled by one kernel, then compute_mel + encoder over the assembled batch.  VAD gating is not part of the set: drive it with
 * wb_vad_detect_batch and push what should be transcribed. */
typedef struct wb_stream_set wb_stream_set;
int wb_stream_set_new(const wb_model* m, int n_streams, size_t chunk_samples, size_t overlap_samples, wb_stream_set** out);
void wb_stream_set_free(wb_stream_set* s);
int wb_stream_set_push(wb_stream_set* s, const int* stream_ids, const float* const* samples, const size_t* n_samples, int count);
int wb_stream_set_ready(const wb_stream_set* s, int* ids_out, int capacity);
int wb_stream_set_encode(wb_stream_set* s, int flush, void* out, wb_dtype out_dtype, int* ids_out, size_t* n_valid_out, int capacity,
                         int* n_chunks_out);
/* test hook: the chunk batch the last wb_stream_set_encode assembled, [n][chunk_samples] f32 */
int wb_debug_stream_set_chunks(const wb_stream_set* s, int n, float* out);

/* ---- test hooks (not part of the reference surface) ------------------------------- */
/* Host restatement of the kernel's FFT factorisation: P[201] = |rfft400(y)|^2.  No GPU needed. */
void wb_debug_fft400_power_host(const float* y400, float* p201);
/* One GEMM through the tcgen05 kernel: out = epilogue(alpha * A[M][K] W[N][K]^T + bias), host buffers
 * (A, W as f32, rounded to bf16 on upload; out as f32).  epilogue: see GemmEpilogue in wb_internal.h. */
int wb_debug_gemm(int device, const float* A, const float* W, const float* bias, const float* resid_or_pe, int M, int N,
                  int K, int epilogue, float alpha, float* out);
/* Times `iters` back-to-back launches of one GEMM shape ([n_batch][rows][K] x [N][K]^T, bf16, device-resident pseudo-random
 * operands, bias on) after 2 warm-up launches; *ms_per_launch receives the mean.  Kernel tuning only. */
int wb_debug_gemm_bench(int device, int n_batch, int rows, int N, int K, int epilogue, int iters, float* ms_per_launch);
/* One attention call: qkv f32 [B][S][3d] (rounded to bf16) -> out f32 [B][S][d]. */
int wb_debug_attention(int device, const float* qkv, int B, int S, int d, int n_heads, float* out);
/* Times `iters` back-to-back attention launches on device-resident pseudo-random bf16 qkv [B][S][3d] (CUDA events on the
 * launching stream, after 2 warm-up launches); *ms_per_launch receives the mean.  Kernel tuning only. */
int wb_debug_attention_bench(int device, int B, int S, int d, int n_heads, int iters, float* ms_per_launch);
/* Encoder::forward_mel truncated after n_layers blocks (n_layers < 0: all), with or without ln_post: stage-wise parity. */
/* Experiment switch (default off; WB_LN_FOLLOW=1 at load): run the LayerNorm that follows a residual GEMM as a concurrent kernel
 * fed by the GEMM's per-32-row completion counters.  Bit-identical results; measured slower under the 1 kW power cap. */
int wb_debug_set_ln_follow(wb_model* m, int on);
int wb_debug_encode(const wb_model* m, const float* mel, size_t mel_len, int n_layers, int ln_post, float* out,
                    size_t out_capacity);
/* Per-kernel timing for bench.py's roofline: while enabled, every launch of the model's path is bracketed by CUDA events on
 * the launching stream; wb_profile_read synchronises and returns the summed milliseconds / launch counts per category
 * {0 mel_stft, 1 mel_finalize, 2 gemm, 3 attention, 4 layernorm, 5 other} and clears the record. */
int wb_profile_enable(wb_model* m, int on);
int wb_profile_read(wb_model* m, float* ms_by_cat, int* launches_by_cat, int n_cat);
/* Decoder test hooks: logits [n_vocab] of Decoder::forward_one after feeding `tokens` (before suppression) for one chunk's states
 * [seq_len][d]; and K, V ([seq_len][d] f32 each) of one decoder layer's cross-attention. */
int wb_debug_decoder_logits(const wb_model* m, const float* states, size_t seq_len, const int32_t* tokens, int n_tokens, float* logits_out);
int wb_debug_cross_kv(const wb_model* m, const float* states, size_t seq_len, int layer, float* k_out, float* v_out);
/* Number of CUDA kernels this library has launched in this process (all models). */
long long wb_launch_count(void);
/* Payload bytes AprReader::load_tensor would read for `name` (src/format/mod.rs:610-628): >= 0, -1 when the tensor is absent or its
 * descriptor points outside the file ("tensor data out of bounds"), -2 when the container does not parse.  No GPU needed. */
long long wb_debug_apr_tensor_bytes(const uint8_t* bytes, size_t n_bytes, const char* name);
/* LayerNorm rows: x f32 [rows][d] -> out f32 (exact f32 result) */
int wb_debug_layernorm(int device, const float* x, const float* gamma, const float* beta, int rows, int d, float* out);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */
