// Internal declarations shared by the translation units of libwhisper_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/whisper_b200.h"

namespace wb {

// ---- the 16-bit operand format of every tensor-core product (GEMM A / W, attention Q / K / V / P) and of every activation that feeds
// one.  IEEE fp16 by default: both formats run tcgen05.mma kind::f16 at the same rate and move the same bytes, but fp16 keeps an 11-bit
// significand against bf16's 8.  With bf16 the weight rounding alone costs 1.5e-2 of the 2e-2 max-abs parity gate after 32 layers
// (profiles/r02_bf16_error_budget.txt: three chunks measured 1.91e-2 .. 2.03e-2); Whisper's value ranges sit far inside fp16's
// (fp16 is the reference implementations' own inference format).  -DWB_OPERANDS_BF16 builds the bf16 variant (the A and B formats of a
// kind::f16 MMA cannot be mixed: bf16 x fp16 raises an illegal-instruction fault on the B200).
#ifdef WB_OPERANDS_BF16
typedef __nv_bfloat16 op16;
constexpr uint32_t kOp16Format = 1;        // F16F32Format::BF16 of the instruction descriptor
constexpr bool kOp16IsFp16 = false;
__host__ __device__ inline float op16_to_float(op16 v) { return __bfloat162float(v); }
__host__ __device__ inline op16 float_to_op16(float v) { return __float2bfloat16_rn(v); }
#else
typedef __half op16;
constexpr uint32_t kOp16Format = 0;        // F16F32Format::F16
constexpr bool kOp16IsFp16 = true;
__host__ __device__ inline float op16_to_float(op16 v) { return __half2float(v); }
__host__ __device__ inline op16 float_to_op16(float v) { return __float2half_rn(v); }
#endif

// ---- error plumbing (mirrors WhisperError::{Audio,Model,Format}, src/error.rs:6-44)
int set_error(int status, const std::string& msg);   // returns status
const char* last_error();
#define WB_CUDA_OK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::wb::set_error(WB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
  } while (0)

// ---- per-device one-time initialisation.  Kernel attributes (opt-in shared memory), __constant__ tables and small device
// tables belong to ONE device's context: a process that drives several GPUs (wb_model_from_apr_devices, or two models on two
// ordinals) must set them up on each.  `run(fn)` calls fn() once per CUDA device, on the calling thread's current device.
constexpr int WB_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= WB_MAX_DEVICES) return 0;
  return dev;
}
struct PerDeviceOnce {
  std::mutex mu;
  std::atomic<bool> done[WB_MAX_DEVICES];
  PerDeviceOnce() {
    for (auto& d : done) d.store(false);
  }
  template <typename F>
  int run(F&& fn) {
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return WB_OK;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev].load(std::memory_order_relaxed)) return WB_OK;
    const int rc = fn(dev);
    if (rc == WB_OK) done[dev].store(true, std::memory_order_release);
    return rc;
  }
};
int device_sm_count();   // SM count of the current device (cached per device)

// every kernel launch of the library bumps this (bench.py reports it as gpu_launches)
extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// ---- GEMM:  out[M][N] = epilogue(alpha * A[M][K] . W[N][K]^T + bias)      (tcgen05 / TMEM / TMA)
enum GemmEpilogue : int {
  EPI_BF16 = 0,        // out bf16 = acc*alpha + bias
  EPI_GELU_BF16 = 1,   // out bf16 = gelu(acc*alpha + bias)
  EPI_RESID_F32 = 2,   // out f32 += acc*alpha + bias            (residual stream, in place)
  EPI_GELU_PE_F32 = 3, // out f32 = gelu(acc*alpha + bias) + pe[row_in_batch][n]
  EPI_F32 = 4,         // out f32 = acc*alpha + bias             (debug / f32 consumers)
};

struct GemmDesc {
  // A operand: bf16, viewed as [n_batch][rows_per_batch][K] with arbitrary (16 B aligned) strides.
  const op16* A;
  long long a_row_stride;     // elements between consecutive rows
  long long a_batch_stride;   // elements between consecutive batch entries
  int rows_per_batch;
  int n_batch;
  // W operand: bf16 [N][K] row-major (the reference's [out][in] layout, attention.rs:33-34)
  const op16* W;
  int N, K;
  // epilogue
  int epilogue;
  float alpha;
  const float* col_scale;     // [N] per-output-column scale or nullptr (int8/int4 weights)
  const float* bias;          // [N] or nullptr
  void* out;                  // row r of batch b lands at row b*out_rows_per_batch + out_row_off + r
  long long ldc;              // output row pitch, elements
  int out_rows_per_batch;
  int out_row_off;
  const float* pe;            // EPI_GELU_PE_F32: [>=rows_per_batch][N] f32
  // EPI_RESID_F32 with n_batch == 1, optional: ready[r / 32] is incremented once per column tile when that tile's reductions into rows
  // [32 (r / 32), +32) have been performed -- a kernel running BESIDE the GEMM (layernorm_follow) picks a row group up as soon as its
  // counter reaches gemm_tiles_n(N) and finds the rows in L2.  nullptr: no signalling.
  unsigned int* ready = nullptr;
  // walk the row tiles from the last to the first: a kernel that runs in the OPPOSITE direction of its producer starts on the rows
  // the producer wrote last, i.e. the ones still in the 126 MB L2 (pipeline.cu alternates the direction from kernel to kernel)
  bool reverse = false;
  // EPI_BF16 / EPI_GELU_BF16: store the 16-bit output as bf16 even in an fp16-operand build (the QKV product feeding a bf16 attention)
  bool out_bf16 = false;
};
int gemm_tiles_n(int N);      // arrivals on ready[] per row group for a residual GEMM of N columns (one per column tile)
int launch_gemm(const GemmDesc& g, cudaStream_t stream);
int gemm_init();   // resolves cuTensorMapEncodeTiled, sets kernel attributes
int make_tmap_op16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1);

// ---- attention: qkv bf16 [B][S][3d] (q | k | v column blocks) -> out bf16 [B][S][d]
// qkv_bf16: q, k, v hold bf16 whatever the build's operand format is (GemmDesc::out_bf16 on the QKV GEMM); out is always op16
int launch_attention(const op16* qkv, op16* out, int B, int S, int d, int n_heads, cudaStream_t stream, bool reverse = false, bool qkv_bf16 = false);
int attention_init();

// ---- mel
struct MelTables {                 // device-resident, built at model load
  const float* window;             // [400] periodic Hann, f32 (mel.rs:215-219)
  const float* filters;            // [n_mels][201] dense, row-major (mel.rs:96-139)
  const int* span_lo;              // [n_mels] first non-zero bin
  const int* span_len;             // [n_mels] bins from first to last non-zero (0 if all-zero row)
  int n_mels;
  // the same spans packed back to back for shared memory: row j covers bins [packed_lo[j], packed_lo[j] + packed_len[j]) with
  // packed_lo = span_lo and packed_len = span_len rounded up to a multiple of 4 (zero weights behind the real span; the kernel keeps
  // power bins 201..203 at zero), stored at packed[span_off[j] ..] with span_off a multiple of 4: weights are read in 16-byte groups.
  // Triangular banks need ~0.6 k (slaney-80) to ~0.8 k (slaney-128) weights; nullptr when the bank does not fit (nnz > 2048 or
  // n_mels > 256) -> dense rows from global memory.
  const float* packed;
  const int* span_off;
  const int* packed_lo;
  const int* packed_len;
  int nnz;
};
// What one mel launch processes.  Regular form: B segments at `audio + b * audio_stride`, n_frames frames each, log-mel rows
// b * n_frames ...  Ragged / view form: per-segment device tables -- seg_off (element offset of the segment's first sample: lets
// chunks be overlapping VIEWS into longer streams), n_valid (samples that exist; later ones read as 0: compute_mel's zero padding),
// n_frames_arr, row_off (first log-mel row) -- and a tile table {segment, first frame} with one entry per 32-frame tile.
struct MelBatch {
  const float* audio = nullptr;
  long long audio_stride = 0;
  const long long* seg_off = nullptr;
  const int* n_valid = nullptr;
  int n_valid_all = 0;
  int hop = 160;
  int B = 0;
  int n_frames = 0;
  const int* n_frames_arr = nullptr;
  const long long* row_off = nullptr;
  const int2* tiles = nullptr;
  int n_tiles = 0;
};
// logmel[row][j] = log10(max(E,1e-10)); chunk_max_key[b] = ordered-int max (must hold key(-inf) on entry: launch_mel_init_keys once,
// launch_mel_finalize re-arms it).
int launch_mel_stft(const MelBatch& job, const MelTables& t, float* logmel, int* chunk_max_key, cudaStream_t stream);
int launch_mel_init_keys(int* keys, int B, cudaStream_t stream);
// clamp/scale/pad: out_f32 [B][T_out][m] (optional), out_bf16 [B][T_out+2][m] with its two zero guard rows (optional).
// done_counter: [B] zero-initialised scratch of the "last block re-arms the key" protocol.
int launch_mel_finalize(const float* logmel, int* chunk_max_key, unsigned int* done_counter, int n_frames, int T_out, int n_mels, int B,
                        float* out_f32, op16* out_bf16_padded, cudaStream_t stream);
// ragged segments, in place allowed (out may alias logmel); keys are NOT re-armed
int launch_mel_finalize_ragged(const float* logmel, const int* max_keys, const int* n_frames_arr, const long long* row_off, int n_mels, int B,
                               long long max_rows, float* out, cudaStream_t stream);
int mel_init();

// ---- elementwise / normalisation
int launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int d, void* out_16, bool out_16_is_bf16,
                     float* out_f32, cudaStream_t stream, bool reverse = false);
// LayerNorm that FOLLOWS a residual GEMM running on another stream: row group g (32 rows) is normalised as soon as ready[g] == need
// (GemmDesc::ready), read through L2, and the counter is re-armed to 0.  Same arithmetic as launch_layernorm (op16 output only).
// wait = true: the concurrent form (gives up after 20 ms without progress); wait = false: the sweep behind the GEMM that normalises
// every group still carrying a full counter.  Launch both: together they cover every row under any kernel schedule.
int launch_layernorm_follow(const float* x, const float* gamma, const float* beta, int rows, int d, op16* out, unsigned int* ready,
                            unsigned int need, bool wait, cudaStream_t stream);
int launch_f32_to_op16(const float* in, op16* out, size_t n, cudaStream_t stream);
int launch_i8_to_op16(const int8_t* in, op16* out, size_t n, cudaStream_t stream);
int launch_i4_to_op16(const uint8_t* in, op16* out, size_t n, cudaStream_t stream);
int launch_i8_to_f32(const int8_t* in, float scale, float* out, size_t n, cudaStream_t stream);
int launch_i4_to_f32(const uint8_t* in, float scale, float* out, size_t n, cudaStream_t stream);
// conv weight [out][in][3] (any of f32/int8/int4 already expanded to bf16) -> [out][3][in]
int launch_conv_repack(const op16* in, op16* out, int c_out, int c_in, cudaStream_t stream);
// mel f32 [B][T][m] -> bf16 [B][T+2][m] with zero rows 0 and T+1
int launch_mel_pad_op16(const float* mel, op16* out, int B, int T, int m, cudaStream_t stream);
int launch_fill_op16_rows(op16* base, long long batch_stride, int B, int row_elems, cudaStream_t stream);

}  // namespace wb
