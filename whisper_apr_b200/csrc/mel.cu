// Fused log-mel front end: framing + periodic Hann + 400-point real FFT + |X|^2 + mel filterbank + log10,
// with a per-chunk running maximum; then the Whisper clamp/scale and the frame-axis padding.
//
// Replaces MelFilterbank::compute (src/audio/mel.rs:233-310), the padding rules of WhisperApr::compute_mel (src/lib.rs:407-443)
// and -- in its ragged form -- BatchPreprocessor::process_batch's per-segment loop (src/audio/batch.rs:157-176).  Layout is
// frame-major [frame][mel] as the reference stores it (mel.rs:298).  All arithmetic is fp32, as in the reference.
//
// Kernel 1 (mel_stft_kernel), round-2 shape: a FRAME NEVER LEAVES ITS WARP.  A CTA of 8 warps owns a tile of 32 consecutive frames
// of one segment; every warp owns 4 of them, 8 lanes per frame, and runs the whole chain on its own with __syncwarp only:
//   load   lane (frame, n2) pulls its 25 sample pairs straight from global memory (coalesced 64 B runs per frame; the 2.5x overlap
//          between neighbouring frames is served by L1) -- no CTA-wide staging buffer, no staging barrier;
//   step A the real frame is packed into a 200-point complex sequence, 200 = 8 x 25: 25-point DFT (two layers of radix-5 butterflies)
//          in registers, then the W200 twiddle, then ONE exchange through a warp-private 6.4 KB shared-memory slab (bank-conflict
//          free for both access patterns at a frame stride of 200 complex values);
//   step B 13 tasks per frame on its 8 lanes (two rounds): two 8-point DFTs (columns k1 and 25-k1) whose outputs are exactly the
//          conjugate-symmetric partners the real-FFT split needs; the power spectrum is formed in registers and written back over
//          the slab;
//   mel    (frame, mel) pairs over the warp's lanes: filterbank over each row's non-zero span only (k ascending like the
//          reference's scalar loop), log10(max(., 1e-10)), coalesced store, running max -> one atomicMax per warp on an
//          order-preserving integer key.
// The round-1 kernel staged all 32 frames' samples and all 32 z rows per CTA (93 KB) and crossed five __syncthreads per tile:
// 16 resident warps per SM that all waited at the same barriers.  Here warps run out of phase with each other (one loads while
// another does butterflies) and the slab is the only per-frame shared memory.
// Segments are described by a MelBatch: fixed-size chunks at a stride, or -- through per-segment offset / length / row tables and
// a tile table -- ragged segments and overlapping chunk VIEWS into longer streams (split_into_chunks without materialising chunks).
// Kernel 2 (mel_finalize_kernel): max(x, gmax-8), (x+4)/4, -1.0 for frames past the computed ones, written as f32 [B][T][m]
//   (API result) and/or bf16 [B][T+2][m] INCLUDING its two zero guard rows (the conv1 GEMM's operand); the block that finishes a
//   chunk last re-arms the chunk's maximum for the next call, so a fused mel step is two launches.
#include "fft400.cuh"
#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {
namespace {

constexpr int NFFT = 400;
constexpr int NFREQ = 201;
constexpr int FT = 32;                 // frames per CTA
constexpr int FW = 4;                  // frames per warp
constexpr int MEL_THREADS = 256;
constexpr int ZS = 200;                // complex values per frame in the exchange slab (2*ZS % 32 == 16: see header)

constexpr int SLAB_FLOATS = FW * ZS * 2;       // 1600 floats = 6.4 KB per warp; the powers ([204 bins][4 frames]) alias its front

__constant__ float2 c_tw25[25];        // exp(-2*pi*i*b*c/25) at [b*5+c]

constexpr int FB_MAX = 2048;           // packed filterbank weights kept in shared memory (triangular banks need ~400-600)
constexpr int MEL_MAX = 256;

struct MelSmem {
  float slab[MEL_THREADS / 32][SLAB_FLOATS];
  float w[NFFT];
  float2 tw200[8 * 25];                // exp(-2*pi*i*n2*k1/200) at [n2*25+k1]
  float2 tw400[NFREQ + 1];             // exp(-2*pi*i*k/400)
  __align__(16) float fbw[FB_MAX];     // packed non-zero spans of the filterbank rows, each padded to a multiple of 4 weights
  int fb_lo[MEL_MAX], fb_len[MEL_MAX], fb_off[MEL_MAX];
};

__device__ __forceinline__ int max_key(float v) {          // order-preserving float -> int
  int i = __float_as_int(v);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float key_to_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF);
}

struct MelKParams {
  const float* audio;
  long long audio_stride;
  const long long* seg_off;
  const int* n_valid;
  int n_valid_all;
  int hop;
  int n_frames;
  const int* n_frames_arr;
  const long long* row_off;
  const int2* tiles;
  int n_tiles, tiles_per_seg;          // total tiles of the launch; regular form: tiles per segment (tile ti = segment ti / tps)
  const float2* tw200_g;
  const float2* tw400_g;
  float* logmel;
  int* max_key;
};

__global__ void __launch_bounds__(MEL_THREADS, 3)
mel_stft_kernel(const MelKParams p, const MelTables tab) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- tables (once per CTA)
  for (int i = tid; i < NFFT; i += MEL_THREADS) s.w[i] = tab.window[i];
  for (int i = tid; i < 200; i += MEL_THREADS) s.tw200[i] = p.tw200_g[i];
  for (int i = tid; i <= NFREQ; i += MEL_THREADS) s.tw400[i] = p.tw400_g[i];
  const bool fb_smem = tab.packed != nullptr;
  if (fb_smem) {
    for (int i = tid; i < tab.nnz; i += MEL_THREADS) s.fbw[i] = __ldg(tab.packed + i);
    for (int i = tid; i < tab.n_mels; i += MEL_THREADS) {
      s.fb_lo[i] = __ldg(tab.packed_lo + i);
      s.fb_len[i] = __ldg(tab.packed_len + i);
      s.fb_off[i] = __ldg(tab.span_off + i);
    }
  }
  __syncthreads();                     // the only CTA-wide barrier

  // ---- persistent CTA: the tables above are loaded once, then the CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
  const int fl = lane >> 3, n2 = lane & 7;
  const int hop = p.hop;
  const int m = tab.n_mels;
  float* slab = s.slab[warp];
  float2* z = reinterpret_cast<float2*>(slab);
  for (int ti = blockIdx.x; ti < p.n_tiles; ti += gridDim.x) {
  int b, f0;
  if (p.tiles) {
    const int2 t = p.tiles[ti];
    b = t.x;
    f0 = t.y;
  } else {
    b = ti / p.tiles_per_seg;
    f0 = (ti - b * p.tiles_per_seg) * FT;
  }
  const int n_frames = p.n_frames_arr ? p.n_frames_arr[b] : p.n_frames;
  const int n_valid = p.n_valid ? p.n_valid[b] : p.n_valid_all;
  const float* seg = p.audio + (p.seg_off ? p.seg_off[b] : static_cast<long long>(b) * p.audio_stride);
  const long long row0 = p.row_off ? p.row_off[b] : static_cast<long long>(b) * p.n_frames;
  const int wf0 = f0 + warp * FW;                        // first frame of the warp
  const int wnf = min(FW, n_frames - wf0);               // live frames of the warp
  if (wnf <= 0) continue;
  const int f = wf0 + fl;                                // this lane group's frame
  const bool live = fl < wnf;

  // ---- load + window + step A: 25-point DFT over n1 of z[8*n1 + n2], then W200^(n2*k1)
  {
    cf v[25];
    const long long base = static_cast<long long>(f) * hop + 2 * n2;
    const float* fp = seg + base;
    // fast path (warp-uniform): every frame of the warp is live, lies inside the valid samples and is 8-byte aligned -> 25 plain
    // float2 loads at immediate offsets; otherwise per-element bounds checks (segment tails, odd view offsets, the padded zone)
    const bool easy = live && (base - 2 * n2 + NFFT <= n_valid) && ((reinterpret_cast<uintptr_t>(fp) & 7) == 0);
    if (__all_sync(0xffffffffu, easy)) {
      float2 xs[25];
#pragma unroll
      for (int n1 = 0; n1 < 25; ++n1) xs[n1] = __ldg(reinterpret_cast<const float2*>(fp + 16 * n1));
#pragma unroll
      for (int n1 = 0; n1 < 25; ++n1) {
        const float2 ws = *reinterpret_cast<const float2*>(&s.w[16 * n1 + 2 * n2]);
        v[n1] = cmulc(cmake(xs[n1].x, xs[n1].y), cmake(ws.x, ws.y));
      }
    } else {
      const bool vec = (reinterpret_cast<uintptr_t>(fp) & 7) == 0;
#pragma unroll
      for (int n1 = 0; n1 < 25; ++n1) {
        const long long i = base + 16 * n1;
        float2 xs = make_float2(0.f, 0.f);
        if (live) {
          if (vec && i + 1 < n_valid) {
            xs = __ldg(reinterpret_cast<const float2*>(seg + i));
          } else {
            if (i < n_valid) xs.x = __ldg(seg + i);
            if (i + 1 < n_valid) xs.y = __ldg(seg + i + 1);
          }
        }
        const float2 ws = *reinterpret_cast<const float2*>(&s.w[16 * n1 + 2 * n2]);
        v[n1] = cmulc(cmake(xs.x, xs.y), cmake(ws.x, ws.y));
      }
    }
    dft25(v, reinterpret_cast<const cf*>(c_tw25));
    float2* zrow = z + fl * ZS + n2 * 25;
#pragma unroll
    for (int k1 = 0; k1 < 25; ++k1) {
      const float2 t = s.tw200[n2 * 25 + k1];
      const cf r = cmul(v[k1], cmake(t.x, t.y));
      zrow[k1] = make_float2(r.x, r.y);
    }
  }
  __syncwarp();

  // ---- step B: 13 tasks per frame on its 8 lanes, two rounds; the powers stay in registers until every z read is done.
  // z holds Z / 2 (the W200 table is pre-halved), so partner bins k and 200-k share E and T (rfft_power_pair).
  float pw0[16], pw1[16];
  {
    const float2* zf = z + fl * ZS;
    {  // round 0: columns 0..7
      const int pcol = n2;
      cf a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float2 t = zf[q * 25 + pcol]; a[q] = cmake(t.x, t.y); }
      dft8(a);
      if (pcol == 0) {
        const float re0 = 2.0f * (a[0].x + a[0].y), re200 = 2.0f * (a[0].x - a[0].y);
        pw0[0] = re0 * re0;
        pw0[8] = re200 * re200;
#pragma unroll
        for (int k2 = 1; k2 < 4; ++k2) {                   // bins 25 k2 and 200 - 25 k2
          const float2 w = s.tw400[25 * k2];
          rfft_power_pair(a[k2], a[8 - k2], cmake(w.x, w.y), pw0[k2], pw0[8 - k2]);
        }
        {
          const float2 w = s.tw400[100];
          float dummy;
          rfft_power_pair(a[4], a[4], cmake(w.x, w.y), pw0[4], dummy);
        }
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2) pw0[8 + k2] = 0.f;
      } else {
        cf c[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float2 t = zf[q * 25 + 25 - pcol]; c[q] = cmake(t.x, t.y); }
        dft8(c);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {                   // bin k = pcol + 25 k2 and its partner 200 - k = (25 - pcol) + 25 (7 - k2)
          const float2 w = s.tw400[pcol + 25 * k2];
          rfft_power_pair(a[k2], c[7 - k2], cmake(w.x, w.y), pw0[k2], pw0[8 + 7 - k2]);
        }
      }
    }
    if (n2 < 5) {  // round 1: columns 8..12 on lanes 0..4 of the frame
      const int pcol = 8 + n2;
      cf a[8], c[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float2 t = zf[q * 25 + pcol]; a[q] = cmake(t.x, t.y); }
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float2 t = zf[q * 25 + 25 - pcol]; c[q] = cmake(t.x, t.y); }
      dft8(a);
      dft8(c);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        const float2 w = s.tw400[pcol + 25 * k2];
        rfft_power_pair(a[k2], c[7 - k2], cmake(w.x, w.y), pw1[k2], pw1[8 + 7 - k2]);
      }
    }
  }
  __syncwarp();                        // every lane has read what it needs from z: the power rows may overwrite the slab
  {
    // powers go back as [bin][frame of the warp]: one 16-byte load in the filterbank gives a bin of all 4 frames, and the 32 lanes of a
    // store instruction (4 frames x 8 columns) hit 32 consecutive words
    float* pf = slab + fl;
    if (n2 == 0) {
      pf[0] = pw0[0];
      pf[200 * FW] = pw0[8];
#pragma unroll
      for (int k2 = 1; k2 < 8; ++k2) pf[25 * k2 * FW] = pw0[k2];
    } else {
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        pf[(n2 + 25 * k2) * FW] = pw0[k2];
        pf[(25 - n2 + 25 * k2) * FW] = pw0[8 + k2];
      }
    }
    if (n2 < 5) {
      const int pcol = 8 + n2;
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        pf[(pcol + 25 * k2) * FW] = pw1[k2];
        pf[(25 - pcol + 25 * k2) * FW] = pw1[8 + k2];
      }
    }
    if (n2 == 7) { pf[201 * FW] = 0.f; pf[202 * FW] = 0.f; pf[203 * FW] = 0.f; }     // zero-weighted padding of a span reads up to bin 203
  }
  __syncwarp();

  // ---- filterbank over non-zero spans, log10, store, running max: (frame, mel) pairs of the warp's frames over its 32 lanes
  float lmax = -INFINITY;
  float* out = p.logmel + (row0 + wf0) * m;
  if (fb_smem) {
    // lane = mel row j (j = lane, lane + 32, ...); every weight of the row's non-zero span meets the bin of all 4 frames of the warp:
    // one 16-byte power load and two packed FMAs per weight, the weights in 16-byte groups (each span starts 16-byte aligned in the
    // packed table and is padded to a multiple of 4 with zeros; bins 201..203 are zero).  k ascending per output, f32 fused
    // multiply-adds, as the reference's loop (mel.rs:290-295); log10 = log2 * log10(2) (MUFU.LG2: absolute error ~1e-7 on values of
    // order 1-10, three orders below the 1e-4 gate).
    for (int j = lane; j < m; j += 32) {
      const int len = s.fb_len[j];
      const float4* fr = reinterpret_cast<const float4*>(&s.fbw[s.fb_off[j]]);
      const float4* pb = reinterpret_cast<const float4*>(slab) + s.fb_lo[j];
      cf e01 = cmake(0.f, 0.f), e23 = cmake(0.f, 0.f);
      for (int k = 0; k < len; k += 4) {
        const float4 w4 = fr[k >> 2];
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 p4 = pb[k + u];
          e01 = caxpy(wv[u], cmake(p4.x, p4.y), e01);
          e23 = caxpy(wv[u], cmake(p4.z, p4.w), e23);
        }
      }
      const float e[FW] = {e01.x, e01.y, e23.x, e23.y};
#pragma unroll
      for (int ff = 0; ff < FW; ++ff) {
        if (ff < wnf) {
          const float v = __log2f(fmaxf(e[ff], 1e-10f)) * 0.30102999566398120f;
          out[ff * m + j] = v;
          lmax = fmaxf(lmax, v);
        }
      }
    }
  } else {
    for (int idx = lane; idx < wnf * m; idx += 32) {
      const int ff = idx / m, j = idx - ff * m;
      const int lo = __ldg(tab.span_lo + j), len = __ldg(tab.span_len + j);
      const float* fr = tab.filters + j * NFREQ + lo;
      const float* pf = slab + lo * FW + ff;
      float e = 0.f;
      for (int k = 0; k < len; ++k) e += __ldg(fr + k) * pf[k * FW];      // k ascending, f32 (mel.rs:290-295)
      const float v = log10f(fmaxf(e, 1e-10f));
      out[idx] = v;
      lmax = fmaxf(lmax, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) atomicMax(p.max_key + b, max_key(lmax));          // one atomic per warp
  __syncwarp();                        // the next tile's z rows overwrite the power rows this warp has just read
  }
}

__global__ void mel_init_max_kernel(int* keys, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) keys[i] = max_key(-INFINITY);
}

// One thread per 4 consecutive values of the padded output INCLUDING the bf16 operand's two guard rows: row r of [0, T_out + 2)
// is a zero guard row for r == 0 and r == T_out + 1, frame r - 1 otherwise.  The last block of a chunk re-arms its max key.
__global__ void __launch_bounds__(256)
mel_finalize_kernel(const float* __restrict__ logmel, int* __restrict__ chunk_max_key, int n_frames, int T_out, int m,
                    float* __restrict__ out_f32, op16* __restrict__ out_bf16, unsigned int* __restrict__ done_counter) {
  const int b = blockIdx.y;
  const long long per_padded = static_cast<long long>(T_out + 2) * m;
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const float floor_v = key_to_float(chunk_max_key[b]) - 8.0f;
  if (i < per_padded) {
    const long long r = i / m;                       // m % 4 == 0: the four values share a row
    const bool guard = r == 0 || r == T_out + 1;
    const long long fi = i - m;                      // index into the [T_out][m] frame block
    const long long n_real = static_cast<long long>(min(n_frames, T_out)) * m;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (!guard) {
      if (fi < n_real) {                             // n_real is a multiple of m, hence of 4: the group is all real or all padding
        const float4 t = *reinterpret_cast<const float4*>(logmel + static_cast<long long>(b) * n_frames * m + fi);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (fmaxf(v[k], floor_v) + 4.0f) / 4.0f;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = -1.0f;     // lib.rs:431-437 pad value
      }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + static_cast<long long>(b) * T_out * m + fi) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (out_bf16) {
      uint2 w;
      w.x = pack_op16x2(v[0], v[1]);
      w.y = pack_op16x2(v[2], v[3]);
      *reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(b) * per_padded + i) = w;
    }
  }
  // re-arm: the block that finishes last for this chunk resets the key (every block has read it above)
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(done_counter + b, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    chunk_max_key[b] = max_key(-INFINITY);
    done_counter[b] = 0;
  }
}

// any n_mels: one value per thread
__global__ void __launch_bounds__(256)
mel_finalize_scalar_kernel(const float* __restrict__ logmel, const int* __restrict__ chunk_max_key, int n_frames, int T_out, int m,
                           float* __restrict__ out_f32) {
  const int b = blockIdx.y;
  const long long per_chunk = static_cast<long long>(T_out) * m;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per_chunk) return;
  const float floor_v = key_to_float(chunk_max_key[b]) - 8.0f;
  const long long n_real = static_cast<long long>(min(n_frames, T_out)) * m;
  float v = -1.0f;
  if (i < n_real) v = (fmaxf(logmel[static_cast<long long>(b) * n_frames * m + i], floor_v) + 4.0f) / 4.0f;
  out_f32[b * per_chunk + i] = v;
}

// ragged: segment b owns rows [row_off[b], row_off[b] + n_frames[b]); in place (out may alias logmel)
__global__ void __launch_bounds__(256)
mel_finalize_ragged_kernel(const float* logmel, const int* __restrict__ max_keys, const int* __restrict__ n_frames_arr,
                           const long long* __restrict__ row_off, int m, float* out) {
  const int b = blockIdx.y;
  const long long n = static_cast<long long>(n_frames_arr[b]) * m;
  const float floor_v = key_to_float(max_keys[b]) - 8.0f;
  const long long base = row_off[b] * m;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[base + i] = (fmaxf(logmel[base + i], floor_v) + 4.0f) / 4.0f;
}

// twiddle tables live on ONE device; a process that drives several GPUs gets a set per device (PerDeviceOnce)
float2* g_tw200[WB_MAX_DEVICES] = {};
float2* g_tw400[WB_MAX_DEVICES] = {};
PerDeviceOnce g_mel_once;

}  // namespace

int mel_init() {
  return g_mel_once.run([](int dev) -> int {
    const double PI = 3.14159265358979323846;
    float2 tw25[25], tw200[200], tw400[NFREQ + 1];
    for (int bb = 0; bb < 5; ++bb)
      for (int c = 0; c < 5; ++c) {
        double a = -2.0 * PI * bb * c / 25.0;
        tw25[bb * 5 + c] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
      }
    for (int n2 = 0; n2 < 8; ++n2)
      for (int k1 = 0; k1 < 25; ++k1) {
        double a = -2.0 * PI * n2 * k1 / 200.0;
        tw200[n2 * 25 + k1] = make_float2(0.5f * static_cast<float>(cos(a)), 0.5f * static_cast<float>(sin(a)));   // pre-halved: rfft_power_pair
      }
    for (int k = 0; k <= NFREQ; ++k) {
      double a = -2.0 * PI * k / 400.0;
      tw400[k] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }
    WB_CUDA_OK(cudaMemcpyToSymbol(c_tw25, tw25, sizeof tw25));
    WB_CUDA_OK(cudaMalloc(&g_tw200[dev], sizeof tw200));
    WB_CUDA_OK(cudaMalloc(&g_tw400[dev], sizeof tw400));
    WB_CUDA_OK(cudaMemcpy(g_tw200[dev], tw200, sizeof tw200, cudaMemcpyHostToDevice));
    WB_CUDA_OK(cudaMemcpy(g_tw400[dev], tw400, sizeof tw400, cudaMemcpyHostToDevice));
    WB_CUDA_OK(cudaFuncSetAttribute(mel_stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem)));
    return WB_OK;
  });
}

int launch_mel_init_keys(int* keys, int B, cudaStream_t stream) {
  if (B <= 0) return WB_OK;
  mel_init_max_kernel<<<(B + 255) / 256, 256, 0, stream>>>(keys, B);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_mel_stft(const MelBatch& job, const MelTables& t, float* logmel, int* chunk_max_key, cudaStream_t stream) {
  int rc = mel_init();
  if (rc != WB_OK) return rc;
  if (job.B <= 0) return WB_OK;
  const int dev = current_device();
  MelKParams p;
  p.audio = job.audio; p.audio_stride = job.audio_stride; p.seg_off = job.seg_off; p.n_valid = job.n_valid; p.n_valid_all = job.n_valid_all;
  p.hop = job.hop; p.n_frames = job.n_frames; p.n_frames_arr = job.n_frames_arr; p.row_off = job.row_off; p.tiles = job.tiles;
  p.tw200_g = g_tw200[dev]; p.tw400_g = g_tw400[dev]; p.logmel = logmel; p.max_key = chunk_max_key;
  if (job.tiles) {
    if (job.n_tiles <= 0) return WB_OK;
    p.n_tiles = job.n_tiles;
    p.tiles_per_seg = 0;
  } else {
    if (job.n_frames <= 0) return WB_OK;
    p.tiles_per_seg = (job.n_frames + FT - 1) / FT;
    p.n_tiles = p.tiles_per_seg * job.B;
  }
  // persistent: 3 resident CTAs per SM (launch bounds), each walks ceil(n_tiles / grid) tiles with the tables loaded once
  const int resident = device_sm_count() * 3;
  const int per_cta = (p.n_tiles + resident - 1) / resident;
  const int grid = (p.n_tiles + per_cta - 1) / per_cta;
  mel_stft_kernel<<<grid, MEL_THREADS, sizeof(MelSmem), stream>>>(p, t);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_mel_finalize(const float* logmel, int* chunk_max_key, unsigned int* done_counter, int n_frames, int T_out, int n_mels, int B,
                        float* out_f32, op16* out_bf16_padded, cudaStream_t stream) {
  if (B <= 0 || T_out <= 0) return WB_OK;
  if (n_mels % 4 == 0) {
    const long long per_padded = static_cast<long long>(T_out + 2) * n_mels;
    dim3 grid(static_cast<unsigned>((per_padded / 4 + 255) / 256), B);
    mel_finalize_kernel<<<grid, 256, 0, stream>>>(logmel, chunk_max_key, n_frames, T_out, n_mels, out_f32, out_bf16_padded, done_counter);
    count_launch();
  } else {
    if (out_bf16_padded || !out_f32) return set_error(WB_ERR_MODEL, "the bf16 conv operand needs n_mels % 4 == 0");
    const long long per_chunk = static_cast<long long>(T_out) * n_mels;
    dim3 grid(static_cast<unsigned>((per_chunk + 255) / 256), B);
    mel_finalize_scalar_kernel<<<grid, 256, 0, stream>>>(logmel, chunk_max_key, n_frames, T_out, n_mels, out_f32);
    count_launch();
    const int rc = launch_mel_init_keys(chunk_max_key, B, stream);     // re-arm for the next call
    if (rc != WB_OK) return rc;
  }
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_mel_finalize_ragged(const float* logmel, const int* max_keys, const int* n_frames_arr, const long long* row_off, int n_mels, int B,
                               long long max_rows, float* out, cudaStream_t stream) {
  if (B <= 0 || max_rows <= 0) return WB_OK;
  long long g = (max_rows * n_mels + 255) / 256;
  if (g > 1024) g = 1024;
  dim3 grid(static_cast<unsigned>(g), B);
  mel_finalize_ragged_kernel<<<grid, 256, 0, stream>>>(logmel, max_keys, n_frames_arr, row_off, n_mels, out);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

}  // namespace wb

// Host restatement of the kernel's FFT index algebra (same fft400.cuh code path), used by the CPU unit tests.
extern "C" void wb_debug_fft400_power_host(const float* y, float* p) {
  using namespace wb;
  const double PI = 3.14159265358979323846;
  cf tw25[25];
  for (int b = 0; b < 5; ++b)
    for (int c = 0; c < 5; ++c) {
      double a = -2.0 * PI * b * c / 25.0;
      tw25[b * 5 + c] = cmake(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }
  cf z[200];   // Y'[n2][k1]
  for (int n2 = 0; n2 < 8; ++n2) {
    cf v[25];
    for (int n1 = 0; n1 < 25; ++n1) v[n1] = cmake(y[16 * n1 + 2 * n2], y[16 * n1 + 2 * n2 + 1]);
    dft25(v, tw25);
    for (int k1 = 0; k1 < 25; ++k1) {
      double a = -2.0 * PI * n2 * k1 / 200.0;
      z[n2 * 25 + k1] = cmul(v[k1], cmake(0.5f * static_cast<float>(cos(a)), 0.5f * static_cast<float>(sin(a))));   // Z / 2, as the kernel
    }
  }
  auto tw400 = [&](int k) {
    double a = -2.0 * PI * k / 400.0;
    return cmake(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
  };
  for (int pcol = 0; pcol < 13; ++pcol) {
    cf a[8];
    for (int n2 = 0; n2 < 8; ++n2) a[n2] = z[n2 * 25 + pcol];
    dft8(a);
    if (pcol == 0) {
      float r0 = 2.0f * (a[0].x + a[0].y), r200 = 2.0f * (a[0].x - a[0].y);
      p[0] = r0 * r0;
      p[200] = r200 * r200;
      for (int k2 = 1; k2 < 4; ++k2) rfft_power_pair(a[k2], a[8 - k2], tw400(25 * k2), p[25 * k2], p[200 - 25 * k2]);
      float dummy;
      rfft_power_pair(a[4], a[4], tw400(100), p[100], dummy);
    } else {
      cf c[8];
      for (int n2 = 0; n2 < 8; ++n2) c[n2] = z[n2 * 25 + 25 - pcol];
      dft8(c);
      for (int k2 = 0; k2 < 8; ++k2) {
        int k = pcol + 25 * k2;
        rfft_power_pair(a[k2], c[7 - k2], tw400(k), p[k], p[200 - k]);
      }
    }
  }
}
