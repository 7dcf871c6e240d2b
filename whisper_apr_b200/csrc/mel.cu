// Fused log-mel front end: framing + periodic Hann + 400-point real FFT + |X|^2 + mel filterbank + log10,
// with a per-chunk running maximum; then the Whisper clamp/scale and the frame-axis padding.
//
// Replaces MelFilterbank::compute (src/audio/mel.rs:233-310) and the padding rules of
// WhisperApr::compute_mel (src/lib.rs:407-443).  Layout is frame-major [frame][mel] as the reference stores it
// (mel.rs:298).  All arithmetic is fp32, as in the reference.
//
// Kernel 1 (mel_stft_kernel): one CTA per tile of 32 consecutive frames of one chunk.
//   * the (31*hop + 400) samples the tile needs are staged in shared memory with coalesced 16-byte loads
//     (each sample is read from HBM once per tile although it belongs to 2.5 frames),
//   * the real frame is packed into a 200-point complex sequence; 200 = 8 x 25:
//       step A  8 threads per frame, each a 25-point DFT (two layers of radix-5 butterflies) in registers,
//               followed by the W200 twiddle, exchanged through shared memory;
//       step B  13 tasks per frame, each two 8-point DFTs (columns k1 and 25-k1) whose outputs are exactly
//               the conjugate-symmetric partners the real-FFT split needs, so the power spectrum is formed in
//               registers and only P[0..200] goes back to shared memory,
//   * the filterbank is applied over each mel row's non-zero span only (391 of 16080 weights for the slaney-80
//     bank), k ascending like the reference's scalar loop; log10(max(.,1e-10)); coalesced store; block max
//     -> one atomicMax per CTA on an order-preserving integer key.
// Kernel 2 (mel_finalize_kernel): max(x, gmax-8), (x+4)/4, -1.0 for frames past the computed ones, written as f32
//   [B][T][m] (API result) and/or bf16 [B][T+2][m] with zero guard rows (the conv1 GEMM's operand).
#include "fft400.cuh"
#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {
namespace {

constexpr int NFFT = 400;
constexpr int NFREQ = 201;
constexpr int FT = 32;                 // frames per CTA
constexpr int MEL_THREADS = 256;
constexpr int MAX_TILE = (FT - 1) * 160 + NFFT;          // 5360 samples for hop <= 160
constexpr int SX_FLOATS = MAX_TILE + 16 * (FT + 3);      // + skew
constexpr int ZSTRIDE = 200;           // complex per frame
constexpr int PSTRIDE = 201;

__constant__ float2 c_tw25[25];        // exp(-2*pi*i*b*c/25) at [b*5+c]

constexpr int FB_MAX = 2048;           // packed filterbank weights kept in shared memory (triangular banks need ~400-600)
constexpr int MEL_MAX = 256;

struct MelSmem {
  union {                              // the staged samples are dead once step A has run; the power spectrum takes their place
    float x[SX_FLOATS];
    float p[FT * PSTRIDE + 4];         // + 4: the zero-weighted padding of the last span may read up to 3 floats past bin 200
  };
  float w[NFFT];
  float2 tw200[8 * 25];                // exp(-2*pi*i*n2*k1/200) at [n2*25+k1]
  float2 tw400[NFREQ + 1];             // exp(-2*pi*i*k/400)
  float2 z[FT * ZSTRIDE];
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

template <int HOP>   // HOP == 160: skewed, conflict-free staging; HOP == 0: any hop, unskewed
__device__ __forceinline__ int sx_index(int i, int hop) {
  if (HOP == 160) return i + 16 * (i / 160);
  return i;
}

template <int HOP>
__global__ void __launch_bounds__(MEL_THREADS, 2)
mel_stft_kernel(const float* __restrict__ audio, long long audio_stride, const int* __restrict__ n_valid_arr, int n_valid_all,
                int hop_rt, int n_frames, int frames_per_tile, MelTables tab, const float2* __restrict__ tw200_g,
                const float2* __restrict__ tw400_g, float* __restrict__ logmel, int* __restrict__ chunk_max_key) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
  const int hop = (HOP == 160) ? 160 : hop_rt;
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * frames_per_tile;
  const int nf = min(frames_per_tile, n_frames - f0);
  const int n_valid = n_valid_arr ? n_valid_arr[b] : n_valid_all;
  const float* chunk = audio + static_cast<long long>(b) * audio_stride;
  const long long s0 = static_cast<long long>(f0) * hop;
  const int tile_len = (nf - 1) * hop + NFFT;

  // ---- stage samples, window, twiddles
  if (HOP == 160) {
    for (int i4 = tid; i4 < (tile_len >> 2); i4 += MEL_THREADS) {
      const int i = i4 << 2;
      const long long g = s0 + i;
      float4 v;
      if (g + 3 < n_valid) {
        v = __ldg(reinterpret_cast<const float4*>(chunk + g));
      } else {
        v.x = g + 0 < n_valid ? chunk[g + 0] : 0.f;
        v.y = g + 1 < n_valid ? chunk[g + 1] : 0.f;
        v.z = g + 2 < n_valid ? chunk[g + 2] : 0.f;
        v.w = g + 3 < n_valid ? chunk[g + 3] : 0.f;
      }
      *reinterpret_cast<float4*>(&s.x[sx_index<HOP>(i, hop)]) = v;    // 160 % 4 == 0: a float4 never straddles a skew step
    }
  } else {
    for (int i = tid; i < tile_len; i += MEL_THREADS) {
      const long long g = s0 + i;
      s.x[i] = g < n_valid ? chunk[g] : 0.f;
    }
  }
  for (int i = tid; i < NFFT; i += MEL_THREADS) s.w[i] = tab.window[i];
  for (int i = tid; i < 200; i += MEL_THREADS) s.tw200[i] = tw200_g[i];
  for (int i = tid; i <= NFREQ; i += MEL_THREADS) s.tw400[i] = tw400_g[i];
  const bool fb_smem = tab.packed != nullptr;
  if (fb_smem) {
    for (int i = tid; i < tab.nnz; i += MEL_THREADS) s.fbw[i] = __ldg(tab.packed + i);
    for (int i = tid; i < tab.n_mels; i += MEL_THREADS) {
      s.fb_lo[i] = __ldg(tab.span_lo + i);
      s.fb_len[i] = __ldg(tab.span_len + i);
      s.fb_off[i] = __ldg(tab.span_off + i);
    }
  }
  __syncthreads();

  // ---- step A: thread (f, n2): 25-point DFT over n1 of z[8*n1 + n2], then W200^(n2*k1)
  {
    const int f = tid >> 3, n2 = tid & 7;
    if (f < nf) {
      cf v[25];
      const int base = f * hop + 2 * n2;
#pragma unroll
      for (int n1 = 0; n1 < 25; ++n1) {
        const int e = 16 * n1 + 2 * n2;            // sample index inside the frame (even)
        const int i = base + 16 * n1;
        float2 xs;
        if (HOP == 160) {
          xs = *reinterpret_cast<const float2*>(&s.x[sx_index<HOP>(i, hop)]);
        } else {
          xs.x = s.x[i];
          xs.y = s.x[i + 1];
        }
        const float2 ws = *reinterpret_cast<const float2*>(&s.w[e]);
        v[n1] = cmake(xs.x * ws.x, xs.y * ws.y);
      }
      dft25(v, reinterpret_cast<const cf*>(c_tw25));
      float2* zrow = &s.z[f * ZSTRIDE + n2 * 25];
#pragma unroll
      for (int k1 = 0; k1 < 25; ++k1) {
        const float2 t = s.tw200[n2 * 25 + k1];
        const cf r = cmul(v[k1], cmake(t.x, t.y));
        zrow[k1] = make_float2(r.x, r.y);
      }
    }
  }
  __syncthreads();

  // ---- step B: task (f, p): 8-point DFTs of columns p and 25-p -> power spectrum in registers
  for (int task = tid; task < nf * 13; task += MEL_THREADS) {
    const int f = task / 13, pcol = task - f * 13;
    const float2* zf = &s.z[f * ZSTRIDE];
    float* pf = &s.p[f * PSTRIDE];
    cf a[8];
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) { const float2 t = zf[n2 * 25 + pcol]; a[n2] = cmake(t.x, t.y); }
    dft8(a);
    if (pcol == 0) {
      const float re0 = a[0].x + a[0].y, re200 = a[0].x - a[0].y;
      pf[0] = re0 * re0;
      pf[200] = re200 * re200;
#pragma unroll
      for (int k2 = 1; k2 < 8; ++k2) {
        const float2 w = s.tw400[25 * k2];
        pf[25 * k2] = rfft_power(a[k2], a[8 - k2], cmake(w.x, w.y));
      }
    } else {
      cf c[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) { const float2 t = zf[n2 * 25 + 25 - pcol]; c[n2] = cmake(t.x, t.y); }
      dft8(c);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        const int k = pcol + 25 * k2;
        const int kk = 25 - pcol + 25 * k2;
        const float2 w1 = s.tw400[k], w2 = s.tw400[kk];
        pf[k] = rfft_power(a[k2], c[7 - k2], cmake(w1.x, w1.y));
        pf[kk] = rfft_power(c[k2], a[7 - k2], cmake(w2.x, w2.y));
      }
    }
  }
  if (tid < 4) s.p[FT * PSTRIDE + tid] = 0.f;      // the zero-weighted padding of the last span reads up to 3 floats past bin 200 of the last frame
  __syncthreads();

  // ---- filterbank over non-zero spans, log10, store, running max
  const int m = tab.n_mels;
  float lmax = -INFINITY;
  float* out = logmel + (static_cast<long long>(b) * n_frames + f0) * m;
  if (fb_smem) {
    // (frame, mel) pairs walked without a division per item; weights and spans from shared memory; log10 = log2 * log10(2)
    // (MUFU.LG2: absolute error ~1e-7 on values of order 1-10, three orders below the 1e-4 gate)
    int f = tid / m, j = tid - f * m;
    const int df = MEL_THREADS / m, dj = MEL_THREADS - df * m;
    for (int idx = tid; idx < nf * m; idx += MEL_THREADS) {
      const int len = s.fb_len[j];
      const float* fr = &s.fbw[s.fb_off[j]];
      const float* pf = &s.p[f * PSTRIDE + s.fb_lo[j]];
      float e = 0.f;
      for (int k = 0; k < len; k += 4) {                          // k ascending, f32 (mel.rs:290-295); padding weights are 0
        const float4 w4 = *reinterpret_cast<const float4*>(fr + k);
        e += w4.x * pf[k];
        e += w4.y * pf[k + 1];
        e += w4.z * pf[k + 2];
        e += w4.w * pf[k + 3];
      }
      const float v = __log2f(fmaxf(e, 1e-10f)) * 0.30102999566398120f;
      out[idx] = v;
      lmax = fmaxf(lmax, v);
      f += df;
      j += dj;
      if (j >= m) { j -= m; ++f; }
    }
  } else {
    for (int idx = tid; idx < nf * m; idx += MEL_THREADS) {
      const int f = idx / m, j = idx - f * m;
      const int lo = __ldg(tab.span_lo + j), len = __ldg(tab.span_len + j);
      const float* fr = tab.filters + j * NFREQ + lo;
      const float* pf = &s.p[f * PSTRIDE + lo];
      float e = 0.f;
      for (int k = 0; k < len; ++k) e += __ldg(fr + k) * pf[k];      // k ascending, f32 (mel.rs:290-295)
      const float v = log10f(fmaxf(e, 1e-10f));
      out[idx] = v;
      lmax = fmaxf(lmax, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0) atomicMax(chunk_max_key + b, max_key(lmax));      // one atomic per warp: no block barrier at the tail
}

__global__ void mel_init_max_kernel(int* keys, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) keys[i] = max_key(-INFINITY);
}

// one thread per 4 consecutive (frame, mel) values of the padded output
__global__ void __launch_bounds__(256)
mel_finalize_kernel(const float* __restrict__ logmel, const int* __restrict__ chunk_max_key, int n_frames, int T_out, int m,
                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.y;
  const long long per_chunk = static_cast<long long>(T_out) * m;
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= per_chunk) return;
  const float floor_v = key_to_float(chunk_max_key[b]) - 8.0f;
  const long long n_real = static_cast<long long>(min(n_frames, T_out)) * m;
  float v[4];
  if (i + 3 < n_real) {
    const float4 t = *reinterpret_cast<const float4*>(logmel + static_cast<long long>(b) * n_frames * m + i);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (fmaxf(v[k], floor_v) + 4.0f) / 4.0f;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i + k < n_real) {
        const float t = logmel[static_cast<long long>(b) * n_frames * m + i + k];
        v[k] = (fmaxf(t, floor_v) + 4.0f) / 4.0f;
      } else {
        v[k] = -1.0f;                         // lib.rs:431-437 pad value
      }
    }
  }
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + b * per_chunk + i) = make_float4(v[0], v[1], v[2], v[3]);
  if (out_bf16) {
    __nv_bfloat16* o = out_bf16 + static_cast<long long>(b) * (T_out + 2) * m + m + i;     // skip guard row 0
    uint2 w;
    w.x = pack_bf16x2(v[0], v[1]);
    w.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(o) = w;
  }
}

// any n_mels: one value per thread
__global__ void __launch_bounds__(256)
mel_finalize_scalar_kernel(const float* __restrict__ logmel, const int* __restrict__ chunk_max_key, int n_frames, int T_out, int m,
                           float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.y;
  const long long per_chunk = static_cast<long long>(T_out) * m;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per_chunk) return;
  const float floor_v = key_to_float(chunk_max_key[b]) - 8.0f;
  const long long n_real = static_cast<long long>(min(n_frames, T_out)) * m;
  float v = -1.0f;
  if (i < n_real) v = (fmaxf(logmel[static_cast<long long>(b) * n_frames * m + i], floor_v) + 4.0f) / 4.0f;
  if (out_f32) out_f32[b * per_chunk + i] = v;
  if (out_bf16) out_bf16[static_cast<long long>(b) * (T_out + 2) * m + m + i] = __float2bfloat16_rn(v);
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
        tw200[n2 * 25 + k1] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
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
    WB_CUDA_OK(cudaFuncSetAttribute(mel_stft_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem)));
    WB_CUDA_OK(cudaFuncSetAttribute(mel_stft_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem)));
    return WB_OK;
  });
}

int launch_mel_stft(const float* audio, long long audio_stride, const int* n_valid, int padded_len, int hop, int n_frames, int B,
                    const MelTables& t, float* logmel, int* chunk_max_key, cudaStream_t stream) {
  int rc = mel_init();
  if (rc != WB_OK) return rc;
  if (B <= 0 || n_frames <= 0) return WB_OK;
  const int dev = current_device();
  mel_init_max_kernel<<<(B + 255) / 256, 256, 0, stream>>>(chunk_max_key, B);
  count_launch();
  int fpt = FT;
  if (hop > 160) {
    fpt = (MAX_TILE - NFFT) / hop + 1;
    if (fpt < 1) fpt = 1;
    if (fpt > FT) fpt = FT;
  }
  dim3 grid((n_frames + fpt - 1) / fpt, B);
  const bool fast = hop == 160 && (audio_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(audio) & 15) == 0);
  if (fast) {
    mel_stft_kernel<160><<<grid, MEL_THREADS, sizeof(MelSmem), stream>>>(audio, audio_stride, n_valid, padded_len, hop, n_frames,
                                                                         fpt, t, g_tw200[dev], g_tw400[dev], logmel, chunk_max_key);
  } else {
    mel_stft_kernel<0><<<grid, MEL_THREADS, sizeof(MelSmem), stream>>>(audio, audio_stride, n_valid, padded_len, hop, n_frames, fpt,
                                                                       t, g_tw200[dev], g_tw400[dev], logmel, chunk_max_key);
  }
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_mel_finalize(const float* logmel, const int* chunk_max_key, int n_frames, int T_out, int n_mels, int B, float* out_f32,
                        __nv_bfloat16* out_bf16_padded, cudaStream_t stream) {
  if (B <= 0 || T_out <= 0) return WB_OK;
  const long long per_chunk = static_cast<long long>(T_out) * n_mels;
  if (n_mels % 4 == 0) {
    dim3 grid(static_cast<unsigned>((per_chunk / 4 + 255) / 256), B);
    mel_finalize_kernel<<<grid, 256, 0, stream>>>(logmel, chunk_max_key, n_frames, T_out, n_mels, out_f32, out_bf16_padded);
    count_launch();
  } else {
    dim3 grid(static_cast<unsigned>((per_chunk + 255) / 256), B);
    mel_finalize_scalar_kernel<<<grid, 256, 0, stream>>>(logmel, chunk_max_key, n_frames, T_out, n_mels, out_f32, out_bf16_padded);
    count_launch();
  }
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
      z[n2 * 25 + k1] = cmul(v[k1], cmake(static_cast<float>(cos(a)), static_cast<float>(sin(a))));
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
      float r0 = a[0].x + a[0].y, r200 = a[0].x - a[0].y;
      p[0] = r0 * r0;
      p[200] = r200 * r200;
      for (int k2 = 1; k2 < 8; ++k2) p[25 * k2] = rfft_power(a[k2], a[8 - k2], tw400(25 * k2));
    } else {
      cf c[8];
      for (int n2 = 0; n2 < 8; ++n2) c[n2] = z[n2 * 25 + 25 - pcol];
      dft8(c);
      for (int k2 = 0; k2 < 8; ++k2) {
        int k = pcol + 25 * k2, kk = 25 - pcol + 25 * k2;
        p[k] = rfft_power(a[k2], c[7 - k2], tw400(k));
        p[kk] = rfft_power(c[k2], a[7 - k2], tw400(kk));
      }
    }
  }
}
