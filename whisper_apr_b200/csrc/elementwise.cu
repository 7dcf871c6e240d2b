// HBM-bound helpers around the tensor-core kernels: LayerNorm, dtype expansion of `.apr` payloads,
// conv-weight repacking and guard-row fills.
#include <algorithm>

#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {
namespace {

// ---------------------------------------------------------------------------------------------
// LayerNorm::forward (src/model/encoder.rs:219-251): per row mean, POPULATION variance (two passes, as the
// reference), 1/sqrt(var + 1e-5), * gamma + beta.  One warp per row; the row lives in registers between the
// passes so x is read from HBM once (4 B/elem in, 2 B/elem out for the bf16 GEMM operand).
// One row by one warp.  CG: read x with ld.global.cg (L2 only) -- the follower reads rows another kernel's TMA reductions have just
// written, which must not be served from a stale L1 line.
template <int NV, bool CG>   // float4 per lane; d == 128 * nv, nv <= NV
__device__ __forceinline__ void ln_row(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long row,
                                       int d, int lane, uint16_t* __restrict__ out_16, bool as_bf16, float* __restrict__ out_f32) {
  const int nv = d >> 7;
  // lane l owns float4 l, l + 32, ...: every load instruction of the warp covers 512 contiguous bytes (a lane reading 32 B runs
  // instead -- tried for 16 B output stores -- touches every sector twice and cost 8 % on this HBM-bound kernel)
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      v[i] = CG ? __ldcg(xr + lane + 32 * i) : xr[lane + 32 * i];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(d);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + e * e);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float inv = 1.0f / sqrtf(sq / static_cast<float>(d) + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < nv) {
      const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
      float4 r;
      r.x = (v[i].x - mean) * inv * g.x + bb.x;
      r.y = (v[i].y - mean) * inv * g.y + bb.y;
      r.z = (v[i].z - mean) * inv * g.z + bb.z;
      r.w = (v[i].w - mean) * inv * g.w + bb.w;
      if (out_16) {
        uint2 w;
        if (as_bf16) {             // the caller asked for bf16 states (WB_BF16): a user-facing format, not the operand format
          w.x = pack_bf16x2(r.x, r.y);
          w.y = pack_bf16x2(r.z, r.w);
        } else {
          w.x = pack_op16x2(r.x, r.y);
          w.y = pack_op16x2(r.z, r.w);
        }
        reinterpret_cast<uint2*>(out_16 + row * d)[lane + 32 * i] = w;
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32 + row * d)[lane + 32 * i] = r;
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int d,
                 uint16_t* __restrict__ out_16, bool as_bf16, float* __restrict__ out_f32, bool reverse) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (reverse) row = rows - 1 - row;                      // last rows first: the ones the producing GEMM left in L2
  ln_row<NV, false>(x, gamma, beta, row, d, threadIdx.x & 31, out_16, as_bf16, out_f32);
}

// The follower: runs on its own stream BESIDE the residual GEMM that produces x (one small block per SM next to the GEMM's CTA: no
// shared memory to speak of, no tensor memory).  Block b takes row groups b, b + gridDim, ...: the GEMM finishes them in ascending
// order.  Thread 0 polls the group's counter (acquire) until all `need` column tiles have reduced their share into the rows, re-arms
// it, and the block's 8 warps normalise 4 rows each straight out of L2 -- the residual stream is never re-read from HBM for its
// LayerNorm and the LayerNorm costs no time of its own beyond the last group's few microseconds.
// Concurrency of two kernels is never guaranteed: WAIT = true gives up after 20 ms without progress (a GEMM takes < 1 ms) and leaves
// its remaining groups; the same kernel with WAIT = false is launched BEHIND the GEMM and normalises whatever groups still carry a
// full counter (normally none: ~3 us).  No schedule can deadlock or skip a row.
template <int NV, bool WAIT>
__global__ void __launch_bounds__(256)
layernorm_follow_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int d,
                        uint16_t* __restrict__ out_16, unsigned int* __restrict__ ready, unsigned int need) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_groups = (rows + 31) >> 5;
  __shared__ int s_go;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    if (threadIdx.x == 0) {
      unsigned int c;
      uint64_t t0 = 0;
      unsigned int spins = 0;
      int ok = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(ready + g) : "memory");
        if (c >= need) { ok = 1; break; }
        if (!WAIT) break;
        __nanosleep(spins < 64 ? 100 : 500);
        if ((++spins & 255u) == 0) {
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 20000000ull) break;
        }
      }
      if (ok) ready[g] = 0u;                            // re-armed for the next residual GEMM (it starts after this kernel has ended)
      s_go = ok;
    }
    __syncthreads();
    const int go = s_go;
    __syncthreads();
    if (!go) {
      if (WAIT) break;
      continue;
    }
    for (int r = warp; r < 32; r += 8) {
      const long long row = static_cast<long long>(g) * 32 + r;
      if (row < rows) ln_row<NV, true>(x, gamma, beta, row, d, lane, out_16, false, nullptr);
    }
  }
}

// Any d: one warp per row, three passes over global/L1.
__global__ void __launch_bounds__(256)
layernorm_generic_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int rows,
                         int d, uint16_t* __restrict__ out_16, bool as_bf16, float* __restrict__ out_f32) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + static_cast<long long>(row) * d;
  float sum = 0.f;
  for (int i = lane; i < d; i += 32) sum += xr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(d);
  float sq = 0.f;
  for (int i = lane; i < d; i += 32) { const float a = xr[i] - mean; sq += a * a; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float inv = 1.0f / sqrtf(sq / static_cast<float>(d) + 1e-5f);
  for (int i = lane; i < d; i += 32) {
    const float r = (xr[i] - mean) * inv * gamma[i] + beta[i];
    if (out_16) {
      if (as_bf16) reinterpret_cast<__nv_bfloat16*>(out_16)[static_cast<long long>(row) * d + i] = __float2bfloat16_rn(r);
      else reinterpret_cast<op16*>(out_16)[static_cast<long long>(row) * d + i] = float_to_op16(r);
    }
    if (out_f32) out_f32[static_cast<long long>(row) * d + i] = r;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, op16* __restrict__ out, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = float_to_op16(in[i]);
}
// int8 payload (format/mod.rs:632-672) -> bf16 integer value (exact: |q| <= 127); the per-tensor scale is applied in
// the GEMM epilogue.
__global__ void i8_to_bf16_kernel(const int8_t* __restrict__ in, op16* __restrict__ out, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = float_to_op16(static_cast<float>(in[i]));
}
// packed int4, even index -> low nibble, two's complement (model/quantized.rs:1887-1969)
__device__ __forceinline__ int unpack_i4(uint8_t byte, size_t idx) {
  int nib = (idx & 1) ? (byte >> 4) : (byte & 0x0F);
  return nib >= 8 ? nib - 16 : nib;
}
__global__ void i4_to_bf16_kernel(const uint8_t* __restrict__ in, op16* __restrict__ out, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = float_to_op16(static_cast<float>(unpack_i4(in[i >> 1], i)));
}
// Vector forms used by the per-layer weight expansion (n % 32 == 0, 16-byte aligned): 16 B of packed payload per thread.
__device__ __forceinline__ uint32_t bf16x2_from_ints(int lo, int hi) {          // exact in either operand format (|q| <= 128)
  return pack_op16x2(static_cast<float>(lo), static_cast<float>(hi));
}
__global__ void __launch_bounds__(256) i8_to_bf16_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n16; i += stride) {
    const uint4 v = __ldg(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = bf16x2_from_ints(static_cast<int8_t>(w[k] & 0xff), static_cast<int8_t>((w[k] >> 8) & 0xff));
      o[2 * k + 1] = bf16x2_from_ints(static_cast<int8_t>((w[k] >> 16) & 0xff), static_cast<int8_t>(w[k] >> 24));
    }
    out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
    out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}
__global__ void __launch_bounds__(256) i4_to_bf16_vec_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n16) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n16; i += stride) {
    const uint4 v = __ldg(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {                  // one 32-bit word = 8 nibbles = 8 weights, low nibble first
      uint32_t o[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int lo = (static_cast<int>(w[k] << (28 - 8 * b))) >> 28;          // sign-extended nibble 2b
        const int hi = (static_cast<int>(w[k] << (24 - 8 * b))) >> 28;          // sign-extended nibble 2b + 1
        o[b] = bf16x2_from_ints(lo, hi);
      }
      out[4 * i + k] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}
__global__ void i8_to_f32_kernel(const int8_t* __restrict__ in, float scale, float* __restrict__ out, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = static_cast<float>(in[i]) * scale;
}
__global__ void i4_to_f32_kernel(const uint8_t* __restrict__ in, float scale, float* __restrict__ out, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = static_cast<float>(unpack_i4(in[i >> 1], i)) * scale;
}
// Conv1d weight [out][in][3] (encoder.rs:96-98) -> [out][3][in]: row `out` becomes the K-major GEMM operand whose
// K index is tap*in + channel, matching the contiguous 3-frame window of the activation view.
__global__ void conv_repack_kernel(const op16* __restrict__ in, op16* __restrict__ out, int c_out, int c_in) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n = static_cast<size_t>(c_out) * c_in * 3;
  if (i >= n) return;
  const int o = static_cast<int>(i / (3 * static_cast<size_t>(c_in)));
  const int rem = static_cast<int>(i - static_cast<size_t>(o) * 3 * c_in);
  const int tap = rem / c_in, ch = rem - tap * c_in;
  out[i] = in[(static_cast<size_t>(o) * c_in + ch) * 3 + tap];
}
// mel f32 [B][T][m] -> bf16 [B][T+2][m], guard rows 0 and T+1 zero
__global__ void mel_pad_bf16_kernel(const float* __restrict__ mel, op16* __restrict__ out, int T, int m) {
  const int b = blockIdx.y;
  const long long per = static_cast<long long>(T + 2) * m;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const long long row = i / m;
  float v = 0.f;
  if (row >= 1 && row <= T) v = mel[static_cast<long long>(b) * T * m + (i - m)];
  out[b * per + i] = float_to_op16(v);
}
__global__ void fill_zero_rows_kernel(op16* base, long long batch_stride, int row_elems) {
  op16* p = base + static_cast<long long>(blockIdx.y) * batch_stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_elems; i += gridDim.x * blockDim.x) p[i] = float_to_op16(0.f);
}

inline unsigned grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  return static_cast<unsigned>(g > 148 * 16 ? 148 * 16 : (g == 0 ? 1 : g));
}

}  // namespace

int launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int d, void* out_16v, bool out_16_is_bf16,
                     float* out_f32, cudaStream_t stream, bool reverse) {
  uint16_t* out_bf16 = static_cast<uint16_t*>(out_16v);
  if (rows <= 0) return WB_OK;
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb);
  if (d % 128 == 0 && d <= 2048) {
    if (d <= 512) layernorm_kernel<4><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, out_bf16, out_16_is_bf16, out_f32, reverse);
    else if (d <= 1280) layernorm_kernel<10><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, out_bf16, out_16_is_bf16, out_f32, reverse);
    else layernorm_kernel<16><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, out_bf16, out_16_is_bf16, out_f32, reverse);
    count_launch();
  } else {
    layernorm_generic_kernel<<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, out_bf16, out_16_is_bf16, out_f32);
    count_launch();
  }
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_layernorm_follow(const float* x, const float* gamma, const float* beta, int rows, int d, op16* out, unsigned int* ready,
                            unsigned int need, bool wait, cudaStream_t stream) {
  if (rows <= 0) return WB_OK;
  if (d % 128 != 0 || d > 1280) return set_error(WB_ERR_MODEL, "layernorm_follow needs d % 128 == 0 and d <= 1280");
  // An SM has ONE L1 / shared-memory split at a time: a resident block that was launched with the default (L1-heavy) carve-out keeps
  // a GEMM CTA that needs 213 KB of shared memory off that SM until it exits -- measured: the follower then never overlaps and times
  // out.  Ask for the GEMM's own split (it reads x with ld.global.cg and has no use for L1 anyway).
  static PerDeviceOnce once;
  int rc = once.run([](int) -> int {
    WB_CUDA_OK((cudaFuncSetAttribute(layernorm_follow_kernel<4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)));
    WB_CUDA_OK((cudaFuncSetAttribute(layernorm_follow_kernel<10, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)));
    return WB_OK;
  });
  if (rc != WB_OK) return rc;
  const int groups = (rows + 31) / 32;
  const int grid = std::min(groups, device_sm_count());          // one block per SM, beside the GEMM's CTA
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  if (wait) {
    if (d <= 512) layernorm_follow_kernel<4, true><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, o, ready, need);
    else layernorm_follow_kernel<10, true><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, o, ready, need);
  } else {
    if (d <= 512) layernorm_follow_kernel<4, false><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, o, ready, need);
    else layernorm_follow_kernel<10, false><<<grid, 256, 0, stream>>>(x, gamma, beta, rows, d, o, ready, need);
  }
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

int launch_f32_to_op16(const float* in, op16* out, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  f32_to_bf16_kernel<<<grid_for(n), 256, 0, s>>>(in, out, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_i8_to_op16(const int8_t* in, op16* out, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  if (n % 16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    i8_to_bf16_vec_kernel<<<grid_for(n / 16), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n / 16);
    count_launch();
    WB_CUDA_OK(cudaGetLastError());
    return WB_OK;
  }
  i8_to_bf16_kernel<<<grid_for(n), 256, 0, s>>>(in, out, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_i4_to_op16(const uint8_t* in, op16* out, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  if (n % 32 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    i4_to_bf16_vec_kernel<<<grid_for(n / 32), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n / 32);
    count_launch();
    WB_CUDA_OK(cudaGetLastError());
    return WB_OK;
  }
  i4_to_bf16_kernel<<<grid_for(n), 256, 0, s>>>(in, out, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_i8_to_f32(const int8_t* in, float scale, float* out, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  i8_to_f32_kernel<<<grid_for(n), 256, 0, s>>>(in, scale, out, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_i4_to_f32(const uint8_t* in, float scale, float* out, size_t n, cudaStream_t s) {
  if (n == 0) return WB_OK;
  i4_to_f32_kernel<<<grid_for(n), 256, 0, s>>>(in, scale, out, n);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_conv_repack(const op16* in, op16* out, int c_out, int c_in, cudaStream_t s) {
  const size_t n = static_cast<size_t>(c_out) * c_in * 3;
  conv_repack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(in, out, c_out, c_in);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_mel_pad_op16(const float* mel, op16* out, int B, int T, int m, cudaStream_t s) {
  if (B <= 0) return WB_OK;
  const long long per = static_cast<long long>(T + 2) * m;
  dim3 grid(static_cast<unsigned>((per + 255) / 256), B);
  mel_pad_bf16_kernel<<<grid, 256, 0, s>>>(mel, out, T, m);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}
int launch_fill_op16_rows(op16* base, long long batch_stride, int B, int row_elems, cudaStream_t s) {
  if (B <= 0 || row_elems <= 0) return WB_OK;
  dim3 grid((row_elems + 255) / 256, B);
  fill_zero_rows_kernel<<<grid, 256, 0, s>>>(base, batch_stride, row_elems);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

}  // namespace wb
