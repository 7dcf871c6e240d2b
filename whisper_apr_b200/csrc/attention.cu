// Non-causal multi-head self-attention on tcgen05 / TMEM, fed by TMA.
//
// Replaces MultiHeadAttention::forward_cross_flash (src/model/attention.rs:894-935) with its per-head
// flash_attention_simd (attention.rs:472-519) of the reference: heads are 64-wide column slices of q, k, v
// (extract_head, attention.rs:1094-1107), scores are scaled by 1/sqrt(64), the softmax is the same online
// (running max / running sum) recurrence, just over KV blocks of 128 instead of 32 -- block size does not
// change the result (the reference asserts this itself, attention.rs:2186-2228).  The encoder passes no mask.
//
// One CTA = one (chunk, head, 128-query tile); 128 threads, thread t owns query row t == TMEM lane t.
//   S  = Q K_j^T    : 4 x tcgen05.mma 128x128x16 (both operands K-major, 128 B swizzled TMA tiles)  -> TMEM cols [0,128)
//   softmax(S) row-wise in registers (tcgen05.ld), P written to shared memory as a K-major swizzled bf16 tile
//   O_j = P V_j     : 8 x tcgen05.mma 128x64x16, V used as an MN-major operand straight from its TMA tile -> TMEM [128,192)
//   o = o * alpha + O_j in registers (fp32); final o / l, bf16 store of the head's 64 columns.
// K/V tiles are double buffered (TMA prefetch two blocks ahead); two CTAs are resident per SM so one CTA's
// softmax overlaps the other's MMAs.
#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {
namespace {

constexpr int BQ = 128, BKV = 128, DH = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;          // 16 KB: one [128][64] bf16 tile
constexpr int ATT_SMEM = 7 * TILE_BYTES + 128;    // Q, K0, K1, V0, V1, P(2 tiles) + barriers
constexpr int ATT_THREADS = 128;
constexpr int ATT_TMEM_COLS = 256;

struct AttnParams {
  int S, d, n_kv_blocks;
  float scale_log2;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + 3 * TILE_BYTES;
  uint8_t* sP = smem + 5 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE_BYTES);
  uint64_t* bar_q = bars;
  uint64_t* bar_kv = bars + 1;      // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int q0 = blockIdx.x * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = p.n_kv_blocks;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;

  if (tid == 0) {
    mbar_expect_tx(bar_q, TILE_BYTES);
    tma_load_3d(sQ, &tmQKV, bar_q, h * DH, q0, b);
    for (int j = 0; j < 2 && j < nkv; ++j) {
      mbar_expect_tx(&bar_kv[j], 2 * TILE_BYTES);
      tma_load_3d(sK + j * TILE_BYTES, &tmQKV, &bar_kv[j], p.d + h * DH, j * BKV, b);
      tma_load_3d(sV + j * TILE_BYTES, &tmQKV, &bar_kv[j], 2 * p.d + h * DH, j * BKV, b);
    }
  }

  constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, DH, 1);     // B = V tile, MN-major

  float o[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const uint32_t p_row = smem_u32(sP) + tid * 128;
  const uint32_t sw = static_cast<uint32_t>(tid & 7);

  for (int j = 0; j < nkv; ++j) {
    const int st = j & 1;
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[st], (j >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * TILE_BYTES);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_f16(tS, umma_desc_sw128(qa + k * 32), umma_desc_sw128(ka + k * 32), idesc_s, k != 0);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, j & 1);
    tc_fence_after_sync();

    const int kv_valid = min(BKV, p.S - j * BKV);
    // pass 1: row max over the valid columns
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tS + lane_sel + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float sv = (c * 32 + i < kv_valid) ? __uint_as_float(v[i]) : -INFINITY;
        mx = fmaxf(mx, sv);
      }
    }
    const float m_new = fmaxf(m_run, mx);
    const float alpha = fast_exp2((m_run - m_new) * p.scale_log2);
    const float mb = m_new * p.scale_log2;
    // pass 2: p = exp2(s*c - m*c), row sum, bf16 P tile (K-major, 128 B swizzle)
    float rs = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tS + lane_sel + c * 32, v);
      tmem_ld_wait();
      float pv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = fast_exp2(__uint_as_float(v[i]) * p.scale_log2 - mb);
        pv[i] = (c * 32 + i < kv_valid) ? e : 0.f;
        rs += pv[i];
      }
      const uint32_t half_base = p_row + static_cast<uint32_t>(c >> 1) * TILE_BYTES;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + g);
        const uint32_t addr = half_base + ((chunk ^ sw) << 4);
        const uint32_t w0 = pack_bf16x2(pv[8 * g + 0], pv[8 * g + 1]);
        const uint32_t w1 = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
        const uint32_t w2 = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
        const uint32_t w3 = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
      }
    }
    l_run = l_run * alpha + rs;
    m_run = m_new;

    tc_fence_before_sync();
    fence_proxy_async_smem();     // generic-proxy P writes -> visible to the tensor core's async proxy
    __syncthreads();
    if (tid == 0) {
      tc_fence_after_sync();
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV + st * TILE_BYTES);
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k) {
        const uint32_t a_addr = pa + (k >> 2) * TILE_BYTES + (k & 3) * 32;
        const uint32_t b_addr = va + k * 16 * 128;
        umma_f16(tO, umma_desc_sw128(a_addr), umma_desc_sw128(b_addr), idesc_o, k != 0);
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, j & 1);
    tc_fence_after_sync();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tO + lane_sel + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();                 // every thread is done with S, O_j, P and (via bar_o) K/V stage `st`
    if (tid == 0 && j + 2 < nkv) {
      mbar_expect_tx(&bar_kv[st], 2 * TILE_BYTES);
      tma_load_3d(sK + st * TILE_BYTES, &tmQKV, &bar_kv[st], p.d + h * DH, (j + 2) * BKV, b);
      tma_load_3d(sV + st * TILE_BYTES, &tmQKV, &bar_kv[st], 2 * p.d + h * DH, (j + 2) * BKV, b);
    }
  }

  // normalise (attention.rs:334-343: divide by the sum, 0 if the sum is <= 1e-10) and store this head's columns
  const int row = q0 + tid;
  if (row < p.S) {
    const float inv = l_run > 1e-10f ? 1.0f / l_run : 0.f;
    uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * p.S + row) * p.d + h * DH);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint4 w;
      w.x = pack_bf16x2(o[8 * g + 0] * inv, o[8 * g + 1] * inv);
      w.y = pack_bf16x2(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
      w.z = pack_bf16x2(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
      w.w = pack_bf16x2(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
      dst[g] = w;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

bool g_att_init = false;

}  // namespace

int attention_init() {
  if (g_att_init) return WB_OK;
  WB_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  g_att_init = true;
  return WB_OK;
}

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int d, int n_heads, cudaStream_t stream) {
  int rc = attention_init();
  if (rc != WB_OK) return rc;
  if (B <= 0 || S <= 0) return WB_OK;
  if (d != n_heads * DH) return set_error(WB_ERR_MODEL, "attention kernel needs d_head == 64 (all Whisper sizes)");
  CUtensorMap tm;
  rc = make_tmap_bf16_3d(&tm, qkv, 3ull * d, S, B, 3ull * d * 2, 3ull * d * 2 * S, DH, 128);
  if (rc != WB_OK) return rc;
  AttnParams p;
  p.S = S;
  p.d = d;
  p.n_kv_blocks = (S + BKV - 1) / BKV;
  p.scale_log2 = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) * log2(e)
  p.out = out;
  dim3 grid((S + BQ - 1) / BQ, n_heads, B);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm, p);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

}  // namespace wb
