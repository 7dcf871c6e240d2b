// Non-causal multi-head self-attention on tcgen05 / TMEM, fed by TMA.
//
// Replaces MultiHeadAttention::forward_cross_flash (src/model/attention.rs:894-935) with its per-head
// flash_attention_simd (attention.rs:472-519) of the reference: heads are 64-wide column slices of q, k, v
// (extract_head, attention.rs:1094-1107), scores are scaled by 1/sqrt(64), the softmax is the same online
// (running max / running sum) recurrence, just over KV blocks of 64 instead of 32 -- block size does not
// change the result (the reference asserts this itself, attention.rs:2186-2228).  The encoder passes no mask.
//
// Warp-specialised kernel, one CTA = one (chunk, head, pair of 128-query tiles), 320 threads:
//   warp 0      TMA producer: Q0/Q1 once, then a 4-stage ring of K_j / V_j tiles ([64][64] bf16, 128 B swizzle)
//   warp 1      tcgen05.mma issuer.  Per KV block j (64 keys) and query tile t in {0,1}:
//                 S_t[j&1] = Q_t K_j^T   4 x (128x64x16), both operands K-major            -> TMEM, double buffered
//                 O_t     += P_t[j&1] V_j 4 x (128x64x16), A = P (smem), B = V_j MN-major   -> TMEM
//               S_{j+1} is issued before the softmax of block j has finished, so the softmax warps never wait
//               for the tensor core in steady state.
//   warps 2-5   softmax warpgroup of tile 0, warps 6-9 of tile 1: thread r owns query row r == TMEM lane r.
//               One tcgen05.ld pass keeps the 64 scores of the row in registers; running max with LAZY rescaling
//               (O_t in TMEM is rescaled only when the max grows by more than 2^8, a rare TMEM read-modify-write);
//               p = exp2(s*c - m*c) -> bf16 -> K-major swizzled P_t tile (double buffered) in shared memory.
// TMEM columns: S_t[b] at t*128 + b*64 (256 total), O_t at 256 + t*64.   smem: Q 32 KB + K/V ring 64 KB + P 64 KB.
// The kernel is MUFU(ex2)-bound by construction: 2 x 8192 exponentials per (256 x 64) block at 16/clk/SM.
#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {
namespace {

constexpr int BQ = 128, BKV = 64, DH = 64;
constexpr int Q_TILE_BYTES = BQ * DH * 2;          // 16 KB
constexpr int KV_TILE_BYTES = BKV * DH * 2;        // 8 KB
constexpr int P_TILE_BYTES = BQ * BKV * 2;         // 16 KB
constexpr int KV_STAGES = 4;
constexpr int WS_THREADS = 320;
constexpr int WS_SMEM = 2 * Q_TILE_BYTES + 2 * KV_STAGES * KV_TILE_BYTES + 4 * P_TILE_BYTES + 256;
constexpr int WS_TMEM_COLS = 512;

struct AttnParams {
  int S, d, n_kv_blocks;
  float scale_log2;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(WS_THREADS, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                                   // [2]
  uint8_t* sK = smem + 2 * Q_TILE_BYTES;                // [KV_STAGES]
  uint8_t* sV = sK + KV_STAGES * KV_TILE_BYTES;         // [KV_STAGES]
  uint8_t* sP = sV + KV_STAGES * KV_TILE_BYTES;         // [tile][buf]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * P_TILE_BYTES);
  uint64_t* q_full = bars;                              // 1
  uint64_t* kv_full = bars + 1;                         // [4]
  uint64_t* kv_empty = bars + 5;                        // [4]
  uint64_t* s_full = bars + 9;                          // [tile*2 + buf]
  uint64_t* p_full = bars + 13;                         // [tile*2 + buf], 128 arrivals
  uint64_t* pv_done = bars + 17;                        // [tile*2 + buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n = p.n_kv_blocks;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, WS_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      mbar_expect_tx(q_full, 2 * Q_TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * DH, q0, b);
      tma_load_3d(sQ + Q_TILE_BYTES, &tmQ, q_full, h * DH, q0 + BQ, b);
      for (int j = 0; j < n; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t use = static_cast<uint32_t>(j / KV_STAGES);
        mbar_wait(&kv_empty[st], (use & 1u) ^ 1u);
        mbar_expect_tx(&kv_full[st], 2 * KV_TILE_BYTES);
        tma_load_3d(sK + st * KV_TILE_BYTES, &tmKV, &kv_full[st], p.d + h * DH, j * BKV, b);
        tma_load_3d(sV + st * KV_TILE_BYTES, &tmKV, &kv_full[st], 2 * p.d + h * DH, j * BKV, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, DH, 1);     // B = V tile, MN-major
      auto issue_s = [&](int t, int jj) {
        const int st = jj % KV_STAGES, buf = jj & 1;
        const uint32_t qa = smem_u32(sQ + t * Q_TILE_BYTES), ka = smem_u32(sK + st * KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_f16(tmem_base + t * 128 + buf * 64, umma_desc_sw128(qa + k * 32), umma_desc_sw128(ka + k * 32), idesc_s, k != 0);
        umma_commit(&s_full[t * 2 + buf]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after_sync();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < n; ++j) {
        const int st = j % KV_STAGES, buf = j & 1;
        if (j + 1 < n) {
          // S buffer (j+1)&1 of both tiles was released by p_full of block j-1, waited for in the previous iteration
          mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
          tc_fence_after_sync();
          issue_s(0, j + 1);
          issue_s(1, j + 1);
        }
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t * 2 + buf], (j >> 1) & 1);      // P_t(j) is in smem, S_t[buf] has been consumed
          tc_fence_after_sync();
          const uint32_t pa = smem_u32(sP + (t * 2 + buf) * P_TILE_BYTES), va = smem_u32(sV + st * KV_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k)
            umma_f16(tmem_base + 256 + t * 64, umma_desc_sw128(pa + k * 32), umma_desc_sw128(va + k * 16 * 128), idesc_o, (j | k) != 0);
          umma_commit(&pv_done[t * 2 + buf]);
        }
        umma_commit(&kv_empty[st]);     // all MMAs reading K_j / V_j were issued before this commit
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax warpgroups
    const int t = (warp - 2) >> 2;                        // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_sel;
    const uint32_t tO = tmem_base + 256 + t * 64 + lane_sel;
    const uint32_t p_row = smem_u32(sP + t * 2 * P_TILE_BYTES) + r * 128;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const float c = p.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n; ++j) {
      const int buf = j & 1;
      mbar_wait(&s_full[t * 2 + buf], (j >> 1) & 1);
      tc_fence_after_sync();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(tS + buf * 64, s0);
      tmem_ld_32x32b_x32(tS + buf * 64 + 32, s1);
      tmem_ld_wait();
      const int kv_valid = p.S - j * BKV;                 // >= 64 except in the last block
      if (kv_valid < BKV) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) s0[i] = 0xff800000u;         // -inf
          if (32 + i >= kv_valid) s1[i] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
      // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
      const bool need = (mx - m_ref) * c > 8.0f;
      float alpha = 1.0f;
      if (need) {
        alpha = fast_exp2((m_ref - mx) * c);              // 0 on the first block (m_ref = -inf)
        m_ref = mx;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        mbar_wait(&pv_done[t * 2 + (buf ^ 1)], ((j - 1) >> 1) & 1);       // PV_{j-1} retired: O_t is stable
        tc_fence_after_sync();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tO + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32b_x32(tO + cc * 32, v);
        }
        tmem_st_wait();
      }
      if (j >= 2) mbar_wait(&pv_done[t * 2 + buf], ((j - 2) >> 1) & 1);    // PV_{j-2} retired: P_t[buf] is free
      const float mb = m_ref * c;
      float rs = 0.f;
      const uint32_t pb = p_row + static_cast<uint32_t>(buf) * P_TILE_BYTES;
      auto emit = [&](const uint32_t (&sv)[32], int cgrp) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = fast_exp2(__uint_as_float(sv[8 * g + i]) * c - mb);   // exp2(-inf) == 0: masked
          rs += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          const uint32_t chunk = static_cast<uint32_t>(cgrp * 4 + g);
          const uint32_t addr = pb + ((chunk ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(e[0], e[1])), "r"(pack_bf16x2(e[2], e[3])),
                       "r"(pack_bf16x2(e[4], e[5])), "r"(pack_bf16x2(e[6], e[7]))
                       : "memory");
        }
      };
      emit(s0, 0);
      emit(s1, 1);
      l_run = l_run * alpha + rs;
      fence_proxy_async_smem();                           // generic-proxy P writes -> visible to the tensor core (async proxy)
      tc_fence_before_sync();
      mbar_arrive(&p_full[t * 2 + buf]);
    }
    // epilogue: O_t / l  (attention.rs:334-343: 0 when the sum is <= 1e-10)
    mbar_wait(&pv_done[t * 2 + ((n - 1) & 1)], ((n - 1) >> 1) & 1);
    tc_fence_after_sync();
    const int row = q0 + t * BQ + r;
    const float inv = l_run > 1e-10f ? 1.0f / l_run : 0.f;
    uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * p.S + row) * p.d + h * DH);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tO + cc * 32, v);
      tmem_ld_wait();
      if (row < p.S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[8 * g + 0]) * inv, __uint_as_float(v[8 * g + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv);
          dst[cc * 4 + g] = w;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, WS_TMEM_COLS);
  }
}

bool g_att_init = false;

}  // namespace

int attention_init() {
  if (g_att_init) return WB_OK;
  WB_CUDA_OK(cudaFuncSetAttribute(attention_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM));
  g_att_init = true;
  return WB_OK;
}

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int d, int n_heads, cudaStream_t stream) {
  int rc = attention_init();
  if (rc != WB_OK) return rc;
  if (B <= 0 || S <= 0) return WB_OK;
  if (d != n_heads * DH) return set_error(WB_ERR_MODEL, "attention kernel needs d_head == 64 (all Whisper sizes)");
  CUtensorMap tq, tkv;
  rc = make_tmap_bf16_3d(&tq, qkv, 3ull * d, S, B, 3ull * d * 2, 3ull * d * 2 * S, DH, BQ);
  if (rc != WB_OK) return rc;
  rc = make_tmap_bf16_3d(&tkv, qkv, 3ull * d, S, B, 3ull * d * 2, 3ull * d * 2 * S, DH, BKV);
  if (rc != WB_OK) return rc;
  AttnParams p;
  p.S = S;
  p.d = d;
  p.n_kv_blocks = (S + BKV - 1) / BKV;
  p.scale_log2 = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) * log2(e)
  p.out = out;
  dim3 grid((S + 2 * BQ - 1) / (2 * BQ), n_heads, B);
  attention_ws_kernel<<<grid, WS_THREADS, WS_SMEM, stream>>>(tq, tkv, p);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

}  // namespace wb
