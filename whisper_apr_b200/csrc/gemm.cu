// tcgen05 / TMEM / TMA GEMM for the encoder's linear layers and (as implicit-im2col views) its conv stem.
//
//   out[M][N] = epilogue(alpha * A[M][K] . W[N][K]^T + bias[N])
//
// replaces LinearWeights::forward / forward_simd (src/model/attention.rs:143-219: y = x W^T + b with
// W stored [out][in]) and Conv1d::forward (src/model/encoder.rs:72-110) of the reference.  Both operands are
// K-major bf16, accumulation is fp32 in tensor memory.
//
// Kernel shape: persistent, a cluster of two CTAs (the two SMs of a TPC, tcgen05 cta_group::2) per 256 x BN tile, 192 or 320 threads per CTA:
//   warp 0      TMA producer: this CTA's 128 rows of A and its half of the W tile per 64-wide k-block, 128-byte swizzle
//   warp 1      TMEM allocator; in the leader CTA the single-thread issuer of tcgen05.mma.cta_group::2 (256 x BN x 16)
//   warps 2..   epilogue: 8 warps (two per TMEM lane quarter, half of the tile's columns each) for 16-bit / GELU outputs, 4 for the
//               residual reduce-add (Gemm2Cfg): tcgen05.ld 32 lanes x 32 columns (double
//               buffered), per-column scale + bias from shared memory,
//               GELU; bf16 results and the fp32 residual update leave through shared-memory staging tiles and TMA
//               (cp.async.bulk.tensor store / cp.reduce.async.bulk.tensor .add) -- no row-per-lane global accesses
// Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA, multicast commits), two TMEM accumulator buffers (MMA <-> epilogue),
// static persistent tile schedule with N fastest so CTA pairs running together share the same A rows through L2.
//
// The A operand is addressed through a 3-D tensor map [batch][rows][K] with free strides, which is what lets
// conv1 / conv2 run as GEMMs without materialising im2col: row t of the conv1 operand is the 3*n_mels
// contiguous values starting at padded frame t (row stride n_mels), and row p of the conv2 operand is the
// 3*d values starting at padded position 2p (row stride 2d).
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 bytes = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_MAX_THREADS = 320;     // TMA warp, MMA warp, up to 8 epilogue warps

struct GemmKParams {
  int rows_per_batch, n_batch, N, K;
  int tiles_m_per_batch, tiles_n;
  int out_rows_per_batch, out_row_off;
  long long ldc;
  float alpha;
  const float* col_scale;
  const float* bias;
  void* out;
  const float* pe;
  unsigned int* ready;          // EPI_RESID_F32: per-32-row completion counters for a follower kernel (or nullptr)
  int reverse;                  // row tiles are walked from the last to the first (GemmDesc::reverse)
  int out_bf16;                 // 16-bit outputs are packed as bf16 whatever the operand format (GemmDesc::out_bf16)
  uint32_t idesc;               // tcgen05 instruction descriptor
};

// GELU (tanh form, encoder.rs:314-318) for bf16 outputs: one MUFU (tanh.approx.f32, relative error 2^-11) instead of the two
// (ex2 + rcp) of gelu_fast -- the result is rounded to bf16 (2^-9) anyway; fc1 0.476 -> 0.463 ms per launch, same parity.
__device__ __forceinline__ float gelu_tanh_approx(float x) {
  const float c = 0.7978846f, k = 0.044715f * 0.7978846f;
  const float x2 = x * x;
  const float u = x * fmaf(k, x2, c);
  const float h = 0.5f * x;
  return fmaf(h, fast_tanh(u), h);
}
__device__ __forceinline__ float gelu_fast(float x) {
  // 0.5x(1+tanh(u)) == x * sigmoid(2u), u = 0.7978846(x + 0.044715x^3)   (encoder.rs:314-318)
  const float c2 = 2.0f * 0.7978846f * 1.4426950408889634f;   // 2 * sqrt(2/pi) * log2(e)
  float u = x + 0.044715f * x * x * x;
  float e = fast_exp2(-c2 * u);
  return __fdividef(x, 1.0f + e);
}

// Epilogue of one 32-column chunk for the CTA-pair kernel: acc = v * scale[n] + bias[n] with scale (alpha x per-column
// dequantisation scale) and bias staged in shared memory once per tile (a global load per chunk stalled the epilogue warps
// for an L2 round trip eight times per tile), and the residual row prefetched one chunk ahead.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, const uint32_t (&v)[32], long long out_row, int r_in_batch, int n0,
                                               uint32_t s_scale, uint32_t s_bias) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 sc = lds_f4(s_scale + 16 * i), bi = lds_f4(s_bias + 16 * i);
    acc[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), sc.x, bi.x);
    acc[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), sc.y, bi.y);
    acc[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), sc.z, bi.z);
    acc[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), sc.w, bi.w);
  }
  if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_PE_F32) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = gelu_fast(acc[i]);
  }
  if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) {
    uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<op16*>(p.out) + out_row * p.ldc + n0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 w;
      w.x = pack_op16x2(acc[8 * i + 0], acc[8 * i + 1]);
      w.y = pack_op16x2(acc[8 * i + 2], acc[8 * i + 3]);
      w.z = pack_op16x2(acc[8 * i + 4], acc[8 * i + 5]);
      w.w = pack_op16x2(acc[8 * i + 6], acc[8 * i + 7]);
      o4[i] = w;
    }
  } else {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldc + n0);
    static_assert(EPI != EPI_RESID_F32, "the residual epilogue goes through the TMA reduction");
    if constexpr (EPI == EPI_GELU_PE_F32) {
      const float4* pe4 = reinterpret_cast<const float4*>(p.pe + static_cast<long long>(r_in_batch) * p.N + n0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 e = __ldg(pe4 + i);
        o4[i] = make_float4(acc[4 * i + 0] + e.x, acc[4 * i + 1] + e.y, acc[4 * i + 2] + e.z, acc[4 * i + 3] + e.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(acc[4 * i + 0], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs (the two SMs of a TPC) computes one 256 x BN tile.  Each CTA stages
// its own 128 rows of A and HALF of the W tile (BN/2 rows) per k-block, the leader CTA issues 256 x BN x 16 MMAs that read
// both halves, and each CTA's TMEM receives its own 128 accumulator rows.  Per SM this halves the W traffic through shared
// memory and L2 (the 1-CTA kernel needs 96 B/clk of operand reads plus 96 B/clk of TMA writes against a 128 B/clk shared-memory
// port), which is what bounds the 1-CTA kernel at ~75 % of the cuBLAS rate.
template <int BN, int EPI>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;             // 16 KB: this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;       // this CTA's half of the W tile
  // Epilogue warps: 8 (two per TMEM lane quarter) where the epilogue computes and converts (16-bit outputs, GELU), 4 for the residual
  // reduce-add, whose epilogue is a scale + bias and whose mainloops are the long ones: there the sixth pipeline stage is worth more
  // than the extra warps (8 warps + 5 stages: out_proj 120 -> 136 us, fc2 429 -> 426 us under ncu; 4 warps + 6 stages kept).
  static constexpr int EPI_WARPS = (EPI == EPI_RESID_F32) ? 4 : 8;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int STAGES = (BN == 256 ? 5 : 7) + (EPI_WARPS == 4 ? 1 : 0);
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_BYTES = 2 * 2 * BN * 4;        // per-tile scale and bias vectors, double buffered
  // One 4 KB staging tile per epilogue warp (32 x 32 f32 for the TMA reduce-add, 32 x 64 op16 for the TMA store), EIGHT epilogue
  // warps: two per TMEM lane quarter (= per scheduler), each taking half of the tile's columns.  With four warps -- one per scheduler,
  // in-order, every chunk waiting out the previous store's read of its staging tile and the latency of its own GELU chain -- a tile's
  // epilogue took ~5 us against 2.2 us of MMAs at K = 512 and about as long as the whole K = 1280 mainloop: the reason the
  // 16-bit-output GEMMs ran the tensor pipe at 85 % where the residual ones reached 97 % (profiles/r02z_*).
  static constexpr int STAGE_C_BYTES = EPI_WARPS * 32 * 32 * 4;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + STAGE_C_BYTES + BAR_BYTES + EPI_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_MAX_THREADS, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                const GemmKParams p) {
  using Cfg = Gemm2Cfg<BN, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int EW = Cfg::EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* sC = sB + STAGES * Cfg::B_BYTES;             // [4 epilogue warps][32 rows][128 B], 128 B swizzle (1024 B aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + Cfg::STAGE_C_BYTES);
  uint64_t* full = bars;                       // used in the leader only (both CTAs' TMA bytes land here)
  uint64_t* empty = bars + STAGES;             // per CTA, signalled by the leader's multicast commit
  uint64_t* tfull = bars + 2 * STAGES;         // per CTA, multicast commit
  uint64_t* tempty = tfull + 2;                // leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 2 * EW); }      // every epilogue warp of both CTAs
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_BF16 || EPI == EPI_GELU_BF16) tma_prefetch_desc(&tmC);
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                          // barriers of both CTAs are initialised before any remote arrive / multicast
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int tiles_m = (p.rows_per_batch + 2 * BM - 1) / (2 * BM);      // 256-row tiles per batch entry
  const int num_tiles = p.n_batch * tiles_m * p.tiles_n;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform registers); lane 0 issues.
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs) {
      const int nb = tile % p.tiles_n;
      const int mb = p.reverse ? p.n_batch * tiles_m - 1 - tile / p.tiles_n : tile / p.tiles_n;
      const int b = mb / tiles_m;
      const int mt = mb - b * tiles_m;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
          tma_load_3d_2sm(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * BK, mt * 2 * BM + static_cast<int>(rank) * BM, b);
          tma_load_3d_2sm(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * BK, nb * BN + static_cast<int>(rank) * (BN / 2), 0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ------------------------------------------------------------ MMA issuer (leader CTA only; lane 0 issues)
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs, ++it) {
        const uint32_t buf = it & 1u;
        const uint32_t use = it >> 1;
        mbar_wait(&tempty[buf], (use & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after_sync();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_f16_2sm(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty[stage]);      // frees the stage in both CTAs once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit_2sm(&tfull[buf]);         // accumulators ready for both CTAs' epilogues
        __syncwarp();
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..9 of both CTAs)
    const int q = warp & 3;                                 // TMEM lane quarter this warp may read (warp id % 4)
    // ONE elected thread per warp issues this warp's TMA stores / reductions and waits on their bulk groups (groups belong to the
    // issuing thread, so the election happens once); see elect_one() for why not `lane == 0`
    const bool leader = elect_one();
    const int half = (warp - 2) >> 2;                       // which share of the tile's column chunks this warp takes (EW / 4 shares)
    const int et = static_cast<int>(threadIdx.x) - 64;      // 0 .. 32 EW - 1 among the epilogue threads
    float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);    // [2][BN]
    float* s_bias = s_scale + 2 * BN;                                                                 // [2][BN]
    constexpr int NCH = BN / 32;
    constexpr bool kResid = EPI == EPI_RESID_F32;
    uint32_t it = 0;
    constexpr int NI = (BN + 32 * EW - 1) / (32 * EW);        // columns per epilogue thread when staging scale / bias
    float pre_s[NI], pre_b[NI];
    auto load_sb = [&](int tile_) {
      const int nb_ = tile_ % p.tiles_n;
#pragma unroll
      for (int ii = 0; ii < NI; ++ii) {
        const int i = ii * 32 * EW + et;
        const int nn = nb_ * BN + i;
        const bool ok = i < BN && nn < p.N;
        pre_s[ii] = ok ? p.alpha * (p.col_scale != nullptr ? __ldg(p.col_scale + nn) : 1.0f) : 0.f;      // columns past N get 0
        pre_b[ii] = (ok && p.bias != nullptr) ? __ldg(p.bias + nn) : 0.f;
      }
    };
    auto store_sb = [&](uint32_t b_) {
#pragma unroll
      for (int ii = 0; ii < NI; ++ii) {
        const int i = ii * 32 * EW + et;
        if (i < BN) {
          sts_f1(smem_u32(s_scale + b_ * BN) + 4 * i, pre_s[ii]);
          sts_f1(smem_u32(s_bias + b_ * BN) + 4 * i, pre_b[ii]);
        }
      }
    };
    if (pair < num_tiles) {
      load_sb(pair);
      store_sb(0);
    }
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++it) {
      const uint32_t buf = it & 1u;
      const uint32_t use = it >> 1;
      const int nb = tile % p.tiles_n;
      const int mb = p.reverse ? p.n_batch * tiles_m - 1 - tile / p.tiles_n : tile / p.tiles_n;
      const int b = mb / tiles_m;
      const int mt = mb - b * tiles_m;
      const int r_in_batch = mt * 2 * BM + static_cast<int>(rank) * BM + q * 32 + lane;
      const bool row_ok = r_in_batch < p.rows_per_batch;
      const long long out_row = static_cast<long long>(b) * p.out_rows_per_batch + p.out_row_off + r_in_batch;
      // this tile's per-column scale (alpha x dequantisation scale) and bias were stored at the end of the previous tile (or before
      // the loop); the NEXT tile's are fetched now, so that their L2 round trip runs under this tile's work instead of in front of it
      // (at K = 512 a tile is 2 us of MMAs: 0.5 us of load latency per tile showed)
      const uint32_t sc = smem_u32(s_scale + buf * BN);      // shared-window addresses: LDS / STS, not generic loads (see lds_f4)
      const uint32_t bi = smem_u32(s_bias + buf * BN);
      const bool has_next = tile + n_pairs < num_tiles;
      if (has_next) load_sb(tile + n_pairs);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");          // scale / bias visible to every epilogue warp
      mbar_wait(&tfull[buf], use & 1u);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);
      // EPI_RESID_F32: the 32 x 32 f32 chunk is staged in shared memory (128 B swizzle: conflict-free 16 B stores) and added to
      // the residual stream by a TMA reduction -- the L2 does the read-modify-write, rows past the end of the batch entry are
      // clipped by the tensor map.  (Row-per-lane global loads/stores cost 32 L1 wavefronts per instruction: 16k cycles per
      // tile, more than the whole K = 1280 mainloop.)
      uint8_t* const stage = sC + (warp - 2) * (32 * 32 * 4);          // this warp's staging tile
      const uint32_t stage_row = smem_u32(stage) + lane * 128;
      const int row0_in_batch = mt * 2 * BM + static_cast<int>(rank) * BM + q * 32;
      auto process = [&](const uint32_t (&v)[32], int c) {
        const int n0 = nb * BN + c * 32;
        if (n0 >= p.N) return;
        if constexpr (kResid) {
          if (row0_in_batch >= p.rows_per_batch) return;      // warp-uniform: nothing of this warp's rows exists
          const uint32_t sc4 = sc + c * 128, bi4 = bi + c * 128;
          if (leader) tma_store_wait_read<0>();            // the previous chunk's reduction has read the staging tile
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 s4 = lds_f4(sc4 + 16 * i), b4 = lds_f4(bi4 + 16 * i);
            const float x0 = fmaf(__uint_as_float(v[4 * i + 0]), s4.x, b4.x), x1 = fmaf(__uint_as_float(v[4 * i + 1]), s4.y, b4.y);
            const float x2 = fmaf(__uint_as_float(v[4 * i + 2]), s4.z, b4.z), x3 = fmaf(__uint_as_float(v[4 * i + 3]), s4.w, b4.w);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + ((static_cast<uint32_t>(i) ^ (lane & 7u)) << 4)), "f"(x0),
                         "f"(x1), "f"(x2), "f"(x3)
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (leader) {
            tma_reduce_add_3d(&tmC, stage, n0, row0_in_batch, b);
            tma_store_commit();
          }
        } else {
          if (row_ok) epilogue_chunk<EPI>(p, v, out_row, r_in_batch, n0, sc + c * 128, bi + c * 128);
        }
      };
      // bf16 outputs: two 32-column chunks make one 32 x 64 bf16 (128 B per row) staging tile, written with conflict-free 16 B
      // shared-memory stores and sent out by one TMA store.  Row-per-lane STG.128 costs 32 L1 wavefronts per instruction, and the
      // L1 / shared-memory data pipe is already ~full with the TMA operand writes and the tensor core's operand reads.
      constexpr bool kBf16Out = EPI == EPI_BF16 || EPI == EPI_GELU_BF16;
      auto half_bf16 = [&](const uint32_t (&v)[32], int c, int hh) {
        const uint32_t sc4 = sc + c * 128, bi4 = bi + c * 128;
        // all of the chunk's scale / bias vectors first: the volatile shared-memory loads keep program order, so loading them group by
        // group in front of their FMAs exposed one LDS latency per 8 columns to an in-order warp
        float4 s4[8], b4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s4[i] = lds_f4(sc4 + 16 * i); b4[i] = lds_f4(bi4 + 16 * i); }
        float a[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), s4[i].x, b4[i].x);
          a[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), s4[i].y, b4[i].y);
          a[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), s4[i].z, b4[i].z);
          a[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), s4[i].w, b4[i].w);
        }
        if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
          for (int k = 0; k < 32; ++k) a[k] = gelu_tanh_approx(a[k]);     // 32 independent chains: the MUFU.TANH latency hides in the batch
        }
        uint32_t pk[16];
        if (p.out_bf16) {                                     // warp-uniform
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(a[2 * i], a[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_op16x2(a[2 * i], a[2 * i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {                         // 8 columns -> one 16 B store
          const uint32_t j = static_cast<uint32_t>(hh * 4 + i);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + ((j ^ (lane & 7u)) << 4)), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                       "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3])
                       : "memory");
        }
      };
      uint32_t v0[32], v1[32];
      constexpr int NCW = NCH / (EW / 4);                       // chunks per warp
      const int c_lo = half * NCW, c_hi = c_lo + NCW;
      tmem_ld_32x32b_x32(t_row + c_lo * 32, v0);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32b_x32(t_row + (c + 1) * 32, v1);
        if constexpr (kBf16Out) {
          const bool live = nb * BN + c * 32 < p.N && row0_in_batch < p.rows_per_batch;       // warp-uniform
          if (live) {
            if (leader) tma_store_wait_read<0>();          // the previous pair's store has read the staging tile
            __syncwarp();
            half_bf16(v0, c, 0);
          }
          tmem_ld_wait();
          if (c + 2 < c_hi) tmem_ld_32x32b_x32(t_row + (c + 2) * 32, v0);
          if (live) {
            half_bf16(v1, c + 1, 1);                          // columns past N (N % 64 == 32) are clipped by the tensor map
            fence_proxy_async_smem();
            __syncwarp();
            if (leader) {
              tma_store_3d(&tmC, stage, nb * BN + c * 32, row0_in_batch, b);
              tma_store_commit();
            }
          }
        } else {
          process(v0, c);
          tmem_ld_wait();
          if (c + 2 < c_hi) tmem_ld_32x32b_x32(t_row + (c + 2) * 32, v0);
          process(v1, c + 1);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (leader) mbar_arrive_leader(&tempty[buf]);
      if (has_next) store_sb(buf ^ 1u);                       // every warp is past this tile's bar.sync: the other buffer is free
      if constexpr (kResid) {
        // tell the follower (layernorm_follow_kernel on another stream) that this column tile's share of rows row0 .. row0 + 31 is in
        // the residual stream: the reductions are complete (wait_group, not .read), ordered before the counter by the fences
        if (p.ready != nullptr && row0_in_batch < p.rows_per_batch && leader) {
          tma_store_wait_all();
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence();
          atomicAdd(p.ready + (row0_in_batch >> 5), 1u);
        }
      }
    }
    if constexpr (kResid || EPI == EPI_BF16 || EPI == EPI_GELU_BF16) {
      if (leader) tma_store_wait_all();     // every store / reduction of this warp has been performed before the CTA exits
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                          // neither CTA exits (or frees TMEM) while its peer may still touch it
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
std::atomic<PFN_tmapEncodeTiled> g_encode{nullptr};

template <int BN, int EPI>
int launch_variant2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmKParams& kp, int num_tiles2, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, EPI>;
  static PerDeviceOnce once;                               // the shared-memory opt-in is a per-device (per-context) attribute
  int rc = once.run([](int) -> int {
    WB_CUDA_OK(cudaFuncSetAttribute(gemm2_tn_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    return WB_OK;
  });
  if (rc != WB_OK) return rc;
  int pairs = device_sm_count() / 2;
  if (num_tiles2 < pairs) pairs = num_tiles2;
  gemm2_tn_kernel<BN, EPI><<<2 * pairs, Cfg::THREADS, Cfg::SMEM, stream>>>(ta, tb, tc, kp);
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

template <int BN>
int launch_bn2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmKParams& kp, int num_tiles2, cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_variant2<BN, EPI_BF16>(ta, tb, tc, kp, num_tiles2, s);
    case EPI_GELU_BF16: return launch_variant2<BN, EPI_GELU_BF16>(ta, tb, tc, kp, num_tiles2, s);
    case EPI_RESID_F32: return launch_variant2<BN, EPI_RESID_F32>(ta, tb, tc, kp, num_tiles2, s);
    case EPI_GELU_PE_F32: return launch_variant2<BN, EPI_GELU_PE_F32>(ta, tb, tc, kp, num_tiles2, s);
    case EPI_F32: return launch_variant2<BN, EPI_F32>(ta, tb, tc, kp, num_tiles2, s);
  }
  return set_error(WB_ERR_MODEL, "unknown GEMM epilogue");
}

}  // namespace

int device_sm_count() {
  static std::atomic<int> sms[WB_MAX_DEVICES];
  const int dev = current_device();
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  sms[dev].store(n, std::memory_order_relaxed);
  return n;
}

int gemm_init() {
  if (g_encode.load(std::memory_order_acquire) != nullptr) return WB_OK;      // a driver entry point: process-wide, not per device
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  WB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return set_error(WB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode.store(reinterpret_cast<PFN_tmapEncodeTiled>(fn), std::memory_order_release);
  return WB_OK;
}

int make_tmap_op16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  int rc = gemm_init();
  if (rc != WB_OK) return rc;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode.load(std::memory_order_acquire)(out, kOp16IsFp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu strides %llu %llu box %u x %u",
             static_cast<int>(r), (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1);
    return set_error(WB_ERR_CUDA, buf);
  }
  return WB_OK;
}

// f32 [d2][d1][d0] output view for the TMA reduction of the residual epilogue (128 B swizzle, box0 * 4 == 128 bytes)
static int make_tmap_f32_3d(CUtensorMap* out, void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                            uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, d2 > 1 ? stride2_bytes : stride1_bytes * d1};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode.load(std::memory_order_acquire)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(WB_ERR_CUDA, "cuTensorMapEncodeTiled failed for the f32 output view (" + std::to_string(static_cast<int>(r)) + ")");
  return WB_OK;
}

int gemm_tiles_n(int N) {                                   // arrivals per 32-row group on GemmDesc::ready: one epilogue warp per column tile
  const int BN = (N % 256 == 0) ? 256 : 128;
  return (N + BN - 1) / BN;
}

int launch_gemm(const GemmDesc& g, cudaStream_t stream) {
  int rc = gemm_init();
  if (rc != WB_OK) return rc;
  if (g.N % 32 != 0 || g.K % 8 != 0) return set_error(WB_ERR_MODEL, "GEMM needs N % 32 == 0 and K % 8 == 0");
  if (g.rows_per_batch <= 0 || g.n_batch <= 0) return WB_OK;
  const int BN = (g.N % 256 == 0) ? 256 : 128;
  CUtensorMap ta, tb;
  uint64_t batch_stride = g.n_batch > 1 ? static_cast<uint64_t>(g.a_batch_stride) * 2 : static_cast<uint64_t>(g.a_row_stride) * 2;
  rc = make_tmap_op16_3d(&ta, g.A, g.K, g.rows_per_batch, g.n_batch, static_cast<uint64_t>(g.a_row_stride) * 2, batch_stride,
                         BK, BM);
  if (rc != WB_OK) return rc;
  rc = make_tmap_op16_3d(&tb, g.W, g.K, g.N, 1, static_cast<uint64_t>(g.K) * 2, static_cast<uint64_t>(g.K) * 2 * g.N, BK, BN / 2);
  if (rc != WB_OK) return rc;
  GemmKParams kp;
  kp.rows_per_batch = g.rows_per_batch;
  kp.n_batch = g.n_batch;
  kp.N = g.N;
  kp.K = g.K;
  kp.tiles_m_per_batch = (g.rows_per_batch + BM - 1) / BM;
  kp.tiles_n = (g.N + BN - 1) / BN;
  kp.out_rows_per_batch = g.out_rows_per_batch;
  kp.out_row_off = g.out_row_off;
  kp.ldc = g.ldc;
  kp.alpha = g.alpha;
  kp.col_scale = g.col_scale;
  kp.bias = g.bias;
  kp.out = g.out;
  kp.pe = g.pe;
  kp.ready = (g.epilogue == EPI_RESID_F32 && g.n_batch == 1) ? g.ready : nullptr;
  kp.reverse = g.reverse ? 1 : 0;
  kp.out_bf16 = g.out_bf16 ? 1 : 0;
  kp.idesc = umma_idesc_op16(2 * BM, BN, 0);
  const int num_tiles2 = kp.n_batch * ((g.rows_per_batch + 2 * BM - 1) / (2 * BM)) * kp.tiles_n;
  CUtensorMap tc = ta;                                    // the f32 debug / positional-embedding epilogues do not read it
  if (g.epilogue == EPI_BF16 || g.epilogue == EPI_GELU_BF16) {
    op16* base = static_cast<op16*>(g.out) + static_cast<long long>(g.out_row_off) * g.ldc;
    rc = make_tmap_op16_3d(&tc, base, g.N, g.rows_per_batch, g.n_batch, static_cast<uint64_t>(g.ldc) * 2,
                           g.n_batch > 1 ? static_cast<uint64_t>(g.out_rows_per_batch) * g.ldc * 2 : static_cast<uint64_t>(g.ldc) * 2 * g.rows_per_batch,
                           64, 32);
    if (rc != WB_OK) return rc;
  }
  if (g.epilogue == EPI_RESID_F32) {
    float* base = static_cast<float*>(g.out) + static_cast<long long>(g.out_row_off) * g.ldc;
    rc = make_tmap_f32_3d(&tc, base, g.N, g.rows_per_batch, g.n_batch, static_cast<uint64_t>(g.ldc) * 4,
                          static_cast<uint64_t>(g.out_rows_per_batch) * g.ldc * 4, 32, 32);
    if (rc != WB_OK) return rc;
  }
  if (BN == 256) return launch_bn2<256>(g.epilogue, ta, tb, tc, kp, num_tiles2, stream);
  return launch_bn2<128>(g.epilogue, ta, tb, tc, kp, num_tiles2, stream);
}

}  // namespace wb
