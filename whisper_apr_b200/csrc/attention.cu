// Non-causal multi-head self-attention on tcgen05 with the probabilities kept in TENSOR MEMORY.
//
// Replaces MultiHeadAttention::forward_cross_flash (src/model/attention.rs:894-935) with its per-head
// flash_attention_simd (attention.rs:472-519) of the reference: heads are 64-wide column slices of q, k, v
// (extract_head, attention.rs:1094-1107), scores are scaled by 1/sqrt(64), the softmax is the same online
// (running max / running sum) recurrence, just over KV blocks of 128 instead of 32 -- the block size does not
// change the result (the reference asserts this itself, attention.rs:2186-2228).  The encoder passes no mask.
//
// The exponential (MUFU.EX2, 16/clk/SM) is the binding pipe for d_head = 64 -- 2x the tensor time -- so everything that is
// not an exponential has to hide behind one.  (The round's first kernel -- 64-key blocks, P staged through shared memory --
// spent as long in barrier / fence / staging latency per block as in the exponentials: 0.72 ms per launch against 0.52 here.)
//   * the KV block is 128 keys: half as many synchronisation rounds per exponential;
//   * P never touches shared memory: the softmax warps write bf16 P straight back to TMEM (tcgen05.st) and the PV MMA takes
//     its A operand from TMEM -- no st.shared, no generic->async proxy fence;
//   * a softmax thread holds its whole 128-score row in registers (setmaxnreg moves registers from the four single-thread
//     role warps to the two softmax warpgroups), so S_t is released right after the tcgen05.ld and the next QK^T runs
//     under the exponentials of the current block;
//   * the two softmax warpgroups (one per 128-query tile) alternate their exponential phases through a named-barrier token,
//     so while one is MUFU-bound the other does its waits, loads, row max and P store.
//
// One persistent CTA per SM, work item = (chunk, head, pair of 128-query tiles), 384 threads:
//   warp 0      TMA producer: Q0/Q1 (double buffered across items), ring of K_j / V_j tiles ([128][64] bf16, 128 B swizzle)
//   warp 1      S_t = Q_t K_j^T issuer (4 x 128x128x16 per tile)
//   warps 2,3   O_t += P_t V_j issuers (8 x 128x64x16, A = P_t in TMEM, B = V_j MN-major from its TMA tile)
//   warps 4-7   softmax warpgroup of tile 0, warps 8-11 of tile 1; thread r owns query row r == TMEM lane r
// TMEM columns: S_t at t*128 (256), O_t at 256 + t*64 (128), P_t at 384 + t*64 (128; 128 keys x bf16 = 64 columns).
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "ptx.cuh"
#include "wb_internal.h"

namespace wb {
namespace {

constexpr int BQ = 128, BKV = 128, DH = 64;
constexpr int Q_TILE_BYTES = BQ * DH * 2;          // 16 KB
constexpr int KV_TILE_BYTES = BKV * DH * 2;        // 16 KB
constexpr int KV_STAGES = 4;
constexpr int TAIL_KEYS = 96;                       // a last KV block with <= 96 keys runs as a 96-key block (1500 = 11 * 128 + 92)
constexpr int TM_THREADS = 384;
constexpr int TM_SMEM = 4 * Q_TILE_BYTES + 2 * KV_STAGES * KV_TILE_BYTES + 256;
constexpr uint32_t TM_COLS = 512, COL_S = 0, COL_O = 256, COL_P = 384;
constexpr int REGS_ROLE = 40, REGS_SOFTMAX = 232;   // 128*40 + 256*232 == 384*168

struct AttnTmParams {
  int S, d, n_kv_blocks;
  int n_qpairs, n_heads, n_items;     // work items = B * n_heads * n_qpairs, q-pair fastest (neighbours share K/V through L2)
  float scale_log2;
  op16* out;
  int use_token;
  int n_chunks, reverse;              // reverse: chunks are walked from the last to the first (the producer's most recent rows first)
  int tail_keys;                      // a last KV block with at most this many keys runs as a 96-key block (0: never)
  uint32_t rt_zero;                   // 0, but only known at run time (exp_row's ordering trick)
};

__device__ __forceinline__ void item_coords(const AttnTmParams& p, int item, int& q0, int& h, int& b) {
  const int qp = item % p.n_qpairs;
  const int r = item / p.n_qpairs;
  h = r % p.n_heads;
  b = p.reverse ? p.n_chunks - 1 - r / p.n_heads : r / p.n_heads;
  q0 = qp * 2 * BQ;
}

// exponentials of one 128-score row -> packed bf16 pairs (what tcgen05.st writes as P) and the row sum (four partial sums).
// Measured alternatives that lost (tools/attn_bench.py, 32 x 20 x 1500 x 1500, ms per launch; this form: 0.517):
// packed FFMA2/FADD2 in batches of 16 (0.625) or pair by pair (one FFMA2 + one FADD2 per pair, 5 instead of 7 instructions: 0.597), scale+exponentials only under the token with the pack/sum after it (0.627),
// a second warpgroup per tile on half the columns (0.69), 1-3 of every 8 exponentials as a polynomial on the FMA pipe (+2..+19 %),
// a run-time loop over 32-column chunks re-read from TMEM with the pack/sum one iteration behind the exponentials (0.60).
// a streaming block (exponentials against the previous blocks' reference max, chunk by chunk under the tcgen05.ld of the next chunk,
// no token, deferred O rescale) was correct and 5 % slower (0.553).
// Handing the token over EARLY (after 48 / 32 / 16 of the 64 pairs, so that the other warpgroup's exponentials start under the tail of
// this one's; token as an mbarrier whose arrival the later pairs depend on, placement checked in the SASS) is slower the larger the
// overlap: 0.531 / 0.546 / 0.595 against 0.499 with exclusive phases and 0.599 with no token at all (profiles/r02l_attn_handover.txt):
// two warps of one scheduler inside the MUFU phase at the same time cost more than the hand-over they hide.
// Part of the row as polynomials on the FMA pipe BEFORE the warpgroup takes the token (8 / 16 / 24 of the 64 pairs computed while it
// would otherwise wait for the other warpgroup's MUFU phase, ordered in front of the barrier through its id operand; the rest on the MUFU
// behind it) is slower in every form -- 0.616 / 0.547 / 0.609 ms with ptxas free to place the MUFU part, 0.543 / 0.569 / 0.645 with that
// part pinned behind the barrier, 0.566 pinned with no polynomial at all, against 0.508 (profiles/r02an_attn_prepoly.txt): the waiting
// warpgroup's FMA work takes issue slots and FMA-pipe cycles from the warpgroup inside the phase, and the placement ptxas finds for the
// plain form (22 exponentials above the barrier, the hand-over behind the last one) is better than any the dependences can force.
// ptxas also hoists register-only work above the token's bar.sync; pinning the phase behind a post-barrier shared-memory load
// made it slower (0.549), and made the two-warpgroups-per-tile variant 0.62 instead of 0.73 -- still behind this form.
// Where the time is (measured): with every tcgen05.mma skipped the kernel takes 0.460 ms, i.e. the softmax + synchronisation
// structure alone runs at ~2530 cycles per 256 x 128 scores = two alternating phases at the lone-warp MUFU rate (microbenchmarks
// tools/micro/mufu.cu, tmem.cu: 14.5 ex2/clk/SM with one warp per scheduler, 15.9 with two; tcgen05.ld 213 B/clk/SM; neither a
// concurrent tcgen05.ld stream nor mbarrier polling costs the MUFU rate more than 4 %); the MMAs add 11 %.  Getting under ~0.46 ms
// needs two warps per scheduler inside the phase, and that variant's phase ran at 13.6 cycles per exponential with MIO-throttle
// stalls (ncu).  The mechanism (SASS): ptxas hoists every FADD / F2FP to the slot right behind the MUFU pair that produces its
// operands -- also when the source issues the exponentials as `asm volatile` groups and consumes them a group later -- so a warp
// waits out the MUFU latency (~22 cycles alone, ~35 behind a second warp's queue) once per pair: 11 cycles per exponential with one
// warp per scheduler, 13 with two.  The microbenchmark keeps 8 exponentials between producer and consumer and reaches the pipe rate.
// ncu: a lone warp per scheduler issues back-to-back MUFU.EX2 every ~9.5 cycles (8 with two warps), and ptxas places each
// FADD two instructions behind the MUFU pair it consumes, so the phase runs at ~11 cycles per exponential (72 % of the pipe).
// 2^x on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + f with the round-to-nearest magic constant, a degree-3 polynomial
// for 2^f on [-0.5, 0.5] (max relative error 1.9e-4, under half an ulp of the fp16 P it is rounded to) and the exponent added as an
// integer.  9 issue slots against one MUFU.EX2 that occupies its pipe for 8 cycles: worth it for a FRACTION of the row, where
// the independent FMA work fills the slots a lone warp otherwise spends waiting on the MUFU results it has just issued.
__device__ __forceinline__ float exp2_fma_pipe(float x) {
  x = fmaxf(x, -125.0f);                                   // exponent stays in range; masked keys (-inf) -> 2^-125 * ~1 ~ 0 in fp16
  const float t = x + 12582912.0f;                         // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.055875536f, 0.24229462f);
  p = fmaf(p, f, 0.6931273f);
  p = fmaf(p, f, 0.99994826f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// POLY 0: every exponential on the MUFU; n > 0: one pair in n goes to the FMA pipe instead; n < 0: MUFU only, consumers lag.
// NP: pairs of the row that exist (64 = a full 128-key block; 48 = the short tail block, whose last 32 columns are never computed).
template <int POLY, int NP, bool PBF>     // PBF: P (and Q, K, V) are bf16 -- see launch_attention
__device__ __forceinline__ float exp_row(const uint32_t (&s)[128], uint32_t (&pk)[64], float c, float mb, uint32_t rt_zero) {
  float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = NP; i < 64; ++i) pk[i] = 0u;               // stored with the rest, never read by the shortened PV product
  if (POLY >= 0) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float x0 = fmaf(__uint_as_float(s[2 * i]), c, -mb), x1 = fmaf(__uint_as_float(s[2 * i + 1]), c, -mb);
      float e0, e1;
      constexpr int PERIOD = POLY > 0 ? POLY : 1;
      if (POLY > 0 && (i % PERIOD) == PERIOD - 1) {
        e0 = exp2_fma_pipe(x0);
        e1 = exp2_fma_pipe(x1);
      } else {
        e0 = fast_exp2(x0);                                  // exp2(-inf) == 0: masked keys contribute nothing
        e1 = fast_exp2(x1);
      }
      rs4[i & 3] += e0 + e1;
      pk[i] = PBF ? pack_bf16x2(e0, e1) : pack_op16x2(e0, e1);
    }
  } else {
    // consumers LAG = -POLY pairs behind their exponentials, with a true dependence that keeps ptxas from re-pairing them: the sum of
    // pair i takes its second operand through a select on a value of pair i + LAG (always false: exponentials are never negative)
    constexpr int LAG = -POLY;
    float e[128];
#pragma unroll
    for (int i = 0; i < NP + LAG; ++i) {
      if (i < NP) {
        e[2 * i] = fast_exp2(fmaf(__uint_as_float(s[2 * i]), c, -mb));
        e[2 * i + 1] = fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), c, -mb));
      }
      if (i >= LAG) {
        const int j = i - LAG;
        float b = e[2 * j + 1];
        // b | (bits(e of pair i) & 0): one LOP3 whose third operand is a RUN-TIME zero, so neither nvcc nor ptxas can drop the
        // dependence -- the sum of pair j is ordered after the exponentials of pair i = j + LAG
        if (i < NP) b = __uint_as_float(__float_as_uint(b) | (__float_as_uint(e[2 * i + 1]) & rt_zero));
        rs4[j & 3] += e[2 * j] + b;
        pk[j] = PBF ? pack_bf16x2(e[2 * j], e[2 * j + 1]) : pack_op16x2(e[2 * j], e[2 * j + 1]);
      }
    }
  }
  return (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
}

template <int POLY, bool PBF = false>
__global__ void __launch_bounds__(TM_THREADS, 1)
attention_tm_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                                   // [qbuf][tile]
  uint8_t* sK = smem + 4 * Q_TILE_BYTES;                // [KV_STAGES]
  uint8_t* sV = sK + KV_STAGES * KV_TILE_BYTES;         // [KV_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * KV_TILE_BYTES);
  uint64_t* q_full = bars;                              // [2]
  uint64_t* q_empty = bars + 2;                         // [2]
  uint64_t* k_full = bars + 4;                          // [KV_STAGES]
  uint64_t* v_full = bars + 8;                          // [KV_STAGES]
  uint64_t* kv_empty = bars + 12;                       // [KV_STAGES]  3 arrivals: S issuer + two PV issuers
  uint64_t* s_full = bars + 16;                         // [tile]
  uint64_t* s_free = bars + 18;                         // [tile]  4 arrivals (one per softmax warp): S_t is in registers
  uint64_t* p_full = bars + 20;                         // [tile]  4 arrivals: P_t is in TMEM
  uint64_t* pv_done = bars + 22;                        // [tile]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n = p.n_kv_blocks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 3); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp < 4) {
    reg_dealloc<REGS_ROLE>();
    if (warp == 0) {
      if (elect_one()) {                                  // one thread, known to the compiler: descriptors go to uniform registers without waterfall loops
        // ------------------------------------------------------------------ TMA producer
        uint32_t g = 0;                                   // running KV block counter across items
        int it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
          int q0, h, b;
          item_coords(p, item, q0, h, b);
          const int qb = it & 1;
          mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qb], 2 * Q_TILE_BYTES);
          tma_load_3d(sQ + (qb * 2) * Q_TILE_BYTES, &tmQKV, &q_full[qb], h * DH, q0, b);
          tma_load_3d(sQ + (qb * 2 + 1) * Q_TILE_BYTES, &tmQKV, &q_full[qb], h * DH, q0 + BQ, b);
          for (int j = 0; j < n; ++j, ++g) {
            const uint32_t st = g % KV_STAGES;
            mbar_wait(&kv_empty[st], ((g / KV_STAGES) & 1u) ^ 1u);
            mbar_expect_tx(&k_full[st], KV_TILE_BYTES);
            tma_load_3d(sK + st * KV_TILE_BYTES, &tmQKV, &k_full[st], p.d + h * DH, j * BKV, b);
            mbar_expect_tx(&v_full[st], KV_TILE_BYTES);
            tma_load_3d(sV + st * KV_TILE_BYTES, &tmQKV, &v_full[st], 2 * p.d + h * DH, j * BKV, b);
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ------------------------------------------------------------------ S_t = Q_t K^T issuer (both tiles)
        constexpr uint32_t kFmt = PBF ? 1u : kIdescOp16Fmt;
        constexpr uint32_t idesc_s = umma_idesc_16(BQ, BKV, 0, kFmt);
        constexpr uint32_t idesc_s_tail = umma_idesc_16(BQ, TAIL_KEYS, 0, kFmt);     // short tail block: 96 key columns instead of 128
        const bool short_tail = p.S - (n - 1) * BKV <= p.tail_keys;
        uint32_t g = 0;
        int it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
          const int qb = it & 1;
          mbar_wait(&q_full[qb], (it >> 1) & 1);
          const uint64_t qd[2] = {umma_desc_sw128(smem_u32(sQ + (qb * 2) * Q_TILE_BYTES)),
                                  umma_desc_sw128(smem_u32(sQ + (qb * 2 + 1) * Q_TILE_BYTES))};
          for (int j = 0; j < n; ++j, ++g) {
            const uint32_t st = g % KV_STAGES;
            mbar_wait(&k_full[st], (g / KV_STAGES) & 1u);
            const uint64_t kd = umma_desc_sw128(smem_u32(sK + st * KV_TILE_BYTES));
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (g >= 1) mbar_wait(&s_free[t], (g - 1) & 1u);         // the softmax warpgroup holds S_t(g-1) in registers
              tc_fence_after_sync();
              const uint32_t idesc = (short_tail && j == n - 1) ? idesc_s_tail : idesc_s;
#pragma unroll
              for (int k = 0; k < DH / 16; ++k) umma_f16(tmem_base + COL_S + t * 128, qd[t] + 2 * k, kd + 2 * k, idesc, k != 0);
              umma_commit(&s_full[t]);
            }
            umma_commit(&kv_empty[st]);
          }
          umma_commit(&q_empty[qb]);                                    // every S MMA of this item has been issued
        }
      }
    } else {
      if (elect_one()) {
        // ------------------------------------------------------------------ O_t += P_t V issuer, one warp per tile
        constexpr uint32_t idesc_o = umma_idesc_16(BQ, DH, 1, PBF ? 1u : kIdescOp16Fmt);        // B = V tile, MN-major
        const int t = warp - 2;
        const uint32_t tO = tmem_base + COL_O + t * 64;
        const uint32_t tP = tmem_base + COL_P + t * 64;
        const bool short_tail = p.S - (n - 1) * BKV <= p.tail_keys;
        uint32_t g = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
          for (int j = 0; j < n; ++j, ++g) {
            const uint32_t st = g % KV_STAGES;
            mbar_wait(&v_full[st], (g / KV_STAGES) & 1u);
            mbar_wait(&p_full[t], g & 1u);                              // P_t(g) is in TMEM
            tc_fence_after_sync();
            const uint64_t vd = umma_desc_sw128(smem_u32(sV + st * KV_TILE_BYTES));
            const int ksteps = (short_tail && j == n - 1) ? TAIL_KEYS / 16 : BKV / 16;    // keys past the tail's 96 are never multiplied
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              if (k < ksteps) umma_f16_ts(tO, tP + 8 * k, vd + 128 * k, idesc_o, (j | k) != 0);
            umma_commit(&pv_done[t]);
            umma_commit(&kv_empty[st]);
          }
        }
      }
    }
  } else {
    reg_alloc<REGS_SOFTMAX>();
    // -------------------------------------------------------------------- softmax warpgroups
    const int t = (warp - 4) >> 2;                        // query tile of this warpgroup
    const int r = (warp & 3) * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + COL_S + t * 128 + lane_sel;
    const uint32_t tO = tmem_base + COL_O + t * 64 + lane_sel;
    const uint32_t tP = tmem_base + COL_P + t * 64 + lane_sel;
    const float c = p.scale_log2;
    const int my_items = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t total_blocks = static_cast<uint32_t>(my_items) * n;
    if (p.use_token && t == 1) asm volatile("bar.arrive %0, 256;" ::"r"(1) : "memory");      // warpgroup 0 owns the first token

    uint32_t g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int q0, h, b;
      item_coords(p, item, q0, h, b);
      float m_ref = -INFINITY, l_run = 0.f;
      // one KV block; `kind` selects the variant: 0 = a full block (no compare/select instructions in the common path), 128 = a
      // masked tail block, 96 = a masked tail block with at most 96 keys: its last 32 score columns are neither computed by the
      // QK^T product, nor loaded, nor exponentiated, nor multiplied into O (S = 1500: 92 keys, a quarter of the tail block's work)
      auto block_body = [&](int j, auto kind) {
        constexpr int KIND = decltype(kind)::value;
        constexpr int NCOL = KIND == 0 ? BKV : KIND;
        mbar_wait(&s_full[t], g & 1u);
        tc_fence_after_sync();
        uint32_t s[128];
#pragma unroll
        for (int q = 0; q < NCOL / 32; ++q) tmem_ld_32x32b_x32(tS + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[q * 32]));
        bool pv_ok = g == 0;                                 // P_t / O_t are free once PV of the previous block retired (waited late)
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);              // S_t(g+1) may now overwrite S_t
        if constexpr (KIND != 0) {
          const int kv_valid = p.S - j * BKV;
#pragma unroll
          for (int i = 0; i < NCOL; ++i)
            if (i >= kv_valid) s[i] = 0xff800000u;           // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < NCOL / 2; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], fmaxf(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
        const bool need = (mx - m_ref) * c > 8.0f;
        float alpha = 1.0f;
        if (need) {
          alpha = fast_exp2((m_ref - mx) * c);              // 0 on the first block (m_ref = -inf)
          m_ref = mx;
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          if (!pv_ok) { mbar_wait(&pv_done[t], (g - 1) & 1u); pv_ok = true; }     // O_t is stable
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
        const float mb = m_ref * c;
        // ping-pong token: MUFU phases of the two warpgroups alternate
        if (p.use_token) asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
        uint32_t pk[64];
        const float rsum = exp_row<POLY, NCOL / 2, PBF>(s, pk, c, mb, p.rt_zero);
        if (p.use_token && !(t == 1 && g + 1 == total_blocks)) asm volatile("bar.arrive %0, 256;" ::"r"(2 - t) : "memory");   // hand the token over
        l_run = l_run * alpha + rsum;
        if (!pv_ok) mbar_wait(&pv_done[t], (g - 1) & 1u);    // PV of the previous block has read P_t
        tc_fence_after_sync();
        tmem_st_32x32b_x32(tP, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
        tmem_st_32x32b_x32(tP + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      };
      for (int j = 0; j < n; ++j, ++g) {
        const int kv_left = p.S - j * BKV;
        if (kv_left >= BKV) block_body(j, std::integral_constant<int, 0>{});
        else if (kv_left <= p.tail_keys) block_body(j, std::integral_constant<int, TAIL_KEYS>{});
        else block_body(j, std::integral_constant<int, BKV>{});
      }
      // epilogue: O_t / l  (attention.rs:334-343: 0 when the sum is <= 1e-10)
      mbar_wait(&pv_done[t], (g - 1) & 1u);
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
          for (int gg = 0; gg < 4; ++gg) {
            uint4 w;
            w.x = pack_op16x2(__uint_as_float(v[8 * gg + 0]) * inv, __uint_as_float(v[8 * gg + 1]) * inv);
            w.y = pack_op16x2(__uint_as_float(v[8 * gg + 2]) * inv, __uint_as_float(v[8 * gg + 3]) * inv);
            w.z = pack_op16x2(__uint_as_float(v[8 * gg + 4]) * inv, __uint_as_float(v[8 * gg + 5]) * inv);
            w.w = pack_op16x2(__uint_as_float(v[8 * gg + 6]) * inv, __uint_as_float(v[8 * gg + 7]) * inv);
            dst[cc * 4 + gg] = w;
          }
        }
      }
      tc_fence_before_sync();       // the next item's first PV (issued after this warpgroup's next p_full) overwrites O_t
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

PerDeviceOnce g_tm_once;

}  // namespace

int attention_init() {
  return g_tm_once.run([](int) -> int {
    WB_CUDA_OK(cudaFuncSetAttribute(attention_tm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
    WB_CUDA_OK(cudaFuncSetAttribute(attention_tm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
    WB_CUDA_OK(cudaFuncSetAttribute(attention_tm_kernel<-2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
    WB_CUDA_OK((cudaFuncSetAttribute(attention_tm_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM)));
    return WB_OK;
  });
}

int launch_attention(const op16* qkv, op16* out, int B, int S, int d, int n_heads, cudaStream_t stream, bool reverse, bool qkv_bf16) {
  int rc = attention_init();
  if (rc != WB_OK) return rc;
  if (B <= 0 || S <= 0) return WB_OK;
  if (d != n_heads * DH) return set_error(WB_ERR_MODEL, "attention kernel needs d_head == 64 (all Whisper sizes)");
  CUtensorMap tm;      // one map serves Q, K and V tiles: box = 64 columns x 128 rows of the [B][S][3d] qkv buffer
  rc = make_tmap_op16_3d(&tm, qkv, 3ull * d, S, B, 3ull * d * 2, 3ull * d * 2 * S, DH, BQ);
  if (rc != WB_OK) return rc;
  AttnTmParams p;
  p.S = S;
  p.d = d;
  p.n_kv_blocks = (S + BKV - 1) / BKV;
  p.n_qpairs = (S + 2 * BQ - 1) / (2 * BQ);
  p.n_heads = n_heads;
  p.n_items = B * n_heads * p.n_qpairs;
  p.n_chunks = B;
  p.reverse = reverse ? 1 : 0;
  p.scale_log2 = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) * log2(e)
  p.out = out;
  static const int no_token = getenv("WB_ATTN_NOTOKEN") != nullptr;       // tuning switch
  p.use_token = no_token ? 0 : 1;
  static const int no_short_tail = getenv("WB_ATTN_NO_SHORT_TAIL") != nullptr;   // A/B switch
  p.tail_keys = no_short_tail ? 0 : TAIL_KEYS;
  p.rt_zero = 0;
  const int sms = device_sm_count();
  const int grid = p.n_items < sms ? p.n_items : sms;
  // Tuning switch, round 2 (profiles/r02g_attention_variants.txt, ms per launch alone, 32 x 20 x 1500 x 1500): default 0.514;
  // WB_ATTN_POLY=8 / 6 / 4 (one pair in 8 / 6 / 4 through exp2_fma_pipe): 0.530 / 0.539 / 0.561; -2 / -4 (consumers forced 2 / 4 pairs
  // behind their exponentials by a true dependence): 0.520 / 0.519.  Neither fewer MUFU operations nor a longer producer-consumer
  // distance shortens the phase: it runs at ~81 % of the MUFU pipe's rate at the clock of the run (2530 of 2048 cycles per 256 x 128
  // scores) and the remainder is the hand-over between the two alternating warpgroups, not the instruction mix inside the phase.
  static const int poly = getenv("WB_ATTN_POLY") ? atoi(getenv("WB_ATTN_POLY")) : 0;
  if (qkv_bf16 && poly == 0) {
    // Q, K, V arrive as bf16 and P is packed as bf16: both products run as bf16 MMAs, the output is still written in the operand format
    // (it feeds out_proj).  In an fp16-operand build this is the one place where bf16 costs nothing in accuracy -- no weights are involved;
    // CPU emulation at depth 32: max-abs 2.79e-3 against 2.58e-3 -- and its MMAs draw less power under the board's cap.
    attention_tm_kernel<0, true><<<grid, TM_THREADS, TM_SMEM, stream>>>(tm, p);
    count_launch();
    WB_CUDA_OK(cudaGetLastError());
    return WB_OK;
  }
  if (qkv_bf16 && kOp16IsFp16) return set_error(WB_ERR_MODEL, "the attention tuning variants take operand-format q, k, v");
  switch (poly) {
    case 8: attention_tm_kernel<8><<<grid, TM_THREADS, TM_SMEM, stream>>>(tm, p); break;
    case -2: attention_tm_kernel<-2><<<grid, TM_THREADS, TM_SMEM, stream>>>(tm, p); break;
    default: attention_tm_kernel<0><<<grid, TM_THREADS, TM_SMEM, stream>>>(tm, p); break;
  }
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

}  // namespace wb
