// The decoder's front half on the GPU: cross-attention K/V of the encoder states computed ONCE per chunk, then the
// batched greedy loop with a KV cache, WhisperTokenSuppressor and argmax on the device -- only token ids go back to the host.
//
// Replaces, for B chunks decoded side by side:
//   Decoder::forward_one / forward_block_cached / compute_attention_cached     src/model/decoder.rs:2125-2172, 2241-2325, 2414-2459
//     (the non-cached path recomputes the encoder K/V projections per token, decoder.rs:2017-2040; the cached path computes them on the
//      first token -- here they are two tcgen05 GEMMs per layer straight from the encoder's op16 output buffer)
//   Decoder::project_to_vocab (weight-tied logits)                             src/model/decoder.rs:1794-1806
//   WhisperApr::decode (feed unseen tokens, suppress, pick)                    src/lib.rs:529-598
//   WhisperTokenSuppressor::{new, apply}                                       src/inference/processors.rs:60-147
//   GreedyDecoder::{argmax, decode}                                            src/inference/greedy.rs:83-146
//
// Precision: the per-token path is f32 end to end (f32 weights, f32 KV cache, f32 logits) -- one row per chunk, so it is bound by
// weight bandwidth and launch latency, not by math, and f32 keeps the argmax on the reference's side of every near-tie.  Only the
// cross-attention K/V (B x 1500 rows per layer, the GEMM-shaped part) are op16 tensor-core outputs.
#include "loader.h"
#include "ptx.cuh"

namespace wb {

struct DecodeState {
  int cap_B = 0, cap_S = 0, cap_T = 0;
  DevBuf<op16> states;         // [B*S][d]   op16 copy of host-supplied states
  DevBuf<op16> kv_cross;       // [L][B*S][2d]
  DevBuf<float> kv_self;       // [L][B][T][2d]
  DevBuf<float> x, xn, qkv, att, hid, q, logits;
  DevBuf<int> tokens;          // [B][T]
  DevBuf<int> lens, finished, pos;     // [B], [B], [1]
  DevBuf<unsigned long long> part_key;   // [argmax groups][32]: packed (ordered logit, ~token id) maxima, zero = "nothing yet"
  DevBuf<int> n_finished;      // [1]
  int argmax_groups = 0;
  // CUDA graphs of one token step (forward_one + pick + advance: 14 launches per layer, launch-bound for every Whisper size).  The
  // position is device-resident, so the same graph serves every position; keyed by the step's shape, captured at the second sighting.
  struct StepGraph {
    int B, S, T, pick, sup;
    const float* logits;
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
  };
  std::vector<StepGraph> graphs;
  uintptr_t buf_sig = 0;       // changes when a scratch buffer was reallocated: captured launches hold the old pointers
  void drop_graphs() {
    for (auto& g : graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  ~DecodeState() { drop_graphs(); }
};

namespace {

constexpr int EOT = 50257, SOT = 50258, LANG_BASE = 50259, TRANSLATE = 50358, TRANSCRIBE = 50359;
constexpr int SPEAKER_TURN = 50360, PREV = 50361, NO_SPEECH = 50362, NO_TIMESTAMPS = 50363, TIMESTAMP_BASE = 50364;
constexpr int DH = 64;

// ------------------------------------------------------------------------------------------------------------------
// x[b] = token_embedding[token_b] + positional_embedding[pos]           (decoder.rs:2147-2154)
__global__ void dec_embed_kernel(const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, const int* __restrict__ tokens,
                                 int T, const int* __restrict__ pos_p, int d, int n_vocab, float* __restrict__ x) {
  const int b = blockIdx.x, pos = *pos_p;
  int tok = tokens[b * T + pos];
  tok = min(max(tok, 0), n_vocab - 1);
  for (int i = threadIdx.x; i < d; i += blockDim.x)
    x[static_cast<size_t>(b) * d + i] = tok_emb[static_cast<size_t>(tok) * d + i] + pos_emb[static_cast<size_t>(pos) * d + i];
}

// reduce-scatter of BT per-lane partial sums over the 32 lanes of a warp: afterwards lane l holds the total of value (l % BT)
template <int BT>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[BT], int lane) {
#pragma unroll
  for (int off = BT / 2; off >= 1; off >>= 1) {
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const bool upper = (lane & off) != 0;
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float r = v[0];
#pragma unroll
  for (int off = BT; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// Skinny f32 linear layer for the token path:  y[b][n] (=|+=) act(x[b] . W[n] + bias[n]),  b < B <= 32, W [N][K] row-major
// (LinearWeights::forward, attention.rs:143-167).  One BLOCK per group of NC output columns, its threads split K: thread t owns the
// float4 columns t, t + blockDim, ... of the NC weight rows (coalesced, the next step's weights in flight under this one's FMAs) and
// keeps BT x NC partial sums; x (B x K floats, the same for every block) is read through L1.  Partial sums meet in a warp
// reduce-scatter, then across the warps through shared memory in a fixed order (deterministic).  Decoder matrices are small (d x d
// is 0.6 MB for tiny): splitting K over the block instead of giving a warp the whole row turns 3-12 dependent k-steps into 1-2 and
// spreads a d x d product over d / 4 blocks instead of d / 32.
// History: v1 staged a [32][512] slice of x in shared memory with a scalar, division-indexed copy loop (29 us per slice: 48 dependent
// L2 round trips); v2 (one warp per NC columns over all of K) 12 us for d x d, 26 us for 4d x d.
// MODE 0: store; 1: GELU then store; 2: accumulate into y (residual);  3: logits -> suppression -> packed running argmax (stores y if given)
constexpr int PICK_GROUP = 64;        // MODE 3: this many consecutive blocks share one row of packed argmax keys

__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
  // order-preserving float -> unsigned, token id complemented: the maximum key is the largest logit, ties -> the SMALLEST id, which is
  // GreedyDecoder::argmax's first strict maximum (greedy.rs:83-96)
  const uint32_t u = __float_as_uint(v);
  const uint32_t ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(ord) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(idx));
}

template <int BT, int NC, int MODE>
__global__ void __launch_bounds__(256) dec_linear_kernel(const float* __restrict__ x, int B, int K, const float* __restrict__ W,
                                                         const float* __restrict__ bias, int N, float* __restrict__ y, int ldy,
                                                         const uint8_t* __restrict__ suppress, unsigned long long* __restrict__ part_key) {
  __shared__ float part[8][NC][BT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = blockDim.x, nw = NT >> 5;
  const int n0 = blockIdx.x * NC;
  float acc[NC][BT];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[c][b] = 0.f;
  {
    const float4* wrow[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) wrow[c] = reinterpret_cast<const float4*>(W + static_cast<size_t>(min(n0 + c, N - 1)) * K);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int K4 = K >> 2;
    float4 w[NC], wn[NC];
    int k4 = tid;
    if (k4 < K4) {
#pragma unroll
      for (int c = 0; c < NC; ++c) w[c] = __ldg(wrow[c] + k4);
    }
    for (; k4 < K4; k4 += NT) {
      const bool more = k4 + NT < K4;
      if (more) {
#pragma unroll
        for (int c = 0; c < NC; ++c) wn[c] = __ldg(wrow[c] + k4 + NT);
      }
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const float4 xv = __ldg(x4 + static_cast<size_t>(min(b, B - 1)) * K4 + k4);      // rows past B repeat row B - 1 (never stored)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc[c][b] = fmaf(w[c].x, xv.x, acc[c][b]);
          acc[c][b] = fmaf(w[c].y, xv.y, acc[c][b]);
          acc[c][b] = fmaf(w[c].z, xv.z, acc[c][b]);
          acc[c][b] = fmaf(w[c].w, xv.w, acc[c][b]);
        }
      }
      if (more) {
#pragma unroll
        for (int c = 0; c < NC; ++c) w[c] = wn[c];
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float r = warp_reduce_scatter<BT>(acc[c], lane);   // lane l: this warp's total for row l % BT
    if (lane < BT) part[warp][c][lane] = r;
  }
  __syncthreads();
  for (int i = tid; i < NC * BT; i += NT) {
    const int c = i / BT, b = i - c * BT, n = n0 + c;
    if (n >= N || b >= B) continue;
    float r = part[0][c][b];
    for (int w = 1; w < nw; ++w) r += part[w][c][b];
    r += bias ? bias[n] : 0.f;
    if (MODE == 1) r = gelu_tanh(r);
    if (MODE == 2) y[static_cast<size_t>(b) * ldy + n] += r;
    else if (MODE == 3) {
      if (y) y[static_cast<size_t>(b) * ldy + n] = r;
      // WhisperTokenSuppressor::apply (processors.rs:126-147): suppressed ids never win; neither do -inf / NaN logits
      if (!suppress[n] && r > -INFINITY) atomicMax(part_key + (blockIdx.x / PICK_GROUP) * 32 + b, argmax_key(r, n));
    } else {
      y[static_cast<size_t>(b) * ldy + n] = r;
    }
  }
}

// Self-attention of the new position over the cache (forward_block_cached + compute_attention_cached, decoder.rs:2250-2266,
// 2414-2459): one block per (chunk, head); the block appends its own head slice of k_new / v_new first.
__global__ void __launch_bounds__(128) dec_self_attn_kernel(const float* __restrict__ qkv, float* __restrict__ kv_self, int T, int d,
                                                           const int* __restrict__ pos_p, float* __restrict__ out) {
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, pos = *pos_p;
  const int len = pos + 1;
  extern __shared__ float sm[];                // q[64] | p[T]
  float* sq = sm;
  float* sp = sm + DH;
  const float* row = qkv + static_cast<size_t>(b) * 3 * d;
  float* cache = kv_self + static_cast<size_t>(b) * T * 2 * d;
  if (tid < DH) {
    sq[tid] = row[h * DH + tid];
    cache[static_cast<size_t>(pos) * 2 * d + h * DH + tid] = row[d + h * DH + tid];
    cache[static_cast<size_t>(pos) * 2 * d + d + h * DH + tid] = row[2 * d + h * DH + tid];
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int t = tid; t < len; t += blockDim.x) {
    const float4* k4 = reinterpret_cast<const float4*>(cache + static_cast<size_t>(t) * 2 * d + h * DH);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) {
      const float4 kv = k4[i];
      s = fmaf(kv.x, sq[4 * i], s); s = fmaf(kv.y, sq[4 * i + 1], s); s = fmaf(kv.z, sq[4 * i + 2], s); s = fmaf(kv.w, sq[4 * i + 3], s);
    }
    s *= 0.125f;                                // 1 / sqrt(64)
    sp[t] = s;
    lmax = fmaxf(lmax, s);
  }
  __shared__ float red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  const float mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float lsum = 0.f;
  for (int t = tid; t < len; t += blockDim.x) {
    const float e = expf(sp[t] - mx);
    sp[t] = e;
    lsum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((tid & 31) == 0) red[tid >> 5] = lsum;
  __syncthreads();
  const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
  // out[i] = sum_t p[t] v[t][i]: two halves of the keys per output dim
  const int i = tid & 63, half = tid >> 6;
  float acc = 0.f;
  for (int t = half; t < len; t += 2) acc = fmaf(sp[t], cache[static_cast<size_t>(t) * 2 * d + d + h * DH + i], acc);
  __shared__ float part[2][DH];
  part[half][i] = acc;
  __syncthreads();
  if (tid < DH) out[static_cast<size_t>(b) * d + h * DH + tid] = (part[0][tid] + part[1][tid]) * inv;
}

// Cross-attention of one query row over the chunk's S cached encoder keys / values (op16 [B*S][2d]: K at column h*64, V at
// column d + h*64).  One block per (chunk, head), 256 threads.  Both passes over the cache keep several independent 128-byte row
// loads in flight per warp (the first version walked one key per thread / one row per warp step by step: 98 us per layer for 73 MB).
__global__ void __launch_bounds__(256) dec_cross_attn_kernel(const float* __restrict__ q, const op16* __restrict__ kv, int S, int d,
                                                             float* __restrict__ out) {
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ float sm[];                // p[S]
  float* sp = sm;
  const op16* base = kv + static_cast<size_t>(b) * S * 2 * d;
  // scores: a key's 64-dim row (128 B) is read by 8 lanes, 16 B each; a warp covers 4 keys per step, the block 32
  const int sub = lane & 7, g = lane >> 3;
  float qv[8];
  {
    const float4* q4 = reinterpret_cast<const float4*>(q + static_cast<size_t>(b) * d + h * DH + sub * 8);
    const float4 a = __ldg(q4), c = __ldg(q4 + 1);
    qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = c.x; qv[5] = c.y; qv[6] = c.z; qv[7] = c.w;
  }
  float lmax = -INFINITY;
  constexpr int UN = 8;
  for (int t0 = warp * 4 + g; t0 < S; t0 += 32 * UN) {
    uint4 u[UN];
#pragma unroll
    for (int r = 0; r < UN; ++r) {
      const int t = t0 + 32 * r;
      u[r] = t < S ? __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(t) * 2 * d + h * DH) + sub) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int r = 0; r < UN; ++r) {
      const uint32_t w[4] = {u[r].x, u[r].y, u[r].z, u[r].w};
      float sc = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc = fmaf(unpack_op16_lo(w[j]), qv[2 * j], sc);
        sc = fmaf(unpack_op16_hi(w[j]), qv[2 * j + 1], sc);
      }
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      sc *= 0.125f;                                // 1 / sqrt(64)
      const int t = t0 + 32 * r;
      if (t < S) {
        if (sub == 0) sp[t] = sc;
        lmax = fmaxf(lmax, sc);
      }
    }
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float lsum = 0.f;
  for (int t = tid; t < S; t += blockDim.x) {
    const float e = expf(sp[t] - mx);
    sp[t] = e;
    lsum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float inv = 1.0f / tot;
  // PV: warp w takes keys w, w + 8, ...; a lane owns two output dims (one coalesced 128 B row of V per key), 16 rows in flight
  float a0 = 0.f, a1 = 0.f;
  constexpr int UV = 16;
  const uint32_t* vbase = reinterpret_cast<const uint32_t*>(base + d + h * DH) + lane;
  const size_t vstride = static_cast<size_t>(d);           // 2d op16 per key = d 32-bit words
  for (int t0 = warp; t0 < S; t0 += 8 * UV) {
    uint32_t u[UV];
#pragma unroll
    for (int r = 0; r < UV; ++r) {
      const int t = t0 + 8 * r;
      u[r] = t < S ? __ldg(vbase + static_cast<size_t>(t) * vstride) : 0u;
    }
#pragma unroll
    for (int r = 0; r < UV; ++r) {
      const int t = t0 + 8 * r;
      const float p = t < S ? sp[t] : 0.f;
      a0 = fmaf(p, unpack_op16_lo(u[r]), a0);
      a1 = fmaf(p, unpack_op16_hi(u[r]), a1);
    }
  }
  __shared__ float part[8][DH];
  part[warp][2 * lane] = a0;
  part[warp][2 * lane + 1] = a1;
  __syncthreads();
  if (tid < DH) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += part[w][tid];
    out[static_cast<size_t>(b) * d + h * DH + tid] = r * inv;
  }
}

// Reduce the per-block argmax partials, then append the token (greedy.rs:118-146): a chunk that emitted EOT stays finished.
// step + 1 < n_init: the next token is the prompt's, nothing to pick.
__global__ void __launch_bounds__(1024) dec_pick_kernel(unsigned long long* __restrict__ part_key, int n_groups, int B, int T,
                                int* __restrict__ tokens, int* __restrict__ lens, int* __restrict__ finished, int* __restrict__ n_finished,
                                const int* __restrict__ pos_p) {
  // lane = chunk b, warp w scans key rows w, w + 32, ... and clears them for the next position
  const int b = threadIdx.x & 31, w = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  __shared__ unsigned long long sk[32][32];
  unsigned long long best = 0ull;
  for (int k = w; k < n_groups; k += n_warps) {
    const unsigned long long v = part_key[k * 32 + b];
    part_key[k * 32 + b] = 0ull;
    best = v > best ? v : best;
  }
  sk[w][b] = best;
  __syncthreads();
  if (w != 0 || b >= B) return;
  for (int k = 1; k < n_warps; ++k) best = sk[k][b] > best ? sk[k][b] : best;
  const int pos = *pos_p;
  if (finished[b]) return;
  // no key at all: every logit was -inf, NaN or suppressed -> argmax returns index 0 (greedy.rs:84-85)
  const int bi = best == 0ull ? 0 : static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(best & 0xFFFFFFFFull));
  tokens[b * T + pos + 1] = bi;
  lens[b] = pos + 2;
  if (bi == EOT) {
    finished[b] = 1;
    atomicAdd(n_finished, 1);
  }
}
__global__ void dec_advance_kernel(int* pos_p) { *pos_p += 1; }
__global__ void dec_init_kernel(int* tokens, int T, int B, const int* init, int n_init, int* lens, int* finished, int* pos, int* n_finished,
                                unsigned long long* part_key, int n_keys) {
  const int b = blockIdx.x;
  for (int i = b * blockDim.x + threadIdx.x; i < n_keys; i += gridDim.x * blockDim.x) part_key[i] = 0ull;   // a logits-only call may have left keys
  for (int i = threadIdx.x; i < T; i += blockDim.x) tokens[b * T + i] = i < n_init ? init[i] : EOT;
  if (threadIdx.x == 0) {
    lens[b] = n_init;
    finished[b] = 0;
    if (b == 0) { *pos = 0; *n_finished = 0; }
  }
}

template <int MODE>
int launch_dec_linear(const float* x, int B, int K, const float* W, const float* bias, int N, float* y, int ldy, const uint8_t* suppress,
                      unsigned long long* part_key, int* groups_out, cudaStream_t st) {
  if (K % 4 != 0) return set_error(WB_ERR_MODEL, "decoder width must be a multiple of 4");
  constexpr int NC = 4;
  const int blocks = (N + NC - 1) / NC;                       // one block per NC columns; its threads split K
  // small matrices: split K over the whole block (1-2 k-steps per thread, d / 4 blocks).  The vocabulary projection has blocks to spare
  // (13 k for 51865 tokens), and there the cross-warp reduction costs more than it saves (168 us against 120 us): one warp per block.
  const int threads = blocks >= 4096 ? 32 : std::min(256, std::max(32, ((K / 4 + 31) / 32) * 32));
  if (groups_out) *groups_out = (blocks + PICK_GROUP - 1) / PICK_GROUP;
  auto go = [&](auto bt) -> int {
    constexpr int BT = decltype(bt)::value;
    dec_linear_kernel<BT, NC, MODE><<<blocks, threads, 0, st>>>(x, B, K, W, bias, N, y, ldy, suppress, part_key);
    return WB_OK;
  };
  int rc;
  if (B <= 4) rc = go(std::integral_constant<int, 4>{});
  else if (B <= 8) rc = go(std::integral_constant<int, 8>{});
  else if (B <= 16) rc = go(std::integral_constant<int, 16>{});
  else rc = go(std::integral_constant<int, 32>{});
  if (rc != WB_OK) return rc;
  count_launch();
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

// rows of packed argmax keys the vocabulary projection needs (one per PICK_GROUP blocks of 4 columns)
static size_t key_rows(int n_vocab) { return (static_cast<size_t>(n_vocab) / 4 + PICK_GROUP) / PICK_GROUP + 1; }

int ensure_decode_state(Replica* m, int B, int S, int T) {
  if (!m->dstate) m->dstate = new DecodeState();
  DecodeState& s = *m->dstate;
  const DecoderW& w = m->dec;
  const size_t d = w.d, L = w.n_layers, b = B;
  int rc;
  if ((rc = s.kv_cross.ensure(L * b * S * 2 * d)) || (rc = s.kv_self.ensure(L * b * T * 2 * d)) || (rc = s.x.ensure(b * d)) ||
      (rc = s.xn.ensure(b * d)) || (rc = s.qkv.ensure(b * 3 * d)) || (rc = s.att.ensure(b * d)) || (rc = s.hid.ensure(b * 4 * d)) ||
      (rc = s.q.ensure(b * d)) || (rc = s.tokens.ensure(b * T)) || (rc = s.lens.ensure(b)) || (rc = s.finished.ensure(b)) ||
      (rc = s.pos.ensure(1)) || (rc = s.n_finished.ensure(1)) || (rc = s.part_key.ensure(key_rows(w.n_vocab) * 32)))
    return rc;
  s.cap_B = B; s.cap_S = S; s.cap_T = T;
  uintptr_t sig = 0;
  for (const void* q : {static_cast<const void*>(s.kv_cross.p), static_cast<const void*>(s.kv_self.p), static_cast<const void*>(s.x.p),
                        static_cast<const void*>(s.xn.p), static_cast<const void*>(s.qkv.p), static_cast<const void*>(s.att.p),
                        static_cast<const void*>(s.hid.p), static_cast<const void*>(s.q.p), static_cast<const void*>(s.tokens.p),
                        static_cast<const void*>(s.lens.p), static_cast<const void*>(s.finished.p), static_cast<const void*>(s.pos.p),
                        static_cast<const void*>(s.n_finished.p), static_cast<const void*>(s.part_key.p)})
    sig = sig * 1000003u + reinterpret_cast<uintptr_t>(q);
  if (sig != s.buf_sig) {
    s.drop_graphs();
    s.buf_sig = sig;
  }
  return WB_OK;
}

// Decoder::forward_one for the B rows at the device-resident position; want_logits: run the vocabulary projection + argmax partials
int forward_one(Replica* m, int B, int S, int T, bool want_logits, int suppress_set, float* logits_out) {
  DecodeState& s = *m->dstate;
  const DecoderW& w = m->dec;
  const int d = w.d, H = w.n_heads;
  cudaStream_t st = m->stream;
  int rc;
  dec_embed_kernel<<<B, 128, 0, st>>>(w.tok_emb, w.pos_emb, s.tokens.p, T, s.pos.p, d, w.n_vocab, s.x.p);
  count_launch();
  for (int l = 0; l < w.n_layers; ++l) {
    const DecLayerW& lw = w.layers[l];
    // self-attention over the cache
    if ((rc = launch_layernorm(s.x.p, lw.ln1_g, lw.ln1_b, B, d, nullptr, false, s.xn.p, st)) != WB_OK) return rc;
    if ((rc = launch_dec_linear<0>(s.xn.p, B, d, lw.sa_wqkv, lw.sa_bqkv, 3 * d, s.qkv.p, 3 * d, nullptr, nullptr, nullptr, st)) != WB_OK) return rc;
    float* cache = s.kv_self.p + static_cast<size_t>(l) * B * T * 2 * d;
    dec_self_attn_kernel<<<dim3(B, H), 128, (DH + T) * sizeof(float), st>>>(s.qkv.p, cache, T, d, s.pos.p, s.att.p);
    count_launch();
    if ((rc = launch_dec_linear<2>(s.att.p, B, d, lw.sa_wo, lw.sa_bo, d, s.x.p, d, nullptr, nullptr, nullptr, st)) != WB_OK) return rc;
    // cross-attention over the precomputed encoder K/V
    if ((rc = launch_layernorm(s.x.p, lw.ln2_g, lw.ln2_b, B, d, nullptr, false, s.xn.p, st)) != WB_OK) return rc;
    if ((rc = launch_dec_linear<0>(s.xn.p, B, d, lw.ca_wq, lw.ca_bq, d, s.q.p, d, nullptr, nullptr, nullptr, st)) != WB_OK) return rc;
    const op16* kv = s.kv_cross.p + static_cast<size_t>(l) * B * S * 2 * d;
    dec_cross_attn_kernel<<<dim3(B, H), 256, S * sizeof(float), st>>>(s.q.p, kv, S, d, s.att.p);
    count_launch();
    if ((rc = launch_dec_linear<2>(s.att.p, B, d, lw.ca_wo, lw.ca_bo, d, s.x.p, d, nullptr, nullptr, nullptr, st)) != WB_OK) return rc;
    // FFN
    if ((rc = launch_layernorm(s.x.p, lw.ln3_g, lw.ln3_b, B, d, nullptr, false, s.xn.p, st)) != WB_OK) return rc;
    if ((rc = launch_dec_linear<1>(s.xn.p, B, d, lw.w1, lw.b1, 4 * d, s.hid.p, 4 * d, nullptr, nullptr, nullptr, st)) != WB_OK) return rc;
    if ((rc = launch_dec_linear<2>(s.hid.p, B, 4 * d, lw.w2, lw.b2, d, s.x.p, d, nullptr, nullptr, nullptr, st)) != WB_OK) return rc;
  }
  if (want_logits) {
    if ((rc = launch_layernorm(s.x.p, w.ln_g, w.ln_b, B, d, nullptr, false, s.xn.p, st)) != WB_OK) return rc;
    if ((rc = launch_dec_linear<3>(s.xn.p, B, d, w.tok_emb, nullptr, w.n_vocab, logits_out, w.n_vocab, w.suppress[suppress_set], s.part_key.p,
                                   &s.argmax_groups, st)) != WB_OK)
      return rc;
  }
  WB_CUDA_OK(cudaGetLastError());
  return WB_OK;
}

// One token step for the B rows -- forward_one, the pick of the next token (when this position emits one) and the position advance --
// replayed from a CUDA graph from its third occurrence on (first: eager, which also runs the per-device kernel attribute set-up;
// second: captured).  Eager when graphs are disabled (WB_NO_GRAPH) or the per-kernel profile is on.
int token_step(Replica* m, int B, int S, int T, bool pick, int sup_set, float* d_logits) {
  DecodeState& s = *m->dstate;
  cudaStream_t st = m->stream;
  auto eager = [&]() -> int {
    int rc = forward_one(m, B, S, T, pick, sup_set, d_logits);
    if (rc != WB_OK) return rc;
    if (pick) {
      dec_pick_kernel<<<1, 1024, 0, st>>>(s.part_key.p, s.argmax_groups, B, T, s.tokens.p, s.lens.p, s.finished.p,
                                                     s.n_finished.p, s.pos.p);
      count_launch();
    }
    dec_advance_kernel<<<1, 1, 0, st>>>(s.pos.p);
    count_launch();
    return WB_OK;
  };
  if (!m->use_graphs || m->prof_on) return eager();
  DecodeState::StepGraph* g = nullptr;
  for (auto& e : s.graphs)
    if (e.B == B && e.S == S && e.T == T && e.pick == static_cast<int>(pick) && e.sup == sup_set && e.logits == d_logits) g = &e;
  if (!g) {
    if (s.graphs.size() >= 16) s.drop_graphs();
    DecodeState::StepGraph e;
    e.B = B; e.S = S; e.T = T; e.pick = pick; e.sup = sup_set; e.logits = d_logits; e.seen = 1;
    s.graphs.push_back(e);
    return eager();
  }
  if (g->exec) {
    WB_CUDA_OK(cudaGraphLaunch(g->exec, st));
    count_launch(static_cast<int>(g->launches));
    return WB_OK;
  }
  const long long before = g_launch_count.load();
  WB_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = eager();
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != WB_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    m->use_graphs = false;
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
  WB_CUDA_OK(cudaGraphLaunch(exec, st));
  return WB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
int load_decoder(Replica* m, const AprFile& f, Uploader& up) {
  DecoderW& w = m->dec;
  w.loaded = false;
  const size_t d = m->cfg.n_text_state, L = m->cfg.n_text_layer, V = m->cfg.n_vocab, C = m->cfg.n_text_ctx;
  const bool has_emb = up.present("decoder.embed_tokens.weight") || up.present("decoder.token_embedding");
  if (!has_emb || d == 0 || L == 0 || V == 0 || C == 0) return WB_OK;          // encoder-only file: the decode entry points refuse
  if (d % DH != 0 || m->cfg.n_text_head * DH != d) return WB_OK;               // only d_head == 64 decoders (every Whisper size)
  w.d = static_cast<int>(d); w.n_heads = static_cast<int>(m->cfg.n_text_head); w.n_layers = static_cast<int>(L);
  w.n_vocab = static_cast<int>(V); w.n_ctx = static_cast<int>(C);
  cudaStream_t st = m->stream;
  int rc;
  auto f32_param = [&](const std::string& name, size_t count, float dflt, float** out) -> int {
    int r = dev_alloc(m, count, out);
    if (r != WB_OK) return r;
    if ((r = launch_fill_f32(*out, count, dflt, st)) != WB_OK) return r;
    return up.load_f32(name, *out, count);
  };
  // token / positional embeddings: HF names first, then the OpenAI ones (lib.rs:851-871); zeros when absent (decoder.rs:1518-1522)
  if ((rc = f32_param(up.present("decoder.embed_tokens.weight") ? "decoder.embed_tokens.weight" : "decoder.token_embedding", V * d, 0.f, &w.tok_emb)) != WB_OK) return rc;
  if ((rc = f32_param(up.present("decoder.embed_positions.weight") ? "decoder.embed_positions.weight" : "decoder.positional_embedding", C * d, 0.f, &w.pos_emb)) != WB_OK) return rc;
  w.layers.resize(L);
  DevBuf<float> tmp;                                          // f32 staging of [k_proj; v_proj] before the op16 conversion
  if ((rc = tmp.ensure(2 * d * d)) != WB_OK) return rc;
  for (size_t i = 0; i < L; ++i) {
    DecLayerW& lw = w.layers[i];
    const std::string p = "decoder.layers." + std::to_string(i);
    if ((rc = f32_param(p + ".self_attn_layer_norm.weight", d, 1.f, &lw.ln1_g)) || (rc = f32_param(p + ".self_attn_layer_norm.bias", d, 0.f, &lw.ln1_b)) ||
        (rc = f32_param(p + ".encoder_attn_layer_norm.weight", d, 1.f, &lw.ln2_g)) || (rc = f32_param(p + ".encoder_attn_layer_norm.bias", d, 0.f, &lw.ln2_b)) ||
        (rc = f32_param(p + ".final_layer_norm.weight", d, 1.f, &lw.ln3_g)) || (rc = f32_param(p + ".final_layer_norm.bias", d, 0.f, &lw.ln3_b)))
      return rc;
    // self-attention: q | k | v rows stacked (three separate projections in decoder.rs:2253-2255)
    if ((rc = dev_alloc(m, 3 * d * d, &lw.sa_wqkv)) || (rc = dev_alloc(m, 3 * d, &lw.sa_bqkv))) return rc;
    if ((rc = launch_fill_f32(lw.sa_wqkv, 3 * d * d, 0.f, st)) || (rc = launch_fill_f32(lw.sa_bqkv, 3 * d, 0.f, st))) return rc;
    const char* proj[3] = {".q_proj", ".k_proj", ".v_proj"};
    for (int k = 0; k < 3; ++k) {
      if ((rc = up.load_f32(p + ".self_attn" + proj[k] + ".weight", lw.sa_wqkv + k * d * d, d * d)) != WB_OK) return rc;
      if ((rc = up.load_f32(p + ".self_attn" + proj[k] + ".bias", lw.sa_bqkv + k * d, d)) != WB_OK) return rc;
    }
    if ((rc = f32_param(p + ".self_attn.out_proj.weight", d * d, 0.f, &lw.sa_wo)) || (rc = f32_param(p + ".self_attn.out_proj.bias", d, 0.f, &lw.sa_bo))) return rc;
    // cross-attention
    if ((rc = f32_param(p + ".encoder_attn.q_proj.weight", d * d, 0.f, &lw.ca_wq)) || (rc = f32_param(p + ".encoder_attn.q_proj.bias", d, 0.f, &lw.ca_bq)) ||
        (rc = f32_param(p + ".encoder_attn.out_proj.weight", d * d, 0.f, &lw.ca_wo)) || (rc = f32_param(p + ".encoder_attn.out_proj.bias", d, 0.f, &lw.ca_bo)))
      return rc;
    if ((rc = launch_fill_f32(tmp.p, 2 * d * d, 0.f, st)) != WB_OK) return rc;
    if ((rc = up.load_f32(p + ".encoder_attn.k_proj.weight", tmp.p, d * d)) || (rc = up.load_f32(p + ".encoder_attn.v_proj.weight", tmp.p + d * d, d * d))) return rc;
    if ((rc = dev_alloc(m, 2 * d * d, &lw.ca_wkv)) != WB_OK) return rc;
    if ((rc = launch_f32_to_op16(tmp.p, lw.ca_wkv, 2 * d * d, st)) != WB_OK) return rc;
    if ((rc = dev_alloc(m, 2 * d, &lw.ca_bkv)) || (rc = launch_fill_f32(lw.ca_bkv, 2 * d, 0.f, st))) return rc;
    if ((rc = up.load_f32(p + ".encoder_attn.k_proj.bias", lw.ca_bkv, d)) || (rc = up.load_f32(p + ".encoder_attn.v_proj.bias", lw.ca_bkv + d, d))) return rc;
    // FFN
    if ((rc = f32_param(p + ".fc1.weight", 4 * d * d, 0.f, &lw.w1)) || (rc = f32_param(p + ".fc1.bias", 4 * d, 0.f, &lw.b1)) ||
        (rc = f32_param(p + ".fc2.weight", 4 * d * d, 0.f, &lw.w2)) || (rc = f32_param(p + ".fc2.bias", d, 0.f, &lw.b2)))
      return rc;
  }
  if ((rc = f32_param("decoder.layer_norm.weight", d, 1.f, &w.ln_g)) || (rc = f32_param("decoder.layer_norm.bias", d, 0.f, &w.ln_b))) return rc;
  // WhisperTokenSuppressor (processors.rs:60-84, 126-147): the seven specials, every language token, and -- in set 0 -- every
  // timestamp token from TIMESTAMP_BASE up
  std::vector<uint8_t> mask(V, 0);
  auto sup = [&](int id) { if (id >= 0 && static_cast<size_t>(id) < V) mask[id] = 1; };
  for (int id : {SOT, NO_SPEECH, TRANSLATE, TRANSCRIBE, PREV, SPEAKER_TURN, NO_TIMESTAMPS}) sup(id);
  for (int id = LANG_BASE; id < TRANSLATE; ++id) sup(id);
  for (int set = 1; set >= 0; --set) {
    if (set == 0)
      for (size_t id = TIMESTAMP_BASE; id < V; ++id) mask[id] = 1;
    if ((rc = dev_alloc(m, V, &w.suppress[set])) != WB_OK) return rc;
    WB_CUDA_OK(cudaMemcpyAsync(w.suppress[set], mask.data(), V, cudaMemcpyHostToDevice, st));
    WB_CUDA_OK(cudaStreamSynchronize(st));                   // `mask` is reused / leaves scope
  }
  WB_CUDA_OK(cudaStreamSynchronize(st));                     // `tmp` leaves scope
  w.loaded = true;
  return WB_OK;
}

void free_decode_state(Replica* m) {
  delete m->dstate;
  m->dstate = nullptr;
}

// K/V of the cross-attention for every decoder layer: [k_proj; v_proj] (2d x d, op16) applied to the B*S encoder rows by the
// tcgen05 GEMM, bias in the epilogue, op16 out [B*S][2d].  d_states: op16 [B*S][d] on this device.
static int cross_kv(Replica* m, const op16* d_states, int B, int S) {
  DecodeState& s = *m->dstate;
  const DecoderW& w = m->dec;
  const int d = w.d;
  const long long M = static_cast<long long>(B) * S;
  for (int l = 0; l < w.n_layers; ++l) {
    GemmDesc g{};
    g.A = d_states; g.a_row_stride = d; g.a_batch_stride = M * d; g.rows_per_batch = static_cast<int>(M); g.n_batch = 1;
    g.W = w.layers[l].ca_wkv; g.N = 2 * d; g.K = d;
    g.epilogue = EPI_BF16; g.alpha = 1.f; g.col_scale = nullptr; g.bias = w.layers[l].ca_bkv;
    g.out = s.kv_cross.p + static_cast<size_t>(l) * M * 2 * d; g.ldc = 2 * d; g.out_rows_per_batch = static_cast<int>(M); g.out_row_off = 0;
    g.pe = nullptr;
    int rc = launch_gemm(g, m->stream);
    if (rc != WB_OK) return rc;
  }
  return WB_OK;
}

int decoder_cross_kv(Replica* m, const op16* d_states, int B) {
  if (!m->dec.loaded) return set_error(WB_ERR_MODEL, "the .apr file carries no decoder tensors");
  int rc = ensure_decode_state(m, B, N_POS_30S, std::max(m->dstate ? m->dstate->cap_T : 0, 8));
  if (rc != WB_OK) return rc;
  return cross_kv(m, d_states, B, N_POS_30S);
}

// WhisperApr::decode with GreedyDecoder for B chunks side by side.  d_states: op16 [B][S][d] on this device.  tokens_out
// [B][max_tokens] (host, padded with EOT), lens_out [B].  The caller holds the replica lock and has set the device.
int decoder_greedy_s(Replica* m, const op16* d_states, int B, int S, const int* initial_tokens, int n_init, int max_tokens,
                     int suppress_timestamps, int* tokens_out, int* lens_out, float* logits_last_host) {
  if (!m->dec.loaded) return set_error(WB_ERR_MODEL, "the .apr file carries no decoder tensors");
  const DecoderW& w = m->dec;
  if (static_cast<size_t>(w.d) != m->cfg.n_audio_state) return set_error(WB_ERR_MODEL, "decoder width differs from the encoder's");
  if (B < 1 || B > 32) return set_error(WB_ERR_MODEL, "decode batch must be 1..32 chunks per call");
  if (n_init < 1 || !initial_tokens) return set_error(WB_ERR_MODEL, "initial tokens required");
  if (S < 1 || S > 4096) return set_error(WB_ERR_MODEL, "bad encoder sequence length");
  const int T = std::min(max_tokens, w.n_ctx);              // GreedyDecoder::new(max_tokens) with max_tokens = n_text_ctx (lib.rs:538)
  if (T < n_init) return set_error(WB_ERR_MODEL, "max_tokens smaller than the initial sequence");
  for (int i = 0; i < n_init; ++i)
    if (initial_tokens[i] < 0 || initial_tokens[i] >= w.n_vocab)
      return set_error(WB_ERR_MODEL, "token " + std::to_string(initial_tokens[i]) + " out of vocabulary range " + std::to_string(w.n_vocab));
  int rc = ensure_decode_state(m, B, S, T);
  if (rc != WB_OK) return rc;
  DecodeState& s = *m->dstate;
  cudaStream_t st = m->stream;
  DevBuf<int> d_init;
  if ((rc = d_init.ensure(n_init)) != WB_OK) return rc;
  WB_CUDA_OK(cudaMemcpyAsync(d_init.p, initial_tokens, n_init * sizeof(int), cudaMemcpyHostToDevice, st));
  dec_init_kernel<<<B, 128, 0, st>>>(s.tokens.p, T, B, d_init.p, n_init, s.lens.p, s.finished.p, s.pos.p, s.n_finished.p, s.part_key.p,
                                     static_cast<int>(key_rows(m->dec.n_vocab) * 32));
  count_launch();
  if ((rc = cross_kv(m, d_states, B, S)) != WB_OK) return rc;
  const int sup_set = suppress_timestamps ? 0 : 1;
  DevBuf<float> d_logits;
  if (logits_last_host && (rc = d_logits.ensure(static_cast<size_t>(B) * w.n_vocab)) != WB_OK) return rc;
  // position p consumes tokens[p]; from p = n_init - 1 on, its logits pick tokens[p + 1]   (while tokens.len() < max_tokens)
  int h_finished = 0;
  for (int p = 0; p + 1 < T && T > n_init; ++p) {
    const bool pick = p >= n_init - 1;
    if ((rc = token_step(m, B, S, T, pick, sup_set, logits_last_host ? d_logits.p : nullptr)) != WB_OK) return rc;
    if (pick && ((p - n_init + 2) % 16 == 0)) {              // every 16 generated tokens: has every chunk emitted EOT?
      WB_CUDA_OK(cudaMemcpyAsync(&h_finished, s.n_finished.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      WB_CUDA_OK(cudaStreamSynchronize(st));
      if (h_finished >= B) break;
    }
  }
  std::vector<int> h_tok(static_cast<size_t>(B) * T);
  WB_CUDA_OK(cudaMemcpyAsync(h_tok.data(), s.tokens.p, h_tok.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  WB_CUDA_OK(cudaMemcpyAsync(lens_out, s.lens.p, B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (logits_last_host) WB_CUDA_OK(cudaMemcpyAsync(logits_last_host, d_logits.p, static_cast<size_t>(B) * w.n_vocab * 4, cudaMemcpyDeviceToHost, st));
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_error(WB_ERR_CUDA, std::string("decoder kernels failed: ") + cudaGetErrorString(e));
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < max_tokens; ++t) tokens_out[static_cast<size_t>(b) * max_tokens + t] = t < T ? h_tok[static_cast<size_t>(b) * T + t] : EOT;
  return WB_OK;
}

// test hook: K and V of one decoder layer's cross-attention for one chunk of S rows, as f32 [S][d] each
int decoder_debug_cross_kv(Replica* m, const op16* d_states, int S, int layer, float* k_out, float* v_out) {
  if (!m->dec.loaded) return set_error(WB_ERR_MODEL, "the .apr file carries no decoder tensors");
  if (layer < 0 || layer >= m->dec.n_layers) return set_error(WB_ERR_MODEL, "decoder layer out of range");
  int rc = ensure_decode_state(m, 1, S, 8);
  if (rc != WB_OK) return rc;
  if ((rc = cross_kv(m, d_states, 1, S)) != WB_OK) return rc;
  const int d = m->dec.d;
  std::vector<op16> h(static_cast<size_t>(S) * 2 * d);
  WB_CUDA_OK(cudaMemcpyAsync(h.data(), m->dstate->kv_cross.p + static_cast<size_t>(layer) * S * 2 * d, h.size() * 2, cudaMemcpyDeviceToHost, m->stream));
  cudaError_t e = cudaStreamSynchronize(m->stream);
  if (e != cudaSuccess) return set_error(WB_ERR_CUDA, std::string("cross K/V GEMM failed: ") + cudaGetErrorString(e));
  for (int t = 0; t < S; ++t)
    for (int i = 0; i < d; ++i) {
      k_out[static_cast<size_t>(t) * d + i] = op16_to_float(h[static_cast<size_t>(t) * 2 * d + i]);
      v_out[static_cast<size_t>(t) * d + i] = op16_to_float(h[static_cast<size_t>(t) * 2 * d + d + i]);
    }
  return WB_OK;
}

int decoder_greedy(Replica* m, const op16* d_states, int B, const int* initial_tokens, int n_init, int max_tokens, int suppress_timestamps,
                   int* tokens_out, int* lens_out) {
  return decoder_greedy_s(m, d_states, B, N_POS_30S, initial_tokens, n_init, max_tokens, suppress_timestamps, tokens_out, lens_out, nullptr);
}

}  // namespace wb
