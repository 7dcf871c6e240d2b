/* whisper_ref.c -- plain-C restatement of whisper.apr's mel + encoder CPU path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path and the timed CPU baseline of
 * bench.py.  The reference itself is Rust and cannot be compiled here (no cargo/rustc, crates not vendored), so this
 * file keeps the reference's loop structure, operation order and threading model function by function:
 *
 *   wref_mel_compute     MelFilterbank::compute            src/audio/mel.rs:233-310  (FFT: rustfft 6.4.1 planner,
 *                                                          call site mel.rs:256-257,279 -- restated as a mixed-radix
 *                                                          2/5 complex FFT of the windowed frame)
 *   wref_compute_mel     WhisperApr::compute_mel           src/lib.rs:407-443
 *   wref_conv1d          Conv1d::forward                   src/model/encoder.rs:72-110 (direct 4-nested loops)
 *   wref_gelu            gelu                              src/model/encoder.rs:314-318
 *   wref_layernorm       LayerNorm::forward                src/model/encoder.rs:219-251
 *   wref_linear_scalar   LinearWeights::forward            src/model/attention.rs:143-167 (scalar; what the FFN uses,
 *                                                          encoder.rs:286,294)
 *   wref_linear_simd     LinearWeights::forward_simd       src/model/attention.rs:181-219 -> trueno Matrix::matmul
 *                                                          (un-vendored crate trueno 0.10.1; restated as a cache-blocked,
 *                                                          compiler-vectorised A[MxK] . Wt[KxN] + broadcast bias)
 *   wref_flash_attention flash_attention{,_simd}           src/model/attention.rs:360-406, 472-519 (KV block 32)
 *   wref_mha             forward_cross_flash               src/model/attention.rs:894-935 (parallel_map over heads,
 *                                                          src/parallel.rs:82-98 -> OpenMP when threads > 1)
 *   wref_encoder_layer   EncoderBlock::forward             src/model/encoder.rs:346-361
 *   wref_conv_stem       ConvFrontend::forward + pos-emb   src/model/encoder.rs:161-175, 464-469
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define N_FFT 400
#define N_FREQ 201

/* ------------------------------------------------------------------------------------------- FFT */
typedef struct { float re, im; } cpx;

/* recursive mixed-radix decimation-in-time FFT, radices 2 and 5 (400 = 2^4 * 5^2), forward, un-normalised */
static void fft_rec(const cpx* in, cpx* out, int n, int stride, const cpx* tw, int tw_stride) {
  if (n == 1) { out[0] = in[0]; return; }
  int r = (n % 2 == 0) ? 2 : 5;
  int m = n / r;
  for (int q = 0; q < r; ++q) fft_rec(in + q * stride, out + q * m, m, stride * r, tw, tw_stride * r);
  cpx tmp[5];
  for (int k = 0; k < m; ++k) {
    for (int q = 0; q < r; ++q) {
      cpx w = tw[(size_t)q * k * tw_stride % N_FFT];
      cpx v = out[q * m + k];
      tmp[q].re = v.re * w.re - v.im * w.im;
      tmp[q].im = v.re * w.im + v.im * w.re;
    }
    /* r-point DFT of the twiddled column; results staged so the inputs stay intact */
    cpx col[5];
    for (int p = 0; p < r; ++p) {
      float sr = 0.f, si = 0.f;
      for (int q = 0; q < r; ++q) {
        cpx w = tw[(size_t)((p * q) % r) * (N_FFT / r) % N_FFT];
        sr += tmp[q].re * w.re - tmp[q].im * w.im;
        si += tmp[q].re * w.im + tmp[q].im * w.re;
      }
      col[p].re = sr; col[p].im = si;
    }
    for (int p = 0; p < r; ++p) out[k + p * m] = col[p];
  }
}

static void make_twiddles(cpx* tw) {
  for (int i = 0; i < N_FFT; ++i) {
    double a = -2.0 * 3.14159265358979323846 * i / N_FFT;
    tw[i].re = (float)cos(a);
    tw[i].im = (float)sin(a);
  }
}

/* ------------------------------------------------------------------------------------------- mel */
/* returns number of frames; out [n_frames][n_mels] frame-major; hop == 0 -> -1 (WhisperError::Audio) */
long wref_mel_compute(const float* audio, long n, const float* filters, int n_mels, long hop, float* out) {
  if (n == 0) return 0;
  if (hop == 0) return -1;
  long n_frames = n >= N_FFT ? (n - N_FFT) / hop + 1 : 0;
  if (n_frames == 0) return 0;
  float window[N_FFT];
  const float pi = 3.14159265358979323846f;
  for (int i = 0; i < N_FFT; ++i) window[i] = 0.5f * (1.0f - cosf(2.0f * pi * (float)i / (float)N_FFT));
  cpx tw[N_FFT];
  make_twiddles(tw);                                       /* "planner" work, once per call (mel.rs:256-257) */
  cpx fin[N_FFT], fout[N_FFT];
  float power[N_FREQ];
  for (long f = 0; f < n_frames; ++f) {
    long start = f * hop;
    for (int i = 0; i < N_FFT; ++i) {
      float s = (start + i < n) ? audio[start + i] : 0.0f;
      fin[i].re = s * window[i];
      fin[i].im = 0.0f;
    }
    fft_rec(fin, fout, N_FFT, 1, tw, 1);
    for (int k = 0; k < N_FREQ; ++k) power[k] = fout[k].re * fout[k].re + fout[k].im * fout[k].im;
    for (int j = 0; j < n_mels; ++j) {
      float e = 0.0f;
      const float* fr = filters + (size_t)j * N_FREQ;
      for (int k = 0; k < N_FREQ; ++k) e += fr[k] * power[k];          /* dense 201-MAC dot, mel.rs:290-295 */
      out[f * n_mels + j] = log10f(e > 1e-10f ? e : 1e-10f);
    }
  }
  float mx = -INFINITY;
  for (long i = 0; i < n_frames * n_mels; ++i) mx = out[i] > mx ? out[i] : mx;
  for (long i = 0; i < n_frames * n_mels; ++i) {
    float v = out[i] > mx - 8.0f ? out[i] : mx - 8.0f;
    out[i] = (v + 4.0f) / 4.0f;
  }
  return n_frames;
}

/* out [3000][n_mels] */
void wref_compute_mel(const float* audio, long n, const float* filters, int n_mels, float* out) {
  const long NS = 480000, NF = 3000;
  float* padded = (float*)calloc(NS, sizeof(float));
  memcpy(padded, audio, (size_t)(n < NS ? n : NS) * sizeof(float));
  float* mel = (float*)malloc((size_t)NF * n_mels * sizeof(float));
  long frames = wref_mel_compute(padded, NS, filters, n_mels, 160, mel);
  for (long i = 0; i < NF * n_mels; ++i) out[i] = -1.0f;
  memcpy(out, mel, (size_t)(frames < NF ? frames : NF) * n_mels * sizeof(float));
  free(mel);
  free(padded);
}

/* --------------------------------------------------------------------------------------- encoder */
static inline float gelu1(float x) {
  const float c = 0.7978846f, k = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(c * (x + k * x * x * x)));
}
void wref_gelu(float* x, long n) {
  for (long i = 0; i < n; ++i) x[i] = gelu1(x[i]);
}

/* input [seq][cin], weight [cout][cin][K=3], out [out_len][cout]; returns out_len */
long wref_conv1d(const float* in, long seq, int cin, const float* w, const float* bias, int cout, int stride, int padding,
                 float* out) {
  const int K = 3;
  long out_len = (seq + 2 * padding - K) / stride + 1;
  for (long p = 0; p < out_len; ++p) {
    long in_start = p * stride - padding;
    for (int oc = 0; oc < cout; ++oc) {
      float sum = bias[oc];
      for (int k = 0; k < K; ++k) {
        long ip = in_start + k;
        if (ip >= 0 && ip < seq)
          for (int ic = 0; ic < cin; ++ic) sum += w[((size_t)oc * cin + ic) * K + k] * in[ip * cin + ic];
      }
      out[p * cout + oc] = sum;
    }
  }
  return out_len;
}

void wref_layernorm(const float* x, long rows, int d, const float* g, const float* b, float* out) {
  for (long s = 0; s < rows; ++s) {
    const float* r = x + s * d;
    float sum = 0.f;
    for (int i = 0; i < d; ++i) sum += r[i];
    float mean = sum / (float)d, var = 0.f;
    for (int i = 0; i < d; ++i) var += (r[i] - mean) * (r[i] - mean);
    var /= (float)d;
    float inv = 1.0f / sqrtf(var + 1e-5f);
    for (int i = 0; i < d; ++i) out[s * d + i] = (r[i] - mean) * inv * g[i] + b[i];
  }
}

/* y[s][o] = b[o] + sum_i x[s][i] w[o][i]   (scalar triple loop) */
void wref_linear_scalar(const float* x, long rows, int in_f, const float* w, const float* b, int out_f, float* y) {
  for (long s = 0; s < rows; ++s)
    for (int o = 0; o < out_f; ++o) {
      float sum = b[o];
      const float* xr = x + s * in_f;
      const float* wr = w + (size_t)o * in_f;
      for (int i = 0; i < in_f; ++i) sum += xr[i] * wr[i];
      y[s * out_f + o] = sum;
    }
}

/* wt [in_f][out_f] = transpose of w (finalize_weights, attention.rs:96-105); y = x . wt + b, blocked and vectorisable */
void wref_linear_simd(const float* x, long rows, int in_f, const float* wt, const float* b, int out_f, float* y) {
  const int KB = 256;
  for (long s = 0; s < rows; ++s) memset(y + s * out_f, 0, (size_t)out_f * sizeof(float));
  for (int k0 = 0; k0 < in_f; k0 += KB) {
    int k1 = k0 + KB < in_f ? k0 + KB : in_f;
    for (long s = 0; s < rows; ++s) {
      float* yr = y + s * out_f;
      for (int k = k0; k < k1; ++k) {
        float a = x[s * in_f + k];
        const float* wr = wt + (size_t)k * out_f;
        for (int o = 0; o < out_f; ++o) yr[o] += a * wr[o];
      }
    }
  }
  for (long s = 0; s < rows; ++s)
    for (int o = 0; o < out_f; ++o) y[s * out_f + o] += b[o];
}

/* q [S][dh], k,v [KV][dh] contiguous per head; out [S][dh] */
void wref_flash_attention(const float* q, const float* k, const float* v, long S, long KV, int dh, int block, float* out) {
  float scale = 1.0f / sqrtf((float)dh);
  float* row_max = (float*)malloc((size_t)S * sizeof(float));
  float* row_sum = (float*)calloc(S, sizeof(float));
  float* scores = (float*)malloc((size_t)block * sizeof(float));
  memset(out, 0, (size_t)S * dh * sizeof(float));
  for (long i = 0; i < S; ++i) row_max[i] = -INFINITY;
  for (long k0 = 0; k0 < KV; k0 += block) {
    long k1 = k0 + block < KV ? k0 + block : KV;
    for (long qi = 0; qi < S; ++qi) {
      float bmax = -INFINITY;
      for (long kj = k0; kj < k1; ++kj) {
        float dot = 0.f;
        for (int d = 0; d < dh; ++d) dot += q[qi * dh + d] * k[kj * dh + d];
        scores[kj - k0] = dot * scale;
        bmax = scores[kj - k0] > bmax ? scores[kj - k0] : bmax;
      }
      float prev = row_max[qi], nm = prev > bmax ? prev : bmax;
      float sp = expf(prev - nm);
      row_sum[qi] *= sp;
      float* o = out + qi * dh;
      for (int d = 0; d < dh; ++d) o[d] *= sp;
      for (long kj = k0; kj < k1; ++kj) {
        float e = expf(scores[kj - k0] - nm);
        row_sum[qi] += e;
        const float* vr = v + kj * dh;
        for (int d = 0; d < dh; ++d) o[d] += e * vr[d];
      }
      row_max[qi] = nm;
    }
  }
  for (long qi = 0; qi < S; ++qi) {
    float inv = row_sum[qi] > 1e-10f ? 1.0f / row_sum[qi] : 0.0f;
    for (int d = 0; d < dh; ++d) out[qi * dh + d] *= inv;
  }
  free(scores); free(row_sum); free(row_max);
}

typedef struct {
  const float *q, *k, *v;
  float* concat;
  long S;
  int d, dh, h_next, h_end;
} head_job;

/* one rayon-style worker: claims heads until none are left */
static void* head_worker(void* arg) {
  head_job* j = (head_job*)arg;
  const long S = j->S;
  const int d = j->d, dh = j->dh;
  for (;;) {
    int h = __atomic_fetch_add(&j->h_next, 1, __ATOMIC_RELAXED);
    if (h >= j->h_end) break;
    float* qh = (float*)malloc((size_t)S * dh * sizeof(float));      /* extract_head copies (attention.rs:1094-1107) */
    float* kh = (float*)malloc((size_t)S * dh * sizeof(float));
    float* vh = (float*)malloc((size_t)S * dh * sizeof(float));
    float* oh = (float*)malloc((size_t)S * dh * sizeof(float));
    for (long s = 0; s < S; ++s)
      for (int c = 0; c < dh; ++c) {
        qh[s * dh + c] = j->q[s * d + h * dh + c];
        kh[s * dh + c] = j->k[s * d + h * dh + c];
        vh[s * dh + c] = j->v[s * d + h * dh + c];
      }
    wref_flash_attention(qh, kh, vh, S, S, dh, 32, oh);
    for (long s = 0; s < S; ++s)
      for (int c = 0; c < dh; ++c) j->concat[s * d + h * dh + c] = oh[s * dh + c];   /* concat_heads */
    free(qh); free(kh); free(vh); free(oh);
  }
  return NULL;
}

/* self-attention; wq_t.. are TRANSPOSED weights [d][d] (forward_simd path); heads [h0, h1) only (sampling) */
void wref_mha(const float* x, long S, int d, int n_heads, const float* wq_t, const float* bq, const float* wk_t, const float* bk,
              const float* wv_t, const float* bv, const float* wo_t, const float* bo, int threads, int h0, int h1, float* out) {
  int dh = d / n_heads;
  float* q = (float*)malloc((size_t)S * d * sizeof(float));
  float* k = (float*)malloc((size_t)S * d * sizeof(float));
  float* v = (float*)malloc((size_t)S * d * sizeof(float));
  float* concat = (float*)calloc((size_t)S * d, sizeof(float));
  wref_linear_simd(x, S, d, wq_t, bq, d, q);
  wref_linear_simd(x, S, d, wk_t, bk, d, k);
  wref_linear_simd(x, S, d, wv_t, bv, d, v);
  head_job job = {q, k, v, concat, S, d, dh, h0, h1};
  int nt = threads > 1 ? threads : 1;
  if (nt > h1 - h0) nt = h1 - h0 > 0 ? h1 - h0 : 1;
  if (nt <= 1) {
    head_worker(&job);
  } else {                                                  /* parallel_map over heads (src/parallel.rs:82-98) */
    pthread_t* th = (pthread_t*)malloc((size_t)nt * sizeof(pthread_t));
    for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, head_worker, &job);
    for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
    free(th);
  }
  wref_linear_simd(concat, S, d, wo_t, bo, d, out);
  free(q); free(k); free(v); free(concat);
}

/* FeedForward::forward on rows [0, rows) (scalar path, as the reference's encoder) */
void wref_ffn(const float* x, long rows, int d, const float* w1, const float* b1, const float* w2, const float* b2, float* out) {
  float* hid = (float*)malloc((size_t)rows * 4 * d * sizeof(float));
  wref_linear_scalar(x, rows, d, w1, b1, 4 * d, hid);
  wref_gelu(hid, rows * 4 * d);
  wref_linear_scalar(hid, rows, 4 * d, w2, b2, d, out);
  free(hid);
}

/* EncoderBlock::forward, in place on x [S][d].  Attention weights transposed, FFN weights as stored. */
void wref_encoder_layer(float* x, long S, int d, int n_heads, const float* ln1g, const float* ln1b, const float* wq_t, const float* bq,
                        const float* wk_t, const float* bk, const float* wv_t, const float* bv, const float* wo_t, const float* bo,
                        const float* ln2g, const float* ln2b, const float* w1, const float* b1, const float* w2, const float* b2,
                        int threads) {
  float* n = (float*)malloc((size_t)S * d * sizeof(float));
  float* t = (float*)malloc((size_t)S * d * sizeof(float));
  wref_layernorm(x, S, d, ln1g, ln1b, n);
  wref_mha(n, S, d, n_heads, wq_t, bq, wk_t, bk, wv_t, bv, wo_t, bo, threads, 0, n_heads, t);
  for (long i = 0; i < S * d; ++i) x[i] += t[i];
  wref_layernorm(x, S, d, ln2g, ln2b, n);
  wref_ffn(n, S, d, w1, b1, w2, b2, t);
  for (long i = 0; i < S * d; ++i) x[i] += t[i];
  free(n); free(t);
}

/* ConvFrontend::forward + positional embedding; mel [T][m] -> x [S][d]; returns S */
long wref_conv_stem(const float* mel, long T, int m, int d, const float* w1, const float* b1, const float* w2, const float* b2,
                    const float* pe, float* x) {
  float* c1 = (float*)malloc((size_t)T * d * sizeof(float));
  long t1 = wref_conv1d(mel, T, m, w1, b1, d, 1, 1, c1);
  wref_gelu(c1, t1 * d);
  long S = wref_conv1d(c1, t1, d, w2, b2, d, 2, 1, x);
  wref_gelu(x, S * d);
  for (long i = 0; i < S * d; ++i) x[i] += pe[i];
  free(c1);
  return S;
}

int wref_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);     /* rayon's default pool size (src/parallel.rs:34-60) */
  return n > 0 ? (int)n : 1;
}
