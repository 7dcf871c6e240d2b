// 400-point real FFT building blocks for the fused mel kernel.
//
// The reference calls rustfft's planner for a 400-point complex FFT of the
// windowed frame (src/audio/mel.rs:256-257,279) and keeps bins 0..200.  Here the
// real frame y[0..400) is packed as z[n] = y[2n] + i*y[2n+1] (n < 200), a
// 200-point complex DFT is taken with the factorisation 200 = 8 x 25
// (25 = 5 x 5, radix-5 butterflies in registers), and the real spectrum is
// recovered with the usual even/odd split.  Every function is host+device so the
// index algebra is unit-tested on the CPU (tests/test_abi.py).
//
// Device code path: every complex operation is ONE packed fp32x2 instruction (Blackwell's FADD2 / FMUL2 / FFMA2 on a 64-bit register
// pair; ptxas folds the component swaps, per-half negations and scalar broadcasts of -i, conj and real scales into operand
// modifiers).  The packed forms run at half the scalar instruction rate (profiles/r02m_f32x2.txt: same flops per clock), so the FP
// pipe time is unchanged -- what they halve is the ISSUE SLOTS, and the mel kernel is issue-bound (profiles/r02o: issue-active 68 %,
// FP pipe 41 %).  The rounding of every result is the scalar one (FFMA2 is two independent fused multiply-adds).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WB_HD __host__ __device__ __forceinline__
#else
#define WB_HD inline
#endif

namespace wb {

struct __attribute__((aligned(8))) cf { float x, y; };

WB_HD cf cmake(float a, float b) { cf r; r.x = a; r.y = b; return r; }

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float2 cf2(cf a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cf f2c(float2 a) { return cmake(a.x, a.y); }
#endif

WB_HD cf cadd(cf a, cf b) {
#if defined(__CUDA_ARCH__)
  return f2c(__fadd2_rn(cf2(a), cf2(b)));
#else
  return cmake(a.x + b.x, a.y + b.y);
#endif
}
WB_HD cf csub(cf a, cf b) {
#if defined(__CUDA_ARCH__)
  return f2c(__fadd2_rn(cf2(a), make_float2(-b.x, -b.y)));
#else
  return cmake(a.x - b.x, a.y - b.y);
#endif
}
// a * b: (a.x b.x - a.y b.y, a.x b.y + a.y b.x) as a multiply and a fused multiply-add per component
WB_HD cf cmul(cf a, cf b) {
#if defined(__CUDA_ARCH__)
  const float2 t = __fmul2_rn(make_float2(a.x, a.x), cf2(b));
  return f2c(__ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), t));
#else
  return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
// component-wise product (a.x b.x, a.y b.y): two real samples times two window values
WB_HD cf cmulc(cf a, cf b) {
#if defined(__CUDA_ARCH__)
  return f2c(__fmul2_rn(cf2(a), cf2(b)));
#else
  return cmake(a.x * b.x, a.y * b.y);
#endif
}
WB_HD cf cscale(cf a, float s) {
#if defined(__CUDA_ARCH__)
  return f2c(__fmul2_rn(cf2(a), make_float2(s, s)));
#else
  return cmake(a.x * s, a.y * s);
#endif
}
// a + s * b (s real)
WB_HD cf caxpy(float s, cf b, cf a) {
#if defined(__CUDA_ARCH__)
  return f2c(__ffma2_rn(make_float2(s, s), cf2(b), cf2(a)));
#else
  return cmake(a.x + s * b.x, a.y + s * b.y);
#endif
}
WB_HD cf cmul_negi(cf a) { return cmake(a.y, -a.x); }   // a * (-i)
WB_HD cf cconj(cf a) { return cmake(a.x, -a.y); }

// In-place forward 5-point DFT (kernel exp(-2*pi*i*n*k/5)).
WB_HD void dft5(cf& x0, cf& x1, cf& x2, cf& x3, cf& x4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  const cf t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const cf a1 = caxpy(c2, t2, caxpy(c1, t1, x0));
  const cf a2 = caxpy(c1, t2, caxpy(c2, t1, x0));
  const cf b1 = cmul_negi(caxpy(s2, t4, cscale(t3, s1)));
  const cf b2 = cmul_negi(caxpy(-s1, t4, cscale(t3, s2)));
  x0 = cadd(x0, cadd(t1, t2));
  x1 = cadd(a1, b1);
  x4 = csub(a1, b1);
  x2 = cadd(a2, b2);
  x3 = csub(a2, b2);
}

// Forward 25-point DFT, in place: v[n] (n = 5a + b) -> v[k] (natural order).
// tw25[b*5 + c] = exp(-2*pi*i*b*c/25).
WB_HD void dft25(cf (&v)[25], const cf* __restrict__ tw25) {
  // stage 1: for each b, 5-point DFT over a (elements v[5a + b]) -> T[b][c] stored at v[5c + b]
#pragma unroll
  for (int b = 0; b < 5; ++b) dft5(v[b], v[5 + b], v[10 + b], v[15 + b], v[20 + b]);
  // twiddle T[b][c] *= W25^(b*c)
#pragma unroll
  for (int b = 1; b < 5; ++b) {
#pragma unroll
    for (int c = 1; c < 5; ++c) v[5 * c + b] = cmul(v[5 * c + b], tw25[b * 5 + c]);
  }
  // stage 2: for each c, 5-point DFT over b (elements v[5c + b]) -> Y[c + 5e] stored at v[5c + e]
#pragma unroll
  for (int c = 0; c < 5; ++c) dft5(v[5 * c], v[5 * c + 1], v[5 * c + 2], v[5 * c + 3], v[5 * c + 4]);
  // now Y[c + 5e] sits at v[5c + e]: transpose to natural order
  cf t[25];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
#pragma unroll
    for (int e = 0; e < 5; ++e) t[c + 5 * e] = v[5 * c + e];
  }
#pragma unroll
  for (int i = 0; i < 25; ++i) v[i] = t[i];
}

// In-place forward 8-point DFT, natural order in and out.
WB_HD void dft8(cf (&v)[8]) {
  const float h = 0.70710678118654752f;
  // radix-2 decimation in time on even/odd
  const cf e0 = cadd(v[0], v[4]), e1 = csub(v[0], v[4]);
  const cf e2 = cadd(v[2], v[6]), e3 = cmul_negi(csub(v[2], v[6]));
  const cf E0 = cadd(e0, e2), E2 = csub(e0, e2), E1 = cadd(e1, e3), E3 = csub(e1, e3);   // DFT4 of (v0,v2,v4,v6)
  const cf o0 = cadd(v[1], v[5]), o1 = csub(v[1], v[5]);
  const cf o2 = cadd(v[3], v[7]), o3 = cmul_negi(csub(v[3], v[7]));
  const cf O0 = cadd(o0, o2), O2 = csub(o0, o2), O1 = cadd(o1, o3), O3 = csub(o1, o3);   // DFT4 of (v1,v3,v5,v7)
  // twiddles W8^k: 1, (1-i)h, -i, (-1-i)h:   O1 (1 - i) h = (O1 + (-i) O1) h,   O3 (-1 - i) h = ((-i) O3 - O3) h
  const cf T1 = cscale(cadd(O1, cmul_negi(O1)), h);
  const cf T2 = cmul_negi(O2);
  const cf T3 = cscale(csub(cmul_negi(O3), O3), h);
  v[0] = cadd(E0, O0); v[4] = csub(E0, O0);
  v[1] = cadd(E1, T1); v[5] = csub(E1, T1);
  v[2] = cadd(E2, T2); v[6] = csub(E2, T2);
  v[3] = cadd(E3, T3); v[7] = csub(E3, T3);
}

// Power of real-spectrum bin k (0 < k < 200) from the packed transform:
// zk = Z[k], zm = Z[200-k], w = exp(-2*pi*i*k/400).
WB_HD float rfft_power(cf zk, cf zm, cf w) {
  cf e = cmake(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));          // (Z[k] + conj Z[N-k]) / 2
  cf d = cmake(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));          // (Z[k] - conj Z[N-k]) / 2
  cf o = cmul(w, cmul_negi(d));                                      // w * (-i) * d
  float re = e.x + o.x, im = e.y + o.y;
  return re * re + im * im;
}

// Powers of the partner bins k and 200-k (0 < k < 200, k != 100) in one go.  zk = Z[k] / 2, zm = Z[200-k] / 2 (the packed transform
// already halved: the 1/2 of the even/odd split is folded into the W200 twiddle table), w = exp(-2*pi*i*k/400):
//   E = zk + conj zm,  T = w * (-i) * (zk - conj zm);   X[k] = E + T,   X[200-k] = conj(E - T)
// -- half the arithmetic of two rfft_power calls, which recompute E and T for each partner.
WB_HD void rfft_power_pair(cf zk, cf zm, cf w, float& pk, float& pm) {
  const cf zc = cconj(zm);
  const cf e = cadd(zk, zc), d = csub(zk, zc);
  const cf t = cmul(cmul_negi(d), w);               // (d.y w.x + d.x w.y, d.y w.y - d.x w.x)
  const cf a = cadd(e, t), b = csub(e, t);
  pk = a.x * a.x + a.y * a.y;
  pm = b.x * b.x + b.y * b.y;
}

}  // namespace wb
