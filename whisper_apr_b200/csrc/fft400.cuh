// 400-point real FFT building blocks for the fused mel kernel.
//
// The reference calls rustfft's planner for a 400-point complex FFT of the
// windowed frame (src/audio/mel.rs:256-257,279) and keeps bins 0..200.  Here the
// real frame y[0..400) is packed as z[n] = y[2n] + i*y[2n+1] (n < 200), a
// 200-point complex DFT is taken with the factorisation 200 = 8 x 25
// (25 = 5 x 5, radix-5 butterflies in registers), and the real spectrum is
// recovered with the usual even/odd split.  Every function is host+device so the
// index algebra is unit-tested on the CPU (tests/test_fft_math.py).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WB_HD __host__ __device__ __forceinline__
#else
#define WB_HD inline
#endif

namespace wb {

struct cf { float x, y; };

WB_HD cf cmake(float a, float b) { cf r; r.x = a; r.y = b; return r; }
WB_HD cf cadd(cf a, cf b) { return cmake(a.x + b.x, a.y + b.y); }
WB_HD cf csub(cf a, cf b) { return cmake(a.x - b.x, a.y - b.y); }
WB_HD cf cmul(cf a, cf b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
WB_HD cf cscale(cf a, float s) { return cmake(a.x * s, a.y * s); }
WB_HD cf cmul_negi(cf a) { return cmake(a.y, -a.x); }   // a * (-i)
WB_HD cf cconj(cf a) { return cmake(a.x, -a.y); }

// In-place forward 5-point DFT (kernel exp(-2*pi*i*n*k/5)).
WB_HD void dft5(cf& x0, cf& x1, cf& x2, cf& x3, cf& x4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  cf t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  cf a1 = cmake(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
  cf a2 = cmake(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
  cf b1 = cmul_negi(cmake(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
  cf b2 = cmul_negi(cmake(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
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
  cf e0 = cadd(v[0], v[4]), e1 = csub(v[0], v[4]);
  cf e2 = cadd(v[2], v[6]), e3 = cmul_negi(csub(v[2], v[6]));
  cf E0 = cadd(e0, e2), E2 = csub(e0, e2), E1 = cadd(e1, e3), E3 = csub(e1, e3);   // DFT4 of (v0,v2,v4,v6)
  cf o0 = cadd(v[1], v[5]), o1 = csub(v[1], v[5]);
  cf o2 = cadd(v[3], v[7]), o3 = cmul_negi(csub(v[3], v[7]));
  cf O0 = cadd(o0, o2), O2 = csub(o0, o2), O1 = cadd(o1, o3), O3 = csub(o1, o3);   // DFT4 of (v1,v3,v5,v7)
  // twiddles W8^k: 1, (1-i)h, -i, (-1-i)h
  cf T1 = cmake((O1.x + O1.y) * h, (O1.y - O1.x) * h);
  cf T2 = cmul_negi(O2);
  cf T3 = cmake((O3.y - O3.x) * h, -(O3.x + O3.y) * h);
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
  const float ex = zk.x + zm.x, ey = zk.y - zm.y;
  const float dx = zk.x - zm.x, dy = zk.y + zm.y;
  const float tx = w.x * dy + w.y * dx, ty = w.y * dy - w.x * dx;
  const float ax = ex + tx, ay = ey + ty, bx = ex - tx, by = ey - ty;
  pk = ax * ax + ay * ay;
  pm = bx * bx + by * by;
}

}  // namespace wb
