// Microbenchmark: issue rate of packed fp32 arithmetic (FADD2 / FFMA2, Blackwell's f32x2 forms) against the scalar forms, 1..8 warps per
// SMSP, 16 independent accumulators per thread.  Question for the mel kernel's butterflies: does one FFMA2 cost one issue slot AND run at
// the scalar instruction rate (2x the flops per slot), or is it a half-rate instruction (same flops, fewer slots)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float2 v[16];
  for (int i = 0; i < 16; ++i) v[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 a = make_float2(0.999f, 1.001f), b = make_float2(1e-3f, -1e-3f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { v[i].x = fmaf(v[i].x, a.x, b.x); v[i].y = fmaf(v[i].y, a.y, b.y); }       // 2 FFMA
      else if (MODE == 1) v[i] = __ffma2_rn(v[i], a, b);                                            // 1 FFMA2
      else if (MODE == 2) { v[i].x += b.x; v[i].y += b.y; }                                         // 2 FADD
      else if (MODE == 3) v[i] = __fadd2_rn(v[i], b);                                               // 1 FADD2
      else { v[i] = __ffma2_rn(v[i], a, b); v[(i + 8) & 15].x = fmaf(v[(i + 8) & 15].x, a.x, b.y); }   // FFMA2 + FFMA mixed
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, double flops_per_iter, double instr_per_iter, float* d, long long* dc) {
  for (int threads : {128, 256, 512, 1024}) {
    int iters = 4096;
    k<MODE><<<148, threads>>>(d, 16, dc);
    k<MODE><<<148, threads>>>(d, iters, dc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%-14s warps/SMSP %d: %.1f lane-results/clk/SM, %.2f warp-instructions/clk/SMSP\n", name, threads / 128,
           flops_per_iter * threads * iters / c, instr_per_iter * (threads / 32) * iters / c / 4);
  }
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  long long* dc; cudaMalloc(&dc, 8);
  run<0>("2x FFMA", 32, 32, d, dc);
  run<1>("FFMA2", 32, 16, d, dc);
  run<2>("2x FADD", 32, 32, d, dc);
  run<3>("FADD2", 32, 16, d, dc);
  run<4>("FFMA2 + FFMA", 48, 32, d, dc);
  return 0;
}
