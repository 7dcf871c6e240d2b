// Microbenchmark: MUFU.EX2 throughput per SM (ops/clk) with 1, 2, 4, 8 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.5f - 1.0f;
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int threads : {128, 256, 512, 1024}) {
    int iters = 4096;
    k<<<148, threads>>>(d, 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<<<148, threads>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = 148.0 * threads * iters * 8;
    printf("threads/SM %4d: %.3f ms, %.1f Gex2/s, %.2f ex2/clk/SM at %d MHz nominal (fma count equal)\n", threads, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  }
  return 0;
}
