// Microbenchmark: MUFU.EX2 throughput per SM (exponentials/clk) with 1, 2, 4, 8 warps per SMSP, for the f32 form and the
// packed half-precision forms (ex2.approx.ftz.f16x2 / .bf16x2: two exponentials per issued instruction if the pipe is 2-wide).
// Clock measured with clock64 inside the kernel (the nominal clock misleads under DVFS).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
  unsigned u[8];
  for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 1e-3f + i; u[i] = 0x3c003c00u + threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.5f - 1.0f;
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] & 0x3fff3fffu) | 0x30003000u;
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] & 0x3fff3fffu) | 0x30003000u;
    } else if (MODE == 3) {          // the attention inner loop's mix: 8 exponentials + 4 bf16x2 packs (+ 8 FMAs, 4 adds)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.5f - 1.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += __uint_as_float(u[i] & 0x3f800000u);
    } else {                         // the same without the packs
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.5f - 1.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] += v[i + 4] * 0.25f;
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, float* d, long long* dc) {
  for (int threads : {128, 256, 512, 1024}) {
    int iters = 4096;
    k<MODE><<<148, threads>>>(d, 16, dc);
    k<MODE><<<148, threads>>>(d, iters, dc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double instr = (double)threads * iters * 8;                       // thread-instructions per SM
    double per = (MODE == 1 || MODE == 2) ? 2.0 : 1.0;
    printf("%-22s warps/SMSP %d: %.2f MUFU lanes/clk/SM = %.2f exponentials/clk/SM\n", name, threads / 128, instr / c, per * instr / c);
  }
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  long long* dc; cudaMalloc(&dc, 8);
  run<0>("ex2.approx.ftz.f32", d, dc);
  run<1>("ex2.approx.ftz.f16x2", d, dc);
  run<2>("ex2.approx.ftz.bf16x2", d, dc);
  run<4>("ex2 f32 + fma + add", d, dc);
  run<3>("ex2 f32 + cvt.bf16x2", d, dc);
  return 0;
}
