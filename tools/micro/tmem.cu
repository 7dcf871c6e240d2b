// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, and what a concurrent LDTM stream does to the MUFU.EX2 rate of
// other warps on the same schedulers (the attention kernel's situation: one warpgroup reads S while the other runs exponentials).
//   block = 256 threads: warps 0-3 stream tcgen05.ld.32x32b.x32 (4 KB per instruction per warp), warps 4-7 stream ex2.
//   mode 0: LDTM only   mode 1: MUFU only   mode 2: both
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) k(int mode, int iters, long long* cyc, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = *reinterpret_cast<volatile uint32_t*>(&slot);
  float acc = 0.f;
  long long t0 = clock64();
  __shared__ unsigned long long never;        // an mbarrier nobody ever completes
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)) : "memory");
  __syncthreads();
  t0 = clock64();
  if (warp < 4 && mode >= 3) {
    // a single-thread role warp polling a barrier that is not ready (mode 3: plain try_wait; mode 4: with a 1 ms suspend hint)
    if ((threadIdx.x & 31) == 0) {
      for (int it = 0; it < iters * 3; ++it) {
        uint32_t ok;
        if (mode == 3)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)), "r"(0u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)), "r"(0u), "r"(20000u) : "memory");
        acc += ok;
      }
    }
  } else if (warp < 4) {
    if (mode != 1) {
      const uint32_t addr = base + ((uint32_t)(warp * 32) << 16);
      for (int it = 0; it < iters; ++it) {
        uint32_t r[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(addr + q * 32)
              : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[it & 31]);
      }
    }
  } else if (mode != 0) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 8; ++rep) {                 // 64 exponentials per 16 KB of TMEM read by the partner warp: the attention ratio is 128 per 16 KB
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] * 0.5f - 1.0f;
      }
    }
    for (int i = 0; i < 8; ++i) acc += v[i];
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
  sink[blockIdx.x * 256 + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
int main() {
  long long* dc; cudaMalloc(&dc, 64); float* sink; cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 5; ++mode) {
    k<<<148, 256>>>(mode, 16, dc, sink);
    k<<<148, 256>>>(mode, iters, dc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long c[8]; cudaMemcpy(c, dc, 64, cudaMemcpyDeviceToHost);
    const char* names[5] = {"LDTM only", "MUFU only", "LDTM + MUFU", "poll + MUFU", "poll(hint)+MUFU"};
    double ld_bytes = 4.0 * 4096 * 4 * iters;             // 4 warps x 4 x 4 KB per iteration
    double exps = 4.0 * 32 * 64 * iters;
    printf("%-12s", names[mode]);
    if (mode != 1 && mode < 3) printf("  LDTM: %.1f B/clk/SM (%lld cycles)", ld_bytes / c[0], c[0]);
    if (mode != 0) printf("  MUFU: %.2f ex2/clk/SM (%lld cycles)", exps / c[4], c[4]);
    printf("\n");
  }
  return 0;
}
