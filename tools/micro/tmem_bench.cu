// Micro-benchmark: tcgen05.ld / tcgen05.st throughput and latency per SM (clock64), 4 and 8 warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_bench tools/micro/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// mode 0: ld + wait per iteration (latency); 1: 4 lds then one wait (throughput); 2: st + wait; 3: 4 st then wait
__global__ void __launch_bounds__(256) bench(int mode, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
  uint32_t r[32], acc = 0;
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  for (int c = 0; c < 8; ++c) st32(base + c * 32, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      ld32(base + (it & 7) * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] ^ r[7] ^ r[13] ^ r[31];
    } else if (mode == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      ld32(base, a); ld32(base + 32, b); ld32(base + 64, c); ld32(base + 96, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += a[0] ^ a[31] ^ b[1] ^ b[30] ^ c[2] ^ c[29] ^ d[3] ^ d[28];
    } else if (mode == 2) {
      r[0] = acc + it;
      st32(base + (it & 7) * 32, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      r[0] = acc + it;
      st32(base, r); st32(base + 32, r); st32(base + 64, r); st32(base + 96, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "r"(512) : "memory");
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4096);
  const char* names[4] = {"ld x32 + wait (latency-bound)", "4 x ld x32 then wait", "st x32 + wait", "4 x st x32 then wait"};
  for (int warps = 4; warps <= 8; warps += 4)
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 2000;
      bench<<<148, warps * 32, 0>>>(mode, iters, d_out, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
      const double per_it = double(cyc) / iters;
      const double bytes = (mode & 1 ? 4.0 : 1.0) * warps * 32 * 32 * 4;   // bytes moved per iteration per SM
      printf("%d warps  %-34s %8.1f cycles/iter  %7.1f B/clk/SM  (%s)\n", warps, names[mode], per_it, bytes / per_it, cudaGetErrorString(e));
    }
  return 0;
}
