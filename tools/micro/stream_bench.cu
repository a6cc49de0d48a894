// Micro-benchmark for the decode weight stream: 148 persistent CTAs, one producer warp issuing cp.async.bulk into an
// mbarrier ring, one consumer warp that only frees the stages.  Compares the current row-major layout (16 row copies of
// 2 KB per 32 KB stage, rows K*2 bytes apart) with a tile-major pre-packed layout (one contiguous 32 KB copy per stage).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/stream_bench tools/micro/stream_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int STAGE_ROWS = 16, KS = 1024;                 // 16 rows x 1024 halves = 32 KB

// mode 0: row-major W[N][K]: stage = 16 copies of 2 KB;  mode 1: packed: stage = one 32 KB copy;  mode 2: packed, 2 x 16 KB
__global__ void __launch_bounds__(64, 1) stream_kernel(const char* __restrict__ W, int K, int rows_per_cta, int mode, int stages, long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 8;
  uint8_t* ws = smem + 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ntile = rows_per_cta / STAGE_ROWS, nks = K / KS;
  const int items = ntile * nks;
  const size_t stage_bytes = static_cast<size_t>(STAGE_ROWS) * KS * 2;
  const long long t0 = clock64();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int s = 0; uint32_t ph = 0;
    const char* base = W + static_cast<size_t>(blockIdx.x) * rows_per_cta * K * 2;
    for (int it = 0; it < items; ++it) {
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) mbar_expect_tx(&full[s], static_cast<uint32_t>(stage_bytes));
      __syncwarp();
      uint8_t* dst = ws + static_cast<size_t>(s) * stage_bytes;
      if (mode == 0) {
        const int tile = it / nks, ks = it % nks;
        if (lane < STAGE_ROWS)
          bulk_g2s(dst + lane * KS * 2, base + (static_cast<size_t>(tile * STAGE_ROWS + lane) * K + ks * KS) * 2, KS * 2, &full[s]);
      } else if (mode == 1) {
        if (lane == 0) bulk_g2s(dst, base + static_cast<size_t>(it) * stage_bytes, static_cast<uint32_t>(stage_bytes), &full[s]);
      } else {
        if (lane < 2) bulk_g2s(dst + lane * (stage_bytes / 2), base + static_cast<size_t>(it) * stage_bytes + lane * (stage_bytes / 2), static_cast<uint32_t>(stage_bytes / 2), &full[s]);
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < items; ++it) {
      mbar_wait(&full[s], ph);
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const int K = 4096, rows_per_cta = 192, ncta = 148;      // 148 x 192 rows x 8 KB = 233 MB (the gate/up matrix of one layer)
  const size_t bytes = static_cast<size_t>(ncta) * rows_per_cta * K * 2;
  char* W; cudaMalloc(&W, bytes * 4);                      // 4 distinct copies: no L2 reuse between repetitions
  cudaMemset(W, 1, bytes * 4);
  long long* d_c; cudaMalloc(&d_c, ncta * 8);
  const char* names[3] = {"row-major, 16 x 2 KB copies per stage", "packed, 1 x 32 KB copy per stage", "packed, 2 x 16 KB copies per stage"};
  for (int stages = 3; stages <= 6; ++stages)
    for (int mode = 0; mode < 3; ++mode) {
      const size_t smem = 128 + static_cast<size_t>(stages) * STAGE_ROWS * KS * 2;
      cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      float best = 1e9f;
      for (int rep = 0; rep < 8; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        stream_kernel<<<ncta, 64, smem>>>(W + (rep % 4) * bytes, K, rows_per_cta, mode, stages, d_c);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
      }
      printf("stages=%d  %-40s %7.1f us  %6.0f GB/s  (%s)\n", stages, names[mode], best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
