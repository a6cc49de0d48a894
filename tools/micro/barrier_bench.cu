// Micro-benchmark (round-2 preparation): latency of a grid-wide barrier across 148 persistent CTAs, the operation the
// fused decode step performs 161 times per token (17-19 us of every 102 us layer are spent in it).
//   v0  one counter: red.release.gpu.add + ld.acquire.gpu spin by thread 0 (what decode_mega.cu does)
//   v1  per-CTA flags: st.release of an epoch into flag[cta]; warp 0 polls all flags (5 per lane) -- no atomics, no hot word
//   v2  cluster hierarchy: barrier.cluster (hardware) inside clusters of 4, v0 between the 37 cluster leaders
//   v3  one counter, atom.add.acq_rel; the last arriver publishes the epoch in a separate flag word everyone else polls
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/barrier_bench tools/micro/barrier_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int V>
__global__ void __launch_bounds__(256, 1) barrier_kernel(unsigned* words, int iters, long long* cycles) {
  unsigned* counter = words;          // v0 / v2 / v3
  unsigned* epoch_w = words + 32;     // v3
  unsigned* flags = words + 64;       // v1: [gridDim.x]
  const int ncta = gridDim.x;
  unsigned target = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    __syncthreads();
    if (V == 0) {
      if (threadIdx.x == 0) {
        target += ncta;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        while (ld_acquire(counter) < target) { }
      }
    } else if (V == 1) {
      if (threadIdx.x == 0) st_release(&flags[blockIdx.x], static_cast<unsigned>(it));
      if (threadIdx.x < 32) {
        bool done;
        do {
          done = true;
          for (int i = threadIdx.x; i < ncta; i += 32) done &= ld_acquire(&flags[i]) >= static_cast<unsigned>(it);
          done = __all_sync(0xffffffffu, done);
        } while (!done);
      }
    } else if (V == 2) {
      cg::cluster_group cl = cg::this_cluster();
      cl.sync();                                         // everybody in the cluster has arrived
      if (cl.block_rank() == 0 && threadIdx.x == 0) {
        target += ncta / cl.num_blocks();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        while (ld_acquire(counter) < target) { }
      }
      cl.sync();                                         // leader releases its cluster
    } else {
      if (threadIdx.x == 0) {
        unsigned old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
        if (old == static_cast<unsigned>(it) * ncta - 1) st_release(epoch_w, static_cast<unsigned>(it));
        else while (ld_acquire(epoch_w) < static_cast<unsigned>(it)) { }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int V>
void run(const char* name, unsigned* words, long long* d_c, int iters, int cluster) {
  cudaMemset(words, 0, 4096);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(256);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = cluster > 1 ? 1 : 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t e = cudaLaunchKernelEx(&cfg, barrier_kernel<V>, words, iters, d_c);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-64s %7.2f us per barrier  (%s / %s)\n", name, ms * 1e3 / iters, cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  unsigned* words; long long* d_c;
  cudaMalloc(&words, 4096); cudaMalloc(&d_c, 148 * 8);
  const int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    run<0>("v0 one counter, red.release + ld.acquire spin", words, d_c, iters, 1);
    run<1>("v1 per-CTA flags, warp poll", words, d_c, iters, 1);
    run<2>("v2 clusters of 4 (barrier.cluster) + leaders on one counter", words, d_c, iters, 4);
    run<2>("v2 clusters of 2", words, d_c, iters, 2);
    run<3>("v3 atom.acq_rel, last arriver publishes an epoch word", words, d_c, iters, 1);
  }
  return 0;
}
