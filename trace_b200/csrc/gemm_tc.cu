// Tensor-core GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] · W[N,K]^T )
//
// A (activations) and W (nn.Linear weight, [out,in]) are both K-major fp16, exactly the layouts the
// reference's F.linear call sites use (HF CLIP q/k/v/out/fc1/fc2, HF Mistral q/k/v/o/gate/up/down,
// SpatialSlotPool.readout multimodal_projector/builder.py:467), so no transposes are ever materialised.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D boxes {64 x 128} (A) and {64 x BN} (W), SWIZZLE_128B,
//              STAGES-deep ring guarded by full/empty mbarriers
//   warp 1   : UMMA issuer   -- one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16)
//              from shared-memory descriptors; accumulators live in TMEM (2 stages x BN fp32 columns)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      -- tcgen05.ld 32x32b.x32 -> registers -> bias / quick_gelu / residual / SwiGLU ->
//              fp16 -> 16-byte global stores; overlaps with the next tile's main loop through the
//              tmem_full / tmem_empty barrier pair.
// Roofline: tensor pipe. 2*M*N*K flop per launch.
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace tb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 halves = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;   // warps 0-3: TMA / UMMA / TMEM alloc / spare ; warps 4-11: epilogue (2 per TMEM lane quarter)
constexpr int EPI_WARPS = 8;

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;       // two accumulator stages
  static constexpr int EPI_BYTES = 8 * 2048;   // per-epilogue-warp transpose buffers
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
};

struct EpiParams {
  const __half* bias;      // [N] or null
  const __half* residual;  // [M, ldr] or null (added after bias/activation)
  __half* out;             // [M, ldo]
  int ldr, ldo;
  int act;                 // 0 none, 1 quick_gelu
  int swiglu;              // 1: columns are (gate,up) pairs -> N/2 outputs silu(gate)*up
};

// x * sigmoid(a x) with ex2.approx + rcp.approx (relative error ~1e-7, far below the fp16 rounding of the output); the
// IEEE division costs ~8 extra issue slots per element and the K=1024 ViT GEMMs are epilogue-issue bound
__device__ __forceinline__ float quick_gelu(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// bias / activation / SwiGLU + fp16 rounding + per-warp transpose + coalesced (residual-adding) stores of one 32-column chunk
__device__ __forceinline__ void epilogue_chunk(const EpiParams& ep, float (&v)[32], uint8_t* wbuf, int lane, int row_base, int col0, int M, int N) {
        // residual rows in the (post-transpose) store layout: requested first so their latency hides behind the math below
        uint4 rv[4];
        if (ep.residual != nullptr && !ep.swiglu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int grow = row_base + i * 8 + (lane >> 2), gcol = col0 + (lane & 3) * 8;
            rv[i] = (grow < M && gcol < N) ? *reinterpret_cast<const uint4*>(ep.residual + static_cast<size_t>(grow) * ep.ldr + gcol)
                                           : make_uint4(0, 0, 0, 0);
          }
        }
        if (ep.bias != nullptr) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (col0 + j8 * 8 < N) {
              uint4 b = *reinterpret_cast<const uint4*>(ep.bias + col0 + j8 * 8);
              float2 f0 = unpack_half2(b.x), f1 = unpack_half2(b.y), f2 = unpack_half2(b.z), f3 = unpack_half2(b.w);
              v[j8 * 8 + 0] += f0.x; v[j8 * 8 + 1] += f0.y; v[j8 * 8 + 2] += f1.x; v[j8 * 8 + 3] += f1.y;
              v[j8 * 8 + 4] += f2.x; v[j8 * 8 + 5] += f2.y; v[j8 * 8 + 6] += f3.x; v[j8 * 8 + 7] += f3.y;
            }
          }
        }
        if (ep.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
        }
        // Transpose through a per-warp 2 KB staging buffer so that every global store instruction writes whole 64-byte
        // row segments (8 rows x 64 B) instead of 32 scattered 16-byte pieces of 32 different rows; the fp16 rounding
        // of the linear output happens here, the residual is added to the rounded value (the reference's two roundings).
                if (ep.swiglu) {
          // (gate, up) column pairs -> 16 outputs (32 B per row) at column col0/2
#pragma unroll
          for (int j8 = 0; j8 < 2; ++j8) {
            uint4 o;
            o.x = pack_half2(silu(v[j8 * 16 + 0]) * v[j8 * 16 + 1], silu(v[j8 * 16 + 2]) * v[j8 * 16 + 3]);
            o.y = pack_half2(silu(v[j8 * 16 + 4]) * v[j8 * 16 + 5], silu(v[j8 * 16 + 6]) * v[j8 * 16 + 7]);
            o.z = pack_half2(silu(v[j8 * 16 + 8]) * v[j8 * 16 + 9], silu(v[j8 * 16 + 10]) * v[j8 * 16 + 11]);
            o.w = pack_half2(silu(v[j8 * 16 + 12]) * v[j8 * 16 + 13], silu(v[j8 * 16 + 14]) * v[j8 * 16 + 15]);
            *reinterpret_cast<uint4*>(wbuf + lane * 32 + j8 * 16) = o;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int rr = i * 16 + (lane >> 1), j = lane & 1;
            const uint4 o = *reinterpret_cast<const uint4*>(wbuf + rr * 32 + j * 16);
            const int grow = row_base + rr;
            if (grow < M && col0 + j * 16 < N)
              *reinterpret_cast<uint4*>(ep.out + static_cast<size_t>(grow) * ep.ldo + (col0 >> 1) + j * 8) = o;
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            uint4 o;
            o.x = pack_half2(v[j8 * 8 + 0], v[j8 * 8 + 1]);
            o.y = pack_half2(v[j8 * 8 + 2], v[j8 * 8 + 3]);
            o.z = pack_half2(v[j8 * 8 + 4], v[j8 * 8 + 5]);
            o.w = pack_half2(v[j8 * 8 + 6], v[j8 * 8 + 7]);
            *reinterpret_cast<uint4*>(wbuf + lane * 64 + ((j8 ^ ((lane >> 1) & 3)) << 4)) = o;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2), j = lane & 3;
            uint4 o = *reinterpret_cast<const uint4*>(wbuf + rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4));
            const int grow = row_base + rr, gcol = col0 + j * 8;
            if (grow < M && gcol < N) {
              if (ep.residual != nullptr) {
                __half2* oh = reinterpret_cast<__half2*>(&o);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) oh[k] = __hadd2(oh[k], rh[k]);
              }
              *reinterpret_cast<uint4*>(ep.out + static_cast<size_t>(grow) * ep.ldo + gcol) = o;
            }
          }
          __syncwarp();
        }
}

// Grouped rasterisation: consecutive tile ids walk all N-tiles of a band of `gm` M-tiles before moving to the next
// band, so one wave of CTAs re-uses each A tile tiles_n times and each W tile gm times out of L2 (a plain m-fastest
// order re-streams the whole A matrix from HBM once per N-tile when A > L2, e.g. the 151 MB fc2 input).
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int gm, int& m, int& n) {
  const int per_band = gm * tiles_n;
  const int band = tile / per_band;
  const int in_band = tile - band * per_band;
  const int m0 = band * gm;
  const int rows = min(gm, tiles_m - m0);
  m = m0 + in_band % rows;
  n = in_band / rows;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               int M, int N, int K, EpiParams ep) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_buf = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  int gm = static_cast<int>(gridDim.x) / tiles_n;
  gm = gm < 1 ? 1 : gm;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_ptr, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, gm, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 halves = 32 B along K inside the 128-B swizzle atom: +2 in the (addr>>4) field
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                       // TMEM lane quarter == warp_id % 4
    const int half = (warp - 4) >> 2;             // which half of the tile's columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int tm, tn;
      tile_coords(tile, tiles_m, tiles_n, gm, tm, tn);
      const int m0 = tm * BM, n0 = tn * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c == (half + 1) * (BN / 64) - 1) {    // this warp's TMEM reads of the stage are done -> release its share
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int col0 = n0 + c * 32;
        if (col0 >= N) continue;                  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        epilogue_chunk(ep, v, epi_buf + (warp - 4) * 2048, lane, m0 + q * 32, col0, M, N);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant: a cluster of two CTAs (one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages only its 128 rows of A and its 128 rows of W per k-block (32 KB instead of 48 KB), i.e. the L2->SM
// operand traffic per flop drops from (128+256) to (128+128) rows per 128x256 outputs -- the single-CTA tile is bound by
// exactly that traffic (85 flop/B) -- and the ring deepens from 4 to 6 stages.  The leader CTA (rank 0) issues the MMAs;
// both CTAs run a TMA producer (completing on the leader's barrier) and an epilogue over their own 128 TMEM lanes.
// ------------------------------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int BN2 = 256;
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = BM * BK * 2;              // 128 rows of A per CTA
  static constexpr int B_BYTES = 128 * BK * 2;             // 128 of the tile's 256 W rows per CTA
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN2;
  static constexpr int EPI_BYTES = 8 * 2048;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_BYTES;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  // executed by both CTAs; the peer bit of the barrier address is cleared so the bytes complete on the leader's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrive on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {   // arrive on `bar` of CTA `cta` of the cluster
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                int M, int N, int K, EpiParams ep) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]  (the leader's copy is the one that is used)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_buf = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM);
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  int gm = num_clusters / tiles_n;
  gm = gm < 1 ? 1 : gm;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);   // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_base_ptr, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, gm, tm, tn);
        const int m0 = tm * 2 * BM + static_cast<int>(rank) * BM;
        const int n0 = tn * BN + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);   // both CTAs' bytes land on this barrier
          tma_load_2d_2sm(sa, &tmap_a, &full_bar[stage], kb * BK, m0);
          tma_load_2d_2sm(sb, &tmap_b, &full_bar[stage], kb * BK, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== UMMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_f16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int tm, tn;
      tile_coords(tile, tiles_m, tiles_n, gm, tm, tn);
      const int m0 = tm * 2 * BM + static_cast<int>(rank) * BM, n0 = tn * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c == (half + 1) * (BN / 64) - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&tmem_empty[acc], 0);   // the leader's MMA warp waits on the leader's barrier
        }
        const int col0 = n0 + c * 32;
        if (col0 >= N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        epilogue_chunk(ep, v, epi_buf + (warp - 4) * 2048, lane, m0 + q * 32, col0, M, N);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

// 2-D K-major fp16 tensor [rows, cols] with leading dimension ld (elements); box {64, box_rows}, 128B swizzle.
// Descriptors are cached by (pointer, shape): the engine's weights and workspaces are allocated once, so after the first
// clip every GEMM launch finds both of its maps here instead of calling into the driver twice (~900 launches per clip).
struct TmapKey {
  const void* ptr;
  int rows, cols, ld, box_rows;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (static_cast<size_t>(k.rows) << 32) ^ static_cast<size_t>(k.cols) ^ (static_cast<size_t>(k.ld) << 17) ^ (static_cast<size_t>(k.box_rows) << 50);
    return h;
  }
};
int make_tmap_uncached(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows);
int make_tmap(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  const TmapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *map = it->second;
    return 0;
  }
  if (make_tmap_uncached(map, ptr, rows, cols, ld, box_rows)) return 1;
  if (cache.size() > 8192) cache.clear();       // callers with ever-changing pointers (tests) must not grow it without bound
  cache.emplace(key, *map);
  return 0;
}
int make_tmap_uncached(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  TB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  TB_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base must be 16-byte aligned");
  TB_CHECK((static_cast<int64_t>(ld) * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
  return 0;
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb_, int M, int N, int K, const EpiParams& ep, int num_sms,
           cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    TB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb_, M, N, K, ep);
  TB_LAUNCH_OK();
  return 0;
}

int launch2(const CUtensorMap& ta, const CUtensorMap& tb_, int M, int N, int K, const EpiParams& ep, int num_sms, cudaStream_t stream) {
  using Cfg = Gemm2Cfg;
  static bool attr_set = false;
  if (!attr_set) {
    TB_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = ceil_div(M, 2 * BM) * ceil_div(N, Cfg::BN2);
  const int clusters = tiles < num_sms / 2 ? tiles : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel, ta, tb_, M, N, K, ep));
  return 0;
}

bool use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TRACE_B200_GEMM_2CTA");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, __half* out, int ldo, int M, int N, int K,
             const __half* bias, const __half* residual, int ldr, int act, int swiglu, int num_sms,
             cudaStream_t stream) {
  TB_CHECK(M > 0 && N > 0 && K > 0, "empty GEMM");
  TB_CHECK(N % 8 == 0, "N must be a multiple of 8");
  TB_CHECK(!(swiglu && (N % 16 != 0)), "SwiGLU epilogue needs N % 16 == 0");
  TB_CHECK(ldo % 8 == 0 && (residual == nullptr || ldr % 8 == 0), "output / residual pitch must be 16-byte aligned");
  TB_CHECK((reinterpret_cast<uintptr_t>(out) & 15) == 0, "output must be 16-byte aligned");
  // Wide tiles when they fill the machine, narrower ones for small problems (fewer idle SMs).
  const bool wide = (static_cast<int64_t>(ceil_div(M, BM)) * ceil_div(N, 256) >= num_sms) && (N % 256 == 0 || N > 1024);
  CUtensorMap ta, tw;
  if (make_tmap(&ta, A, M, K, lda, BM)) return 1;
  EpiParams ep{bias, residual, out, ldr, ldo, act, swiglu};
  if (wide && use_2cta() && N % 256 == 0) {      // 256 x 256 tiles on CTA pairs
    if (make_tmap(&tw, W, N, K, ldw, 128)) return 1;
    return launch2(ta, tw, M, N, K, ep, num_sms, stream);
  }
  if (make_tmap(&tw, W, N, K, ldw, wide ? 256 : 128)) return 1;
  return wide ? launch<256>(ta, tw, M, N, K, ep, num_sms, stream) : launch<128>(ta, tw, M, N, K, ep, num_sms, stream);
}

}  // namespace tb
