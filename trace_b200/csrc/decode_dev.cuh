// Device-side building blocks of the fused decode step (decode_stream.cu): constants of the CTA shape, the TMA bulk copy,
// RMSNorm-fused activation staging from a dense buffer, and the (k-phase, tile, k-stage) enumeration of a CTA's weight rows.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int GV_CONS = 256;            // 8 consumer warps (MMA + epilogue)
constexpr int GV_THREADS = GV_CONS + 32; // + 1 producer warp (TMA bulk copies)
constexpr int GV_WARPS = 8;
constexpr int GV_TILE = 16;              // rows per MMA tile
constexpr int GV_MAXB = 8;
constexpr int GV_MAXSTAGES = 8;
constexpr int XS_PAD = 32;               // halves; row pitch = KP*2 + 64 bytes -> conflict-free 16-byte fragment reads
constexpr int WS_PAD = 32;               // same trick for the weight stages

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // consumer warps only

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// RMSNorm weights of this thread's columns: a weight, so it can be fetched BEFORE the activations exist (before the
// grid barrier / griddepcontrol.wait) and its HBM latency disappears from the critical path.
struct GammaRegs { uint4 g[2]; };
__device__ __forceinline__ GammaRegs load_gamma(const __half* gamma, int K) {
  GammaRegs r;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = threadIdx.x + j * GV_CONS;
    r.g[j] = (i < K / 8) ? *reinterpret_cast<const uint4*>(gamma + i * 8) : make_uint4(0, 0, 0, 0);
  }
  return r;
}

// HF MistralRMSNorm (fp32 mean of squares over the full row, x*rstd rounded to fp16, then weight * that in fp16)
// fused into the staging: ONE read of x (kept in registers across the reduction), one barrier pair.  K <= 4096.
template <int NB>
__device__ __forceinline__ void stage_x_rmsnorm_nb(__half* xs, const __half* x, int ldx, int B, int K, const GammaRegs& gm, float eps,
                                                   float* s_red) {
  const int tid = threadIdx.x;
  const int pitch = K + XS_PAD;
  uint4 xr[NB][2];
  float ss[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    ss[b] = 0.f;
    if (b < B) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = tid + j * GV_CONS;
        xr[b][j] = (i < K / 8) ? __ldcg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * ldx + i * 8)) : make_uint4(0, 0, 0, 0);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < B) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint4 u = xr[b][j];
        float2 a = unpack_half2(u.x), c = unpack_half2(u.y), d = unpack_half2(u.z), e = unpack_half2(u.w);
        ss[b] += a.x * a.x + a.y * a.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y + e.x * e.x + e.y * e.y;
      }
      ss[b] = warp_sum(ss[b]);
    }
  }
  const int w = tid >> 5, l = tid & 31;
  if (l == 0)
#pragma unroll
    for (int b = 0; b < NB; ++b) s_red[w * GV_MAXB + b] = ss[b];
  cons_sync();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < B) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < GV_WARPS; ++i) t += s_red[i * GV_MAXB + b];     // same order in every thread: deterministic
      const float rstd = rsqrtf(t / K + eps);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = tid + j * GV_CONS;
        if (i < K / 8) {
          uint4 u = xr[b][j];
          uint32_t* up = &u.x;
          const uint32_t* gp = &gm.g[j].x;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 v = unpack_half2(up[q]);
            float2 gg = unpack_half2(gp[q]);
            up[q] = pack_half2(gg.x * __half2float(__float2half_rn(v.x * rstd)), gg.y * __half2float(__float2half_rn(v.y * rstd)));
          }
          *reinterpret_cast<uint4*>(xs + b * pitch + i * 8) = u;
        }
      }
    }
  }
}

// Enumerates the (k-phase, tile, k-stage) items of one matrix for one CTA, skipping tiles without a needed row.
struct GvItems {
  int K, KP, KS, r0, ntile, vlo, vhi;
  int k0, tile, ks;
  bool done;
  __device__ __forceinline__ void init(int K_, int KP_, int KS_, int r0_, int r1_, int vlo_, int vhi_) {
    K = K_; KP = KP_; KS = KS_; r0 = r0_; vlo = vlo_; vhi = vhi_;
    ntile = (r1_ - r0_ + GV_TILE - 1) / GV_TILE;
    k0 = 0; tile = 0; ks = 0; done = false;
    skip();
  }
  __device__ __forceinline__ bool tile_ok() const {
    const int tr0 = r0 + tile * GV_TILE;
    return max(tr0, vlo) < min(tr0 + GV_TILE, vhi);
  }
  __device__ __forceinline__ void skip() {      // move to the next tile that has a needed row
    while (!done && !tile_ok()) {
      ks = 0;
      if (++tile == ntile) { tile = 0; k0 += KP; if (k0 >= K) done = true; }
    }
  }
  __device__ __forceinline__ void next() {
    ks += KS;
    if (ks >= KP) {
      ks = 0;
      if (++tile == ntile) { tile = 0; k0 += KP; if (k0 >= K) { done = true; return; } }
      skip();
    }
  }
};

__device__ __forceinline__ void head_range(int h, int vocab, int& lo, int& hi) {
  if (h <= 0) { lo = 0; hi = vocab + 1; }                 // text head (0) or text-only mode (-1)
  else if (h == 1) { lo = vocab + 1; hi = vocab + 14; }   // time head
  else { lo = vocab + 14; hi = vocab + 27; }              // score head
}


}  // namespace
}  // namespace tb
