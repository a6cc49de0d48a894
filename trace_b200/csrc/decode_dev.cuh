// Device-side building blocks shared by the per-op decode kernels (decode.cu) and the fused decode-step kernel
// (decode_mega.cu): activation staging, the TMA-bulk weight stream (producer warp + mbarrier ring) and the
// mma.sync consumer.  See decode.cu for the design notes.
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

// Balanced static partition: rows are dealt to the CTAs in units of 2 (SwiGLU pairs stay together), so every
// CTA streams the same number of weight bytes (+-2 rows) whatever N is -- no wave quantisation.
__device__ __forceinline__ void cta_rows(int N, int& r0, int& r1) {
  const long long U = N / 2;
  r0 = static_cast<int>(U * blockIdx.x / gridDim.x) * 2;
  r1 = static_cast<int>(U * (blockIdx.x + 1) / gridDim.x) * 2;
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- activation staging into shared memory: xs[b][k], b < B, k in [k0, k0+KP) --------------------------------
// executed by the 256 consumer threads only.  Activations are read with ld.global.cg (other CTAs rewrite them
// between the phases of the fused decode step).
__device__ __forceinline__ void stage_x_plain(__half* xs, const __half* x, int ldx, int B, int k0, int KP) {
  const int pitch = KP + XS_PAD;
  for (int b = 0; b < B; ++b)
    for (int i = threadIdx.x; i < KP / 8; i += GV_CONS)
      *reinterpret_cast<uint4*>(xs + b * pitch + i * 8) = __ldcg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * ldx + k0 + i * 8));
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

__device__ __forceinline__ void stage_x_rmsnorm(__half* xs, const __half* x, int ldx, int B, int K, const GammaRegs& gm, float eps,
                                                float* s_red) {
  if (B == 1) stage_x_rmsnorm_nb<1>(xs, x, ldx, B, K, gm, eps, s_red);        // keeps the common case free of spills
  else if (B <= 2) stage_x_rmsnorm_nb<2>(xs, x, ldx, B, K, gm, eps, s_red);
  else if (B <= 4) stage_x_rmsnorm_nb<4>(xs, x, ldx, B, K, gm, eps, s_red);
  else stage_x_rmsnorm_nb<8>(xs, x, ldx, B, K, gm, eps, s_red);
}

// Shared-memory carve-up of a GEMV CTA.
struct GvSmem {
  __half* xs;          // [B][KP + XS_PAD]
  float* s_red;        // rmsnorm scratch
  float* s_part;       // [tiles][8 warps][16 rows][8 batch]
  uint64_t* full;      // [stages]
  uint64_t* empty;     // [stages]
  __half* ws;          // [stages][16][KS + WS_PAD]
  int ws_stride;       // halves between consecutive stages (constant for the life of the ring)
};
__device__ __forceinline__ GvSmem gv_carve(uint8_t* smem, int B, int KP, int KS, int max_tiles, int stages) {
  GvSmem m;
  size_t off = 0;
  m.ws = reinterpret_cast<__half*>(smem + off); off += static_cast<size_t>(stages) * GV_TILE * (KS + WS_PAD) * 2;
  m.xs = reinterpret_cast<__half*>(smem + off); off += static_cast<size_t>(B) * (KP + XS_PAD) * 2;
  m.s_part = reinterpret_cast<float*>(smem + off); off += static_cast<size_t>(max_tiles) * GV_WARPS * GV_TILE * 8 * 4;
  m.s_red = reinterpret_cast<float*>(smem + off); off += (GV_WARPS * GV_MAXB + GV_MAXB) * 4;
  m.full = reinterpret_cast<uint64_t*>(smem + off); off += GV_MAXSTAGES * 8;
  m.empty = reinterpret_cast<uint64_t*>(smem + off);
  m.ws_stride = GV_TILE * (KS + WS_PAD);
  return m;
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
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

// ---- producer warp: streams rows [vlo,vhi) of the tiles of [r0,r1) in (phase, tile, k-stage) order ----------
// One bulk copy per valid row per stage (lanes 0..15 issue them in one instruction); the copies do not depend on
// the previous kernel / phase, so this warp never waits for activations and runs up to `stages` stages ahead in
// shared memory plus `pf_ahead` further stages ahead in L2 (cp.async.bulk.prefetch.L2): while the consumers sit in a
// barrier the HBM pipe keeps draining the next rows into L2, and the ring refills from L2 when they resume.
__device__ __forceinline__ void gemv_producer(const __half* __restrict__ W, int K, int KP, int KS, int r0, int r1, int vlo,
                                              int vhi, const GvSmem& m, int stages, int& s, uint32_t& ph, int pf_ahead = 0,
                                              int max_inflight = 0) {
  const int lane = threadIdx.x & 31;
  const int pitch = KS + WS_PAD;
  GvItems it, pf;
  it.init(K, KP, KS, r0, r1, vlo, vhi);
  pf = it;
  int lead = 0;                                  // how many items the prefetch iterator is ahead of the copy iterator
  while (!it.done) {
    while (lead < stages + pf_ahead && !pf.done) {
      if (lead >= stages) {                      // the first `stages` items go straight into the ring
        const int row = r0 + pf.tile * GV_TILE + lane;
        if (lane < GV_TILE && row >= vlo && row < vhi && row < r0 + (pf.tile + 1) * GV_TILE)
          bulk_prefetch_l2(W + static_cast<size_t>(row) * K + pf.k0 + pf.ks, KS * 2);
      }
      pf.next();
      ++lead;
    }
    const int tr0 = r0 + it.tile * GV_TILE;
    const int lo = max(tr0, vlo), hi = min(tr0 + GV_TILE, vhi);
    mbar_wait(&m.empty[s], ph ^ 1);
    if (max_inflight > 0) {
      // request throttle: at most `max_inflight` stages of this SM queued in the memory system.  Every byte beyond the
      // bandwidth-delay product only lengthens the queue that the latency-critical traffic of the step (barrier words,
      // activation vectors, K/V rows) has to wait in.
      const int sd = s >= max_inflight ? s - max_inflight : s - max_inflight + stages;
      mbar_wait(&m.full[sd], s >= max_inflight ? ph : ph ^ 1);
    }
    if (lane == 0) mbar_expect_tx(&m.full[s], static_cast<uint32_t>(hi - lo) * KS * 2);
    __syncwarp();
    const int row = tr0 + lane;
    if (lane < GV_TILE && row >= lo && row < hi)
      bulk_g2s(m.ws + static_cast<size_t>(s) * m.ws_stride + static_cast<size_t>(lane) * pitch,
               W + static_cast<size_t>(row) * K + it.k0 + it.ks, KS * 2, &m.full[s]);
    if (++s == stages) { s = 0; ph ^= 1; }
    it.next();
    --lead;
  }
}

// ---- consumer warps: MMA over the staged tiles; partial sums per (tile, warp) in shared memory ---------------
__device__ __forceinline__ void gemv_consume_phase(int k0, int KP, int KS, int r0, int r1, int vlo, int vhi, const GvSmem& m,
                                                   int B, int stages, int& s, uint32_t& ph) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntile = (r1 - r0 + GV_TILE - 1) / GV_TILE;
  const int ksw = KS / GV_WARPS;                  // columns of a stage owned by this warp (multiple of 32)
  const int nch = ksw / 32;
  const int wpitch = KS + WS_PAD;
  const bool xok = g < B;
  for (int tile = 0; tile < ntile; ++tile) {
    const int tr0 = r0 + tile * GV_TILE;
    float* p = m.s_part + (static_cast<size_t>(tile) * GV_WARPS + warp) * (GV_TILE * 8);
    if (max(tr0, vlo) >= min(tr0 + GV_TILE, vhi)) {
      if (k0 == 0) { p[g * 8 + 2 * t] = 0.f; p[g * 8 + 2 * t + 1] = 0.f; p[(g + 8) * 8 + 2 * t] = 0.f; p[(g + 8) * 8 + 2 * t + 1] = 0.f; }
      continue;
    }
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ks = 0; ks < KP; ks += KS) {
      mbar_wait(&m.full[s], ph);
      const __half* wa = m.ws + static_cast<size_t>(s) * m.ws_stride + static_cast<size_t>(g) * wpitch + warp * ksw + t * 8;
      const __half* wb = wa + 8 * wpitch;
      const __half* xp = m.xs + g * (KP + XS_PAD) + ks + warp * ksw + t * 8;
#pragma unroll 4
      for (int ch = 0; ch < nch; ++ch) {
        const uint4 a = *reinterpret_cast<const uint4*>(wa + ch * 32);
        const uint4 b = *reinterpret_cast<const uint4*>(wb + ch * 32);
        const uint4 xv = xok ? *reinterpret_cast<const uint4*>(xp + ch * 32) : make_uint4(0, 0, 0, 0);
        mma_16816(c, a.x, b.x, a.y, b.y, xv.x, xv.y);
        mma_16816(c, a.z, b.z, a.w, b.w, xv.z, xv.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&m.empty[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    if (k0 > 0) {
      p[g * 8 + 2 * t] += c[0]; p[g * 8 + 2 * t + 1] += c[1];
      p[(g + 8) * 8 + 2 * t] += c[2]; p[(g + 8) * 8 + 2 * t + 1] += c[3];
    } else {
      p[g * 8 + 2 * t] = c[0]; p[g * 8 + 2 * t + 1] = c[1];
      p[(g + 8) * 8 + 2 * t] = c[2]; p[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
  }
}

// sum of the 8 warps' partials for output `idx` = tile*128 + row*8 + n
__device__ __forceinline__ float gemv_out(const float* s_part, int idx) {
  const int tile = idx >> 7, rn = idx & 127;
  const float* p = s_part + tile * (GV_WARPS * GV_TILE * 8) + rn;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < GV_WARPS; ++w) v += p[w * (GV_TILE * 8)];
  return v;
}

__device__ __forceinline__ void gv_init_barriers(const GvSmem& m, int stages) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&m.full[i], 1);
      mbar_init(&m.empty[i], GV_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
}

__device__ __forceinline__ void head_range(int h, int vocab, int& lo, int& hi) {
  if (h <= 0) { lo = 0; hi = vocab + 1; }                 // text head (0) or text-only mode (-1)
  else if (h == 1) { lo = vocab + 1; hi = vocab + 14; }   // time head
  else { lo = vocab + 14; hi = vocab + 27; }              // score head
}


}  // namespace
}  // namespace tb
