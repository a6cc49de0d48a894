// Flash-style attention forward for the two tensor-core attention sites of the path:
//   * ViT:     HF CLIPAttention (non-causal, 16 heads x 64, n = 577)            -- clip_encoder.py:50
//   * prefill: HF MistralAttention (causal, GQA 32 q / 8 kv heads x 128)        -- trace_mistral.py:178
// softmax(Q K^T * scale) V with an online (running max / sum) softmax in fp32, P rounded to fp16 before
// the PV product as in the reference's eager path (softmax fp32 -> .to(fp16) -> matmul).
//
// Round-1 implementation: mma.sync m16n8k16 (legacy tensor path), cp.async double-buffered K/V tiles,
// XOR-swizzled shared memory + ldmatrix.  64 query rows per CTA (4 warps x 16 rows), 64 keys per tile.
// The tcgen05/TMEM version (S and O accumulators in TMEM) is the planned round-2 replacement; attention is
// 8.6 % of the ViT flops and 3.6 % of the prefill flops.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int BQ = 64;
constexpr int BKV = 64;
constexpr int ATT_THREADS = 128;

template <int HD>
__device__ __forceinline__ uint32_t sw_addr(uint32_t base, int row, int chunk) {
  // row-major [rows][HD] halves, 16-byte chunks XOR-swizzled with the low 3 row bits
  return base + static_cast<uint32_t>(row * (HD * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int HD>
__device__ __forceinline__ void load_tile(uint32_t sbase, const __half* g, int64_t row_stride, int row0, int nrows_valid) {
  constexpr int CH = HD / 8;
  constexpr int TOTAL = 64 * CH;
#pragma unroll
  for (int i = 0; i < TOTAL / ATT_THREADS; ++i) {
    const int idx = threadIdx.x + i * ATT_THREADS;
    const int r = idx / CH, c = idx % CH;
    const bool ok = (row0 + r) < nrows_valid;
    const __half* src = g + static_cast<int64_t>(ok ? (row0 + r) : 0) * row_stride + c * 8;
    uint32_t dst = sw_addr<HD>(sbase, r, c);
    uint32_t sz = ok ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
  }
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const __half* __restrict__ Q, const __half* __restrict__ K, const __half* __restrict__ V,
                __half* __restrict__ O, int seqlen, int q_stride, int kv_stride, int o_stride, int64_t q_bstride,
                int64_t kv_bstride, int64_t o_bstride, int group, float scale_log2) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int TILE_BYTES = 64 * HD * 2;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + TILE_BYTES;           // 2 stages
  const uint32_t sV = sK + 2 * TILE_BYTES;       // 2 stages

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = qt * BQ;

  const __half* Qp = Q + b * q_bstride + static_cast<int64_t>(h) * HD;
  const __half* Kp = K + b * kv_bstride + static_cast<int64_t>(h / group) * HD;
  const __half* Vp = V + b * kv_bstride + static_cast<int64_t>(h / group) * HD;

  int n_tiles = (seqlen + BKV - 1) / BKV;
  if (CAUSAL) {
    const int last_q = min(q0 + BQ, seqlen) - 1;
    n_tiles = min(n_tiles, last_q / BKV + 1);
  }

  load_tile<HD>(sQ, Qp, q_stride, q0, seqlen);
  load_tile<HD>(sK, Kp, kv_stride, 0, seqlen);
  load_tile<HD>(sV, Vp, kv_stride, 0, seqlen);
  cp_async_commit();

  float o_acc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[HD / 16][4];

  for (int kt = 0; kt < n_tiles; ++kt) {
    const int st = kt & 1;
    if (kt + 1 < n_tiles) {
      load_tile<HD>(sK + (st ^ 1) * TILE_BYTES, Kp, kv_stride, (kt + 1) * BKV, seqlen);
      load_tile<HD>(sV + (st ^ 1) * TILE_BYTES, Vp, kv_stride, (kt + 1) * BKV, seqlen);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        ldmatrix_x4(qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], sw_addr<HD>(sQ, warp * 16 + (lane & 15), ks * 2 + (lane >> 4)));
    }
    // ---- S = Q K^T ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    const uint32_t kb = sK + st * TILE_BYTES;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {          // pairs of 8-key n-tiles
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(b0, b1, b2, b3, sw_addr<HD>(kb, np * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)));
        mma_16816(s[np * 2], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_16816(s[np * 2 + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    // ---- mask + online softmax ----
    const int kbase = kt * BKV;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const bool need_mask = (kbase + BKV > seqlen) || (CAUSAL && (kbase + BKV - 1 > q0 + warp * 16));
    if (need_mask) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = kbase + i * 8 + 2 * t;
        if (c >= seqlen || (CAUSAL && c > r0)) s[i][0] = -INFINITY;
        if (c + 1 >= seqlen || (CAUSAL && c + 1 > r0)) s[i][1] = -INFINITY;
        if (c >= seqlen || (CAUSAL && c > r1)) s[i][2] = -INFINITY;
        if (c + 1 >= seqlen || (CAUSAL && c + 1 > r1)) s[i][3] = -INFINITY;
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
    const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0 * scale_log2;   // rows with nothing visible yet
    const float ms1 = (mn1 == -INFINITY) ? 0.f : mn1 * scale_log2;
    const float corr0 = exp2f(m_run[0] * scale_log2 - ms0);          // exp2(-inf) = 0 on the first tile
    const float corr1 = exp2f(m_run[1] * scale_log2 - ms1);
    m_run[0] = mn0; m_run[1] = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(s[i][0] * scale_log2 - ms0), p1 = exp2f(s[i][1] * scale_log2 - ms0);
      const float p2 = exp2f(s[i][2] * scale_log2 - ms1), p3 = exp2f(s[i][3] * scale_log2 - ms1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      pf[i][0] = pack_half2(p0, p1);
      pf[i][1] = pack_half2(p2, p3);
    }
    l_run[0] = l_run[0] * corr0 + rs0;
    l_run[1] = l_run[1] * corr1 + rs1;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o_acc[i][0] *= corr0; o_acc[i][1] *= corr0; o_acc[i][2] *= corr1; o_acc[i][3] *= corr1; }
    // ---- O += P V ----
    const uint32_t vb = sV + st * TILE_BYTES;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {            // 16 keys per step
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {    // pairs of 8-dim n-tiles
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(b0, b1, b2, b3, sw_addr<HD>(vb, kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)));
        mma_16816(o_acc[dp * 2], pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1], b0, b1);
        mma_16816(o_acc[dp * 2 + 1], pf[kk * 2][0], pf[kk * 2][1], pf[kk * 2 + 1][0], pf[kk * 2 + 1][1], b2, b3);
      }
    }
    __syncthreads();   // all warps done with stage st before it is refilled
  }

  // ---- finalize: quad-reduce the row sums, normalise, store ----
  float l0 = l_run[0], l1 = l_run[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __half* Op = O + b * o_bstride + static_cast<int64_t>(h) * HD;
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    const int c = i * 8 + 2 * t;
    if (r0 < seqlen) *reinterpret_cast<uint32_t*>(Op + static_cast<int64_t>(r0) * o_stride + c) = pack_half2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
    if (r1 < seqlen) *reinterpret_cast<uint32_t*>(Op + static_cast<int64_t>(r1) * o_stride + c) = pack_half2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
  }
}

template <int HD, bool CAUSAL>
int launch_attn(const __half* Q, const __half* K, const __half* V, __half* O, int seqlen, int q_stride, int kv_stride,
                int o_stride, int64_t q_bstride, int64_t kv_bstride, int64_t o_bstride, int heads, int group, int batch,
                float scale, cudaStream_t stream) {
  constexpr int SMEM = 5 * 64 * HD * 2;
  static bool attr_set = false;
  if (!attr_set) {
    TB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(seqlen, BQ), heads, batch);
  attn_fwd_kernel<HD, CAUSAL><<<grid, ATT_THREADS, SMEM, stream>>>(Q, K, V, O, seqlen, q_stride, kv_stride, o_stride,
                                                                  q_bstride, kv_bstride, o_bstride, group,
                                                                  scale * 1.4426950408889634f);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace

int attention_fwd(const __half* Q, const __half* K, const __half* V, __half* O, int seqlen, int q_stride,
                  int kv_stride, int o_stride, int64_t q_bstride, int64_t kv_bstride, int64_t o_bstride, int heads,
                  int group, int batch, int head_dim, int causal, float scale, cudaStream_t stream) {
  TB_CHECK(head_dim == 64 || head_dim == 128, "attention head_dim must be 64 (ViT) or 128 (LLM)");
  TB_CHECK(q_stride % 8 == 0 && kv_stride % 8 == 0 && o_stride % 2 == 0, "attention strides must keep 16-byte alignment");
  TB_CHECK(group >= 1 && heads % group == 0, "bad GQA grouping");
  if (seqlen == 0 || batch == 0) return 0;
#define TB_ATT(HD, C) launch_attn<HD, C>(Q, K, V, O, seqlen, q_stride, kv_stride, o_stride, q_bstride, kv_bstride, o_bstride, heads, group, batch, scale, stream)
  if (head_dim == 64) return causal ? TB_ATT(64, true) : TB_ATT(64, false);
  return causal ? TB_ATT(128, true) : TB_ATT(128, false);
#undef TB_ATT
}

}  // namespace tb
