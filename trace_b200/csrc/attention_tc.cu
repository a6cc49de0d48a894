// Blackwell-native attention forward: softmax(Q K^T * scale) V with both contractions on tcgen05 tensor cores,
// operands fed by TMA, S = Q K^T and P V accumulators in TMEM.
//   * ViT:     HF CLIPAttention (non-causal, 16 heads x 64, n = 577)            -- clip_encoder.py:50
//   * prefill: HF MistralAttention (causal, GQA 32 q / 8 kv heads x 128)        -- trace_mistral.py:178
//
// One CTA = one 128-row query tile of one (batch, head).  Roles (192 threads):
//   warp 4 : TMA producer -- Q once, then K / V tiles of 128 keys into a 2-stage ring (boxes {64 x 128}, SWIZZLE_128B,
//            straight out of the fused qkv activation: no transposes, no per-head copies)
//   warp 5 : UMMA issuer  -- S[128x128] = Q K^T   (tcgen05.mma kind::f16, M=128 N=128, A/B K-major)        -> TMEM cols [0,128)
//                            PV[128xd]  = P V     (M=128 N=d K=128; P from smem (K-major), V as an MN-major B operand,
//                                                  i.e. the [key][d] tile exactly as TMA wrote it)          -> TMEM cols [128,128+d)
//   warps 0-3 : softmax + output (thread = query row): tcgen05.ld S -> scale / mask -> running max / sum (fp32, exp2)
//            -> P (fp16, as the reference rounds softmax before the PV matmul) written to smem in the swizzled K-major
//            layout UMMA expects -> fence.proxy.async -> after the PV MMA: O = O*corr + PV in registers.
// The d=64 (ViT) instantiation uses 112 KB of smem and 256 TMEM columns, so two CTAs share an SM and one CTA's softmax
// overlaps the other's MMAs; d=128 runs one CTA per SM.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int AT_BQ = 128;
constexpr int AT_BKV = 128;
constexpr int AT_THREADS = 192;
constexpr int ATOM_BYTES = 128 * 128;   // one {64 halves x 128 rows} swizzle-128B tile

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// MN-major operand tile, 128-byte swizzle: rows (the K index: keys) are 128 B apart, 8-row groups 1024 B apart (SBO),
// 64-element column atoms `lbo_bytes` apart (LBO).  cute::UMMA::make_umma_desc<Major::MN>.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ O, int seqlen, int o_stride, int64_t o_bstride,
               int q_col0, int k_col0, int v_col0, int group, float scale_log2) {
  constexpr int NA = HD / 64;                       // 64-column atoms per head row
  constexpr int Q_BYTES = NA * ATOM_BYTES;
  constexpr int KV_BYTES = NA * ATOM_BYTES;
  constexpr int TMEM_COLS = 256;                    // S: [0,128)  PV: [128, 128+HD)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                       // 2 stages
  uint8_t* sV = sK + 2 * KV_BYTES;                  // 2 stages
  uint8_t* sP = sV + 2 * KV_BYTES;                  // 2 atoms (128 keys)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATOM_BYTES);
  uint64_t* q_full = bars;                          // [1]
  uint64_t* kv_full = bars + 1;                     // [2]
  uint64_t* kv_empty = bars + 3;                    // [2]
  uint64_t* s_full = bars + 5;                      // [1]
  uint64_t* p_full = bars + 6;                      // [1]
  uint64_t* pv_full = bars + 7;                     // [1]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * AT_BQ;
  const int row_base = b * seqlen;                  // first row of this batch element in the fused activation
  int n_tiles = (seqlen + AT_BKV - 1) / AT_BKV;
  if (CAUSAL) n_tiles = min(n_tiles, (min(q0 + AT_BQ, seqlen) - 1) / AT_BKV + 1);

  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_base_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a) tma_load_2d(sQ + a * ATOM_BYTES, &tmap, q_full, q_col0 + h * HD + a * 64, row_base + q0);
      const int kvh = h / group;
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * KV_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          tma_load_2d(sK + st * KV_BYTES + a * ATOM_BYTES, &tmap, &kv_full[st], k_col0 + kvh * HD + a * 64, row_base + j * AT_BKV);
          tma_load_2d(sV + st * KV_BYTES + a * ATOM_BYTES, &tmap, &kv_full[st], v_col0 + kvh * HD + a * 64, row_base + j * AT_BKV);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc_s = idesc_f16(128, AT_BKV, 0);       // S = Q K^T : both operands K-major
    constexpr uint32_t idesc_pv = idesc_f16(128, HD, 1);          // PV = P V  : V is MN-major ([key][d] rows)
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t ad = umma_desc_sw128(smem_u32(sQ + (kk / 4) * ATOM_BYTES)) + 2 * (kk % 4);
          const uint64_t bd = umma_desc_sw128(smem_u32(sK + st * KV_BYTES + (kk / 4) * ATOM_BYTES)) + 2 * (kk % 4);
          umma_f16(tmem_base, ad, bd, idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);                                   // P(j) is in shared memory, S(j) has been read
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < AT_BKV / 16; ++kk) {                 // 16 keys per MMA
          const uint64_t ad = umma_desc_sw128(smem_u32(sP + (kk / 4) * ATOM_BYTES)) + 2 * (kk % 4);
          const uint64_t bd = umma_desc_mn_sw128(smem_u32(sV + st * KV_BYTES) + kk * 2048, ATOM_BYTES);
          umma_f16(tmem_base + 128, ad, bd, idesc_pv, kk != 0 ? 1u : 0u);
        }
        umma_commit(pv_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + output (thread = query row) =====================
    const int r = warp * 32 + lane;                               // TMEM lane == row of the tile
    const int qrow = q0 + r;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float o_acc[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * AT_BKV;
      // ---- pass A: row maximum ----
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = kbase + c * 32 + i;
          const bool ok = col < seqlen && (!CAUSAL || col <= qrow);
          if (ok) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float ms = (m_new == -INFINITY) ? 0.f : m_new * scale_log2;
      const float corr = exp2f(m_run * scale_log2 - ms);
      // ---- pass B: P = exp2(S*scale - max) -> fp16 -> swizzled K-major smem ----
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = kbase + c * 32 + 2 * i;
          const bool ok0 = col < seqlen && (!CAUSAL || col <= qrow);
          const bool ok1 = col + 1 < seqlen && (!CAUSAL || col + 1 <= qrow);
          const float p0 = ok0 ? exp2f(__uint_as_float(v[2 * i]) * scale_log2 - ms) : 0.f;
          const float p1 = ok1 ? exp2f(__uint_as_float(v[2 * i + 1]) * scale_log2 - ms) : 0.f;
          rs += p0 + p1;
          pk[i] = pack_half2(p0, p1);
        }
        uint8_t* prow = sP + (c >> 1) * ATOM_BYTES + r * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (c & 1) * 4 + q4;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
        }
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async();                                        // generic-proxy smem writes -> visible to the UMMA
      mbar_arrive(p_full);
#pragma unroll
      for (int i = 0; i < HD; ++i) o_acc[i] *= corr;
      // ---- O += P V ----
      mbar_wait(pv_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    if (qrow < seqlen) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __half* op = O + b * o_bstride + static_cast<int64_t>(qrow) * o_stride + h * HD;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        uint4 o;
        o.x = pack_half2(o_acc[i * 8 + 0] * inv, o_acc[i * 8 + 1] * inv);
        o.y = pack_half2(o_acc[i * 8 + 2] * inv, o_acc[i * 8 + 3] * inv);
        o.z = pack_half2(o_acc[i * 8 + 4] * inv, o_acc[i * 8 + 5] * inv);
        o.w = pack_half2(o_acc[i * 8 + 6] * inv, o_acc[i * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(op + i * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

EncodeTiledFn at_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int HD, bool CAUSAL>
int launch_tc(const CUtensorMap& tm, __half* O, int seqlen, int o_stride, int64_t o_bstride, int q_col0, int k_col0, int v_col0,
              int heads, int group, int batch, float scale, cudaStream_t stream) {
  constexpr int SMEM = (HD / 64) * ATOM_BYTES * 5 + 2 * ATOM_BYTES + 1024 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    TB_CUDA(cudaFuncSetAttribute(attn_tc_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(seqlen, AT_BQ), heads, batch);
  attn_tc_kernel<HD, CAUSAL><<<grid, AT_THREADS, SMEM, stream>>>(tm, O, seqlen, o_stride, o_bstride, q_col0, k_col0, v_col0, group,
                                                                scale * 1.4426950408889634f);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace

// qkv: fused activation [batch*seqlen, width] fp16 (row pitch `width`), Q/K/V heads at column offsets q_col0/k_col0/v_col0.
int attention_tc_fwd(const __half* qkv, int width, __half* O, int seqlen, int o_stride, int64_t o_bstride, int q_col0, int k_col0,
                     int v_col0, int heads, int group, int batch, int head_dim, int causal, float scale, cudaStream_t stream) {
  TB_CHECK(head_dim == 64 || head_dim == 128, "attention head_dim must be 64 (ViT) or 128 (LLM)");
  TB_CHECK(width % 8 == 0 && o_stride % 8 == 0, "row pitches must be 16-byte aligned");
  TB_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "16-byte alignment");
  if (seqlen == 0 || batch == 0) return 0;
  EncodeTiledFn fn = at_encode_fn();
  TB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tm;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(batch) * seqlen};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(width) * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(qkv), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
#define TB_ATC(HD, C) launch_tc<HD, C>(tm, O, seqlen, o_stride, o_bstride, q_col0, k_col0, v_col0, heads, group, batch, scale, stream)
  if (head_dim == 64) return causal ? TB_ATC(64, true) : TB_ATC(64, false);
  return causal ? TB_ATC(128, true) : TB_ATC(128, false);
#undef TB_ATC
}

}  // namespace tb
