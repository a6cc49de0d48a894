// Blackwell-native attention forward: softmax(Q K^T * scale) V with both contractions on tcgen05 tensor cores,
// operands fed by TMA; S = Q K^T, the probabilities P and the output accumulator O all live in TMEM.
//   * ViT:     HF CLIPAttention (non-causal, 16 heads x 64, n = 577)            -- clip_encoder.py:50
//   * prefill: HF MistralAttention (causal, GQA 32 q / 8 kv heads x 128)        -- trace_mistral.py:178
//
// One CTA = one 128-row query tile of one (batch, head).  Roles (192 threads):
//   warp 4 : TMA producer -- Q once, then K and V tiles of 128 keys into two 2-stage rings (boxes {64 x 128}, SWIZZLE_128B,
//            straight out of the fused qkv activation: no transposes, no per-head copies)
//   warp 5 : UMMA issuer  -- S[128x128] = Q K^T   (tcgen05.mma kind::f16, M=128 N=128, A/B K-major)        -> TMEM cols [0,128)
//                            O[128xd] += P V      (M=128 N=d K=128; P as the A operand from TMEM cols [128+d,192+d), V as an
//                                                  MN-major B operand, i.e. the [key][d] tile exactly as TMA wrote it) -> TMEM cols [128,128+d)
//            software-pipelined: S(j) is issued as soon as the softmax warps hold S(j-1) in registers, i.e. it runs
//            under the exponentials of tile j-1; O += P(j-1) V(j-1) follows when P(j-1) is in TMEM.
//   warps 0-3 : softmax (thread = query row): ONE tcgen05.ld of the 128 scores into registers -> release the S buffer ->
//            running max with LAZY rescale (O and the row sum are only rescaled when the max grows by more than 2^8, so
//            the accumulator stays in TMEM and is touched by these warps only then and in the epilogue) -> ex2 ->
//            P (fp16, as the reference rounds softmax before the PV matmul) back into TMEM (tcgen05.st: row = lane, two keys
//            per 32-bit column), from where the TS form of tcgen05.mma takes it as the A operand.
// The d=64 (ViT) instantiation uses 80 KB of smem and 256 TMEM columns, so two CTAs share an SM (the ViT tile loop is only 5
// long: the second CTA hides the first one's load / drain bubbles); d=128 (512 TMEM columns) runs one CTA per SM.
// Measured bounds and the variants that were tried: DESIGN.md section 5.
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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, HD == 64 ? 2 : 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ O, int seqlen, int o_stride, int64_t o_bstride,
               int q_col0, int k_col0, int v_col0, int group, float scale_log2) {
  constexpr int NA = HD / 64;                       // 64-column atoms per head row
  constexpr int Q_BYTES = NA * ATOM_BYTES;
  constexpr int KV_BYTES = NA * ATOM_BYTES;
  constexpr int P_COL = 128 + HD;                   // P (fp16 pairs): 64 columns
  constexpr int TMEM_COLS = HD == 64 ? 256 : 512;   // S: [0,128)  O: [128, 128+HD)  P: [128+HD, 192+HD)
  // 1024-byte alignment (SWIZZLE_128B atoms) comes from the declaration, not from a padded request: with the pad the
  // d = 64 instantiation needed 113.1 KB + 1 KB reserved per CTA and exactly missed two CTAs per SM (228 KB)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                       // 2 stages
  uint8_t* sV = sK + 2 * KV_BYTES;                  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * KV_BYTES);
  uint64_t* q_full = bars;                          // [1]
  uint64_t* k_full = bars + 1;                      // [2]
  uint64_t* k_empty = bars + 3;                     // [2]
  uint64_t* v_full = bars + 5;                      // [2]
  uint64_t* v_empty = bars + 7;                     // [2]
  uint64_t* s_full = bars + 9;                      // [1] S(j) complete in TMEM
  uint64_t* s_free = bars + 10;                     // [1] S(j) is in the softmax warps' registers
  uint64_t* p_full = bars + 11;                     // [1] P(j) in shared memory, O rescaled
  uint64_t* pv_done = bars + 12;                    // [1] O += P(j) V(j) complete
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * AT_BQ;
  const int row_base = b * seqlen;                  // first row of this batch element in the fused activation
  int n_tiles = (seqlen + AT_BKV - 1) / AT_BKV;
  if (CAUSAL) n_tiles = min(n_tiles, (min(q0 + AT_BQ, seqlen) - 1) / AT_BKV + 1);

  if (warp == 4 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
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
        const uint32_t par = ((j >> 1) & 1) ^ 1;
        mbar_wait(&k_empty[st], par);
        mbar_expect_tx(&k_full[st], KV_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_2d(sK + st * KV_BYTES + a * ATOM_BYTES, &tmap, &k_full[st], k_col0 + kvh * HD + a * 64, row_base + j * AT_BKV);
        mbar_wait(&v_empty[st], par);
        mbar_expect_tx(&v_full[st], KV_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_2d(sV + st * KV_BYTES + a * ATOM_BYTES, &tmap, &v_full[st], v_col0 + kvh * HD + a * 64, row_base + j * AT_BKV);
      }
    }
  } else if (warp == 5) {
    // ===================== UMMA issuer =====================
    constexpr uint32_t idesc_s = idesc_f16(128, AT_BKV, 0);       // S = Q K^T : both operands K-major
    constexpr uint32_t idesc_pv = idesc_f16(128, HD, 1);          // O += P V  : V is MN-major ([key][d] rows)
    mbar_wait(q_full, 0);
    for (int j = 0; j <= n_tiles; ++j) {
      if (j < n_tiles) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);                // S(j-1) has left TMEM
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint64_t ad = umma_desc_sw128(smem_u32(sQ + (kk / 4) * ATOM_BYTES)) + 2 * (kk % 4);
            const uint64_t bd = umma_desc_sw128(smem_u32(sK + st * KV_BYTES + (kk / 4) * ATOM_BYTES)) + 2 * (kk % 4);
            umma_f16(tmem_base, ad, bd, idesc_s, kk != 0 ? 1u : 0u);
          }
          umma_commit(s_full);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int i = j - 1, st = i & 1;
        mbar_wait(p_full, i & 1);                                 // P(i) is in shared memory, O is rescaled
        mbar_wait(&v_full[st], (i >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int kk = 0; kk < AT_BKV / 16; ++kk) {              // 16 keys per MMA
            const uint64_t bd = umma_desc_mn_sw128(smem_u32(sV + st * KV_BYTES) + kk * 2048, ATOM_BYTES);
            umma_f16_ts(tmem_base + 128, tmem_base + P_COL + kk * 8, bd, idesc_pv, (i > 0 || kk != 0) ? 1u : 0u);
          }
          umma_commit(pv_done);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax (thread = query row) =====================
    const int r = warp * 32 + lane;                               // TMEM lane == row of the tile
    const int qrow = q0 + r;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;                         // m_run: (possibly stale) reference maximum of the row
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * AT_BKV;
      const bool masked = (kbase + AT_BKV > seqlen) || (CAUSAL && kbase + AT_BKV - 1 > q0);
      // the whole score row of the tile in registers: ONE round of tcgen05.ld (168 registers; two CTAs of 192 threads share an SM at d = 64)
      uint32_t v[128];
      float mx;
      {
        tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_32x32(taddr + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
        tmem_ld_32x32(taddr + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);                                      // the issuer may overwrite S with tile j+1
        if (masked) {
#pragma unroll
          for (int i = 0; i < 128; ++i) {
            const int col = kbase + i;
            if (!(col < seqlen && (!CAUSAL || col <= qrow))) v[i] = __float_as_uint(-INFINITY);
          }
        }
        // eight independent chains instead of one of 128 dependent FMNMX
        float m8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m8[k] = __uint_as_float(v[k]);
#pragma unroll
        for (int i = 8; i < 128; i += 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) m8[k] = fmaxf(m8[k], __uint_as_float(v[i + k]));
        }
        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      }
      // lazy rescale: keep the old reference maximum unless the new one is more than 2^8 above it (P stays <= 256 in fp16,
      // the row sum and O stay consistent with the reference that was used for every term)
      bool grow = false;
      float corr = 1.f;
      if (j == 0) {
        m_run = mx;
      } else if ((mx - m_run) * scale_log2 > 8.f) {
        grow = true;
        corr = ex2_approx((m_run - mx) * scale_log2);
        m_run = mx;
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);                          // P buffer is free again, O holds tiles < j
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          // rare (the maximum of some row of this warp grew by more than 2^8): 16-column chunks keep the temporaries small
          // next to the 128 live scores
#pragma unroll 1
          for (int c = 0; c < HD / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(taddr + 128 + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
            tmem_st_32x16(taddr + 128 + c * 16, o);
          }
          tmem_st_wait();
          l_run *= corr;
        }
      }
      const float ms = (m_run == -INFINITY) ? 0.f : m_run * scale_log2;
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {                               // 32 keys -> 16 columns of fp16 pairs in the row's TMEM lane
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[c * 32 + 2 * q]), scale_log2, -ms));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[c * 32 + 2 * q + 1]), scale_log2, -ms));
          rs += p0 + p1;
          pk[q] = pack_half2(p0, p1);
        }
        tmem_st_32x16(taddr + P_COL + c * 16, pk);
      }
      tmem_st_wait();
      l_run += rs;
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> fp16 ----
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    __half* op = O + b * o_bstride + static_cast<int64_t>(qrow) * o_stride + h * HD;
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(taddr + 128 + c * 32, o);
      tmem_ld_wait();
      if (qrow < seqlen) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_half2(__uint_as_float(o[i * 8 + 0]) * inv, __uint_as_float(o[i * 8 + 1]) * inv);
          w.y = pack_half2(__uint_as_float(o[i * 8 + 2]) * inv, __uint_as_float(o[i * 8 + 3]) * inv);
          w.z = pack_half2(__uint_as_float(o[i * 8 + 4]) * inv, __uint_as_float(o[i * 8 + 5]) * inv);
          w.w = pack_half2(__uint_as_float(o[i * 8 + 6]) * inv, __uint_as_float(o[i * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c * 32 + i * 8) = w;
        }
      }
    }
    tc_fence_before();
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
  constexpr int SMEM = (HD / 64) * ATOM_BYTES * 5 + 128;
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
