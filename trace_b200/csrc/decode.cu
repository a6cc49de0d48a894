// The per-token decode loop (trace_mistral.py:114-264 forward + :317-347 head switch + HF greedy loop +
// trace_arch.py:345-375 decode-step embedding dispatch) as HBM-streaming kernels with all generation
// state on the device (no host round trip per token; the reference does one int(tensor) sync per row per step).
//
//   gemv_f16        out[b,n] = epi( W[n,:] · pro(x[b,:]) )         -- q/k/v, o, gate/up, down projections
//   attn_decode     RoPE + KV append + split-KV attention + last-CTA combine
//   head_gemv       final RMSNorm + active-head logits (fp16-rounded, then fp32 as trace_mistral.py:191-200)
//                   + -inf mask of inactive heads (:244-252) + per-CTA argmax candidates
//   decode_finalize argmax (first max), EOS/pad bookkeeping, head switch {V:1, V+1:2, V+14:0}, next-token
//                   embedding dispatch (text / sync / time / score tables), pos++.
//
// GEMV design: HBM-bound, weights are read exactly once per step with 128-bit streaming loads
// (ld.global.nc.L1::no_allocate), 16 rows per CTA, K split over the 8 warps of the CTA.  The batch (<= 8 rows)
// rides in the N=8 slot of mma.sync.m16n8k16 so B=1..8 cost the same instruction stream; each lane feeds the
// MMA straight from its 16-byte weight load (the k-slot permutation is applied identically to the activation
// fragment, so no shuffles or shared-memory transposes are needed).  Activations are staged (and RMS-normed)
// once per CTA in shared memory.  Algorithmic bytes per launch = N*K*2 (+ B*(K+N)*2).
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = 8;
constexpr int GV_ROWS = 16;
constexpr int GV_MAXB = 8;
constexpr int XS_PAD = 32;   // halves; row pitch = KP*2 + 64 bytes -> conflict-free 16-byte fragment reads

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// ---- activation staging (+ optional RMSNorm) into shared memory: xs[b][k], b < B, k in [k0, k0+KP) ----
template <int PRO>
__device__ __forceinline__ void stage_x(__half* xs, const __half* x, int ldx, int B, int k0, int KP, int K,
                                        const __half* gamma, float eps, float* s_red) {
  const int tid = threadIdx.x;
  const int pitch = KP + XS_PAD;
  if (PRO == 1) {
    // HF MistralRMSNorm: fp32 mean of squares over the full row (KP == K)
    float ss[GV_MAXB];
#pragma unroll
    for (int b = 0; b < GV_MAXB; ++b) {
      ss[b] = 0.f;
      if (b < B) {
        for (int i = tid; i < K / 8; i += GV_THREADS) {
          uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * ldx + i * 8);
          float2 a = unpack_half2(u.x), c = unpack_half2(u.y), d = unpack_half2(u.z), e = unpack_half2(u.w);
          ss[b] += a.x * a.x + a.y * a.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y + e.x * e.x + e.y * e.y;
        }
        ss[b] = warp_sum(ss[b]);
      }
    }
    const int w = tid >> 5, l = tid & 31;
    if (l == 0)
#pragma unroll
      for (int b = 0; b < GV_MAXB; ++b) s_red[w * GV_MAXB + b] = ss[b];
    __syncthreads();
    if (tid < GV_MAXB) {
      float t = 0.f;
      for (int i = 0; i < GV_WARPS; ++i) t += s_red[i * GV_MAXB + tid];
      s_red[GV_WARPS * GV_MAXB + tid] = rsqrtf(t / K + eps);
    }
    __syncthreads();
  }
  for (int b = 0; b < B; ++b) {
    const float rstd = (PRO == 1) ? s_red[GV_WARPS * GV_MAXB + b] : 1.f;
    for (int i = tid; i < KP / 8; i += GV_THREADS) {
      uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * ldx + k0 + i * 8);
      if (PRO == 1) {
        uint4 gm = *reinterpret_cast<const uint4*>(gamma + k0 + i * 8);
        uint32_t* up = &u.x;
        const uint32_t* gp = &gm.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 v = unpack_half2(up[j]);
          float2 gg = unpack_half2(gp[j]);
          // weight * (x*rstd).to(fp16) : two roundings as in the reference
          up[j] = pack_half2(gg.x * __half2float(__float2half_rn(v.x * rstd)), gg.y * __half2float(__float2half_rn(v.y * rstd)));
        }
      }
      *reinterpret_cast<uint4*>(xs + b * pitch + i * 8) = u;
    }
  }
}

// ---- main loop over one staged phase: 16 rows x KP columns, K split over the 8 warps ----
__device__ __forceinline__ void gemv_phase(float (&c)[4], const __half* __restrict__ W, int K, int row0, int k0,
                                           int KP, const __half* xs, int B) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wk = KP / GV_WARPS;                 // columns per warp (multiple of 32)
  const int nch = wk / 32;
  const __half* ra = W + static_cast<size_t>(row0 + g) * K + k0 + warp * wk + t * 8;
  const __half* rb = ra + static_cast<size_t>(8) * K;
  const __half* xp = xs + g * (KP + XS_PAD) + warp * wk + t * 8;
  const bool xok = g < B;
  int ch = 0;
  for (; ch + 4 <= nch; ch += 4) {
    uint4 wa[4], wb[4], xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      wa[i] = ldg_nc_v4(ra + (ch + i) * 32);
      wb[i] = ldg_nc_v4(rb + (ch + i) * 32);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = xok ? *reinterpret_cast<const uint4*>(xp + (ch + i) * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mma_16816(c, wa[i].x, wb[i].x, wa[i].y, wb[i].y, xv[i].x, xv[i].y);
      mma_16816(c, wa[i].z, wb[i].z, wa[i].w, wb[i].w, xv[i].z, xv[i].w);
    }
  }
  for (; ch < nch; ++ch) {
    uint4 wa = ldg_nc_v4(ra + ch * 32), wb = ldg_nc_v4(rb + ch * 32);
    uint4 xv = xok ? *reinterpret_cast<const uint4*>(xp + ch * 32) : make_uint4(0, 0, 0, 0);
    mma_16816(c, wa.x, wb.x, wa.y, wb.y, xv.x, xv.y);
    mma_16816(c, wa.z, wb.z, wa.w, wb.w, xv.z, xv.w);
  }
}

// cross-warp reduction; returns the value for (row = tid>>3, n = tid&7), valid for tid < 128
__device__ __forceinline__ float gemv_reduce(const float (&c)[4], float* s_part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float* p = s_part + warp * (GV_ROWS * 8);
  p[g * 8 + 2 * t] = c[0];
  p[g * 8 + 2 * t + 1] = c[1];
  p[(g + 8) * 8 + 2 * t] = c[2];
  p[(g + 8) * 8 + 2 * t + 1] = c[3];
  __syncthreads();
  float v = 0.f;
  if (threadIdx.x < 128) {
#pragma unroll
    for (int w = 0; w < GV_WARPS; ++w) v += s_part[w * (GV_ROWS * 8) + threadIdx.x];
  }
  return v;
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(GV_THREADS)
gemv_kernel(const __half* __restrict__ W, const __half* __restrict__ x, int ldx, __half* __restrict__ out, int ldo,
            int B, int N, int K, int KP, const __half* __restrict__ gamma, float eps) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* xs = reinterpret_cast<__half*>(smem);
  float* s_part = reinterpret_cast<float*>(smem + static_cast<size_t>(B) * (KP + XS_PAD) * 2);
  float* s_red = s_part + GV_WARPS * GV_ROWS * 8;
  const int row0 = blockIdx.x * GV_ROWS;
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += KP) {
    if (k0 > 0) __syncthreads();
    stage_x<PRO>(xs, x, ldx, B, k0, KP, K, gamma, eps, s_red);
    __syncthreads();
    gemv_phase(c, W, K, row0, k0, KP, xs, B);
  }
  const float v = gemv_reduce(c, s_part);
  const int row = threadIdx.x >> 3, n = threadIdx.x & 7;
  if (threadIdx.x < 128) {
    if (EPI == 2) {
      const float partner = __shfl_xor_sync(0xffffffffu, v, 8);   // row ^ 1 (the `up` value for even rows)
      if ((row & 1) == 0 && n < B) out[static_cast<size_t>(n) * ldo + ((row0 + row) >> 1)] = __float2half_rn(silu_f(v) * partner);
    } else if (n < B) {
      __half* o = out + static_cast<size_t>(n) * ldo + row0 + row;
      *o = (EPI == 1) ? __float2half_rn(v + __half2float(*o)) : __float2half_rn(v);
    }
  }
}

int pick_kp(int K) {
  // largest multiple of 256 that divides K and is <= 4096
  for (int kp = 4096; kp >= 256; kp -= 256)
    if (K % kp == 0) return kp;
  return 0;
}

size_t gemv_smem(int KP, int B) { return static_cast<size_t>(B) * (KP + XS_PAD) * 2 + (GV_WARPS * GV_ROWS * 8 + GV_WARPS * GV_MAXB + GV_MAXB) * 4; }

template <typename Kern>
int ensure_smem(Kern k, size_t bytes, size_t* cached) {
  if (bytes > *cached) {
    TB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    *cached = bytes;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// head GEMV: logits + mask + argmax candidates
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void head_range(int h, int vocab, int& lo, int& hi) {
  if (h <= 0) { lo = 0; hi = vocab + 1; }                 // text head (0) or text-only mode (-1)
  else if (h == 1) { lo = vocab + 1; hi = vocab + 14; }   // time head
  else { lo = vocab + 14; hi = vocab + 27; }              // score head
}

__global__ void __launch_bounds__(GV_THREADS)
head_gemv_kernel(const __half* __restrict__ W, const __half* __restrict__ x, int ldx, int B, int n_valid, int K,
                 const __half* __restrict__ gamma, float eps, int vocab, const int32_t* __restrict__ heads,
                 float* __restrict__ cand_val, int32_t* __restrict__ cand_idx, float* __restrict__ logits_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* xs = reinterpret_cast<__half*>(smem);
  float* s_part = reinterpret_cast<float*>(smem + static_cast<size_t>(B) * (K + XS_PAD) * 2);
  float* s_red = s_part + GV_WARPS * GV_ROWS * 8;
  __shared__ float s_cv[4][8];
  __shared__ int s_ci[4][8];
  const int row0 = blockIdx.x * GV_ROWS;
  const int ncta = gridDim.x;
  int lo[GV_MAXB], hi[GV_MAXB];
  bool needed = false;
#pragma unroll
  for (int b = 0; b < GV_MAXB; ++b) {
    lo[b] = hi[b] = 0;
    if (b < B) {
      head_range(heads[b], vocab, lo[b], hi[b]);
      needed = needed || (row0 < hi[b] && row0 + GV_ROWS > lo[b]);
    }
  }
  if (!needed) {   // no row decodes from this slice of the vocabulary at this step: skip the weight read
    if (threadIdx.x < B) {
      cand_val[threadIdx.x * ncta + blockIdx.x] = -INFINITY;
      cand_idx[threadIdx.x * ncta + blockIdx.x] = 0x7fffffff;
    }
    if (logits_out != nullptr && threadIdx.x < 128) {
      const int row = threadIdx.x >> 3, n = threadIdx.x & 7;
      if (n < B && row0 + row < n_valid) logits_out[static_cast<size_t>(n) * n_valid + row0 + row] = -INFINITY;
    }
    return;
  }
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  stage_x<1>(xs, x, ldx, B, 0, K, K, gamma, eps, s_red);
  __syncthreads();
  gemv_phase(c, W, K, row0, 0, K, xs, B);
  float v = gemv_reduce(c, s_part);
  const int row = threadIdx.x >> 3, n = threadIdx.x & 7;
  int idx = 0x7fffffff;
  if (threadIdx.x < 128) {
    const int r = row0 + row;
    v = __half2float(__float2half_rn(v));          // Linear output is fp16, then .float() (trace_mistral.py:191-200)
    int l = 0, h = 0;
#pragma unroll
    for (int b = 0; b < GV_MAXB; ++b) if (b == n) { l = lo[b]; h = hi[b]; }
    if (n >= B || r < l || r >= h) v = -INFINITY;  // masked_fill of inactive ranges (:244-252)
    else idx = r;
    if (logits_out != nullptr && n < B && r < n_valid) logits_out[static_cast<size_t>(n) * n_valid + r] = v;
    // argmax over the rows held by this warp (lanes n, n+8, n+16, n+24), first index on ties
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) < 8) { s_cv[threadIdx.x >> 5][n] = v; s_ci[threadIdx.x >> 5][n] = idx; }
  }
  __syncthreads();
  if (threadIdx.x < B) {
    float bv = s_cv[0][threadIdx.x];
    int bi = s_ci[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 4; ++w) {
      const float ov = s_cv[w][threadIdx.x];
      const int oi = s_ci[w][threadIdx.x];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    cand_val[threadIdx.x * ncta + blockIdx.x] = bv;
    cand_idx[threadIdx.x * ncta + blockIdx.x] = bi;
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: greedy token, bookkeeping, head switch, next embedding
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finalize_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx, int ncand, int B, int vocab,
                int eos, int pad, const __half* __restrict__ embed, const __half* __restrict__ time_tab,
                const __half* __restrict__ score_tab, const __half* __restrict__ sync_tab, int H, DecodeState st,
                __half* __restrict__ x_next, int advance_pos) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_tok[GV_MAXB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int step = *st.step;
  for (int b = 0; b < B; ++b) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < ncand; i += 256) {
      const float v = cand_val[b * ncand + i];
      const int ix = cand_idx[b * ncand + i];
      if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
      int tok = st.finished[b] ? pad : bi;                        // finished rows keep emitting pad (HF greedy)
      if (step < st.max_new) st.out_ids[static_cast<int64_t>(b) * st.max_new + step] = tok;
      if (!st.finished[b] && tok == eos) st.finished[b] = 1;
      const int h = st.heads[b];
      if (h >= 0) {                                               // trace_mistral.py:86-88, 336-344
        if (tok == vocab) st.heads[b] = 1;
        else if (tok == vocab + 1) st.heads[b] = 2;
        else if (tok == vocab + 14) st.heads[b] = 0;
      }
      if (advance_pos) st.pos[b] += 1;
      s_tok[b] = tok;
    }
    __syncthreads();
  }
  if (tid == 0) {
    int all = 1;
    for (int b = 0; b < B; ++b) all &= st.finished[b];
    *st.step = step + 1;
    if (all && *st.done_step == 0) *st.done_step = step + 1;
  }
  // next-step input embedding (trace_arch.py:349-373): id%vocab text row, overwritten by sync/time/score rows
  for (int b = 0; b < B; ++b) {
    const int tok = s_tok[b];
    const __half* src;
    if (tok == vocab) src = sync_tab;
    else if (tok > vocab && tok < vocab + 14) src = time_tab + static_cast<size_t>(tok - vocab - 1) * H;
    else if (tok >= vocab + 14) src = score_tab + static_cast<size_t>(tok - vocab - 14) * H;
    else src = embed + static_cast<size_t>(tok % vocab) * H;
    for (int i = tid; i < H / 8; i += 256)
      reinterpret_cast<uint4*>(x_next + static_cast<size_t>(b) * H)[i] = reinterpret_cast<const uint4*>(src)[i];
  }
}

// ------------------------------------------------------------------------------------------------
// decode attention
// ------------------------------------------------------------------------------------------------
constexpr int AD_THREADS = 128;
constexpr int AD_G = 4;      // query heads per kv head (Mistral 32/8)
constexpr int AD_HD = 128;

__global__ void __launch_bounds__(AD_THREADS)
attn_decode_kernel(const __half* __restrict__ qkv, int heads, int kvh, const float* __restrict__ cos_tab,
                   const float* __restrict__ sin_tab, __half* __restrict__ k_pool, __half* __restrict__ v_pool,
                   const int32_t* __restrict__ block_table, int max_pages, int page_size, DecodeState st, int nsplit,
                   int chunk_max, float* __restrict__ part_o, float* __restrict__ part_ml,
                   int32_t* __restrict__ counters, __half* __restrict__ attn_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* qs = reinterpret_cast<float*>(smem);              // [4][128]
  float* knew = qs + AD_G * AD_HD;                         // [128]
  float* vnew = knew + AD_HD;                              // [128]
  float* red = vnew + AD_HD;                               // [4 warps][4][128]
  float* sc = red + 4 * AD_G * AD_HD;                      // [4][chunk_max + 1]
  __shared__ int s_last;
  const int split = blockIdx.x, kv = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t = st.pos[b];                                 // cached keys; also the position id of the new token
  const int slot = st.slot_ids[b];
  const int32_t* table = block_table + static_cast<size_t>(slot) * max_pages;
  const int row_w = (heads + 2 * kvh) * AD_HD;
  const __half* row = qkv + static_cast<size_t>(b) * row_w;
  const float scale = rsqrtf(static_cast<float>(AD_HD));
  const int cpitch = chunk_max + 1;

  // ---- RoPE on the 4 query heads (and on the new key, split 0) ----
  const float* cs = cos_tab + static_cast<size_t>(t) * (AD_HD / 2);
  const float* sn = sin_tab + static_cast<size_t>(t) * (AD_HD / 2);
  for (int i = tid; i < AD_G * (AD_HD / 2); i += AD_THREADS) {
    const int h = i / (AD_HD / 2), d = i % (AD_HD / 2);
    const __half* qp = row + (kv * AD_G + h) * AD_HD;
    const float a = __half2float(qp[d]), c2 = __half2float(qp[d + AD_HD / 2]);
    qs[h * AD_HD + d] = __half2float(__float2half_rn(a * cs[d] - c2 * sn[d])) * scale;
    qs[h * AD_HD + d + AD_HD / 2] = __half2float(__float2half_rn(c2 * cs[d] + a * sn[d])) * scale;
  }
  if (split == 0) {
    const int page = table[t / page_size], poff = t % page_size;
    const size_t dst = ((static_cast<size_t>(page) * kvh + kv) * page_size + poff) * AD_HD;
    if (tid < AD_HD / 2) {
      const __half* kp = row + (heads + kv) * AD_HD;
      const float a = __half2float(kp[tid]), c2 = __half2float(kp[tid + AD_HD / 2]);
      const __half ra = __float2half_rn(a * cs[tid] - c2 * sn[tid]);
      const __half rb = __float2half_rn(c2 * cs[tid] + a * sn[tid]);
      k_pool[dst + tid] = ra;
      k_pool[dst + tid + AD_HD / 2] = rb;
      knew[tid] = __half2float(ra);
      knew[tid + AD_HD / 2] = __half2float(rb);
    } else {
      const int d = (tid - AD_HD / 2) * 2;
      const __half* vp = row + (heads + kvh + kv) * AD_HD;
      const __half2 v2 = *reinterpret_cast<const __half2*>(vp + d);
      *reinterpret_cast<__half2*>(v_pool + dst + d) = v2;
      vnew[d] = __low2float(v2);
      vnew[d + 1] = __high2float(v2);
    }
  }
  __syncthreads();

  // ---- this CTA's key range ----
  int chunk = (t + nsplit - 1) / nsplit;
  chunk = (chunk + 31) & ~31;
  const int start = split * chunk;
  const int end = min(start + chunk, t);
  const int nloc = max(end - start, 0);

  // ---- phase A: scores, lane <-> key ----
  for (int base = warp * 32; base < nloc; base += AD_THREADS) {
    const int j = base + lane;
    const bool valid = j < nloc;
    const int key = start + (valid ? j : 0);
    const int page = table[key / page_size];
    const __half* kr = k_pool + ((static_cast<size_t>(page) * kvh + kv) * page_size + key % page_size) * AD_HD;
    float acc[AD_G] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int cI = 0; cI < AD_HD / 8; ++cI) {
      const uint4 u = *reinterpret_cast<const uint4*>(kr + cI * 8);
      float kf[8];
      float2 p0 = unpack_half2(u.x), p1 = unpack_half2(u.y), p2 = unpack_half2(u.z), p3 = unpack_half2(u.w);
      kf[0] = p0.x; kf[1] = p0.y; kf[2] = p1.x; kf[3] = p1.y; kf[4] = p2.x; kf[5] = p2.y; kf[6] = p3.x; kf[7] = p3.y;
#pragma unroll
      for (int h = 0; h < AD_G; ++h) {
        const float4 q0 = *reinterpret_cast<const float4*>(qs + h * AD_HD + cI * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(qs + h * AD_HD + cI * 8 + 4);
        acc[h] += q0.x * kf[0] + q0.y * kf[1] + q0.z * kf[2] + q0.w * kf[3] + q1.x * kf[4] + q1.y * kf[5] + q1.z * kf[6] + q1.w * kf[7];
      }
    }
    if (valid) {
#pragma unroll
      for (int h = 0; h < AD_G; ++h) sc[h * cpitch + j] = acc[h];
    }
  }
  int ntot = nloc;
  if (split == 0) {   // the token being decoded attends to itself: key from registers, not from the cache
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) d += qs[warp * AD_HD + lane * 4 + i] * knew[lane * 4 + i];
    d = warp_sum(d);
    if (lane == 0) sc[warp * cpitch + nloc] = d;
    ntot = nloc + 1;
  }
  __syncthreads();

  // ---- softmax statistics, warp h <-> head h ----
  float m = -INFINITY, l = 0.f;
  {
    float* s = sc + warp * cpitch;
    for (int j = lane; j < ntot; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    for (int j = lane; j < ntot; j += 32) {
      const float p = __expf(s[j] - m);
      s[j] = p;
      l += p;
    }
    l = warp_sum(l);
  }
  __syncthreads();

  // ---- phase B: O = P V, lane <-> 4 dims, warps stride over keys ----
  float o[AD_G][4];
#pragma unroll
  for (int h = 0; h < AD_G; ++h) o[h][0] = o[h][1] = o[h][2] = o[h][3] = 0.f;
  for (int j = warp; j < nloc; j += 4) {
    const int key = start + j;
    const int page = table[key / page_size];
    const __half* vr = v_pool + ((static_cast<size_t>(page) * kvh + kv) * page_size + key % page_size) * AD_HD + lane * 4;
    const uint2 u = *reinterpret_cast<const uint2*>(vr);
    const float2 v0 = unpack_half2(u.x), v1 = unpack_half2(u.y);
#pragma unroll
    for (int h = 0; h < AD_G; ++h) {
      const float p = sc[h * cpitch + j];
      o[h][0] += p * v0.x; o[h][1] += p * v0.y; o[h][2] += p * v1.x; o[h][3] += p * v1.y;
    }
  }
  if (split == 0 && warp == 0) {
#pragma unroll
    for (int h = 0; h < AD_G; ++h) {
      const float p = sc[h * cpitch + nloc];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[h][i] += p * vnew[lane * 4 + i];
    }
  }
#pragma unroll
  for (int h = 0; h < AD_G; ++h)
    *reinterpret_cast<float4*>(red + (warp * AD_G + h) * AD_HD + lane * 4) = make_float4(o[h][0], o[h][1], o[h][2], o[h][3]);
  __syncthreads();
  const size_t pbase = (static_cast<size_t>(b) * kvh + kv) * nsplit + split;
  {
    float4 a = *reinterpret_cast<const float4*>(red + tid * 4);
#pragma unroll
    for (int w = 1; w < 4; ++w) {
      const float4 x4 = *reinterpret_cast<const float4*>(red + w * AD_G * AD_HD + tid * 4);
      a.x += x4.x; a.y += x4.y; a.z += x4.z; a.w += x4.w;
    }
    *reinterpret_cast<float4*>(part_o + pbase * (AD_G * AD_HD) + tid * 4) = a;
    if (lane == 0) {
      part_ml[pbase * (AD_G * 2) + warp * 2] = m;
      part_ml[pbase * (AD_G * 2) + warp * 2 + 1] = l;
    }
  }

  // ---- the last CTA of this (row, kv head) combines the split partials ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&counters[b * kvh + kv], 1) == nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const size_t gb = (static_cast<size_t>(b) * kvh + kv) * nsplit;
  const int h = tid >> 5;              // 32 threads x 4 dims per head
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(part_ml + (gb + s) * (AD_G * 2) + h * 2));
  float L = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nsplit; ++s) {
    const float ms = __ldcg(part_ml + (gb + s) * (AD_G * 2) + h * 2);
    const float ls = __ldcg(part_ml + (gb + s) * (AD_G * 2) + h * 2 + 1);
    if (ls == 0.f) continue;
    const float w = __expf(ms - M);
    L += w * ls;
    const float4 po = __ldcg(reinterpret_cast<const float4*>(part_o + (gb + s) * (AD_G * AD_HD) + tid * 4));
    acc.x += w * po.x; acc.y += w * po.y; acc.z += w * po.z; acc.w += w * po.w;
  }
  const float inv = 1.f / L;
  __half* op = attn_out + static_cast<size_t>(b) * heads * AD_HD + (kv * AD_G) * AD_HD + tid * 4;
  uint2 pk;
  pk.x = pack_half2(acc.x * inv, acc.y * inv);
  pk.y = pack_half2(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(op) = pk;
  if (tid == 0) counters[b * kvh + kv] = 0;
}

}  // namespace

int gemv_f16(const __half* W, const __half* x, int ldx, __half* out, int ldo, int B, int N, int K, int pro,
             const __half* gamma, float eps, int epi, cudaStream_t stream) {
  TB_CHECK(B >= 1 && B <= GV_MAXB, "decode batch must be in [1, 8]");
  TB_CHECK(N % GV_ROWS == 0, "GEMV rows must be a multiple of 16");
  const int KP = pick_kp(K);
  TB_CHECK(KP > 0, "GEMV K must be a multiple of 256");
  TB_CHECK(!(pro == 1 && KP != K), "fused RMSNorm needs K <= 4096");
  TB_CHECK(ldx % 8 == 0, "activation pitch must be 16-byte aligned");
  const size_t smem = gemv_smem(KP, B);
  const int grid = N / GV_ROWS;
#define TB_GV(P, E)                                                                                   \
  do {                                                                                                \
    static size_t cached = 0;                                                                         \
    if (ensure_smem(gemv_kernel<P, E>, smem, &cached)) return 1;                                      \
    gemv_kernel<P, E><<<grid, GV_THREADS, smem, stream>>>(W, x, ldx, out, ldo, B, N, K, KP, gamma, eps); \
  } while (0)
  if (pro == 0 && epi == 0) TB_GV(0, 0);
  else if (pro == 0 && epi == 1) TB_GV(0, 1);
  else if (pro == 1 && epi == 0) TB_GV(1, 0);
  else if (pro == 1 && epi == 2) TB_GV(1, 2);
  else if (pro == 0 && epi == 2) TB_GV(0, 2);
  else if (pro == 1 && epi == 1) TB_GV(1, 1);
  else TB_CHECK(false, "unsupported GEMV prologue/epilogue combination");
#undef TB_GV
  TB_LAUNCH_OK();
  return 0;
}

int attn_decode(const __half* qkv, int B, int heads, int kvh, int hd, const float* cos_tab, const float* sin_tab,
                __half* k_pool, __half* v_pool, const int32_t* block_table, int max_pages, int page_size,
                const DecodeState& st, int nsplit, int max_ctx, float* part_o, float* part_ml, int32_t* counters,
                __half* attn_out, cudaStream_t stream) {
  TB_CHECK(hd == AD_HD && heads == kvh * AD_G, "decode attention is specialised for head_dim 128 and 4 q heads per kv head");
  TB_CHECK(nsplit >= 1 && nsplit <= 64, "bad split count");
  int chunk_max = (ceil_div(max_ctx, nsplit) + 31) & ~31;
  const size_t smem = (static_cast<size_t>(AD_G) * AD_HD + 2 * AD_HD + 4 * AD_G * AD_HD + static_cast<size_t>(AD_G) * (chunk_max + 1)) * 4;
  TB_CHECK(smem <= 200 * 1024, "decode attention chunk does not fit shared memory; raise nsplit");
  static size_t cached = 0;
  if (ensure_smem(attn_decode_kernel, smem, &cached)) return 1;
  attn_decode_kernel<<<dim3(nsplit, kvh, B), AD_THREADS, smem, stream>>>(qkv, heads, kvh, cos_tab, sin_tab, k_pool, v_pool,
                                                                         block_table, max_pages, page_size, st, nsplit,
                                                                         chunk_max, part_o, part_ml, counters, attn_out);
  TB_LAUNCH_OK();
  return 0;
}

int head_gemv(const __half* Wh, const __half* x, int ldx, int B, int n_valid, int Npad, int K, const __half* gamma,
              float eps, int vocab, const DecodeState& st, float* cand_val, int32_t* cand_idx, float* logits_out,
              cudaStream_t stream) {
  TB_CHECK(B >= 1 && B <= GV_MAXB, "decode batch must be in [1, 8]");
  TB_CHECK(Npad % GV_ROWS == 0 && K % 256 == 0 && K <= 4096, "head GEMV shape");
  const size_t smem = gemv_smem(K, B);
  static size_t cached = 0;
  if (ensure_smem(head_gemv_kernel, smem, &cached)) return 1;
  head_gemv_kernel<<<Npad / GV_ROWS, GV_THREADS, smem, stream>>>(Wh, x, ldx, B, n_valid, K, gamma, eps, vocab, st.heads,
                                                                 cand_val, cand_idx, logits_out);
  TB_LAUNCH_OK();
  return 0;
}

int decode_finalize(const float* cand_val, const int32_t* cand_idx, int ncand, int B, int vocab, int eos, int pad,
                    const __half* embed, const __half* time_tab, const __half* score_tab, const __half* sync_tab,
                    int H, const DecodeState& st, __half* x_next, int advance_pos, cudaStream_t stream) {
  finalize_kernel<<<1, 256, 0, stream>>>(cand_val, cand_idx, ncand, B, vocab, eos, pad, embed, time_tab, score_tab,
                                         sync_tab, H, st, x_next, advance_pos);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
