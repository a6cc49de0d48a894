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
#include "decode_dev.cuh"

namespace tb {
namespace {

template <int PRO, int EPI>
__global__ void __launch_bounds__(GV_THREADS, 1)
gemv_kernel(const __half* __restrict__ W, const __half* __restrict__ x, int ldx, __half* __restrict__ out, int ldo,
            int B, int N, int K, int KP, int KS, int max_tiles, int stages, const __half* __restrict__ gamma, float eps) {
  extern __shared__ __align__(128) uint8_t smem[];
  const GvSmem m = gv_carve(smem, B, KP, KS, max_tiles, stages);
  int r0, r1;
  cta_rows(N, r0, r1);
  gv_init_barriers(m, stages);
  pdl_launch_dependents();
  if (threadIdx.x >= GV_CONS) {
    int ps = 0;
    uint32_t pph = 0;
    gemv_producer(W, K, KP, KS, r0, r1, r0, r1, m, stages, ps, pph);    // weights never depend on the previous kernel
    return;
  }
  GammaRegs gm;
  if (PRO == 1) gm = load_gamma(gamma, K);                     // a weight: fetched before the dependency wait
  pdl_wait();                                                  // activations do depend on the previous kernel
  int s = 0;
  uint32_t ph = 0;
  for (int k0 = 0; k0 < K; k0 += KP) {
    if (k0 > 0) cons_sync();
    if (PRO == 1) stage_x_rmsnorm(m.xs, x, ldx, B, K, gm, eps, m.s_red);
    else stage_x_plain(m.xs, x, ldx, B, k0, KP);
    cons_sync();
    gemv_consume_phase(k0, KP, KS, r0, r1, r0, r1, m, B, stages, s, ph);
  }
  cons_sync();
  const int nout = ((r1 - r0 + GV_TILE - 1) / GV_TILE) * 128;
  for (int idx = threadIdx.x; idx < nout; idx += GV_CONS) {
    const float v = gemv_out(m.s_part, idx);
    const int row = r0 + (idx >> 7) * GV_TILE + ((idx & 127) >> 3), n = idx & 7;
    if (EPI == 2) {
      const float partner = __shfl_xor_sync(0xffffffffu, v, 8);   // row ^ 1 (the `up` value for even rows)
      if ((row & 1) == 0 && n < B && row < r1) out[static_cast<size_t>(n) * ldo + (row >> 1)] = __float2half_rn(silu_f(v) * partner);
    } else if (n < B && row < r1) {
      __half* o = out + static_cast<size_t>(n) * ldo + row;
      *o = (EPI == 1) ? __float2half_rn(v + __half2float(*o)) : __float2half_rn(v);
    }
  }
}

int pick_kp(int K, int B) {
  // largest multiple of 256 that divides K with the staged activations (B rows) <= 64 KB
  for (int kp = 16384; kp >= 256; kp -= 256)
    if (K % kp == 0 && static_cast<size_t>(B) * (kp + XS_PAD) * 2 <= 72 * 1024) return kp;
  return 0;
}
int pick_ks(int KP) {
  for (int ks = 1024; ks >= 256; ks -= 256)
    if (KP % ks == 0) return ks;
  return 0;
}

int gemv_grid(int N, int num_sms) {
  int g = num_sms;
  if (g > N / 2) g = N / 2;
  return g < 1 ? 1 : g;
}

struct GvPlan { int KP, KS, max_tiles, stages, grid; size_t smem; };
int gemv_plan(int N, int K, int B, int kp_force, int num_sms, GvPlan* p) {
  p->KP = kp_force > 0 ? kp_force : pick_kp(K, B);
  TB_CHECK(p->KP > 0, "GEMV K must be a multiple of 256");
  p->KS = pick_ks(p->KP);
  p->grid = gemv_grid(N, num_sms);
  const int rows = ((N / 2 + p->grid - 1) / p->grid) * 2 + 2;
  p->max_tiles = (rows + GV_TILE - 1) / GV_TILE;
  const size_t fixed = static_cast<size_t>(B) * (p->KP + XS_PAD) * 2 + static_cast<size_t>(p->max_tiles) * GV_WARPS * GV_TILE * 8 * 4 +
                       (GV_WARPS * GV_MAXB + GV_MAXB) * 4 + 2 * GV_MAXSTAGES * 8 + 128;
  const size_t stage = static_cast<size_t>(GV_TILE) * (p->KS + WS_PAD) * 2;
  static size_t budget = 0;
  if (budget == 0) {
    const char* e = getenv("TRACE_B200_GEMV_SMEM_KB");
    budget = static_cast<size_t>((e != nullptr && atoi(e) >= 48 && atoi(e) <= 220) ? atoi(e) : 200) * 1024;
  }
  TB_CHECK(fixed + 2 * stage <= budget, "GEMV shared memory budget exceeded");
  int st = static_cast<int>((budget - fixed) / stage);
  if (st > GV_MAXSTAGES) st = GV_MAXSTAGES;
  p->stages = st;
  p->smem = fixed + st * stage;
  return 0;
}

template <typename Kern>
int ensure_smem(Kern k, size_t bytes, size_t* cached) {
  if (bytes > *cached) {
    TB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    *cached = bytes;
  }
  return 0;
}

// launch with programmatic dependent launch enabled (the kernel calls griddepcontrol.wait before it touches
// anything the previous kernel wrote)
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TB_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// head GEMV: logits + mask + argmax candidates
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GV_THREADS, 1)
head_gemv_kernel(const __half* __restrict__ W, const __half* __restrict__ x, int ldx, int B, int n_valid, int Npad, int K,
                 int KS, int max_tiles, int stages, const __half* __restrict__ gamma, float eps, int vocab,
                 const int32_t* __restrict__ heads, float* __restrict__ cand_val, int32_t* __restrict__ cand_idx,
                 float* __restrict__ logits_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const GvSmem m = gv_carve(smem, B, K, KS, max_tiles, stages);
  __shared__ float s_cv[GV_WARPS][8];
  __shared__ int s_ci[GV_WARPS][8];
  int r0, r1;
  cta_rows(Npad, r0, r1);
  const int ncta = gridDim.x;
  gv_init_barriers(m, stages);
  pdl_launch_dependents();
  pdl_wait();                                  // heads[] is written by the previous step's finalize kernel
  int lo[GV_MAXB], hi[GV_MAXB];
  int ulo = 0x7fffffff, uhi = 0;
#pragma unroll
  for (int b = 0; b < GV_MAXB; ++b) {
    lo[b] = hi[b] = 0;
    if (b < B) {
      head_range(heads[b], vocab, lo[b], hi[b]);
      ulo = min(ulo, lo[b]);
      uhi = max(uhi, hi[b]);
    }
  }
  const int vlo = max(ulo, r0), vhi = min(uhi, r1);   // only rows some batch row decodes from are read at all
  if (threadIdx.x >= GV_CONS) {
    int ps = 0;
    uint32_t pph = 0;
    if (vlo < vhi) gemv_producer(W, K, K, KS, r0, r1, vlo, vhi, m, stages, ps, pph);
    return;
  }
  const int n = threadIdx.x & 7;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const int nout = ((r1 - r0 + GV_TILE - 1) / GV_TILE) * 128;
  if (vlo < vhi) {
    const GammaRegs gm = load_gamma(gamma, K);
    stage_x_rmsnorm(m.xs, x, ldx, B, K, gm, eps, m.s_red);
    cons_sync();
    int s = 0;
    uint32_t ph = 0;
    gemv_consume_phase(0, K, KS, r0, r1, vlo, vhi, m, B, stages, s, ph);
    cons_sync();
    int l = 0, h = 0;
#pragma unroll
    for (int b = 0; b < GV_MAXB; ++b) if (b == n) { l = lo[b]; h = hi[b]; }
    for (int idx = threadIdx.x; idx < nout; idx += GV_CONS) {
      const int r = r0 + (idx >> 7) * GV_TILE + ((idx & 127) >> 3);
      float v = __half2float(__float2half_rn(gemv_out(m.s_part, idx)));   // Linear output is fp16, then .float() (trace_mistral.py:191-200)
      const bool act = (n < B) && r < r1 && r >= l && r < h;               // masked_fill of inactive ranges (:244-252)
      if (!act) v = -INFINITY;
      if (logits_out != nullptr && n < B && r < r1 && r < n_valid) logits_out[static_cast<size_t>(n) * n_valid + r] = v;
      if (act && (v > bv || (v == bv && r < bi))) { bv = v; bi = r; }
    }
  } else if (logits_out != nullptr) {
    for (int idx = threadIdx.x; idx < nout; idx += GV_CONS) {
      const int r = r0 + (idx >> 7) * GV_TILE + ((idx & 127) >> 3);
      if (n < B && r < r1 && r < n_valid) logits_out[static_cast<size_t>(n) * n_valid + r] = -INFINITY;
    }
  }
  // argmax per batch row over the CTA: lanes with equal n (xor 8, 16), then the 8 warps; first index on ties
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) < 8) { s_cv[threadIdx.x >> 5][n] = bv; s_ci[threadIdx.x >> 5][n] = bi; }
  cons_sync();
  if (threadIdx.x < B) {
    float v = s_cv[0][threadIdx.x];
    int i = s_ci[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < GV_WARPS; ++w) {
      const float ov = s_cv[w][threadIdx.x];
      const int oi = s_ci[w][threadIdx.x];
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    cand_val[threadIdx.x * ncta + blockIdx.x] = v;
    cand_idx[threadIdx.x * ncta + blockIdx.x] = i;
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
  pdl_launch_dependents();
  pdl_wait();
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
  pdl_launch_dependents();
  pdl_wait();
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

  // ---- phase A: scores.  Two lanes per key (64 dims each), 16 keys per warp per pass; all 8 loads of a lane
  //      are issued before the first FMA (memory-level parallelism: the KV rows come from HBM, not L2) ----
  for (int base = warp * 16; base < nloc; base += 64) {
    const int j = base + (lane >> 1);
    const int hf = lane & 1;
    const bool valid = j < nloc;
    const int key = start + (valid ? j : 0);
    const int page = table[key / page_size];
    const __half* kr = k_pool + ((static_cast<size_t>(page) * kvh + kv) * page_size + key % page_size) * AD_HD + hf * 64;
    uint4 kreg[8];
#pragma unroll
    for (int cI = 0; cI < 8; ++cI) kreg[cI] = *reinterpret_cast<const uint4*>(kr + cI * 8);
    float acc[AD_G] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int cI = 0; cI < 8; ++cI) {
      float kf[8];
      float2 p0 = unpack_half2(kreg[cI].x), p1 = unpack_half2(kreg[cI].y), p2 = unpack_half2(kreg[cI].z), p3 = unpack_half2(kreg[cI].w);
      kf[0] = p0.x; kf[1] = p0.y; kf[2] = p1.x; kf[3] = p1.y; kf[4] = p2.x; kf[5] = p2.y; kf[6] = p3.x; kf[7] = p3.y;
#pragma unroll
      for (int h = 0; h < AD_G; ++h) {
        const float4 q0 = *reinterpret_cast<const float4*>(qs + h * AD_HD + hf * 64 + cI * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(qs + h * AD_HD + hf * 64 + cI * 8 + 4);
        acc[h] += q0.x * kf[0] + q0.y * kf[1] + q0.z * kf[2] + q0.w * kf[3] + q1.x * kf[4] + q1.y * kf[5] + q1.z * kf[6] + q1.w * kf[7];
      }
    }
#pragma unroll
    for (int h = 0; h < AD_G; ++h) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 1);
    if (valid && hf == 0) {
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
  for (int j0 = warp; j0 < nloc; j0 += 32) {       // 8 keys per pass per warp, loads first
    uint2 vreg[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * 4;
      const int key = start + (j < nloc ? j : 0);
      const int page = table[key / page_size];
      vreg[u] = *reinterpret_cast<const uint2*>(v_pool + ((static_cast<size_t>(page) * kvh + kv) * page_size + key % page_size) * AD_HD + lane * 4);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u * 4;
      if (j < nloc) {
        const float2 v0 = unpack_half2(vreg[u].x), v1 = unpack_half2(vreg[u].y);
#pragma unroll
        for (int h = 0; h < AD_G; ++h) {
          const float p = sc[h * cpitch + j];
          o[h][0] += p * v0.x; o[h][1] += p * v0.y; o[h][2] += p * v1.x; o[h][3] += p * v1.y;
        }
      }
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

static int g_num_sms = 0;
static int gv_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int gemv_f16(const __half* W, const __half* x, int ldx, __half* out, int ldo, int B, int N, int K, int pro,
             const __half* gamma, float eps, int epi, cudaStream_t stream) {
  TB_CHECK(B >= 1 && B <= GV_MAXB, "decode batch must be in [1, 8]");
  TB_CHECK(N % 2 == 0 && N >= 2, "GEMV rows must be even");
  TB_CHECK(ldx % 8 == 0 && K % 256 == 0, "GEMV K must be a multiple of 256 and the activation pitch 16-byte aligned");
  TB_CHECK((reinterpret_cast<uintptr_t>(W) & 15) == 0, "weights must be 16-byte aligned");
  GvPlan p;
  if (gemv_plan(N, K, B, 0, gv_sms(), &p)) return 1;
  TB_CHECK(!(pro == 1 && p.KP != K), "fused RMSNorm needs the whole row staged (K <= 4096)");
#define TB_GV(P, E)                                                                                              \
  do {                                                                                                           \
    static size_t cached = 0;                                                                                    \
    if (ensure_smem(gemv_kernel<P, E>, p.smem, &cached)) return 1;                                               \
    if (launch_pdl(gemv_kernel<P, E>, dim3(p.grid), dim3(GV_THREADS), p.smem, stream, W, x, ldx, out, ldo, B, N, K, p.KP, p.KS, \
                   p.max_tiles, p.stages, gamma, eps)) return 1;                                                 \
  } while (0)
  if (pro == 0 && epi == 0) TB_GV(0, 0);
  else if (pro == 0 && epi == 1) TB_GV(0, 1);
  else if (pro == 1 && epi == 0) TB_GV(1, 0);
  else if (pro == 1 && epi == 2) TB_GV(1, 2);
  else if (pro == 0 && epi == 2) TB_GV(0, 2);
  else if (pro == 1 && epi == 1) TB_GV(1, 1);
  else TB_CHECK(false, "unsupported GEMV prologue/epilogue combination");
#undef TB_GV
  return 0;
}

int head_gemv_ncand() { return gemv_grid(1 << 30, gv_sms()); }

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
  return launch_pdl(attn_decode_kernel, dim3(nsplit, kvh, B), dim3(AD_THREADS), smem, stream, qkv, heads, kvh, cos_tab, sin_tab,
                    k_pool, v_pool, block_table, max_pages, page_size, st, nsplit, chunk_max, part_o, part_ml, counters, attn_out);
}

int head_gemv(const __half* Wh, const __half* x, int ldx, int B, int n_valid, int Npad, int K, const __half* gamma,
              float eps, int vocab, const DecodeState& st, float* cand_val, int32_t* cand_idx, float* logits_out,
              cudaStream_t stream) {
  TB_CHECK(B >= 1 && B <= GV_MAXB, "decode batch must be in [1, 8]");
  TB_CHECK(Npad % 2 == 0 && K % 256 == 0 && K <= 4096, "head GEMV shape");
  GvPlan p;
  if (gemv_plan(Npad, K, B, K, gv_sms(), &p)) return 1;
  TB_CHECK(p.grid == head_gemv_ncand(), "head GEMV grid mismatch");
  static size_t cached = 0;
  if (ensure_smem(head_gemv_kernel, p.smem, &cached)) return 1;
  return launch_pdl(head_gemv_kernel, dim3(p.grid), dim3(GV_THREADS), p.smem, stream, Wh, x, ldx, B, n_valid, Npad, K, p.KS,
                    p.max_tiles, p.stages, gamma, eps, vocab, static_cast<const int32_t*>(st.heads), cand_val, cand_idx, logits_out);
}

int decode_finalize(const float* cand_val, const int32_t* cand_idx, int ncand, int B, int vocab, int eos, int pad,
                    const __half* embed, const __half* time_tab, const __half* score_tab, const __half* sync_tab,
                    int H, const DecodeState& st, __half* x_next, int advance_pos, cudaStream_t stream) {
  return launch_pdl(finalize_kernel, dim3(1), dim3(256), 0, stream, cand_val, cand_idx, ncand, B, vocab, eos, pad, embed, time_tab,
                    score_tab, sync_tab, H, st, x_next, advance_pos);
}

}  // namespace tb
