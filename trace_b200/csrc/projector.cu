// SpatialSlotPool (multimodal_projector/builder.py:427-467) up to the slot-weighted sums:
//   y   = LayerNorm(feat, eps 1e-6) ; y = y*cos + rotate_half(y)*sin   (RoPE over the PATCH index, full-width
//         rotate-half: dims d and d+D/2 are partners, :326-330, :451-455)
//   lg  = y @ slots[D,8] ; p = softmax(lg, over the n patches) ; res[s,:] = sum_n p[n,s] * y[n,:]
// The readout Linear (D -> 4096, :467) runs on the tcgen05 GEMM over res[F*8, D].
//
// HBM/L2-bound and tiny (15 MFLOP per frame), so the only thing that matters is parallelism: the first version ran one
// CTA per frame (32 CTAs on 148 SMs, 1.0 ms per 32 frames).  Now two grid-filling kernels:
//   slot_logits_kernel : warp per patch row (F*n/8 CTAs): LN statistics, rotated row, 8 slot logits -> global scratch
//   slot_sum_kernel    : CTA per (frame, 128-dim slice) (F*D/128 CTAs): softmax over the patches (recomputed per CTA from
//                        the 18 KB of logits), then thread = (partner pair, row group): recompute the rotated values from
//                        the cached statistics and accumulate the 8 weighted sums; the 4 row groups are combined in a
//                        fixed order through shared memory (deterministic).
// Algorithmic bytes: n*D*2 read (twice, second pass from L2) + 8*D*2 written per frame.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int PJ_THREADS = 256;
constexpr int NS = 8;

// ---- kernel 1: one warp per patch row -> logits [F][n][8] fp32 (fp16-rounded values), stats [F][n][2] ----
__global__ void __launch_bounds__(PJ_THREADS)
slot_logits_kernel(const __half* __restrict__ feats, int ld, int frame_rows, int row_off, int n, int D, int F,
                   const __half* __restrict__ ln_w, const __half* __restrict__ ln_b, float eps,
                   const __half* __restrict__ slots, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                   float* __restrict__ logits, float* __restrict__ stats) {
  extern __shared__ __align__(16) uint8_t smem[];
  __half* s_slots = reinterpret_cast<__half*>(smem);               // [D][8]
  __half* s_w = s_slots + D * NS;                                  // [D]
  __half* s_b = s_w + D;                                           // [D]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half_d = D / 2;
  for (int i = tid; i < D * NS / 8; i += PJ_THREADS)
    reinterpret_cast<uint4*>(s_slots)[i] = reinterpret_cast<const uint4*>(slots)[i];
  for (int i = tid; i < D / 8; i += PJ_THREADS) {
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(ln_w)[i];
    reinterpret_cast<uint4*>(s_b)[i] = reinterpret_cast<const uint4*>(ln_b)[i];
  }
  __syncthreads();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (PJ_THREADS / 32) + warp;     // flattened (frame, patch)
  if (row >= static_cast<int64_t>(F) * n) return;
  const int f = static_cast<int>(row / n), r = static_cast<int>(row % n);
  const __half* xr = feats + (static_cast<size_t>(f) * frame_rows + row_off + r) * ld;
  float sum = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    float2 v = unpack_half2(*reinterpret_cast<const uint32_t*>(xr + d));
    sum += v.x + v.y;
  }
  const float mean = warp_sum(sum) / D;
  float sq = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    float2 v = unpack_half2(*reinterpret_cast<const uint32_t*>(xr + d));
    sq += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean);
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
  float lg[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) lg[s2] = 0.f;
  // partner pairs (d, d + D/2): y'_d = y_d*c - y_{d+h}*s ; y'_{d+h} = y_{d+h}*c + y_d*s  (cos/sin tables are
  // the duplicated-frequency form emb = cat(freqs, freqs): index d and d+h share freq d)
  for (int d = lane; d < half_d; d += 32) {
    const float ya = __half2float(__float2half_rn((__half2float(xr[d]) - mean) * rstd * __half2float(s_w[d]) + __half2float(s_b[d])));
    const float yb = __half2float(__float2half_rn((__half2float(xr[d + half_d]) - mean) * rstd * __half2float(s_w[d + half_d]) + __half2float(s_b[d + half_d])));
    const float c = cos_tab[static_cast<size_t>(r) * half_d + d], sn = sin_tab[static_cast<size_t>(r) * half_d + d];
    const float ra = __half2float(__float2half_rn(ya * c - yb * sn));
    const float rb = __half2float(__float2half_rn(yb * c + ya * sn));
    const uint4 wa = *reinterpret_cast<const uint4*>(s_slots + d * NS);
    const uint4 wb = *reinterpret_cast<const uint4*>(s_slots + (d + half_d) * NS);
    float2 a0 = unpack_half2(wa.x), a1 = unpack_half2(wa.y), a2 = unpack_half2(wa.z), a3 = unpack_half2(wa.w);
    float2 b0 = unpack_half2(wb.x), b1 = unpack_half2(wb.y), b2 = unpack_half2(wb.z), b3 = unpack_half2(wb.w);
    lg[0] += ra * a0.x + rb * b0.x; lg[1] += ra * a0.y + rb * b0.y;
    lg[2] += ra * a1.x + rb * b1.x; lg[3] += ra * a1.y + rb * b1.y;
    lg[4] += ra * a2.x + rb * b2.x; lg[5] += ra * a2.y + rb * b2.y;
    lg[6] += ra * a3.x + rb * b3.x; lg[7] += ra * a3.y + rb * b3.y;
  }
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) lg[s2] = warp_sum(lg[s2]);
  if (lane == 0) {
    float* o = logits + row * NS;
#pragma unroll
    for (int s2 = 0; s2 < NS; ++s2) o[s2] = __half2float(__float2half_rn(lg[s2]));       // fp16 matmul output
    stats[row * 2] = mean;
    stats[row * 2 + 1] = rstd;
  }
}

// ---- kernel 2: CTA per (frame, slice of PJ_PAIRS partner pairs); thread = (pair, row group) ----
constexpr int PJ_PAIRS = 64;                       // partner pairs per CTA (128 dims)
constexpr int PJ_GROUPS = PJ_THREADS / PJ_PAIRS;   // 4 row groups

__global__ void __launch_bounds__(PJ_THREADS)
slot_sum_kernel(const __half* __restrict__ feats, int ld, int frame_rows, int row_off, int n, int D,
                const __half* __restrict__ ln_w, const __half* __restrict__ ln_b, const float* __restrict__ cos_tab,
                const float* __restrict__ sin_tab, const float* __restrict__ logits, const float* __restrict__ stats,
                __half* __restrict__ res) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* s_p = reinterpret_cast<float*>(smem);                     // [n][8]  logits -> probabilities
  float* s_stat = s_p + n * NS;                                    // [n][2]
  float* s_red = s_stat + n * 2;                                   // [PJ_GROUPS][PJ_PAIRS][16]
  const int f = blockIdx.y, slice = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half_d = D / 2;
  const float* lgf = logits + static_cast<size_t>(f) * n * NS;
  for (int i = tid; i < n * NS / 4; i += PJ_THREADS) reinterpret_cast<float4*>(s_p)[i] = reinterpret_cast<const float4*>(lgf)[i];
  for (int i = tid; i < n * 2 / 2; i += PJ_THREADS)
    reinterpret_cast<float2*>(s_stat)[i] = reinterpret_cast<const float2*>(stats + static_cast<size_t>(f) * n * 2)[i];
  __syncthreads();
  // softmax over the patch axis, one warp per slot (recomputed by every slice CTA of the frame: 18 KB, negligible)
  if (warp < NS) {
    float mx = -INFINITY;
    for (int r = lane; r < n; r += 32) mx = fmaxf(mx, s_p[r * NS + warp]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int r = lane; r < n; r += 32) sm += __expf(s_p[r * NS + warp] - mx);
    sm = warp_sum(sm);
    const float inv = 1.f / sm;
    for (int r = lane; r < n; r += 32)
      s_p[r * NS + warp] = __half2float(__float2half_rn(__expf(s_p[r * NS + warp] - mx) * inv));  // softmax -> fp16
  }
  __syncthreads();
  const int pair = tid % PJ_PAIRS, grp = tid / PJ_PAIRS;
  const int d = slice * PJ_PAIRS + pair;
  float acc_a[NS], acc_b[NS];
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) acc_a[s2] = acc_b[s2] = 0.f;
  if (d < half_d) {
    const __half* base = feats + (static_cast<size_t>(f) * frame_rows + row_off) * ld;
    const float wa = __half2float(ln_w[d]), ba = __half2float(ln_b[d]);
    const float wb = __half2float(ln_w[d + half_d]), bb = __half2float(ln_b[d + half_d]);
    const int per = (n + PJ_GROUPS - 1) / PJ_GROUPS;
    const int r_end = min(n, (grp + 1) * per);
#pragma unroll 4
    for (int r = grp * per; r < r_end; ++r) {
      const __half* xr = base + static_cast<size_t>(r) * ld;
      const float mean = s_stat[r * 2], rstd = s_stat[r * 2 + 1];
      const float ya = __half2float(__float2half_rn((__half2float(xr[d]) - mean) * rstd * wa + ba));
      const float yb = __half2float(__float2half_rn((__half2float(xr[d + half_d]) - mean) * rstd * wb + bb));
      const float c = cos_tab[static_cast<size_t>(r) * half_d + d], sn = sin_tab[static_cast<size_t>(r) * half_d + d];
      const float ra = __half2float(__float2half_rn(ya * c - yb * sn));
      const float rb = __half2float(__float2half_rn(yb * c + ya * sn));
      const float4 p0 = *reinterpret_cast<const float4*>(s_p + r * NS);
      const float4 p1 = *reinterpret_cast<const float4*>(s_p + r * NS + 4);
      acc_a[0] += p0.x * ra; acc_a[1] += p0.y * ra; acc_a[2] += p0.z * ra; acc_a[3] += p0.w * ra;
      acc_a[4] += p1.x * ra; acc_a[5] += p1.y * ra; acc_a[6] += p1.z * ra; acc_a[7] += p1.w * ra;
      acc_b[0] += p0.x * rb; acc_b[1] += p0.y * rb; acc_b[2] += p0.z * rb; acc_b[3] += p0.w * rb;
      acc_b[4] += p1.x * rb; acc_b[5] += p1.y * rb; acc_b[6] += p1.z * rb; acc_b[7] += p1.w * rb;
    }
  }
  float* mine = s_red + (static_cast<size_t>(grp) * PJ_PAIRS + pair) * 16;
#pragma unroll
  for (int s2 = 0; s2 < NS; ++s2) { mine[s2] = acc_a[s2]; mine[NS + s2] = acc_b[s2]; }
  __syncthreads();
  // fixed-order combine of the row groups: thread = (pair, 4 of the 16 sums)
  if (d < half_d) {
#pragma unroll
    for (int q = 0; q < 16 / PJ_GROUPS; ++q) {
      const int k = grp * (16 / PJ_GROUPS) + q;                    // 0..15: slot s (k < 8: dim d) or slot s (k >= 8: dim d + D/2)
      float v = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < PJ_GROUPS; ++g2) v += s_red[(static_cast<size_t>(g2) * PJ_PAIRS + pair) * 16 + k];
      const int s2 = k & 7;
      res[(static_cast<size_t>(f) * NS + s2) * D + (k < NS ? d : d + half_d)] = __float2half_rn(v);
    }
  }
}

}  // namespace

int slot_pool_scratch_floats(int F, int n) { return F * n * (NS + 2); }

int slot_pool_fused(const __half* feats, int ld, int frame_rows, int row_off, int n, int D, int F,
                    const __half* ln_w, const __half* ln_b, float eps, const __half* slots, const float* cos_tab,
                    const float* sin_tab, float* scratch, __half* res, cudaStream_t stream) {
  TB_CHECK(D % 64 == 0, "projector width must be a multiple of 64");
  TB_CHECK(ld % 2 == 0 && n % 2 == 0, "feature pitch and patch count must be even");
  TB_CHECK(scratch != nullptr, "slot pool scratch missing");
  if (F == 0) return 0;
  float* logits = scratch;                                   // [F][n][8]
  float* stats = scratch + static_cast<size_t>(F) * n * NS;  // [F][n][2]
  const size_t smem1 = static_cast<size_t>(D) * NS * 2 + static_cast<size_t>(D) * 4;
  const size_t smem2 = static_cast<size_t>(n) * NS * 4 + static_cast<size_t>(n) * 8 + static_cast<size_t>(PJ_GROUPS) * PJ_PAIRS * 16 * 4;
  TB_CHECK(smem1 <= 200 * 1024 && smem2 <= 200 * 1024, "projector shared memory exceeds the SM budget");
  static size_t attr1 = 0, attr2 = 0;
  if (smem1 > attr1) {
    TB_CUDA(cudaFuncSetAttribute(slot_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem1)));
    attr1 = smem1;
  }
  if (smem2 > attr2) {
    TB_CUDA(cudaFuncSetAttribute(slot_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
    attr2 = smem2;
  }
  const int64_t rows = static_cast<int64_t>(F) * n;
  slot_logits_kernel<<<static_cast<unsigned>((rows + PJ_THREADS / 32 - 1) / (PJ_THREADS / 32)), PJ_THREADS, smem1, stream>>>(
      feats, ld, frame_rows, row_off, n, D, F, ln_w, ln_b, eps, slots, cos_tab, sin_tab, logits, stats);
  TB_LAUNCH_OK();
  slot_sum_kernel<<<dim3((D / 2 + PJ_PAIRS - 1) / PJ_PAIRS, F), PJ_THREADS, smem2, stream>>>(feats, ld, frame_rows, row_off, n, D, ln_w, ln_b,
                                                                                            cos_tab, sin_tab, logits, stats, res);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
