// Fused SpatialSlotPool (multimodal_projector/builder.py:427-467) up to the slot-weighted sums:
//   y   = LayerNorm(feat, eps 1e-6) ; y = y*cos + rotate_half(y)*sin   (RoPE over the PATCH index, full-width
//         rotate-half: dims d and d+D/2 are partners, :326-330, :451-455)
//   lg  = y @ slots[D,8] ; p = softmax(lg, over the n patches) ; res[s,:] = sum_n p[n,s] * y[n,:]
// The readout Linear (D -> 4096, :467) runs on the tcgen05 GEMM over res[F*8, D].
//
// HBM-bound: one CTA per frame streams the frame's n x D fp16 features twice (second pass hits L2):
// algorithmic bytes = n*D*2 read + 8*D*2 written per frame (1.18 MB + 16 KB at n=576, D=1024).
// Pass 1 (warp per patch row): LN statistics, rotated row, 8 slot logits -> shared memory.
// Softmax over patches per slot (one warp per slot).  Pass 2 (thread per partner-pair of dims):
// recompute the rotated row from the cached statistics and accumulate the 8 weighted sums in registers.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int PJ_THREADS = 256;
constexpr int NS = 8;

__global__ void __launch_bounds__(PJ_THREADS)
slot_pool_kernel(const __half* __restrict__ feats, int ld, int frame_rows, int row_off, int n, int D,
                 const __half* __restrict__ ln_w, const __half* __restrict__ ln_b, float eps,
                 const __half* __restrict__ slots, const float* __restrict__ cos_tab,
                 const float* __restrict__ sin_tab, __half* __restrict__ res) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* s_logit = reinterpret_cast<float*>(smem);                 // [n][8]  (becomes p)
  float* s_mean = s_logit + n * NS;                                // [n]
  float* s_rstd = s_mean + n;                                      // [n]
  __half* s_slots = reinterpret_cast<__half*>(s_rstd + n);         // [D][8]
  __half* s_w = s_slots + D * NS;                                  // [D]
  __half* s_b = s_w + D;                                           // [D]

  const int f = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* base = feats + (static_cast<size_t>(f) * frame_rows + row_off) * ld;
  const int half_d = D / 2;

  for (int i = tid; i < D * NS / 8; i += PJ_THREADS)
    reinterpret_cast<uint4*>(s_slots)[i] = reinterpret_cast<const uint4*>(slots)[i];
  for (int i = tid; i < D / 8; i += PJ_THREADS) {
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(ln_w)[i];
    reinterpret_cast<uint4*>(s_b)[i] = reinterpret_cast<const uint4*>(ln_b)[i];
  }
  __syncthreads();

  // ---------------- pass 1: one warp per patch row ----------------
  for (int r = warp; r < n; r += PJ_THREADS / 32) {
    const __half* xr = base + static_cast<size_t>(r) * ld;
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
    for (int s = 0; s < NS; ++s) lg[s] = 0.f;
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
    for (int s = 0; s < NS; ++s) lg[s] = warp_sum(lg[s]);
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) s_logit[r * NS + s] = __half2float(__float2half_rn(lg[s]));  // fp16 matmul output
      s_mean[r] = mean;
      s_rstd[r] = rstd;
    }
  }
  __syncthreads();

  // ---------------- softmax over the patch axis, one warp per slot ----------------
  if (warp < NS) {
    float mx = -INFINITY;
    for (int r = lane; r < n; r += 32) mx = fmaxf(mx, s_logit[r * NS + warp]);
    mx = warp_max(mx);
    float sm = 0.f;
    for (int r = lane; r < n; r += 32) sm += __expf(s_logit[r * NS + warp] - mx);
    sm = warp_sum(sm);
    const float inv = 1.f / sm;
    for (int r = lane; r < n; r += 32)
      s_logit[r * NS + warp] = __half2float(__float2half_rn(__expf(s_logit[r * NS + warp] - mx) * inv));  // softmax -> fp16
  }
  __syncthreads();

  // ---------------- pass 2: thread per partner pair, 8 running sums each ----------------
  for (int d = tid; d < half_d; d += PJ_THREADS) {
    float acc_a[NS], acc_b[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) acc_a[s] = acc_b[s] = 0.f;
    const float wa = __half2float(s_w[d]), ba = __half2float(s_b[d]);
    const float wb = __half2float(s_w[d + half_d]), bb = __half2float(s_b[d + half_d]);
#pragma unroll 4
    for (int r = 0; r < n; ++r) {
      const __half* xr = base + static_cast<size_t>(r) * ld;
      const float mean = s_mean[r], rstd = s_rstd[r];
      const float ya = __half2float(__float2half_rn((__half2float(xr[d]) - mean) * rstd * wa + ba));
      const float yb = __half2float(__float2half_rn((__half2float(xr[d + half_d]) - mean) * rstd * wb + bb));
      const float c = cos_tab[static_cast<size_t>(r) * half_d + d], sn = sin_tab[static_cast<size_t>(r) * half_d + d];
      const float ra = __half2float(__float2half_rn(ya * c - yb * sn));
      const float rb = __half2float(__float2half_rn(yb * c + ya * sn));
      const float4 p0 = *reinterpret_cast<const float4*>(s_logit + r * NS);
      const float4 p1 = *reinterpret_cast<const float4*>(s_logit + r * NS + 4);
      acc_a[0] += p0.x * ra; acc_a[1] += p0.y * ra; acc_a[2] += p0.z * ra; acc_a[3] += p0.w * ra;
      acc_a[4] += p1.x * ra; acc_a[5] += p1.y * ra; acc_a[6] += p1.z * ra; acc_a[7] += p1.w * ra;
      acc_b[0] += p0.x * rb; acc_b[1] += p0.y * rb; acc_b[2] += p0.z * rb; acc_b[3] += p0.w * rb;
      acc_b[4] += p1.x * rb; acc_b[5] += p1.y * rb; acc_b[6] += p1.z * rb; acc_b[7] += p1.w * rb;
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      __half* o = res + (static_cast<size_t>(f) * NS + s) * D;
      o[d] = __float2half_rn(acc_a[s]);
      o[d + half_d] = __float2half_rn(acc_b[s]);
    }
  }
}

}  // namespace

int slot_pool_fused(const __half* feats, int ld, int frame_rows, int row_off, int n, int D, int F,
                    const __half* ln_w, const __half* ln_b, float eps, const __half* slots, const float* cos_tab,
                    const float* sin_tab, __half* res, cudaStream_t stream) {
  TB_CHECK(D % 64 == 0, "projector width must be a multiple of 64");
  TB_CHECK(ld % 2 == 0 && n % 2 == 0, "feature pitch and patch count must be even");
  if (F == 0) return 0;
  const size_t smem = static_cast<size_t>(n) * NS * 4 + static_cast<size_t>(n) * 8 + static_cast<size_t>(D) * NS * 2 +
                      static_cast<size_t>(D) * 4;
  TB_CHECK(smem <= 200 * 1024, "projector shared memory exceeds the SM budget");
  static size_t attr = 0;
  if (smem > attr) {
    TB_CUDA(cudaFuncSetAttribute(slot_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  slot_pool_kernel<<<F, PJ_THREADS, smem, stream>>>(feats, ld, frame_rows, row_off, n, D, ln_w, ln_b, eps, slots,
                                                    cos_tab, sin_tab, res);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
