// Row-wise LayerNorm / RMSNorm (HBM/L2-bound, one CTA of 128 threads per row, 16-byte vector I/O).
//   LayerNorm : HF CLIP pre_layrnorm / layer_norm1 / layer_norm2 (eps 1e-5) -- fp32 statistics, one rounding.
//   RMSNorm   : HF MistralRMSNorm (modeling_mistral.py): x*rsqrt(mean(x^2)+eps) in fp32 -> fp16, then
//               weight * that in fp16 (two roundings, reproduced).
// Also the ViT embedding assembly + pre-LN (CLIPVisionEmbeddings + pre_layrnorm) and im2col for the
// 14x14/14 patch convolution (clip_encoder.py:50).
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int NT = 128;
constexpr int MAXV = 4;   // up to 4 x 8 halves per thread -> D <= 4096

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += red[i];
  return t;
}

// mode 0: LayerNorm(w,b)   mode 1: RMSNorm(w)
template <int MODE>
__device__ __forceinline__ void norm_row(float (&v)[MAXV][8], int nv, int D, const __half* w, const __half* b,
                                         float eps, __half* out, float* red) {
  const int tid = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv && tid + i * NT < D / 8)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += (MODE == 0) ? v[i][j] : v[i][j] * v[i][j];
  s = block_sum(s, red);
  float mean = 0.f, rstd;
  if (MODE == 0) {
    mean = s / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (i < nv && tid + i * NT < D / 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) { float d = v[i][j] - mean; q += d * d; }
    q = block_sum(q, red);
    rstd = rsqrtf(q / D + eps);
  } else {
    rstd = rsqrtf(s / D + eps);
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * NT;
    if (i < nv && c < D / 8) {
      float wf[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(w + c * 8), wf);
      if (MODE == 0) {
        float bf[8];
        unpack8(*reinterpret_cast<const uint4*>(b + c * 8), bf);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * wf[j] + bf[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = wf[j] * __half2float(__float2half_rn(v[i][j] * rstd));
      }
      *reinterpret_cast<uint4*>(out + c * 8) = pack8(o);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT) norm_kernel(const __half* x, int ldx, const __half* w, const __half* b,
                                                  float eps, __half* out, int ldo, int D) {
  __shared__ float red[NT / 32];
  const int row = blockIdx.x;
  const __half* xr = x + static_cast<size_t>(row) * ldx;
  float v[MAXV][8];
  const int nv = (D / 8 + NT - 1) / NT;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * NT;
    if (i < nv && c < D / 8) unpack8(*reinterpret_cast<const uint4*>(xr + c * 8), v[i]);
  }
  norm_row<MODE>(v, nv, D, w, b, eps, out + static_cast<size_t>(row) * ldo, red);
}

// One WARP per row (D = NV * 256, the row stays packed in NV uint4 registers per lane): no block barriers, 8 rows per CTA.
// The one-CTA-per-row kernel above is latency-bound (four __syncthreads per 2 KB row: 28 us for the 18 464 rows of a ViT
// chunk whose 76 MB of traffic is an L2 hit); this one is bound by the L2 / HBM stream.
template <int MODE, int NV>
__global__ void __launch_bounds__(256) norm_warp_kernel(const __half* __restrict__ x, int ldx, const __half* __restrict__ w,
                                                        const __half* __restrict__ b, float eps, __half* __restrict__ out, int ldo, int rows) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int D = NV * 256;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx);
  uint4 u[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) u[i] = xr[lane + 32 * i];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(u[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += (MODE == 0) ? f[j] : f[j] * f[j];
  }
  s = warp_sum(s);
  float mean = 0.f, rstd;
  if (MODE == 0) {
    mean = s / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; q += d * d; }
    }
    q = warp_sum(q);
    rstd = rsqrtf(q / D + eps);
  } else {
    rstd = rsqrtf(s / D + eps);
  }
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldo);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    float f[8], wf[8], o[8];
    unpack8(u[i], f);
    unpack8(reinterpret_cast<const uint4*>(w)[c], wf);
    if (MODE == 0) {
      float bf[8];
      unpack8(reinterpret_cast<const uint4*>(b)[c], bf);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[j] - mean) * rstd * wf[j] + bf[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = wf[j] * __half2float(__float2half_rn(f[j] * rstd));
    }
    orow[c] = pack8(o);
  }
}

template <int MODE>
bool launch_norm_warp(const __half* x, int ldx, const __half* w, const __half* b, float eps, __half* out, int ldo, int rows, int D,
                      cudaStream_t stream) {
  // 8 rows (warps) per CTA; a prefill has ~2 k rows, i.e. 1.7 CTAs per SM at 8 -- 2 rows per CTA keep every SM's schedulers fed
  const int wpc = rows >= 16384 ? 8 : 2;
  const int grid = (rows + wpc - 1) / wpc;
  switch (D) {
    case 256: norm_warp_kernel<MODE, 1><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    case 512: norm_warp_kernel<MODE, 2><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    case 768: norm_warp_kernel<MODE, 3><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    case 1024: norm_warp_kernel<MODE, 4><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    case 1280: norm_warp_kernel<MODE, 5><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    case 2048: norm_warp_kernel<MODE, 8><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    case 4096: norm_warp_kernel<MODE, 16><<<grid, 32 * wpc, 0, stream>>>(x, ldx, w, b, eps, out, ldo, rows); return true;
    default: return false;
  }
}

__global__ void __launch_bounds__(NT) vit_embed_ln_kernel(const __half* patch, const __half* cls, const __half* pos,
                                                          const __half* ln_w, const __half* ln_b, float eps,
                                                          __half* out, int P, int D) {
  __shared__ float red[NT / 32];
  const int row = blockIdx.x;             // f*(P+1) + t
  const int f = row / (P + 1), t = row % (P + 1);
  const __half* src = (t == 0) ? cls : patch + (static_cast<size_t>(f) * P + (t - 1)) * D;
  const __half* pr = pos + static_cast<size_t>(t) * D;
  float v[MAXV][8];
  const int nv = (D / 8 + NT - 1) / NT;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * NT;
    if (i < nv && c < D / 8) {
      float a[8], p[8];
      unpack8(*reinterpret_cast<const uint4*>(src + c * 8), a);
      unpack8(*reinterpret_cast<const uint4*>(pr + c * 8), p);
      // embeddings = cat(cls, patches) + position_embedding : an fp16 add in the reference -> round once here
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = __half2float(__float2half_rn(a[j] + p[j]));
    }
  }
  norm_row<0>(v, nv, D, ln_w, ln_b, eps, out + static_cast<size_t>(row) * D, red);
}

__global__ void vit_im2col_kernel(const __half* __restrict__ px, __half* __restrict__ out, int image, int ps, int Kp,
                                  int P, int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int col = static_cast<int>(i % Kp);
  const int64_t r = i / Kp;
  const int p = static_cast<int>(r % P);
  const int64_t f = r / P;
  const int side = image / ps;
  const int py = p / side, pxx = p % side;
  __half v = __float2half(0.f);
  if (col < 3 * ps * ps) {
    const int c = col / (ps * ps), rem = col % (ps * ps), ky = rem / ps, kx = rem % ps;
    v = px[((f * 3 + c) * image + (py * ps + ky)) * image + pxx * ps + kx];
  }
  out[i] = v;
}

}  // namespace

int layernorm_f16(const __half* x, int ldx, const __half* w, const __half* b, float eps, __half* out, int ldo,
                  int rows, int D, cudaStream_t stream) {
  TB_CHECK(D % 8 == 0 && D <= NT * 8 * MAXV, "LayerNorm width must be a multiple of 8 and <= 4096");
  TB_CHECK(ldx % 8 == 0 && ldo % 8 == 0, "row pitch must be 16-byte aligned");
  if (rows == 0) return 0;
  if (!launch_norm_warp<0>(x, ldx, w, b, eps, out, ldo, rows, D, stream))
    norm_kernel<0><<<rows, NT, 0, stream>>>(x, ldx, w, b, eps, out, ldo, D);
  TB_LAUNCH_OK();
  return 0;
}

int rmsnorm_f16(const __half* x, int ldx, const __half* w, float eps, __half* out, int ldo, int rows, int D,
                cudaStream_t stream) {
  TB_CHECK(D % 8 == 0 && D <= NT * 8 * MAXV, "RMSNorm width must be a multiple of 8 and <= 4096");
  TB_CHECK(ldx % 8 == 0 && ldo % 8 == 0, "row pitch must be 16-byte aligned");
  if (rows == 0) return 0;
  if (!launch_norm_warp<1>(x, ldx, w, nullptr, eps, out, ldo, rows, D, stream))
    norm_kernel<1><<<rows, NT, 0, stream>>>(x, ldx, w, nullptr, eps, out, ldo, D);
  TB_LAUNCH_OK();
  return 0;
}

int vit_im2col(const __half* pixels, __half* out, int F, int image, int ps, int Kp, cudaStream_t stream) {
  TB_CHECK(image % ps == 0 && Kp >= 3 * ps * ps && Kp % 8 == 0, "bad patch geometry");
  const int P = (image / ps) * (image / ps);
  const int64_t total = static_cast<int64_t>(F) * P * Kp;
  if (total == 0) return 0;
  vit_im2col_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, stream>>>(pixels, out, image, ps, Kp, P,
                                                                                       total);
  TB_LAUNCH_OK();
  return 0;
}

int vit_embed_ln(const __half* patch, const __half* cls, const __half* pos, const __half* ln_w, const __half* ln_b,
                 float eps, __half* out, int F, int P, int D, cudaStream_t stream) {
  TB_CHECK(D % 8 == 0 && D <= NT * 8 * MAXV, "ViT width must be a multiple of 8 and <= 4096");
  if (F == 0) return 0;
  vit_embed_ln_kernel<<<F * (P + 1), NT, 0, stream>>>(patch, cls, pos, ln_w, ln_b, eps, out, P, D);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
