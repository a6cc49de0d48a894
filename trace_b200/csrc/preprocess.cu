// Frame pre-processing on the device: what process_video does on the host per frame after decoding
// (trace/mm_utils.py:424-462) -- expand2square with the CLIP mean colour (:259-270), CLIPImageProcessor.preprocess
// (PIL bicubic shortest-edge resize with antialiasing, centre crop, 1/255 rescale, mean/std normalise) and the driver's
// fp16 cast (evaluate.py:316) -- as two byte-exact integer passes + one normalisation:
//
//   pass H : uint8 [F,H,W,3] (RGB, as decoded) -> uint8 [F,rows,OW,3]   Pillow's horizontal resample, fixed point 22 bit
//   pass V : uint8 [F,rows,OW,3]               -> fp16  [F,3,OH,OW]      Pillow's vertical resample + normalise
//
// Pillow resamples 8-bit images with integer coefficients (ImagingResampleHorizontal_8bpc / Vertical_8bpc:
// ss = 1 << 21; ss += pixel * k; out = clip8(ss >> 22)) and rounds to uint8 BETWEEN the two passes; the coefficient and
// bound tables are produced on the host in float64 exactly like Pillow's precompute_coeffs / normalize_coeffs_8bpc
// (trace_b200/preprocess.py), so the integer stage is bit-exact against PIL and the fp16 output differs from the
// reference only where the reference's own float32 arithmetic does.
// The padded square is never materialised: a source pixel outside the pasted image is the pad colour.
// HBM-bound byte work: 6.2 MB read + 0.7 MB written per 1080p frame.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Pillow Resample.c

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// one thread per (frame, square row, out x); 3 channels per thread
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, int H, int W, int S, int pad_x, int pad_y, int row0, int rows, int OW,
                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ coef, int ksize, int pad_r, int pad_g, int pad_b,
                  uint8_t* __restrict__ tmp) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int yr = blockIdx.y;               // row inside [row0, row0 + rows) of the square
  const int f = blockIdx.z;
  if (ox >= OW) return;
  const int y = row0 + yr - pad_y;         // row in the decoded frame
  const int xmin = bounds[2 * ox], n = bounds[2 * ox + 1];
  const int32_t* k = coef + static_cast<size_t>(ox) * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  const bool row_in = y >= 0 && y < H;
  const uint8_t* line = src + (static_cast<size_t>(f) * H + (row_in ? y : 0)) * W * 3;
  for (int j = 0; j < n; ++j) {
    const int x = xmin + j - pad_x;
    int r = pad_r, g = pad_g, b = pad_b;
    if (row_in && x >= 0 && x < W) { r = line[x * 3]; g = line[x * 3 + 1]; b = line[x * 3 + 2]; }
    const int kk = k[j];
    s0 += r * kk; s1 += g * kk; s2 += b * kk;
  }
  uint8_t* o = tmp + ((static_cast<size_t>(f) * rows + yr) * OW + ox) * 3;
  o[0] = static_cast<uint8_t>(clip8(s0));
  o[1] = static_cast<uint8_t>(clip8(s1));
  o[2] = static_cast<uint8_t>(clip8(s2));
}

// one thread per (frame, out y, out x); writes the three channel planes (NCHW)
__global__ void __launch_bounds__(256)
resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int rows, int row0, int OW, int OH, int crop_x, int crop_y, int C,
                       const int32_t* __restrict__ bounds, const int32_t* __restrict__ coef, int ksize, float m0, float m1, float m2,
                       float d0, float d1, float d2, __half* __restrict__ out) {
  const int cx = blockIdx.x * blockDim.x + threadIdx.x;   // x inside the crop
  const int cy = blockIdx.y;
  const int f = blockIdx.z;
  if (cx >= C) return;
  const int ox = cx + crop_x, oy = cy + crop_y;
  const int ymin = bounds[2 * oy], n = bounds[2 * oy + 1];
  const int32_t* k = coef + static_cast<size_t>(oy) * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  const uint8_t* col = tmp + ((static_cast<size_t>(f) * rows + (ymin - row0)) * OW + ox) * 3;
  for (int j = 0; j < n; ++j) {
    const uint8_t* px = col + static_cast<size_t>(j) * OW * 3;
    const int kk = k[j];
    s0 += px[0] * kk; s1 += px[1] * kk; s2 += px[2] * kk;
  }
  // transformers rescale(): float32(float64(v) * (1/255)); normalize(): (x - mean) / std in float32; driver: .half()
  const double inv = 1.0 / 255.0;
  const float r = static_cast<float>(static_cast<double>(clip8(s0)) * inv);
  const float g = static_cast<float>(static_cast<double>(clip8(s1)) * inv);
  const float b = static_cast<float>(static_cast<double>(clip8(s2)) * inv);
  const size_t plane = static_cast<size_t>(C) * C;
  __half* o = out + static_cast<size_t>(f) * 3 * plane + static_cast<size_t>(cy) * C + cx;
  o[0] = __float2half_rn(__fdiv_rn(r - m0, d0));
  o[plane] = __float2half_rn(__fdiv_rn(g - m1, d1));
  o[2 * plane] = __float2half_rn(__fdiv_rn(b - m2, d2));
}

}  // namespace

int preprocess_frames(const uint8_t* frames, int F, int H, int W, int S, int pad_x, int pad_y, int OW, int OH, int crop, int crop_x,
                      int crop_y, int row0, int rows, const int32_t* hbounds, const int32_t* hcoef, int hk, const int32_t* vbounds,
                      const int32_t* vcoef, int vk, const float* mean, const float* stdv, const int* pad_rgb, uint8_t* tmp, __half* out,
                      cudaStream_t stream) {
  TB_CHECK(F > 0 && H > 0 && W > 0 && OW > 0 && OH > 0 && crop > 0 && rows > 0, "bad pre-processing geometry");
  TB_CHECK(crop_x >= 0 && crop_y >= 0 && crop_x + crop <= OW && crop_y + crop <= OH, "crop outside the resized image");
  TB_CHECK(F <= 65535 && rows <= 65535 && crop <= 65535, "grid dimension overflow");
  resample_h_kernel<<<dim3(ceil_div(OW, 256), rows, F), 256, 0, stream>>>(frames, H, W, S, pad_x, pad_y, row0, rows, OW, hbounds, hcoef, hk,
                                                                          pad_rgb[0], pad_rgb[1], pad_rgb[2], tmp);
  TB_LAUNCH_OK();
  resample_v_norm_kernel<<<dim3(ceil_div(crop, 256), crop, F), 256, 0, stream>>>(tmp, rows, row0, OW, OH, crop_x, crop_y, crop, vbounds, vcoef,
                                                                                 vk, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2], out);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
