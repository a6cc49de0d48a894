// Prefill RoPE + paged KV-cache write (HF MistralAttention: apply_rotary_pos_emb on q,k; DynamicCache.update).
// Operates in place on the fused qkv activation [B*S, (heads + 2*kvh) * hd] produced by the QKV GEMM:
// q and k columns are rotated (rotate_half pairs d, d+hd/2; cos/sin tables hold the fp16-rounded values the
// reference casts to, [max_pos, hd/2]); rotated k and raw v rows are appended to the page pool
//   KV pool layout: ONE pool per layer [page][kv_head][K|V][page_size][hd] fp16, page = block_table[slot][pos / page_size];
//   the 16-byte chunks of a row are stored at chunk ^ (pos & 7).  A (page, kv head) tile is then one contiguous 32 KB block that
//   the fused decode step pulls into shared memory with a single cp.async.bulk, and both the mma.sync B-fragment reads of K
//   and the ldmatrix.trans reads of V are bank-conflict free without padding (kv_tile_offset / kv_chunk_swz in kernels.h).
// One CTA per token; HBM-bound: reads+writes the qkv row once, writes 2*kvh*hd halves to the pool.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

__global__ void __launch_bounds__(128)
rope_kv_kernel(__half* __restrict__ qkv, int S, int heads, int kvh, int hd, const float* __restrict__ cos_tab,
               const float* __restrict__ sin_tab, __half* __restrict__ kv_pool,
               const int32_t* __restrict__ block_table, const int32_t* __restrict__ slot_ids, int max_pages,
               int page_size) {
  const int tok = blockIdx.x;              // b*S + pos
  const int b = tok / S, pos = tok % S;
  const int hh = hd / 2;
  __half* row = qkv + static_cast<int64_t>(tok) * (heads + 2 * kvh) * hd;
  const int slot = slot_ids[b];
  const int page = block_table[static_cast<int64_t>(slot) * max_pages + pos / page_size];
  const int poff = pos % page_size;
  // q and k heads: rotation pairs (d, d + hd/2), eight consecutive d per thread (16-byte loads / stores of both halves, cos / sin as
  // float4 pairs); the arithmetic per element is unchanged: fp32 products, one rounding to fp16
  const int hh8 = hh / 8;
  const int nvec = (heads + kvh) * hh8;
  const float4* ct = reinterpret_cast<const float4*>(cos_tab + static_cast<int64_t>(pos) * hh);
  const float4* st = reinterpret_cast<const float4*>(sin_tab + static_cast<int64_t>(pos) * hh);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int h = i / hh8, d = (i % hh8) * 8;
    __half* p = row + h * hd;
    float a[8], c2[8], c[8], sn[8], ra[8], rb[8];
    unpack8(*reinterpret_cast<const uint4*>(p + d), a);
    unpack8(*reinterpret_cast<const uint4*>(p + d + hh), c2);
    *reinterpret_cast<float4*>(&c[0]) = ct[d / 4];
    *reinterpret_cast<float4*>(&c[4]) = ct[d / 4 + 1];
    *reinterpret_cast<float4*>(&sn[0]) = st[d / 4];
    *reinterpret_cast<float4*>(&sn[4]) = st[d / 4 + 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ra[j] = a[j] * c[j] - c2[j] * sn[j];
      rb[j] = c2[j] * c[j] + a[j] * sn[j];
    }
    const uint4 va = pack8(ra), vb = pack8(rb);
    *reinterpret_cast<uint4*>(p + d) = va;
    *reinterpret_cast<uint4*>(p + d + hh) = vb;
    if (h >= heads) {
      __half* kd = kv_pool + kv_tile_offset(page, kvh, h - heads, 0, page_size, hd) + static_cast<int64_t>(poff) * hd;
      *reinterpret_cast<uint4*>(kd + kv_chunk_swz(d, poff)) = va;
      *reinterpret_cast<uint4*>(kd + kv_chunk_swz(d + hh, poff)) = vb;
    }
  }
  const __half* vsrc = row + (heads + kvh) * hd;
  for (int i = threadIdx.x; i < kvh * hd / 8; i += blockDim.x) {
    const int h = (i * 8) / hd, d = (i * 8) % hd;
    __half* vd = kv_pool + kv_tile_offset(page, kvh, h, 1, page_size, hd) + static_cast<int64_t>(poff) * hd + kv_chunk_swz(d, poff);
    *reinterpret_cast<uint4*>(vd) = *reinterpret_cast<const uint4*>(vsrc + i * 8);
  }
}

}  // namespace

int rope_kv_prefill(__half* qkv, int B, int S, int heads, int kvh, int hd, const float* cos_tab, const float* sin_tab,
                    __half* kv_pool, const int32_t* block_table, const int32_t* slot_ids, int max_pages,
                    int page_size, cudaStream_t stream) {
  TB_CHECK(hd % 16 == 0, "head_dim must be a multiple of 16");
  if (B * S == 0) return 0;
  rope_kv_kernel<<<B * S, 128, 0, stream>>>(qkv, S, heads, kvh, hd, cos_tab, sin_tab, kv_pool, block_table, slot_ids, max_pages,
                                            page_size);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
