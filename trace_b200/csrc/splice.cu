// Prefill activation assembly (trace_arch.py:394-426): one CTA per output row of inputs_embeds [B, S, H],
// S = L - 1 + 14*T.  The single <video> placeholder (-201 / -200) expands to 14 rows per frame:
// 8 slot rows (projector output) then 6 time rows (TimeTower table, ids from format(t,'0>6.1f'),
// trace_arch.py:237-258); -205 rows take the SyncTower row; everything else is embed_tokens[clamp(id,0)].
// Pure gather: H*2 bytes read + written per row.
#include "common.cuh"
#include "kernels.h"

namespace tb {
namespace {

__global__ void __launch_bounds__(128)
splice_kernel(const int64_t* __restrict__ ids, int L, int T, int vpos_hint, const __half* __restrict__ slots,
              const int32_t* __restrict__ time_ids, const __half* __restrict__ embed,
              const __half* __restrict__ time_tab, const __half* __restrict__ sync_tab, int H,
              __half* __restrict__ out, int S) {
  const int b = blockIdx.y, pos = blockIdx.x;
  const int64_t* row_ids = ids + static_cast<int64_t>(b) * L;
  __shared__ int s_vpos;
  if (threadIdx.x == 0) {
    int v = -1;
    for (int i = 0; i < L; ++i)
      if (row_ids[i] == -201 || row_ids[i] == -200) { v = i; break; }
    s_vpos = v < 0 ? 0x7fffffff : v;   // no placeholder (text-only prompt, T == 0): every row is a plain gather
  }
  __syncthreads();
  const int vpos = s_vpos;
  const int nv = 14 * T;
  const __half* src;
  if (pos < vpos) {
    const int64_t id = row_ids[pos];
    src = (id == -205) ? sync_tab : embed + (id < 0 ? 0 : id) * H;
  } else if (pos < vpos + nv) {
    const int j = pos - vpos, f = j / 14, r = j % 14;
    if (r < 8) src = slots + ((static_cast<int64_t>(b) * T + f) * 8 + r) * H;
    else src = time_tab + static_cast<int64_t>(time_ids[(static_cast<int64_t>(b) * T + f) * 6 + (r - 8)]) * H;
  } else {
    const int64_t id = row_ids[pos - nv + 1];
    src = (id == -205) ? sync_tab : embed + (id < 0 ? 0 : id) * H;
  }
  __half* dst = out + (static_cast<int64_t>(b) * S + pos) * H;
  for (int i = threadIdx.x; i < H / 8; i += 128)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

}  // namespace

int splice_prefill(const int64_t* ids, int B, int L, int T, const __half* slots, const int32_t* time_ids,
                   const __half* embed, const __half* time_tab, const __half* sync_tab, int vocab, int H,
                   __half* out, cudaStream_t stream) {
  (void)vocab;
  TB_CHECK(H % 8 == 0, "hidden size must be a multiple of 8");
  const int S = T == 0 ? L : L - 1 + 14 * T;   // T == 0: text-only prompt, embed_tokens(inputs) (trace_mistral.py:306-307)
  if (B == 0) return 0;
  splice_kernel<<<dim3(S, B), 128, 0, stream>>>(ids, L, T, 0, slots, time_ids, embed, time_tab, sync_tab, H, out, S);
  TB_LAUNCH_OK();
  return 0;
}

}  // namespace tb
