// The decode step as ONE persistent cooperative kernel whose HBM traffic is a single uninterrupted stream.
//
//   per layer : RMSNorm+QKV -> RoPE / KV append / split-KV attention -> O(+residual) -> RMSNorm+gate/up(SwiGLU) -> down(+residual)
//   then      : final RMSNorm + active-head logits + masked argmax (or temperature / top-k sampling), and (CTA 0) the token
//               bookkeeping, head switch and next-token embedding dispatch.
//
// One CTA per SM: 8 consumer warps + 1 producer warp.  The producer streams EVERYTHING the step reads from HBM -- the
// CTA's weight rows of every phase and the K/V pages of its attention work item -- through one mbarrier ring with
// cp.async.bulk, in program order, and never waits for an activation; a K/V page tile of one kv head is a single
// contiguous 32 KB block in the cache (kernels.h: kv_tile_offset) and takes one ring stage, exactly like 16 weight rows.
// The consumers run mma.sync on whatever stage is at the ring front: m16n8k16 with the batch in the N slot for the
// projections, and for attention S = Q K^T (the 4 query heads of the kv head in the M slot) and O += P V (m16n8k8,
// V fragments with ldmatrix.trans) on 8 keys per warp, online softmax in registers.
//
// Phases exchange activations WITHOUT grid barriers: every value another CTA needs is written as an 8-byte unit
// {2 halves | 1 float, epoch} with one store, and readers poll the units they need until the epoch matches (the
// flag travels with the data, so there is no fence, no counter, and no second round trip to fetch the data after
// a barrier opened; a CTA only ever waits for the producers of the values it actually reads -- the attention CTAs
// of kv head g wait for q/k/v of g, not for the grid).  Buffer reuse is ordered by the data dependencies themselves
// (DESIGN.md §4).  Every wait is bounded by a watchdog (globaltimer): a protocol error traps instead of hanging.
//
// Replaces: trace_mistral.py:114-264 forward (decoder layers via HF MistralModel, heads :190-200, mask :244-252),
// :336-344 head switch, trace_arch.py:345-375 embedding dispatch, HF greedy / sampling step -- for one generated token.
#include <vector>

#include "common.cuh"
#include "decode_dev.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int DS_G = 4;                 // query heads per kv head
constexpr int DS_HD = 128;
constexpr int DS_PAGE = 64;             // == TRACE_B200_PAGE_SIZE
constexpr int DS_MAX_LAYERS = 48;
constexpr int DS_STAGE_HALVES = GV_TILE * (1024 + WS_PAD);     // 33792 B: 16 weight rows x 1024 columns, or one K|V page tile
constexpr uint32_t DS_KV_TILE_BYTES = 2 * DS_PAGE * DS_HD * 2; // 32768
constexpr unsigned DS_EPOCH_STRIDE = 512;                      // epochs reserved per launch (8 per layer + head)
constexpr int DS_RES_PITCH = 40;                               // residual rows kept per CTA (H/148 = 27.7, + the bandwidth-share skew)
// Block = 2 consumer warpgroups (8 warps) + 1 producer warpgroup (warp 8 streams, warps 9..11 only exist so that
// setmaxnreg -- a warpgroup-wide instruction -- can move registers: 384 x 168 at launch -> consumers 224, producer group 56).
// With 219 KB of the SM's 256 KB configured as shared memory the L1 is a few KB: a spilled register is an L2 round trip.
constexpr int DS_THREADS = GV_CONS + 128;
constexpr int DS_REGS_CONSUMER = 224;
constexpr int DS_REGS_PRODUCER = 56;
// words between the epoch flags of two producer CTAs (dense exchange, batch >= 4).  With the flags packed (stride 1) every CTA polls the
// same five cache lines -- 148 x 148 acquire loads per round into a handful of L2 slices; one flag per 128-byte line spreads them.
#ifndef DS_FLAG_STRIDE
#define DS_FLAG_STRIDE 32
#endif

struct DsLayer {
  const __half *ln1, *qkv_w, *o_w, *ln2, *gateup_w, *down_w;
  __half* kv_pool;
  void* pad;
};

struct DsParams {
  int B, H, qkvw, qd, ffn, layers, heads, kvh, vocab, head_rows, n_valid;
  float eps;
  const DsLayer* L;
  const __half *final_norm, *head_w;
  const __half* x;                      // dense [B,H]: this step's input embedding (prefill / previous step's finalize)
  __half* x_next;                       // dense [B,H]: next step's input embedding (same buffer as x)
  uint2 *x_ll, *qkv_ll, *attn_ll, *act_ll, *part_o_ll;
  uint4 *part_ml_ll, *cand_ll;
  // batch >= 4: the large exchanges are plain fp16 buffers + one epoch flag per producer CTA (fx / fq / fa / fact: [gridDim] each)
  __half *xd, *qkvd, *attnd, *actd;
  unsigned *fx, *fq, *fa, *fact;
  const float *cos_tab, *sin_tab;
  const int32_t* block_table;
  int max_pages;
  int nsplit, ncomb;
  float* logits_out;
  float* samp_logits;
  const __half *embed, *time_tab, *score_tab, *sync_tab;
  DecodeStreamState st;
  unsigned* epoch_ctr;
  int* err;
  int kp[5], ks[5];
  int part_bytes, stages, xs_cap, group_tiles;
  int advance;
  int watchdog_ms;                      // every wait in the kernel traps after this long (a protocol error must not hang the GPU)
  int poll_sleep_ns;                    // back-off between re-polls of an exchange unit that has not arrived (0 = none)
  unsigned long long* dbg;
};

struct DsSmem {
  __half* ws;
  __half* xs;
  float* s_part;
  float* s_red;
  uint64_t* full;
  uint64_t* empty;
};

// ---- watchdog --------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Watch {
  unsigned long long t0, limit_ns;
  int* err;
  unsigned sleep_ns;
};
__device__ __noinline__ void ds_die(int* err, int code) {
  atomicExch(err, code * 1000 + static_cast<int>(blockIdx.x));
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void ds_watch(const Watch& w, unsigned& spins, int code) {
  if ((++spins & 1023u) == 0 && gtimer() - w.t0 > w.limit_ns) ds_die(w.err, code);
}
// between two polls of an exchange unit that has not arrived yet
__device__ __forceinline__ void ds_repoll(const Watch& w, unsigned& spins, int code) {
  ds_watch(w, spins, code);
  if (w.sleep_ns) __nanosleep(w.sleep_ns);
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void ds_mbar_wait(uint64_t* bar, uint32_t parity, const Watch& w, int code) {
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) ds_watch(w, spins, code);
}

// ---- flag-in-data units ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ll_ld16(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void ll_st8(void* p, uint32_t d, uint32_t f) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(d), "r"(f) : "memory");
}
__device__ __forceinline__ void ll_st16(void* p, uint32_t d0, uint32_t d1, uint32_t f) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%2};" ::"l"(p), "r"(d0), "r"(f), "r"(d1) : "memory");
}
__device__ __forceinline__ uint4 ll_wait16(const void* p, unsigned epoch, const Watch& w, int code) {
  uint4 v = ll_ld16(p);
  unsigned spins = 0;
  while (v.y != epoch || v.w != epoch) {
    ds_repoll(w, spins, code);
    v = ll_ld16(p);
  }
  return v;
}
__device__ __forceinline__ uint32_t half_bits(float v) { return static_cast<uint32_t>(__half_as_ushort(__float2half_rn(v))); }

// Rows are dealt to the CTAs in units of `unit` (2: output pairs share one exchange unit; 4: two SwiGLU pairs do): every CTA
// streams the same number of weight bytes (+- one unit) whatever N is -- no wave quantisation.
// (Tried and dropped, profiles/r02_decode_timeline.md: rows proportional to a measured per-SM bandwidth share, open loop from a
// calibration kernel and closed loop from the kernel's own per-SM streaming times.  The systematic part of the per-CTA skew
// is +-2 %, the rest changes with the row ranges themselves; neither variant moved the step time.)
__device__ __forceinline__ void rows_of_cta(int N, int unit, int c, int grid, int& r0, int& r1) {
  const long long U = N / unit;
  r0 = static_cast<int>(U * c / grid) * unit;
  r1 = (c == grid - 1) ? N : static_cast<int>(U * (c + 1) / grid) * unit;
}

// timeline (TRACE_B200_MEGA_TIMING=1): [2 CTAs][layers][32] stamps of CTAs 0 and 77, then [gridDim][32] stamps of EVERY CTA for
// the middle layer (slot 31 = %smid): how unevenly the SMs finish a phase
__device__ __forceinline__ void stamp(const DsParams& p, int l, int k) {
  if (p.dbg != nullptr && threadIdx.x == 0) {
    const unsigned long long t = gtimer();
    if (blockIdx.x == 0 || blockIdx.x == 77) p.dbg[(static_cast<size_t>(blockIdx.x == 0 ? 0 : 1) * p.layers + l) * 32 + k] = t;
    if (l == p.layers / 2) {
      unsigned long long* q = p.dbg + static_cast<size_t>(2) * p.layers * 32 + static_cast<size_t>(blockIdx.x) * 32;
      q[k] = t;
      if (k == 10) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        q[31] = smid;
      }
    }
  }
}

// ---- activation staging ----------------------------------------------------------------------------------------
// ---- dense exchange (batch >= 4) -----------------------------------------------------------------------------------
// At batch 4..8 an activation vector is 32..460 KB per CTA per phase: doubling it with in-band flags and polling it unit by
// unit costs more than it saves.  There the producers write plain fp16 and then release ONE epoch flag per CTA; consumers
// acquire the flags of the CTAs they depend on (one poll per thread), then read the data with wide loads, all in flight.
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// all 256 consumer threads, after the CTA's stores of the exchange (bar.sync + release: cumulative over the CTA's writes)
__device__ __forceinline__ void publish(unsigned* flags, int vid, unsigned epoch) {
  cons_sync();
  if (threadIdx.x == 0) st_release_u32(flags + vid * DS_FLAG_STRIDE, epoch);
}
template <typename Pred>
__device__ __forceinline__ void wait_flags(const unsigned* flags, unsigned epoch, int ncta, Pred need, const Watch& w, int code) {
  if (static_cast<int>(threadIdx.x) < ncta && need(static_cast<int>(threadIdx.x))) {
    unsigned spins = 0;
    while (ld_acquire_u32(flags + threadIdx.x * DS_FLAG_STRIDE) != epoch) ds_repoll(w, spins, code);
  }
  cons_sync();
}
// RMSNorm weights of this thread's element groups: elements 4*(c*256 + tid) .. +3 for the row chunks c = 0..H/1024-1.
// A weight: fetched BEFORE the activations are polled, so its HBM latency hides behind the wait.
struct Gamma4 { uint2 g[4]; };
__device__ __forceinline__ Gamma4 load_gamma4(const __half* gamma, int H) {
  Gamma4 r;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int i = c * GV_CONS + threadIdx.x;
    r.g[c] = (4 * i < H) ? *reinterpret_cast<const uint2*>(gamma + 4 * i) : make_uint2(0, 0);
  }
  return r;
}

// HF MistralRMSNorm (fp32 mean of squares over the row, x*rstd rounded to fp16, then weight * that in fp16) fused into
// the staging of the whole hidden row (H % 1024 == 0, H <= 4096).  Source: the exchange buffer (src_ll) or, for the
// first phase of a launch, the dense input embedding.
template <int NLD>
__device__ __forceinline__ void stage_rms(__half* xs, int pitch, const uint2* src_ll, const __half* dense, int B, int H, const Gamma4& gm,
                                          float eps, float* s_red, unsigned epoch, const Watch& w) {
  const int cpr = H >> 10;                       // 256-thread chunks per row
  const int nchunk = B * cpr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* s4 = reinterpret_cast<const uint4*>(src_ll);
  for (int c0 = 0; c0 < nchunk; c0 += NLD) {
    uint4 v[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int c = c0 + i;
      if (c < nchunk) {
        const int b = c / cpr, j = (c - b * cpr) * GV_CONS + threadIdx.x;
        if (src_ll != nullptr) {
          v[i] = ll_ld16(s4 + static_cast<size_t>(b) * (H >> 2) + j);
        } else {
          const uint2 d = __ldcg(reinterpret_cast<const uint2*>(dense + static_cast<size_t>(b) * H) + j);
          v[i] = make_uint4(d.x, epoch, d.y, epoch);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int c = c0 + i;
      if (c < nchunk) {
        const int b = c / cpr, j = (c - b * cpr) * GV_CONS + threadIdx.x;
        unsigned spins = 0;
        while (v[i].y != epoch || v[i].w != epoch) {
          ds_repoll(w, spins, 22);
          v[i] = ll_ld16(s4 + static_cast<size_t>(b) * (H >> 2) + j);
        }
        *reinterpret_cast<uint2*>(xs + static_cast<size_t>(b) * pitch + 4 * j) = make_uint2(v[i].x, v[i].z);
        const float2 a = unpack_half2(v[i].x), d = unpack_half2(v[i].z);
        const float ss = warp_sum(a.x * a.x + a.y * a.y + d.x * d.x + d.y * d.y);
        if (lane == 0) s_red[c * GV_WARPS + warp] = ss;
      }
    }
  }
  cons_sync();
  for (int b = 0; b < B; ++b) {
    float t = 0.f;
    for (int k = 0; k < cpr * GV_WARPS; ++k) t += s_red[b * cpr * GV_WARPS + k];     // same order in every thread: deterministic
    const float rstd = rsqrtf(t / H + eps);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (cc < cpr) {
        uint2* px = reinterpret_cast<uint2*>(xs + static_cast<size_t>(b) * pitch + 4 * (cc * GV_CONS + threadIdx.x));
        uint2 u = *px;
        const float2 x0 = unpack_half2(u.x), x1 = unpack_half2(u.y), g0 = unpack_half2(gm.g[cc].x), g1 = unpack_half2(gm.g[cc].y);
        u.x = pack_half2(g0.x * __half2float(__float2half_rn(x0.x * rstd)), g0.y * __half2float(__float2half_rn(x0.y * rstd)));
        u.y = pack_half2(g1.x * __half2float(__float2half_rn(x1.x * rstd)), g1.y * __half2float(__float2half_rn(x1.y * rstd)));
        *px = u;
      }
    }
  }
}

// ---- weight stream: producer -----------------------------------------------------------------------------------
// rows [vlo,vhi) of the tiles of [r0,r1) in (k-phase, tile, k-stage) order, one bulk copy per row per stage
__device__ __forceinline__ void produce_matrix(const __half* __restrict__ W, int K, int KP, int KS, int r0, int r1, int vlo, int vhi,
                                               const DsSmem& m, int stages, int& s, uint32_t& ph, const Watch& w, bool k_outer) {
  const int lane = threadIdx.x & 31;
  const int pitch = KS + WS_PAD;
  auto put = [&](int tile, int kcol) {
    const int tr0 = r0 + tile * GV_TILE;
    const int lo = max(tr0, vlo), hi = min(tr0 + GV_TILE, vhi);
    ds_mbar_wait(&m.empty[s], ph ^ 1, w, 31);
    if (lane == 0) mbar_expect_tx(&m.full[s], static_cast<uint32_t>(hi - lo) * KS * 2);
    __syncwarp();
    const int row = tr0 + lane;
    if (lane < GV_TILE && row >= lo && row < hi)
      bulk_g2s(m.ws + static_cast<size_t>(s) * DS_STAGE_HALVES + static_cast<size_t>(lane) * pitch, W + static_cast<size_t>(row) * K + kcol, KS * 2,
               &m.full[s]);
    if (++s == stages) { s = 0; ph ^= 1; }
  };
  if (k_outer) {
    // (k-stage, tile) order: the consumers multiply a 1024-column chunk of the input for all their tiles as soon as that
    // chunk exists (consume_korder), so a late producer of the previous phase only delays the chunk it produces
    const int ntile = (r1 - r0 + GV_TILE - 1) / GV_TILE;
    for (int kcol = 0; kcol < K; kcol += KS)
      for (int tile = 0; tile < ntile; ++tile) put(tile, kcol);
    return;
  }
  GvItems it;
  it.init(K, KP, KS, r0, r1, vlo, vhi);
  while (!it.done) {
    put(it.tile, it.k0 + it.ks);
    it.next();
  }
}

// ---- weight stream: consumers ----------------------------------------------------------------------------------
template <bool FULL>
__device__ __forceinline__ void ds_mma_item(float (&c)[4], const uint32_t (&w)[32], const __half* xp, bool xok, int nch) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (FULL || ch < nch) {
      const uint4 xv = xok ? *reinterpret_cast<const uint4*>(xp + ch * 32) : make_uint4(0, 0, 0, 0);
      mma_16816(c, w[ch * 8 + 0], w[ch * 8 + 4], w[ch * 8 + 1], w[ch * 8 + 5], xv.x, xv.y);
      mma_16816(c, w[ch * 8 + 2], w[ch * 8 + 6], w[ch * 8 + 3], w[ch * 8 + 7], xv.z, xv.w);
    }
  }
}
template <bool FULL>
__device__ __forceinline__ void ds_load_item(uint32_t (&w)[32], const __half* wa, const __half* wb, int nch) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
    if (FULL || ch < nch) {
      a = *reinterpret_cast<const uint4*>(wa + ch * 32);
      b = *reinterpret_cast<const uint4*>(wb + ch * 32);
    }
    w[ch * 8 + 0] = a.x; w[ch * 8 + 1] = a.y; w[ch * 8 + 2] = a.z; w[ch * 8 + 3] = a.w;
    w[ch * 8 + 4] = b.x; w[ch * 8 + 5] = b.y; w[ch * 8 + 6] = b.z; w[ch * 8 + 7] = b.w;
  }
}
// partial sums: [tile of the group][warp][16 rows][NBC batch columns]; the m16n8 accumulator fragment of lane (g, t)
// holds rows g and g+8, columns 2t and 2t+1, of which only n < NBC are kept
template <int NBC>
__device__ __forceinline__ void ds_flush_tile(float* p, const float (&c)[4], int g, int t, bool add) {
  if (NBC == 1) {
    if (t == 0) {
      if (add) { p[g] += c[0]; p[g + 8] += c[2]; }
      else { p[g] = c[0]; p[g + 8] = c[2]; }
    }
  } else if (2 * t < NBC) {
    float* a = p + g * NBC + 2 * t;
    float* b = p + (g + 8) * NBC + 2 * t;
    if (add) { a[0] += c[0]; a[1] += c[1]; b[0] += c[2]; b[1] += c[3]; }
    else { a[0] = c[0]; a[1] = c[1]; b[0] = c[2]; b[1] = c[3]; }
  }
}
template <int NBC>
__device__ __forceinline__ float ds_part_out(const float* s_part, int idx) {
  const int tile = idx / (GV_TILE * NBC), rn = idx % (GV_TILE * NBC);
  const float* p = s_part + tile * (GV_WARPS * GV_TILE * NBC) + rn;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < GV_WARPS; ++w) v += p[w * (GV_TILE * NBC)];
  return v;
}
// tiles [t0,t1) of this CTA for one k-phase; tiles outside [tlo,thi) have no needed row (head phase) and are zeroed
template <int NBC>
__device__ __forceinline__ void consume_tiles(const DsSmem& m, const __half* xs, int KP, int KS, int t0, int t1, int tlo, int thi, bool add, int B,
                                              int stages, int& rs, uint32_t& rph, const Watch& w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ksw = KS / GV_WARPS;
  const int nch = ksw / 32;
  const int wpitch = KS + WS_PAD;
  const int nks = KP / KS;
  const bool xok = g < B;
  const bool full = nch == 4;
  const __half* xbase = xs + g * (KP + XS_PAD) + warp * ksw + t * 8;
  for (int tile = t0; tile < t1; ++tile) {
    float* pt = m.s_part + (static_cast<size_t>(tile - t0) * GV_WARPS + warp) * (GV_TILE * NBC);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    if (tile < tlo || tile >= thi) {
      if (!add) ds_flush_tile<NBC>(pt, c, g, t, false);
      continue;
    }
    for (int j = 0; j < nks; ++j) {
      ds_mbar_wait(&m.full[rs], rph, w, 32);
      const __half* wa = m.ws + static_cast<size_t>(rs) * DS_STAGE_HALVES + static_cast<size_t>(g) * wpitch + warp * ksw + t * 8;
      const __half* wb = wa + 8 * wpitch;
      uint32_t wr[32];
      // stage -> registers, hand the stage back to the producer, then the MMA chain (off the ring's refill cycle)
      if (full) ds_load_item<true>(wr, wa, wb, nch);
      else ds_load_item<false>(wr, wa, wb, nch);
      __syncwarp();
      if (lane == 0) mbar_arrive(&m.empty[rs]);
      if (full) ds_mma_item<true>(c, wr, xbase + j * KS, xok, nch);
      else ds_mma_item<false>(c, wr, xbase + j * KS, xok, nch);
      if (++rs == stages) { rs = 0; rph ^= 1; }
    }
    ds_flush_tile<NBC>(pt, c, g, t, add);
  }
}

// O / down projection, k-outer: for every KS-column chunk of the input, in order, every warp fetches ITS OWN K slice of the
// chunk (KS/8 = 128 columns per batch row), stages it in a warp-private buffer for its MMA fragments (no CTA barrier --
// nobody else reads that slice) and multiplies it into the accumulators of all the CTA's row tiles (<= 3).  The fetch of
// chunk j+1 is in flight while chunk j is multiplied, and a late producer of the previous phase only delays the chunk it
// produces.
//   batch <= 2 : the slice is polled out of the flag-in-data buffer (one 16-byte unit pair per lane and batch row);
//   batch >= 4 : the caller has acquired the producers' epoch flags (wait_flags); the warp loads its slice of the dense buffer with
//                16-byte loads.  (Acquiring only the ~12 producer CTAs of each chunk inside this loop was tried: the per-chunk
//                predicate evaluation and the blocking polls cost more than they saved -- batch 4: 4.0 -> 4.5 ms per step.)
template <int NBC, bool DENSE>
__device__ __forceinline__ void consume_korder(const DsSmem& m, const uint2* src_ll, const __half* src_d, int K, int KS, int ntile, int B,
                                               unsigned epoch, int stages, int& rs, uint32_t& rph, const Watch& w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ksw = KS / GV_WARPS;                   // this warp's columns of a chunk
  const int nch = ksw / 32;
  const int wpitch = KS + WS_PAD;
  const int spitch = ksw + XS_PAD;                 // warp-private slice buffer: [NBC rows][ksw + pad]
  __half* slice = m.xs + static_cast<size_t>(warp) * NBC * spitch;
  const bool xok = g < B;
  const bool full = nch == 4;
  float c[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  constexpr int NL = DENSE ? (NBC + 1) / 2 : NBC;  // 16-byte loads per lane per chunk
  uint4 nxt[NL];
  const int l16 = ksw >> 3;                        // dense: 16-byte loads per row of the slice (16 when KS = 1024)
  auto issue = [&](int kcol) {
    if (DENSE) {
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const int idx = lane + 32 * i, b = idx / l16, j = idx - b * l16;
        if (b < B) nxt[i] = __ldcg(reinterpret_cast<const uint4*>(src_d + static_cast<size_t>(b) * K + kcol + warp * ksw) + j);
      }
    } else {
      const uint4* s4 = reinterpret_cast<const uint4*>(src_ll);
#pragma unroll
      for (int b = 0; b < NL; ++b)
        if (4 * lane < ksw && b < B) nxt[b] = ll_ld16(s4 + static_cast<size_t>(b) * (K >> 2) + ((kcol + warp * ksw) >> 2) + lane);
    }
  };
  issue(0);
  for (int kcol = 0; kcol < K; kcol += KS) {
    __syncwarp();                                  // every lane is done with the previous chunk's fragments
    if (DENSE) {
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const int idx = lane + 32 * i, b = idx / l16, j = idx - b * l16;
        if (b < B) *reinterpret_cast<uint4*>(slice + b * spitch + 8 * j) = nxt[i];
      }
    } else {
      const uint4* s4 = reinterpret_cast<const uint4*>(src_ll);
#pragma unroll
      for (int b = 0; b < NL; ++b) {
        if (4 * lane < ksw && b < B) {
          const uint4* a = s4 + static_cast<size_t>(b) * (K >> 2) + ((kcol + warp * ksw) >> 2) + lane;
          unsigned spins = 0;
          while (nxt[b].y != epoch || nxt[b].w != epoch) {
            ds_repoll(w, spins, 23);
            nxt[b] = ll_ld16(a);
          }
          *reinterpret_cast<uint2*>(slice + b * spitch + 4 * lane) = make_uint2(nxt[b].x, nxt[b].z);
        }
      }
    }
    __syncwarp();
    if (kcol + KS < K) issue(kcol + KS);
    const __half* xp = slice + g * spitch + t * 8;
#pragma unroll
    for (int tile = 0; tile < 3; ++tile) {
      if (tile < ntile) {
        ds_mbar_wait(&m.full[rs], rph, w, 32);
        const __half* wa = m.ws + static_cast<size_t>(rs) * DS_STAGE_HALVES + static_cast<size_t>(g) * wpitch + warp * ksw + t * 8;
        const __half* wb = wa + 8 * wpitch;
        uint32_t wr[32];
        if (full) ds_load_item<true>(wr, wa, wb, nch);
        else ds_load_item<false>(wr, wa, wb, nch);
        __syncwarp();
        if (lane == 0) mbar_arrive(&m.empty[rs]);
        if (full) ds_mma_item<true>(c[tile], wr, xp, xok, nch);
        else ds_mma_item<false>(c[tile], wr, xp, xok, nch);
        if (++rs == stages) { rs = 0; rph ^= 1; }
      }
    }
  }
#pragma unroll
  for (int tile = 0; tile < 3; ++tile)
    if (tile < ntile) ds_flush_tile<NBC>(m.s_part + (static_cast<size_t>(tile) * GV_WARPS + warp) * (GV_TILE * NBC), c[tile], g, t, false);
}

// epilogue of tiles [t0,t1): kind 0 store, 1 residual add (residual rows live in shared memory), 2 SwiGLU pairs.
// Outputs leave as exchange units {2 adjacent outputs of one batch row, epoch}.
template <int NBC, bool DENSE>
__device__ __forceinline__ void epilogue_ll(const DsSmem& m, int r0, int r1, int t0, int t1, int B, int kind, uint2* out_ll, __half* out_d,
                                            int row_units, unsigned epoch, __half* s_res) {
  const int nout = (t1 - t0) * (GV_TILE * NBC);
  for (int base = 0; base < nout; base += GV_CONS) {            // warp-uniform trip count: the shuffles below are full-warp
    const int idx = base + threadIdx.x;
    const bool in = idx < nout;
    const float v = in ? ds_part_out<NBC>(m.s_part, idx) : 0.f;
    const int row = r0 + t0 * GV_TILE + idx / NBC, n = idx % NBC;
    const bool ok = in && n < B && row < r1;
    if (kind == 2) {
      const float up = __shfl_xor_sync(0xffffffffu, v, NBC);                 // the next row of the same batch column
      const uint32_t a = half_bits(silu_f(v) * up);                          // meaningful on even rows
      const uint32_t a2 = __shfl_xor_sync(0xffffffffu, a, 2 * NBC);          // the next SwiGLU output
      if (ok && (row & 3) == 0) {
        if (DENSE) __stcg(reinterpret_cast<unsigned*>(out_d + static_cast<size_t>(n) * (2 * row_units) + (row >> 1)), a | (a2 << 16));
        else ll_st8(out_ll + static_cast<size_t>(n) * row_units + (row >> 2), a | (a2 << 16), epoch);
      }
    } else {
      float r = v;
      if (kind == 1 && ok) r += __half2float(s_res[n * DS_RES_PITCH + (row - r0)]);
      const __half h = __float2half_rn(r);
      if (kind == 1 && ok) s_res[n * DS_RES_PITCH + (row - r0)] = h;
      const uint32_t hb = static_cast<uint32_t>(__half_as_ushort(h));
      const uint32_t hp = __shfl_xor_sync(0xffffffffu, hb, NBC);
      if (ok && (row & 1) == 0) {
        if (DENSE) __stcg(reinterpret_cast<unsigned*>(out_d + static_cast<size_t>(n) * (2 * row_units) + row), hb | (hp << 16));
        else ll_st8(out_ll + static_cast<size_t>(n) * row_units + (row >> 1), hb | (hp << 16), epoch);
      }
    }
  }
}

// ---- attention -------------------------------------------------------------------------------------------------
// Per-step control block of this CTA's (fixed) attention work item (b, kv head, split): key pages [p0, p0+np) of the
// row's block table, resolved once per launch.
struct AttnCtl {
  int valid, b, kv, split, t, p0, np, page_new;
  int pages[64];
  float cs[DS_HD / 2], sn[DS_HD / 2];
};
__device__ __forceinline__ void attn_ctl_init(const DsParams& p, AttnCtl* c, int vid) {
  const int items = p.B * p.kvh * p.nsplit;
  const int it = vid;
  if (threadIdx.x == 0) {
    c->valid = (it < items && p.layers > 0) ? 1 : 0;
    c->np = 0;
    if (c->valid) {
      c->split = it % p.nsplit; c->kv = (it / p.nsplit) % p.kvh; c->b = it / (p.nsplit * p.kvh);
      c->t = p.st.pos[c->b];
      const int npages = (c->t + DS_PAGE - 1) / DS_PAGE;                 // pages holding cached keys (positions < t)
      const int pps = (npages + p.nsplit - 1) / p.nsplit;
      c->p0 = c->split * pps;
      c->np = max(0, min(pps, npages - c->p0));
    }
  }
  __syncthreads();
  if (c->valid) {
    const int32_t* table = p.block_table + static_cast<size_t>(p.st.slot_ids[c->b]) * p.max_pages;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) c->pages[i] = (i < c->np) ? table[c->p0 + i] : 0;
    if (threadIdx.x == 64) c->page_new = table[c->t / DS_PAGE];
    if (threadIdx.x < DS_HD / 2) {
      c->cs[threadIdx.x] = p.cos_tab[static_cast<size_t>(c->t) * (DS_HD / 2) + threadIdx.x];
      c->sn[threadIdx.x] = p.sin_tab[static_cast<size_t>(c->t) * (DS_HD / 2) + threadIdx.x];
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

// Work item (b, kv, split) on the 8 consumer warps.  Scratch lives in the s_part region.
template <bool DENSE>
__device__ __forceinline__ void attn_item(const DsParams& p, const DsLayer& L, const AttnCtl& ctl, const DsSmem& m, float* scratch, int dl,
                                          unsigned ep_qkv, unsigned ep_part, unsigned ep_attn, int stages, int& rs, uint32_t& rph,
                                          const Watch& w) {
  __half* raw = reinterpret_cast<__half*>(scratch);          // [512 q | 128 k | 128 v] as they come out of the QKV projection
  __half* qs = raw + 768;                                    // [4][128] RoPE'd q (fp16, as the reference's q after apply_rotary_pos_emb)
  float* knew = reinterpret_cast<float*>(qs + DS_G * DS_HD);  // [128] RoPE'd key of the token being decoded (fp16-rounded)
  float* vnew = knew + DS_HD;                                // [128]
  float* red_o = vnew + DS_HD;                               // [8 warps][4][128]
  float* red_ml = red_o + GV_WARPS * DS_G * DS_HD;           // [8 warps][4][2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int b = ctl.b, kv = ctl.kv, split = ctl.split, tpos = ctl.t;
  const float sl2 = rsqrtf(static_cast<float>(DS_HD)) * 1.4426950408889634f;    // softmax scale in the exp2 domain

  // ---- q (4 heads), and for split 0 the new k / v row, from the QKV exchange buffer: only this kv head's producers are awaited
  if (DENSE) {
    const int q0 = kv * DS_G * DS_HD, k0r = p.heads * DS_HD + kv * DS_HD, v0r = (p.heads + p.kvh) * DS_HD + kv * DS_HD;
    const int grid = gridDim.x, qkvw = p.qkvw;
    wait_flags(p.fq, ep_qkv, grid, [&](int c) {
      int a0, a1;
      rows_of_cta(qkvw, 2, c, grid, a0, a1);
      bool need = a0 < q0 + DS_G * DS_HD && a1 > q0;
      if (split == 0) need = need || (a0 < k0r + DS_HD && a1 > k0r) || (a0 < v0r + DS_HD && a1 > v0r);
      return need;
    }, w, 41);
    const __half* row = p.qkvd + static_cast<size_t>(b) * qkvw;
    int src = -1, dst = 0;
    if (tid < 64) { src = q0 + tid * 8; dst = tid * 8; }
    else if (split == 0 && tid < 80) { src = k0r + (tid - 64) * 8; dst = 512 + (tid - 64) * 8; }
    else if (split == 0 && tid < 96) { src = v0r + (tid - 80) * 8; dst = 640 + (tid - 80) * 8; }
    if (src >= 0) *reinterpret_cast<uint4*>(raw + dst) = __ldcg(reinterpret_cast<const uint4*>(row + src));
  } else {
    const uint4* row4 = reinterpret_cast<const uint4*>(p.qkv_ll + static_cast<size_t>(b) * (p.qkvw >> 1));
    int src = -1, dst = 0;
    if (tid < 128) { src = (kv * DS_G * DS_HD) / 4 + tid; dst = tid * 4; }
    else if (split == 0 && tid < 160) { src = (p.heads * DS_HD + kv * DS_HD) / 4 + (tid - 128); dst = 512 + (tid - 128) * 4; }
    else if (split == 0 && tid < 192) { src = ((p.heads + p.kvh) * DS_HD + kv * DS_HD) / 4 + (tid - 160); dst = 640 + (tid - 160) * 4; }
    if (src >= 0) {
      const uint4 v = ll_wait16(row4 + src, ep_qkv, w, 41);
      *reinterpret_cast<uint2*>(raw + dst) = make_uint2(v.x, v.z);
    }
  }
  cons_sync();
  {
    const int h = tid >> 6, d = tid & 63;                    // 4 heads x 64 rotation pairs
    const float a = __half2float(raw[h * DS_HD + d]), c2 = __half2float(raw[h * DS_HD + d + 64]);
    qs[h * DS_HD + d] = __float2half_rn(a * ctl.cs[d] - c2 * ctl.sn[d]);
    qs[h * DS_HD + d + 64] = __float2half_rn(c2 * ctl.cs[d] + a * ctl.sn[d]);
    if (split == 0) {                                        // KV append (DynamicCache.update) + the token's own key / value
      __half* kt = L.kv_pool + kv_tile_offset(ctl.page_new, p.kvh, kv, 0, DS_PAGE, DS_HD);
      const int poff = tpos % DS_PAGE;
      if (tid < 64) {
        const float ka = __half2float(raw[512 + tid]), kc = __half2float(raw[512 + tid + 64]);
        const __half ra = __float2half_rn(ka * ctl.cs[tid] - kc * ctl.sn[tid]);
        const __half rb = __float2half_rn(kc * ctl.cs[tid] + ka * ctl.sn[tid]);
        kt[poff * DS_HD + kv_chunk_swz(tid, poff)] = ra;
        kt[poff * DS_HD + kv_chunk_swz(tid + 64, poff)] = rb;
        knew[tid] = __half2float(ra);
        knew[tid + 64] = __half2float(rb);
      } else if (tid < 128) {
        const int d2 = (tid - 64) * 2;
        const uint32_t u = *reinterpret_cast<const uint32_t*>(raw + 640 + d2);
        *reinterpret_cast<uint32_t*>(kt + DS_PAGE * DS_HD + poff * DS_HD + kv_chunk_swz(d2, poff)) = u;
        const float2 f = unpack_half2(u);
        vnew[d2] = f.x;
        vnew[d2 + 1] = f.y;
      }
    }
  }
  cons_sync();
  stamp(p, dl, 20);
  // ---- Q fragments (A operand of m16n8k16: rows 0..3 = the 4 query heads, rows 4..15 zero) ----
  uint32_t aq0[8], aq2[8];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    aq0[ks] = (g < DS_G) ? *reinterpret_cast<const uint32_t*>(qs + g * DS_HD + ks * 16 + 2 * t) : 0u;
    aq2[ks] = (g < DS_G) ? *reinterpret_cast<const uint32_t*>(qs + g * DS_HD + ks * 16 + 8 + 2 * t) : 0u;
  }
  float mrow = -INFINITY, lrow = 0.f;      // online softmax state of head g (exp2 domain), replicated over the quad
  // O^T accumulators (dims in the M slot, the 4 heads in the N slot of m16n8k8): o[mt][0..1] = (dim 16mt+g, heads 2t, 2t+1),
  // o[mt][2..3] = (dim 16mt+8+g, heads 2t, 2t+1); lanes with t >= 2 hold the padding heads 4..7.  Half the registers of
  // O with the heads in the M slot (4 of 16 rows used).
  float o[8][4];
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f;
  const int src_a = 8 * t, src_b = 8 * t + 4;                // lanes holding the softmax state of heads 2t and 2t+1
  for (int i = 0; i < ctl.np; ++i) {
    ds_mbar_wait(&m.full[rs], rph, w, 42);
    const __half* Kt = m.ws + static_cast<size_t>(rs) * DS_STAGE_HALVES;
    const __half* Vt = Kt + DS_PAGE * DS_HD;
    // S = Q K^T for this warp's 8 keys (B operand: key 8*warp + g, 16 dims per k-step; chunk swizzle = chunk ^ (key & 7))
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
    const __half* krow = Kt + (8 * warp + g) * DS_HD + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + (((2 * ks) ^ g) << 3));
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + (((2 * ks + 1) ^ g) << 3));
      mma_16816(sc, aq0[ks], 0u, aq2[ks], 0u, b0, b1);
    }
    const int key0 = (ctl.p0 + i) * DS_PAGE + 8 * warp + 2 * t;
    const float s0 = (key0 < tpos) ? sc[0] * sl2 : -INFINITY;
    const float s1 = (key0 + 1 < tpos) ? sc[1] * sl2 : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float mn = fmaxf(mrow, mx);
    const float mref = (mn == -INFINITY) ? 0.f : mn;
    const float p0 = exp2f(s0 - mref), p1 = exp2f(s1 - mref);
    const float corr = exp2f(mrow - mref);
    float ps = p0 + p1;
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    ps += __shfl_xor_sync(0xffffffffu, ps, 2);
    lrow = lrow * corr + ps;
    mrow = mn;
    const uint32_t bp = (g < DS_G) ? pack_half2(p0, p1) : 0u;   // B operand of m16n8k8: P^T[keys 2t, 2t+1][head g]
    const float ca = __shfl_sync(0xffffffffu, corr, src_a), cb = __shfl_sync(0xffffffffu, corr, src_b);
    // O^T += V^T P^T : V^T fragments of the warp's 8 keys, 32 dims per ldmatrix.x4.trans
    const uint32_t vaddr = smem_u32(Vt + (8 * warp + (lane & 7)) * DS_HD);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      uint32_t v0, v1, v2, v3;
      ldmatrix_x4_trans(v0, v1, v2, v3, vaddr + ((((4 * q4 + (lane >> 3)) ^ (lane & 7)) << 3) << 1));
      o[2 * q4][0] *= ca; o[2 * q4][1] *= cb; o[2 * q4][2] *= ca; o[2 * q4][3] *= cb;
      o[2 * q4 + 1][0] *= ca; o[2 * q4 + 1][1] *= cb; o[2 * q4 + 1][2] *= ca; o[2 * q4 + 1][3] *= cb;
      mma_1688(o[2 * q4], v0, v1, bp);
      mma_1688(o[2 * q4 + 1], v2, v3, bp);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&m.empty[rs]);
    if (++rs == stages) { rs = 0; rph ^= 1; }
  }
  if (split == 0 && warp == 0) {   // the token being decoded attends to itself: key / value from shared memory, not the cache
    float sn4[DS_G];
#pragma unroll
    for (int h = 0; h < DS_G; ++h) {
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) d += __half2float(qs[h * DS_HD + lane * 4 + k]) * knew[lane * 4 + k];
      sn4[h] = warp_sum(d);
    }
    const float s = ((g & 3) == 0 ? sn4[0] : (g & 3) == 1 ? sn4[1] : (g & 3) == 2 ? sn4[2] : sn4[3]) * sl2;
    const float mn = fmaxf(mrow, s);
    const float pr = exp2f(s - mn), corr = exp2f(mrow - mn);
    lrow = lrow * corr + pr;
    mrow = mn;
    const float ca = __shfl_sync(0xffffffffu, corr, src_a), cb = __shfl_sync(0xffffffffu, corr, src_b);
    const float pa = __shfl_sync(0xffffffffu, pr, src_a), pb = __shfl_sync(0xffffffffu, pr, src_b);
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      const float va = vnew[16 * mt + g], vb = vnew[16 * mt + 8 + g];
      o[mt][0] = o[mt][0] * ca + pa * va;
      o[mt][1] = o[mt][1] * cb + pb * va;
      o[mt][2] = o[mt][2] * ca + pa * vb;
      o[mt][3] = o[mt][3] * cb + pb * vb;
    }
  }
  stamp(p, dl, 21);
  // ---- CTA-level merge of the 8 warps' states ----
  if (t < 2) {
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      float* ra = red_o + (warp * DS_G + 2 * t) * DS_HD + 16 * mt + g;
      ra[0] = o[mt][0];
      ra[DS_HD] = o[mt][1];
      ra[8] = o[mt][2];
      ra[DS_HD + 8] = o[mt][3];
    }
  }
  if (g < DS_G && t == 0) { red_ml[(warp * DS_G + g) * 2] = mrow; red_ml[(warp * DS_G + g) * 2 + 1] = lrow; }
  cons_sync();
  const int item = (b * p.kvh + kv) * p.nsplit + split;
  {
    const int h = tid >> 6, d = (tid & 63) * 2;
    float M = -INFINITY;
#pragma unroll
    for (int ww = 0; ww < GV_WARPS; ++ww) M = fmaxf(M, red_ml[(ww * DS_G + h) * 2]);
    const float Mref = (M == -INFINITY) ? 0.f : M;
    float Ls = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int ww = 0; ww < GV_WARPS; ++ww) {
      const float wt = exp2f(red_ml[(ww * DS_G + h) * 2] - Mref);
      Ls += wt * red_ml[(ww * DS_G + h) * 2 + 1];
      const float2 x2 = *reinterpret_cast<const float2*>(red_o + (ww * DS_G + h) * DS_HD + d);
      a0 += wt * x2.x;
      a1 += wt * x2.y;
    }
    ll_st16(p.part_o_ll + static_cast<size_t>(item) * (DS_G * DS_HD) + h * DS_HD + d, __float_as_uint(a0), __float_as_uint(a1), ep_part);
    if ((tid & 63) == 0) ll_st16(p.part_ml_ll + static_cast<size_t>(item) * DS_G + h, __float_as_uint(M), __float_as_uint(Ls), ep_part);
  }
  stamp(p, dl, 22);
  // ---- distributed combine: split s < ncomb merges the partials of 512/ncomb output dims of this (row, kv head) ----
  if (split < p.ncomb) {
    const int cd = (DS_G * DS_HD) / p.ncomb;
    if (tid < cd / 2) {
      const int D = split * cd + 2 * tid, h = D >> 7;
      const size_t gi = static_cast<size_t>(b * p.kvh + kv) * p.nsplit;
      float M = -INFINITY, Ls = 0.f, a0 = 0.f, a1 = 0.f;
      for (int s0 = 0; s0 < p.nsplit; s0 += 4) {
        uint4 ml[4], po[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (s0 + k < p.nsplit) {
            ml[k] = ll_ld16(p.part_ml_ll + (gi + s0 + k) * DS_G + h);
            po[k] = ll_ld16(p.part_o_ll + (gi + s0 + k) * (DS_G * DS_HD) + D);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (s0 + k < p.nsplit) {
            unsigned spins = 0;
            while (ml[k].y != ep_part || ml[k].w != ep_part) { ds_repoll(w, spins, 43); ml[k] = ll_ld16(p.part_ml_ll + (gi + s0 + k) * DS_G + h); }
            while (po[k].y != ep_part || po[k].w != ep_part) { ds_repoll(w, spins, 44); po[k] = ll_ld16(p.part_o_ll + (gi + s0 + k) * (DS_G * DS_HD) + D); }
            const float ms = __uint_as_float(ml[k].x), ls = __uint_as_float(ml[k].z);
            if (ls != 0.f) {
              const float mn = fmaxf(M, ms);
              const float c1 = exp2f(M - mn), c2 = exp2f(ms - mn);
              Ls = Ls * c1 + ls * c2;
              a0 = a0 * c1 + __uint_as_float(po[k].x) * c2;
              a1 = a1 * c1 + __uint_as_float(po[k].z) * c2;
              M = mn;
            }
          }
        }
      }
      const float inv = 1.f / Ls;
      if (DENSE) __stcg(reinterpret_cast<unsigned*>(p.attnd + static_cast<size_t>(b) * p.qd + kv * DS_G * DS_HD + D), pack_half2(a0 * inv, a1 * inv));
      else ll_st8(p.attn_ll + static_cast<size_t>(b) * (p.qd >> 1) + ((kv * DS_G * DS_HD + D) >> 1), pack_half2(a0 * inv, a1 * inv), ep_attn);
    }
  }
  stamp(p, dl, 23);
}

// ---- finalize (CTA 0, 256 consumer threads) --------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned f2ord(float f) {   // order-preserving map float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// HF sampling of one row (GenerationMixin._sample with the default warpers: temperature, top-k; top_p = 1): the
// kept set is {z >= k-th largest z}, p ~ exp((z - max) / T), token = first id whose cumulative mass (ids ascending)
// reaches u * total, u from a counter-based generator (seed, step, row) that the tests reproduce on the host.
__device__ __forceinline__ int sample_row(const DsParams& p, int b, int lo, int hi, float zmax, int step, float* s_f, int* s_i) {
  const int tid = threadIdx.x;
  const float* z = p.samp_logits + static_cast<size_t>(b) * p.n_valid;
  const int n = hi - lo;
  const float invT = 1.f / __int_as_float(p.st.ctrl[4]);
  const int top_k = p.st.ctrl[5];
  unsigned thr = 0u;                                  // keep z with f2ord(z) >= thr
  if (top_k > 0 && top_k < n) {
    // largest threshold with count(z >= thr) >= top_k  ==  the k-th largest value (ties kept, as torch.topk + '<' does)
    unsigned lo_u = 0u, hi_u = 0xffffffffu;
    for (int it = 0; it < 32; ++it) {
      const unsigned mid = lo_u + ((hi_u - lo_u) >> 1) + ((hi_u - lo_u) & 1u);
      int cnt = 0;
      for (int i = tid; i < n; i += GV_CONS) cnt += (f2ord(__ldcg(z + lo + i)) >= mid) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      cons_sync();
      if ((tid & 31) == 0) s_i[tid >> 5] = cnt;
      cons_sync();
      int tot = 0;
#pragma unroll
      for (int k = 0; k < GV_WARPS; ++k) tot += s_i[k];
      if (tot >= top_k) lo_u = mid; else hi_u = mid - 1;
      if (lo_u == hi_u) break;
    }
    thr = lo_u;
  }
  // per-thread mass of a contiguous id chunk (fixed order -> reproducible)
  const int chunk = (n + GV_CONS - 1) / GV_CONS;
  const int c0 = min(n, tid * chunk), c1 = min(n, c0 + chunk);
  float mass = 0.f;
  for (int i = c0; i < c1; ++i) {
    const float v = __ldcg(z + lo + i);
    if (f2ord(v) >= thr) mass += __expf((v - zmax) * invT);
  }
  cons_sync();
  s_f[tid] = mass;
  cons_sync();
  int tok = 0;
  if (tid == 0) {
    float total = 0.f;
    for (int k = 0; k < GV_CONS; ++k) total += s_f[k];
    const unsigned long long seed = (static_cast<unsigned long long>(static_cast<unsigned>(p.st.ctrl[7])) << 32) | static_cast<unsigned>(p.st.ctrl[6]);
    const unsigned long long r = splitmix64(seed ^ (static_cast<unsigned long long>(step) * 0x9E3779B97F4A7C15ull) ^
                                            (static_cast<unsigned long long>(b) * 0xD1B54A32D192ED03ull));
    const float u = (static_cast<float>(r >> 40) + 0.5f) * (1.0f / 16777216.0f);
    const float target = u * total;
    float acc = 0.f;
    int k = 0;
    for (; k < GV_CONS - 1; ++k) {
      if (acc + s_f[k] >= target) break;
      acc += s_f[k];
    }
    const int d0 = min(n, k * chunk), d1 = min(n, d0 + chunk);
    int pick = -1, last = -1;
    for (int i = d0; i < d1; ++i) {
      const float v = __ldcg(z + lo + i);
      if (f2ord(v) >= thr) {
        last = i;
        acc += __expf((v - zmax) * invT);
        if (acc >= target) { pick = i; break; }
      }
    }
    if (pick < 0) pick = last >= 0 ? last : d0;       // rounding at the very end of the mass
    s_i[16] = lo + pick;
  }
  cons_sync();
  tok = s_i[16];
  cons_sync();
  return tok;
}

__device__ __forceinline__ void finalize_step(const DsParams& p, int ncta, unsigned ep_cand, const int* s_hlo, const int* s_hhi,
                                              const Watch& w) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_feed[GV_MAXB], s_cont[GV_MAXB];
  __shared__ float s_f[GV_CONS];
  __shared__ int s_w[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const DecodeStreamState& st = p.st;
  const int step = *st.step;
  // ctrl: 0 row stride of out_ids / forced, 1 eos, 2 pad, 3 do_sample, 8 use_forced, 9 rows of this launch (bit mask: an admission
  // launch of the continuous batcher finalizes ONE row and must leave the others -- their state and next embedding -- alone)
  const int cap = st.ctrl[0], eos = st.ctrl[1], pad = st.ctrl[2], do_sample = st.ctrl[3], use_forced = st.ctrl[8], mask = st.ctrl[9];
  for (int b = 0; b < p.B; ++b) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (tid < ncta) {
      const uint4 c = ll_wait16(p.cand_ll + static_cast<size_t>(b) * ncta + tid, ep_cand, w, 51);
      bv = __uint_as_float(c.x);
      bi = static_cast<int>(c.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    cons_sync();
    for (int k = 0; k < 8; ++k)
      if (s_v[k] > bv || (s_v[k] == bv && s_i[k] < bi)) { bv = s_v[k]; bi = s_i[k]; }
    cons_sync();
    const bool in_launch = (mask >> b) & 1;
    const int rs = st.row_step[b];                        // tokens this row has emitted so far
    int tok = bi;
    if (in_launch && do_sample && !st.finished[b]) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");     // the candidate lines were released after the logits stores
      tok = sample_row(p, b, s_hlo[b], s_hhi[b], bv, rs, s_f, s_w);
    }
    if (tid == 0) {
      int feed = -1;
      if (in_launch) {
        const int budget = st.row_max[b];
        if (st.finished[b]) tok = pad;                     // finished rows keep emitting pad (HF generation loop)
        if (rs < budget) st.out_ids[static_cast<int64_t>(b) * cap + rs] = tok;
        feed = (use_forced && !st.finished[b] && rs < budget) ? static_cast<int>(st.forced[static_cast<int64_t>(b) * cap + rs]) : tok;
        if (!st.finished[b] && feed == eos) st.finished[b] = 1;
        st.row_step[b] = rs + 1;
        s_cont[b] = rs + 1 < budget;
      }
      s_feed[b] = feed;
    }
    cons_sync();
  }
  if (tid == 0) {
    int all = 1;                                           // every row has hit EOS or used up its budget
    for (int b = 0; b < p.B; ++b) all &= (st.finished[b] || st.row_step[b] >= st.row_max[b]) ? 1 : 0;
    for (int b = 0; b < p.B; ++b) {
      const int tok = s_feed[b];
      if (tok < 0) continue;
      // the head switch happens in prepare_inputs_for_generation (trace_mistral.py:336-344), i.e. only for tokens that are
      // fed back: not after the last token of the budget and not once every row has finished
      if (!all && s_cont[b] && st.heads[b] >= 0) {
        if (tok == p.vocab) st.heads[b] = 1;
        else if (tok == p.vocab + 1) st.heads[b] = 2;
        else if (tok == p.vocab + 14) st.heads[b] = 0;
      }
      if (p.advance && s_cont[b]) st.pos[b] += 1;          // a row out of budget stays where it is (its KV slot is full)
    }
    *st.step = step + 1;
    if (all && *st.done_step == 0) *st.done_step = step + 1;
    *p.epoch_ctr += 1;
  }
  // next-step input embedding (trace_arch.py:349-373): id % vocab text row, overwritten by the sync / time / score rows
  for (int b = 0; b < p.B; ++b) {
    const int tok = s_feed[b];
    if (tok < 0) continue;
    const __half* src;
    if (tok == p.vocab) src = p.sync_tab;
    else if (tok > p.vocab && tok < p.vocab + 14) src = p.time_tab + static_cast<size_t>(tok - p.vocab - 1) * p.H;
    else if (tok >= p.vocab + 14) src = p.score_tab + static_cast<size_t>(tok - p.vocab - 14) * p.H;
    else src = p.embed + static_cast<size_t>(tok % p.vocab) * p.H;
    for (int i = tid; i < p.H / 8; i += GV_CONS)
      reinterpret_cast<uint4*>(p.x_next + static_cast<size_t>(b) * p.H)[i] = reinterpret_cast<const uint4*>(src)[i];
  }
}

struct PhaseGeo { int N, K, KP, KS, unit; };
__device__ __forceinline__ PhaseGeo phase_geo(const DsParams& p, int type) {
  PhaseGeo g;
  switch (type) {
    case 0: g.N = p.qkvw; g.K = p.H; g.unit = 2; break;
    case 1: g.N = p.H; g.K = p.qd; g.unit = 2; break;
    case 2: g.N = 2 * p.ffn; g.K = p.H; g.unit = 4; break;
    case 3: g.N = p.H; g.K = p.ffn; g.unit = 2; break;
    default: g.N = p.head_rows; g.K = p.H; g.unit = 2; break;
  }
  g.KP = p.kp[type];
  g.KS = p.ks[type];
  return g;
}

template <int NB>
__global__ void __launch_bounds__(DS_THREADS, 1) decode_stream_kernel(const __grid_constant__ DsParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  DsSmem m;
  {
    size_t off = 0;
    m.ws = reinterpret_cast<__half*>(smem + off); off += static_cast<size_t>(p.stages) * DS_STAGE_HALVES * 2;
    m.xs = reinterpret_cast<__half*>(smem + off); off += (static_cast<size_t>(p.B) * p.xs_cap * 2 + 127) & ~static_cast<size_t>(127);
    m.s_part = reinterpret_cast<float*>(smem + off); off += static_cast<size_t>(p.part_bytes);
    m.s_red = reinterpret_cast<float*>(smem + off); off += GV_MAXB * 4 * GV_WARPS * 4;
    m.full = reinterpret_cast<uint64_t*>(smem + off); off += GV_MAXSTAGES * 8;
    m.empty = reinterpret_cast<uint64_t*>(smem + off);
  }
  __shared__ DsLayer s_layers[DS_MAX_LAYERS];
  __shared__ AttnCtl s_ctl;
  __shared__ int s_hlo[GV_MAXB], s_hhi[GV_MAXB];
  __shared__ __half s_res[GV_MAXB * DS_RES_PITCH];
  __shared__ float s_cv[GV_WARPS][8];
  __shared__ int s_ci[GV_WARPS][8];
  Watch w;
  w.t0 = gtimer();
  w.err = p.err;
  w.sleep_ns = static_cast<unsigned>(p.poll_sleep_ns);
  w.limit_ns = static_cast<unsigned long long>(p.watchdog_ms) * 1000000ull;
  for (int i = threadIdx.x; i < p.layers * static_cast<int>(sizeof(DsLayer) / 8); i += DS_THREADS)
    reinterpret_cast<unsigned long long*>(s_layers)[i] = reinterpret_cast<const unsigned long long*>(p.L)[i];
  // active-head row ranges: read ONCE, before CTA 0's finalize of this step can rewrite heads[]
  if (threadIdx.x < GV_MAXB) {
    int lo = 0, hi = 0;
    if (threadIdx.x < p.B) head_range(p.st.heads[threadIdx.x], p.vocab, lo, hi);
    s_hlo[threadIdx.x] = lo;
    s_hhi[threadIdx.x] = hi;
  }
  const unsigned ep0 = *p.epoch_ctr * DS_EPOCH_STRIDE;          // CTA 0 bumps the counter only after every CTA delivered its candidates
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&m.full[i], 1);
      mbar_init(&m.empty[i], GV_WARPS);
    }
    fence_barrier_init();
  }
  const int vid = blockIdx.x;
  stamp(p, 0, 0);                                               // kernel entry (after the layer table copy was issued)
  if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && p.layers > 0)   // slot 2: gap since the previous launch's last stamp
  {
    // per-step log in the free slots 4 / 5 of the per-CTA rows: row i = decode step i of the call (gap before it, entry -> end)
    const unsigned long long gap = p.dbg[0] - p.dbg[static_cast<size_t>(p.layers - 1) * 32 + 30];
    p.dbg[2] = gap;
    const int si = *p.st.step;
    if (si >= 0 && si < static_cast<int>(gridDim.x)) p.dbg[static_cast<size_t>(2) * p.layers * 32 + static_cast<size_t>(si) * 32 + 5] = gap;
  }
  attn_ctl_init(p, &s_ctl, vid);                                // contains __syncthreads
  stamp(p, 0, 1);                                               // prologue done: the producer starts streaming
  const int ncta = gridDim.x;
  int ulo = 0x7fffffff, uhi = 0;
  for (int b = 0; b < p.B; ++b) { ulo = min(ulo, s_hlo[b]); uhi = max(uhi, s_hhi[b]); }
  const int nseq = p.layers * 4 + 1;
  const int do_sample = p.st.ctrl[3];

  if (threadIdx.x >= GV_CONS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DS_REGS_PRODUCER));
    if (threadIdx.x >= GV_CONS + 32) return;
    // ===================== producer warp: everything the step reads from HBM, in program order =====================
    int s = 0;
    uint32_t ph = 0;
    for (int seq = 0; seq < nseq; ++seq) {
      const int type = seq == nseq - 1 ? 4 : (seq & 3), l = seq >> 2;
      const DsLayer& L = s_layers[type == 4 ? 0 : l];
      const PhaseGeo g = phase_geo(p, type);
      int r0, r1;
      rows_of_cta(g.N, g.unit, vid, ncta, r0, r1);
      const int vlo = type == 4 ? max(ulo, r0) : r0, vhi = type == 4 ? min(uhi, r1) : r1;
      const __half* W = type == 0 ? L.qkv_w : type == 1 ? L.o_w : type == 2 ? L.gateup_w : type == 3 ? L.down_w : p.head_w;
      if (vlo < vhi) produce_matrix(W, g.K, g.KP, g.KS, r0, r1, vlo, vhi, m, p.stages, s, ph, w, type == 1 || type == 3);
      if (type == 0 && s_ctl.valid) {
        // the K/V pages of this CTA's attention work item: cached rows do not depend on this step, so they queue right
        // behind the QKV weights and are in shared memory before q exists
        for (int i = 0; i < s_ctl.np; ++i) {
          ds_mbar_wait(&m.empty[s], ph ^ 1, w, 33);
          if ((threadIdx.x & 31) == 0) {
            mbar_expect_tx(&m.full[s], DS_KV_TILE_BYTES);
            bulk_g2s(m.ws + static_cast<size_t>(s) * DS_STAGE_HALVES, L.kv_pool + kv_tile_offset(s_ctl.pages[i], p.kvh, s_ctl.kv, 0, DS_PAGE, DS_HD),
                     DS_KV_TILE_BYTES, &m.full[s]);
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DS_REGS_CONSUMER));
  int rs = 0;
  uint32_t rph = 0;
  int xr0, xr1;
  rows_of_cta(p.H, 2, vid, ncta, xr0, xr1);
  // residual rows of this CTA (the O and down projections share one row partition): kept in shared memory for the whole step
  for (int i = threadIdx.x; i < p.B * DS_RES_PITCH; i += GV_CONS) {
    const int n = i / DS_RES_PITCH, r = i % DS_RES_PITCH;
    if (xr0 + r < xr1) s_res[i] = p.x[static_cast<size_t>(n) * p.H + xr0 + r];
  }
  cons_sync();
  constexpr int NLD = NB <= 2 ? 8 : 16;     // (16 polls in flight per thread make ptxas spill the batch-1 kernel: a spill is an L2 round trip here)
  constexpr bool DENSE = NB >= 4;
  const int xpitch_h = p.H + XS_PAD;
  const auto all_ctas = [](int) { return true; };

  for (int seq = 0; seq < nseq; ++seq) {
    const int type = seq == nseq - 1 ? 4 : (seq & 3), l = type == 4 ? max(p.layers - 1, 0) : (seq >> 2);
    const DsLayer& L = s_layers[l];
    const PhaseGeo g = phase_geo(p, type);
    const unsigned epl = ep0 + 8u * static_cast<unsigned>(l);
    int r0, r1;
    rows_of_cta(g.N, g.unit, vid, ncta, r0, r1);
    const int vlo = type == 4 ? max(ulo, r0) : r0, vhi = type == 4 ? min(uhi, r1) : r1;
    const int ntile = (r1 - r0 + GV_TILE - 1) / GV_TILE;
    int tlo = 0, thi = 0;
    if (vlo < vhi) { tlo = (vlo - r0) / GV_TILE; thi = (vhi - 1 - r0) / GV_TILE + 1; }
    const bool pro = (type != 1 && type != 3);
    float bv = -INFINITY;                       // head phase: running argmax candidate of this thread's batch column
    int bi = 0x7fffffff;
    const int ncol = threadIdx.x % NB;
    if (vlo < vhi) {
      // ---- input vector of the phase ----
      // x comes from: the dense input embedding (first phase of the launch), the O phase (gate/up), the previous layer's
      // down phase (QKV, head)
      const unsigned epx = type == 2 ? epl + 4 : (ep0 + 8u * static_cast<unsigned>(type == 4 ? p.layers - 1 : l - 1) + 6);
      if (pro && DENSE) {      // (both branches compile for every NB; the dead one is eliminated)
        const GammaRegs gm = load_gamma(type == 0 ? L.ln1 : type == 2 ? L.ln2 : p.final_norm, p.H);
        if (seq > 0) wait_flags(p.fx, epx, ncta, all_ctas, w, 22);
        stage_x_rmsnorm_nb<NB>(m.xs, seq == 0 ? p.x : p.xd, p.H, p.B, p.H, gm, p.eps, m.s_red);
      } else if (pro) {
        const Gamma4 gm = load_gamma4(type == 0 ? L.ln1 : type == 2 ? L.ln2 : p.final_norm, p.H);
        const bool dense = seq == 0;
        stage_rms<NLD>(m.xs, xpitch_h, dense ? nullptr : p.x_ll, p.x, p.B, p.H, gm, p.eps, m.s_red, dense ? 0u : epx, w);
      }
      if (!pro) {
        // O / down projection: k-outer, chunks multiplied as they arrive (consume_korder)
        stamp(p, l, 10 + type * 2);
        if (DENSE) wait_flags(type == 1 ? p.fa : p.fact, type == 1 ? epl + 3 : epl + 5, ncta, all_ctas, w, 21);
        consume_korder<NB, DENSE>(m, type == 1 ? p.attn_ll : p.act_ll, type == 1 ? p.attnd : p.actd, g.K, g.KS, ntile, p.B,
                                  type == 1 ? epl + 3 : epl + 5, p.stages, rs, rph, w);
        cons_sync();
        stamp(p, l, 24 + type);
        epilogue_ll<NB, DENSE>(m, r0, r1, 0, ntile, p.B, 1, p.x_ll, p.xd, p.H >> 1, type == 1 ? epl + 4 : epl + 6, s_res);
      }
      const int nkp = g.K / g.KP;
      for (int k0 = 0; k0 < g.K && pro; k0 += g.KP) {          // (RMSNorm phases stage the whole hidden row: one k-phase)
        const __half* xcur = m.xs;
        cons_sync();                                      // staged activations visible
        if (k0 == 0) stamp(p, l, 10 + type * 2);
        const int gt = nkp > 1 ? ntile : p.group_tiles;
        for (int t0 = 0; t0 < ntile; t0 += gt) {
          const int t1 = min(ntile, t0 + gt);
          consume_tiles<NB>(m, xcur, g.KP, g.KS, t0, t1, tlo, thi, k0 > 0, p.B, p.stages, rs, rph, w);
          if (k0 + g.KP < g.K) continue;         // more k-phases to accumulate
          cons_sync();
          if (t1 == ntile) stamp(p, l, 24 + type);
          if (type == 0) epilogue_ll<NB, DENSE>(m, r0, r1, t0, t1, p.B, 0, p.qkv_ll, p.qkvd, p.qkvw >> 1, epl + 1, s_res);
          else if (type == 1) epilogue_ll<NB, DENSE>(m, r0, r1, t0, t1, p.B, 1, p.x_ll, p.xd, p.H >> 1, epl + 4, s_res);
          else if (type == 2) epilogue_ll<NB, DENSE>(m, r0, r1, t0, t1, p.B, 2, p.act_ll, p.actd, p.ffn >> 1, epl + 5, s_res);
          else if (type == 3) epilogue_ll<NB, DENSE>(m, r0, r1, t0, t1, p.B, 1, p.x_ll, p.xd, p.H >> 1, epl + 6, s_res);
          else {
            // ---- heads: logits of the active ranges (fp16-rounded like the reference's fp16 GEMM) + argmax candidates ----
            const int lo = s_hlo[ncol], hi = s_hhi[ncol];
            const int nout = (t1 - t0) * (GV_TILE * NB);
            for (int idx = threadIdx.x; idx < nout; idx += GV_CONS) {
              const int r = r0 + t0 * GV_TILE + idx / NB;
              float v = __half2float(__float2half_rn(ds_part_out<NB>(m.s_part, idx)));
              const bool act = (ncol < p.B) && r < r1 && r >= lo && r < hi;
              if (!act) v = -INFINITY;
              if (ncol < p.B && r < r1 && r < p.n_valid) {
                if (p.logits_out != nullptr) p.logits_out[static_cast<size_t>(ncol) * p.n_valid + r] = v;
                if (do_sample) __stcg(p.samp_logits + static_cast<size_t>(ncol) * p.n_valid + r, v);
              }
              if (act && (v > bv || (v == bv && r < bi))) { bv = v; bi = r; }
            }
          }
          if (t1 < ntile) cons_sync();           // the next group reuses the partial-sum buffer
        }
      }
      stamp(p, l, 11 + type * 2);
    } else if (type == 4 && p.logits_out != nullptr) {
      for (int idx = threadIdx.x; idx < ntile * GV_TILE * NB; idx += GV_CONS) {
        const int r = r0 + idx / NB;
        if (ncol < p.B && r < r1 && r < p.n_valid) p.logits_out[static_cast<size_t>(ncol) * p.n_valid + r] = -INFINITY;
      }
    }
    if (DENSE && type < 4)                        // every CTA publishes every exchange, rows or not
      publish(type == 0 ? p.fq : type == 2 ? p.fact : p.fx, vid, type == 0 ? epl + 1 : type == 1 ? epl + 4 : type == 2 ? epl + 5 : epl + 6);
    if (type == 0) {
      cons_sync();                               // the partial-sum buffer becomes the attention scratch
      if (s_ctl.valid) attn_item<DENSE>(p, L, s_ctl, m, m.s_part, l, epl + 1, epl + 2, epl + 3, p.stages, rs, rph, w);
      stamp(p, l, 3);
      if (DENSE) publish(p.fa, vid, epl + 3);
      cons_sync();
    } else if (type == 4) {
#pragma unroll
      for (int o = NB; o <= 16; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if ((threadIdx.x & 31) < NB) { s_cv[threadIdx.x >> 5][ncol] = bv; s_ci[threadIdx.x >> 5][ncol] = bi; }
      cons_sync();                               // also orders every thread's logits stores before the release below
      if (threadIdx.x < p.B) {
        float v = s_cv[0][threadIdx.x];
        int i = s_ci[0][threadIdx.x];
#pragma unroll
        for (int ww = 1; ww < GV_WARPS; ++ww) {
          const float ov = s_cv[ww][threadIdx.x];
          const int oi = s_ci[ww][threadIdx.x];
          if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (do_sample) asm volatile("fence.acq_rel.gpu;" ::: "memory");   // sampling reads the logits after the candidate line
        ll_st16(p.cand_ll + static_cast<size_t>(threadIdx.x) * ncta + vid, __float_as_uint(v), static_cast<uint32_t>(i),
                ep0 + 8u * static_cast<unsigned>(p.layers) + 1);
      }
    }
  }
  stamp(p, max(p.layers - 1, 0), 29);             // this CTA's candidates are out
  if (vid == 0) {
    cons_sync();
    finalize_step(p, ncta, ep0 + 8u * static_cast<unsigned>(p.layers) + 1, s_hlo, s_hhi, w);
    stamp(p, max(p.layers - 1, 0), 30);           // next token chosen, next embedding written: the step is over
    if (p.dbg != nullptr && threadIdx.x == 0 && p.layers > 0) {
      const int si = *p.st.step - 1;              // finalize advanced the step counter
      if (si >= 0 && si < static_cast<int>(gridDim.x))
        p.dbg[static_cast<size_t>(2) * p.layers * 32 + static_cast<size_t>(si) * 32 + 4] = p.dbg[static_cast<size_t>(p.layers - 1) * 32 + 30] - p.dbg[0];
    }
  }
}

}  // namespace

// host side ---------------------------------------------------------------------------------------------------
static int ds_pick_ks(int KP) {
  for (int ks = 1024; ks >= 256; ks -= 256)
    if (KP % ks == 0) return ks;
  return 0;
}

size_t decode_stream_exchange_bytes(int H, int qkvw, int qd, int ffn, int num_sms, size_t* offs /*[16]*/) {
  // flag-in-data regions: x | qkv | attn | act | part_o | part_ml | cand; dense regions (batch >= 4): x | qkv | attn | act | 4 flag arrays;
  // every region 256-byte aligned
  size_t o = 0;
  auto put = [&](int i, size_t bytes) { offs[i] = o; o += (bytes + 255) & ~static_cast<size_t>(255); };
  put(0, static_cast<size_t>(GV_MAXB) * (H / 2) * 8);
  put(1, static_cast<size_t>(GV_MAXB) * (qkvw / 2) * 8);
  put(2, static_cast<size_t>(GV_MAXB) * (qd / 2) * 8);
  put(3, static_cast<size_t>(GV_MAXB) * (ffn / 2) * 8);
  put(4, static_cast<size_t>(num_sms) * DS_G * DS_HD * 8);
  put(5, static_cast<size_t>(num_sms) * DS_G * 16);
  put(6, static_cast<size_t>(GV_MAXB) * num_sms * 16);
  put(7, static_cast<size_t>(GV_MAXB) * H * 2);
  put(8, static_cast<size_t>(GV_MAXB) * qkvw * 2);
  put(9, static_cast<size_t>(GV_MAXB) * qd * 2);
  put(10, static_cast<size_t>(GV_MAXB) * ffn * 2);
  put(11, static_cast<size_t>(4) * num_sms * 4 * DS_FLAG_STRIDE);
  offs[12] = o;
  return o;
}

int decode_stream_layer_bytes() { return static_cast<int>(sizeof(DsLayer)); }

static int ds_plan(const DecodeStreamArgs& a, DsParams& p, size_t* smem_out) {
  const int B = a.B;
  TB_CHECK(B >= 1 && B <= GV_MAXB, "decode batch must be in [1, 8]");
  TB_CHECK(a.H % 1024 == 0 && a.H <= 4096, "fused decode step stages the whole hidden row (hidden % 1024 == 0, <= 4096)");
  TB_CHECK(a.layers <= DS_MAX_LAYERS, "fused decode step keeps the layer pointer table in shared memory (<= 48 layers)");
  TB_CHECK(a.heads == a.kvh * DS_G && a.qd == a.heads * DS_HD && a.page_size == DS_PAGE, "fused decode step is specialised for 128-dim heads, 4 q heads per kv head, 64-position pages");
  TB_CHECK(a.ffn % 4 == 0 && a.qkvw % 4 == 0 && a.head_rows % 2 == 0, "row counts must divide into exchange units");
  TB_CHECK(a.num_sms <= GV_CONS, "candidate merge handles <= 256 CTAs");
  // RMSNorm phases (QKV, gate/up, heads) stage the whole hidden row; the O / down projections take their input chunk by
  // chunk into warp-private slices (consume_korder), so the staging buffer is sized by the hidden row alone
  const int Ks[5] = {a.H, a.qd, a.H, a.ffn, a.H};
  for (int i = 0; i < 5; ++i) {
    TB_CHECK(Ks[i] % 256 == 0, "decode GEMV K must be a multiple of 256");
    p.kp[i] = Ks[i];
    p.ks[i] = ds_pick_ks(Ks[i]);
  }
  const int kpmax = a.H;
  const int nbc = B == 1 ? 1 : B == 2 ? 2 : B <= 4 ? 4 : 8;
  const int Ns[5] = {a.qkvw, a.H, 2 * a.ffn, a.H, a.head_rows};
  const int units[5] = {2, 2, 4, 2, 2};
  int max_tiles = 0;
  for (int i = 0; i < 5; ++i) {
    const int rows = ((Ns[i] / units[i] + a.num_sms - 1) / a.num_sms) * units[i] + units[i] + (Ns[i] % units[i]);
    const int t = (rows + GV_TILE - 1) / GV_TILE;
    if (t > max_tiles) max_tiles = t;
    if (i == 1 || i == 3) TB_CHECK(rows <= DS_RES_PITCH && t <= 3, "O / down projection: at most 3 row tiles and 40 residual rows per CTA");
  }
  // partial sums of a tile group: 28 KB, except at batch 4..8 where the region is sized by the attention scratch (20 KB) --
  // that is what leaves room for a 5th / 4th ring stage next to the 33 / 66 KB of staged activations
  int group_tiles = ((nbc >= 4 ? 20 : 28) * 1024) / (GV_WARPS * GV_TILE * nbc * 4);
  if (group_tiles > max_tiles) group_tiles = max_tiles;
  int nsplit = a.num_sms / (B * a.kvh);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 32) nsplit = 32;
  TB_CHECK(B * a.kvh * nsplit <= a.num_sms, "one attention work item per CTA");
  TB_CHECK((a.max_ctx / DS_PAGE + nsplit - 1) / nsplit + 1 <= 64, "attention work item exceeds the per-CTA page table (64 pages)");
  int ncomb = 1;
  while (ncomb * 2 <= nsplit && ncomb * 2 <= 16) ncomb *= 2;
  const size_t attn_scratch = (768 + DS_G * DS_HD) * 2 + (2 * DS_HD + GV_WARPS * DS_G * DS_HD + GV_WARPS * DS_G * 2) * 4;
  size_t part_bytes = static_cast<size_t>(group_tiles) * GV_WARPS * GV_TILE * nbc * 4;
  if (part_bytes < attn_scratch) part_bytes = attn_scratch;
  part_bytes = (part_bytes + 127) & ~static_cast<size_t>(127);
  const int xs_cap = kpmax + XS_PAD;
  const size_t fixed = ((static_cast<size_t>(B) * xs_cap * 2 + 127) & ~static_cast<size_t>(127)) + part_bytes + GV_MAXB * 4 * GV_WARPS * 4 +
                       2 * GV_MAXSTAGES * 8 + 256;
  const size_t stage = static_cast<size_t>(DS_STAGE_HALVES) * 2;
  const size_t budget = 220 * 1024;   // + 6.3 KB static shared memory (layer table, control blocks, finalize scratch) <= 227 KB
  TB_CHECK(fixed + 2 * stage <= budget, "decode step shared memory budget exceeded");
  int st = static_cast<int>((budget - fixed) / stage);
  if (st > GV_MAXSTAGES) st = GV_MAXSTAGES;
  if (a.max_stages > 1 && st > a.max_stages) st = a.max_stages;
  p.nsplit = nsplit;
  p.ncomb = ncomb;
  p.part_bytes = static_cast<int>(part_bytes);
  p.stages = st;
  p.xs_cap = xs_cap;
  p.group_tiles = group_tiles;
  *smem_out = fixed + st * stage;
  return 0;
}

template <int NB>
static int ds_launch(const DsParams& p, int grid, size_t smem, bool cooperative, cudaStream_t stream) {
  auto kern = decode_stream_kernel<NB>;
  static size_t set_smem = 0;
  if (smem > set_smem) {
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    set_smem = smem;
  }
  // every CTA must be resident at once (CTAs wait for each other's data): a cooperative launch makes the driver check it
  int per_sm = 0;
  TB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DS_THREADS, smem));
  TB_CHECK(per_sm >= 1, "fused decode step does not fit one CTA per SM");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = cooperative ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return 0;
}

// what a launch at batch a.B would look like (tests assert that the full-size code paths are the ones exercised)
int decode_stream_plan(const DecodeStreamArgs& a, int t, int32_t* out /*[8]*/) {
  DsParams p{};
  size_t smem = 0;
  if (ds_plan(a, p, &smem)) return 1;
  const int npages = (t + DS_PAGE - 1) / DS_PAGE;
  out[0] = p.nsplit; out[1] = p.ncomb; out[2] = (npages + p.nsplit - 1) / p.nsplit; out[3] = p.kp[3]; out[4] = p.ks[3];
  out[5] = p.stages; out[6] = p.group_tiles; out[7] = static_cast<int32_t>(smem);
  return 0;
}

int decode_stream_step(const DecodeStreamArgs& a, cudaStream_t stream) {
  DsParams p{};
  p.B = a.B; p.H = a.H; p.qkvw = a.qkvw; p.qd = a.qd; p.ffn = a.ffn; p.layers = a.layers; p.heads = a.heads; p.kvh = a.kvh;
  p.vocab = a.vocab; p.head_rows = a.head_rows; p.n_valid = a.n_valid; p.eps = a.eps;
  p.L = reinterpret_cast<const DsLayer*>(a.layer_table_dev);
  p.final_norm = a.final_norm; p.head_w = a.head_w; p.x = a.x; p.x_next = a.x;
  size_t offs[16];
  decode_stream_exchange_bytes(a.H, a.qkvw, a.qd, a.ffn, a.num_sms, offs);
  uint8_t* ex = static_cast<uint8_t*>(a.exchange);
  p.x_ll = reinterpret_cast<uint2*>(ex + offs[0]); p.qkv_ll = reinterpret_cast<uint2*>(ex + offs[1]);
  p.attn_ll = reinterpret_cast<uint2*>(ex + offs[2]); p.act_ll = reinterpret_cast<uint2*>(ex + offs[3]);
  p.part_o_ll = reinterpret_cast<uint2*>(ex + offs[4]); p.part_ml_ll = reinterpret_cast<uint4*>(ex + offs[5]);
  p.cand_ll = reinterpret_cast<uint4*>(ex + offs[6]);
  p.xd = reinterpret_cast<__half*>(ex + offs[7]); p.qkvd = reinterpret_cast<__half*>(ex + offs[8]);
  p.attnd = reinterpret_cast<__half*>(ex + offs[9]); p.actd = reinterpret_cast<__half*>(ex + offs[10]);
  p.fx = reinterpret_cast<unsigned*>(ex + offs[11]); p.fq = p.fx + a.num_sms * DS_FLAG_STRIDE; p.fa = p.fq + a.num_sms * DS_FLAG_STRIDE; p.fact = p.fa + a.num_sms * DS_FLAG_STRIDE;
  p.cos_tab = a.cos_tab; p.sin_tab = a.sin_tab; p.block_table = a.block_table; p.max_pages = a.max_pages;
  p.logits_out = a.logits_out; p.samp_logits = a.samp_logits;
  p.embed = a.embed; p.time_tab = a.time_tab; p.score_tab = a.score_tab; p.sync_tab = a.sync_tab;
  p.st = a.st; p.epoch_ctr = a.epoch_ctr; p.err = a.err; p.advance = a.advance; p.dbg = a.dbg; p.poll_sleep_ns = a.poll_sleep_ns; p.watchdog_ms = a.watchdog_ms > 0 ? a.watchdog_ms : 4000;
  size_t smem = 0;
  if (ds_plan(a, p, &smem)) return 1;
  int rc;
  if (a.B == 1) rc = ds_launch<1>(p, a.num_sms, smem, a.cooperative != 0, stream);
  else if (a.B == 2) rc = ds_launch<2>(p, a.num_sms, smem, a.cooperative != 0, stream);
  else if (a.B <= 4) rc = ds_launch<4>(p, a.num_sms, smem, a.cooperative != 0, stream);
  else rc = ds_launch<8>(p, a.num_sms, smem, a.cooperative != 0, stream);
  return rc;
}

}  // namespace tb
