// The whole decode step as ONE persistent kernel (one CTA per SM: 8 consumer warps + 1 producer warp + 1 poller warp):
//
//   for each of the 32 layers: RMSNorm+QKV -> RoPE/KV-append/attention -> O(+residual) -> RMSNorm+gate/up(SwiGLU) -> down(+residual)
//   then final RMSNorm + active-head logits + masked argmax, and (CTA 0) the token bookkeeping / head switch / next embedding.
//
// Why: per-op kernels leave the HBM pipe idle at every kernel boundary (launch ramp, activation staging, epilogue:
// 5-7 us per boundary x 162 boundaries per token against a 2.2 ms HBM floor).  Here the producer warp of every CTA
// streams that CTA's weight rows with cp.async.bulk into an mbarrier ring and NEVER waits for activations: while the
// consumer warps sit in a grid barrier or stage the next activation vector, the ring keeps filling with the next
// phase's weights, so the weight stream is continuous across phase and layer boundaries.  Phases are separated by a
// grid-wide barrier (atomic counter, consumers only); activations written by other CTAs are re-read with ld.global.cg.
//
// Structure (measured, profiles/r01_decode_ab.md): the body is ONE loop over the 4*layers+1 phases with one copy of each
// building block and the kernel is templated on the batch bucket -- the body is re-executed for every layer, so its SASS
// (146 KB; 354 KB when every phase was an inlined copy) has to stay in the instruction cache.  Consumers copy a ring stage
// into registers, release it, then run the mma.sync chain.  Optional TMEM weight reservoir (off by default): see below.
//
// Replaces: trace_mistral.py:114-264 forward (decoder layers via HF MistralModel, heads :190-200, mask :244-252),
// :336-344 head switch, trace_arch.py:345-375 embedding dispatch, HF greedy argmax -- for one generated token.
#include "common.cuh"
#include "decode_dev.cuh"
#include "kernels.h"

namespace tb {
namespace {

constexpr int MG_G = 4;       // query heads per kv head
constexpr int MG_HD = 128;
constexpr int MG_MAX_LAYERS = 48;

// ---- grid barrier (consumer warps only; the producer warp never synchronises with anything but its ring) ----
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// ------------------------------------------------------------------------------------------------
// TMEM weight reservoir.
//
// The shared-memory ring (4-5 stages x 33 KB) covers only ~3 us of this SM's share of the HBM stream; every grid
// barrier (+ epilogue + activation staging) is a 5-7 us hole in which the ring is full, the producer stalls and
// the HBM pipe drains.  Tensor memory (256 KB per SM) is otherwise unused by this kernel, so the consumer warps use
// it as 8 more stages: while they wait in a grid barrier they keep taking full stages out of the ring, park their
// OWN mma fragments of the stage in TMEM (tcgen05.st: thread-private lane x 32 columns per stage -> no layout
// contract with anybody) and hand the stage back to the producer, which keeps streaming.  After the barrier the
// parked stages are read back with tcgen05.ld and fed to the same mma.sync sequence as a ring stage.
//
// Determinism: a row tile is always accumulated in one register accumulator in k order, whether its stages come
// from TMEM, from the ring or from both; only the ORDER IN WHICH TILES ARE VISITED depends on how many stages
// happened to be parked, so results are bit-identical run to run.
// ------------------------------------------------------------------------------------------------
constexpr int TM_SLOTS = 8;            // parked stages per thread: 8 x 32 columns; warps w and w+4 share a lane quadrant
constexpr int TM_COLS = 512;

struct Stream {                        // consumer-side view of the weight stream (uniform over the 256 consumer threads)
  int s;                               // ring stage of the next item
  uint32_t ph;
  int tm_head, tm_count;               // FIFO of parked stages (always older than anything in the ring)
  int seq, seq_left;                   // phase sequence number of the ring-front item, items of it not yet taken from the ring
  int seq_end;                         // layers*4 + 1
  uint32_t tm_base;                    // this thread's TMEM address: lane quadrant of the warp, column half of the warp
  int enabled;
};
struct StreamTab {                     // per-CTA item geometry of the 5 phase types (qkv, o, gate/up, down, head)
  int cnt[5], ks[5];
  volatile int code[2];
  volatile unsigned arrived;           // consumers -> poller warp: barrier target this CTA has arrived at (0xffffffff = quit)
  volatile unsigned released;          // poller warp -> consumers: last barrier target seen complete
};
constexpr int MG_THREADS = GV_THREADS + 32;   // 8 consumer warps + producer warp + barrier poller warp
__device__ __forceinline__ int seq_type(int seq, int seq_end) { return seq == seq_end - 1 ? 4 : (seq & 3); }
__device__ __forceinline__ void stream_skip_empty(Stream& st, const StreamTab& tab) {
  while (st.seq_left == 0 && st.seq < st.seq_end) {
    ++st.seq;
    st.seq_left = st.seq < st.seq_end ? tab.cnt[seq_type(st.seq, st.seq_end)] : 0;
  }
}
__device__ __forceinline__ void stream_ring_advance(Stream& st, int stages) {
  if (++st.s == stages) { st.s = 0; st.ph ^= 1; }
}

// move the ring-front stage into TMEM (all 256 consumer threads; the caller has seen full[s] complete)
__device__ __forceinline__ void relay_one(Stream& st, const StreamTab& tab, const GvSmem& m, int stages) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int KS = tab.ks[seq_type(st.seq, st.seq_end)];
  const int ksw = KS / GV_WARPS, nch = ksw / 32, wpitch = KS + WS_PAD;
  mbar_wait(&m.full[st.s], st.ph);
  const __half* wa = m.ws + static_cast<size_t>(st.s) * m.ws_stride + static_cast<size_t>(g) * wpitch + warp * ksw + t * 8;
  const __half* wb = wa + 8 * wpitch;
  uint32_t r[32];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
    if (ch < nch) {
      a = *reinterpret_cast<const uint4*>(wa + ch * 32);
      b = *reinterpret_cast<const uint4*>(wb + ch * 32);
    }
    r[ch * 8 + 0] = a.x; r[ch * 8 + 1] = a.y; r[ch * 8 + 2] = a.z; r[ch * 8 + 3] = a.w;
    r[ch * 8 + 4] = b.x; r[ch * 8 + 5] = b.y; r[ch * 8 + 6] = b.z; r[ch * 8 + 7] = b.w;
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(&m.empty[st.s]);          // the stage is in registers: the producer may refill it
  tmem_st_32x32(st.tm_base + static_cast<uint32_t>((st.tm_head + st.tm_count) & (TM_SLOTS - 1)) * 32u, r);
  tmem_st_wait();
  stream_ring_advance(st, stages);
  ++st.tm_count;
  --st.seq_left;
  stream_skip_empty(st, tab);
}

__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target, Stream& st, StreamTab& tab, const GvSmem& m, int stages, bool allow) {
  cons_sync();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    // release-reduction: orders every write this CTA made before the bar.sync above (cumulativity) without a separate
    // membar.gl + atomic round trip; the acquire loads below pair with it
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
  }
  if (!st.enabled) {
    if (threadIdx.x == 0)
      while (ld_acquire_u32(bar) < target) { }
    cons_sync();
    return;
  }
  // wait for the other CTAs; meanwhile park arriving stages in TMEM so that the producer never stops streaming.
  // The barrier word itself is watched by the poller warp (an L2 round trip under full HBM load is 1-2 us: a blocking
  // poll between two parked stages would cap the relay at one stage per round trip); it reports through shared memory.
  if (threadIdx.x == 0) tab.arrived = target;
  int patience = st.enabled == 2 ? 4096 : 0;            // test mode: keep parking after the release until TMEM is full
  for (int it = 0;; ++it) {
    if (threadIdx.x == 0) {
      int code = 0;
      const bool can = allow && st.tm_count < TM_SLOTS && st.seq_left > 0;
      if (tab.released == target && (!can || --patience < 0)) code = 2;   // leave as soon as the grid is through
      else if (can && mbar_test(&m.full[st.s], st.ph)) code = 1;
      if (code == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");   // the poller's acquire, made ours before the CTA barrier
      tab.code[it & 1] = code;
    }
    cons_sync();
    const int code = tab.code[it & 1];
    if (code == 2) break;
    if (code == 1) relay_one(st, tab, m, stages);
  }
}

// the poller warp (one lane): for every barrier the consumers arrive at, wait for the grid and report it
__device__ __forceinline__ void barrier_poller(const unsigned* bar, StreamTab& tab) {
  unsigned done = 0;
  for (;;) {
    unsigned t;
    while ((t = tab.arrived) == done) { }
    if (t == 0xffffffffu) return;
    while (ld_acquire_u32(bar) < t) { }
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    tab.released = t;
    done = t;
  }
}

// One k-phase of one matrix on the consumer side.  The first n_tm = min(tm_count, items) items of the phase are parked
// in TMEM, the rest comes through the ring.  Every warp runs two tile streams with their own accumulators:
//   ring stream : tiles from the one that contains item n_tm onwards -- taken whenever the ring-front stage is full,
//                 so stages go back to the producer as early as possible and the HBM stream never stalls on us;
//   TMEM stream : tiles [0, ts) -- read back (tcgen05.ld) in the slack while the ring-front stage is still in flight.
// The parked head of the straddled tile is folded into the ring accumulator first.  A tile is always accumulated in
// k order in one accumulator, so the result does not depend on how many stages were parked or on the interleaving.
// Warps run decoupled inside a phase (each owns a K slice of every stage) and meet again at the caller's cons_sync.
template <bool FULL>
__device__ __forceinline__ void mma_item(float (&c)[4], const uint32_t (&w)[32], const __half* xp, bool xok, int nch) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (FULL || ch < nch) {
      const uint4 xv = xok ? *reinterpret_cast<const uint4*>(xp + ch * 32) : make_uint4(0, 0, 0, 0);
      mma_16816(c, w[ch * 8 + 0], w[ch * 8 + 4], w[ch * 8 + 1], w[ch * 8 + 5], xv.x, xv.y);
      mma_16816(c, w[ch * 8 + 2], w[ch * 8 + 6], w[ch * 8 + 3], w[ch * 8 + 7], xv.z, xv.w);
    }
  }
}
// the warp's K slice of one ring stage -> registers (FULL: KS = 1024, 4 chunks of 32 columns, no predicates)
template <bool FULL>
__device__ __forceinline__ void load_item(uint32_t (&w)[32], const __half* wa, const __half* wb, int nch) {
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
// Partial sums live in shared memory as [tile][warp][16 rows][NBC batch columns] (NBC = batch bucket 1/2/4/8): the
// m16n8 accumulator fragment of lane (g, t) holds rows g and g+8, columns 2t and 2t+1, of which only n < NBC are kept.
// (With 8 columns for every batch size the B=1 step spent 53 KB of shared memory on zeros -- one ring stage.)
template <int NBC>
__device__ __forceinline__ void flush_tile(float* p, const float (&c)[4], int g, int t, bool add) {
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
// sum of the 8 warps' partials for output idx = (tile*16 + row)*NBC + n
template <int NBC>
__device__ __forceinline__ float part_out(const float* s_part, int idx) {
  const int tile = idx / (GV_TILE * NBC), rn = idx % (GV_TILE * NBC);
  const float* p = s_part + tile * (GV_WARPS * GV_TILE * NBC) + rn;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < GV_WARPS; ++w) v += p[w * (GV_TILE * NBC)];
  return v;
}

template <int NBC>
__device__ __forceinline__ void consume_phase(Stream& st, const StreamTab& tab, int k0, int KP, int KS, int r0, int r1, int vlo,
                                              int vhi, const GvSmem& m, int B, int stages, int early_release) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntile = (r1 - r0 + GV_TILE - 1) / GV_TILE;
  const int ksw = KS / GV_WARPS;                  // columns of a stage owned by this warp (multiple of 32)
  const int nch = ksw / 32;
  const int wpitch = KS + WS_PAD;
  const int nks = KP / KS;
  const bool xok = g < B;
  int tlo = 0, thi = 0;                           // tiles with a needed row: [tlo, thi)
  if (vlo < vhi) { tlo = (vlo - r0) / GV_TILE; thi = (vhi - 1 - r0) / GV_TILE + 1; }
  if (k0 == 0) {
    for (int tile = 0; tile < ntile; ++tile) {
      if (tile >= tlo && tile < thi) continue;
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      flush_tile<NBC>(m.s_part + (static_cast<size_t>(tile) * GV_WARPS + warp) * (GV_TILE * NBC), z, g, t, false);
    }
  }
  // `st` lives in local memory (its address is passed to the shared barrier routine): keep the cursor in registers here
  int rs = st.s;
  uint32_t rph = st.ph;
  const uint32_t tm_base = st.tm_base;
  const int tm_head = st.tm_head;
  const int n_items = (thi - tlo) * nks;
  const int n_tm = min(st.tm_count, n_items);
  const int ts = n_tm / nks;                      // tiles [0, ts) come from TMEM only; tile ts starts in TMEM (maybe with 0 items)
  const int na = ts * nks;                        // items of the TMEM stream
  const __half* xbase = m.xs + g * (KP + XS_PAD) + warp * ksw + t * 8;
  float* pbase = m.s_part + (static_cast<size_t>(tlo) * GV_WARPS + warp) * (GV_TILE * NBC);
  const bool add = k0 > 0;
  float ca[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t w[32];
  // parked head of the straddled tile -> ring accumulator
  const bool full = nch == 4;
  for (int i = na; i < n_tm; ++i) {
    tmem_ld_32x32(tm_base + static_cast<uint32_t>((tm_head + i) & (TM_SLOTS - 1)) * 32u, w);
    tmem_ld_wait();
    mma_item<false>(cb, w, xbase + (i - na) * KS, xok, nch);
  }
  int ia = 0, ib = n_tm;
  int ja = 0, jb = n_tm - na;                     // k-stage index inside the current tile of each stream
  float* pa = pbase;                              // partial-sum block of the current tile of each stream
  float* pb = pbase + static_cast<size_t>(ts) * (GV_WARPS * GV_TILE * NBC);
  while (ia < na || ib < n_items) {
    bool ring;
    if (ib >= n_items) ring = false;
    else if (ia >= na) ring = true;
    else {
      int ok = 0;
      if (lane == 0) ok = mbar_test(&m.full[rs], rph) ? 1 : 0;
      ring = __shfl_sync(0xffffffffu, ok, 0) != 0;
    }
    if (ring) {
      mbar_wait(&m.full[rs], rph);
      const __half* wa = m.ws + static_cast<size_t>(rs) * m.ws_stride + static_cast<size_t>(g) * wpitch + warp * ksw + t * 8;
      const __half* wb = wa + 8 * wpitch;
      const __half* xp = xbase + jb * KS;
      // stage -> registers, hand the stage back to the producer (mbarrier.arrive is a release: it is ordered after the
      // shared-memory reads above), then the MMA chain: the chain is off the ring's refill cycle
      if (full) load_item<true>(w, wa, wb, nch);
      else load_item<false>(w, wa, wb, nch);
      if (early_release) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&m.empty[rs]);
      }
      if (full) mma_item<true>(cb, w, xp, xok, nch);
      else mma_item<false>(cb, w, xp, xok, nch);
      if (!early_release) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&m.empty[rs]);
      }
      if (++rs == stages) { rs = 0; rph ^= 1; }
      if (++jb == nks) {
        flush_tile<NBC>(pb, cb, g, t, add);
        cb[0] = cb[1] = cb[2] = cb[3] = 0.f;
        jb = 0;
        pb += GV_WARPS * GV_TILE * NBC;
      }
      ++ib;
    } else {
      tmem_ld_32x32(tm_base + static_cast<uint32_t>((tm_head + ia) & (TM_SLOTS - 1)) * 32u, w);
      tmem_ld_wait();
      mma_item<false>(ca, w, xbase + ja * KS, xok, nch);
      if (++ja == nks) {
        flush_tile<NBC>(pa, ca, g, t, add);
        ca[0] = ca[1] = ca[2] = ca[3] = 0.f;
        ja = 0;
        pa += GV_WARPS * GV_TILE * NBC;
      }
      ++ia;
    }
  }
  st.s = rs;
  st.ph = rph;
  st.tm_head = (tm_head + n_tm) & (TM_SLOTS - 1);
  st.tm_count -= n_tm;
  st.seq_left -= n_items - n_tm;                  // ring items of a phase are only taken once its parked items are used up
  stream_skip_empty(st, tab);
}

struct MegaLayer {
  const __half *ln1, *qkv_w, *o_w, *ln2, *gateup_w, *down_w;
  __half *k_pool, *v_pool;
};

struct MegaParams {
  int B, H, qkvw, qd, ffn, layers, heads, kvh, vocab, head_rows, n_valid;
  float eps;
  const MegaLayer* L;                 // device array [layers]
  const __half *final_norm, *head_w;
  __half *x, *qkv, *attn, *act;       // decode activations [B, .]
  const float *cos_tab, *sin_tab;
  const int32_t* block_table;
  int max_pages, page_size;
  float *part_o, *part_ml;
  int32_t* counters;
  int nsplit, chunk_max;
  float* cand_val;
  int32_t* cand_idx;
  float* logits_out;
  const __half *embed, *time_tab, *score_tab, *sync_tab;
  int eos, pad;
  DecodeState st;
  unsigned *gbar, *gexit;
  int kp_qkv, ks_qkv, kp_o, ks_o, kp_gu, ks_gu, kp_dn, ks_dn, kp_hd, ks_hd;
  unsigned long long* dbg;            // optional timeline [2 CTAs][layers][10] of %globaltimer stamps
  int pf_ahead;                       // (experimental, off) TMA L2 prefetch distance of the weight stream, in stages
  int part_bytes, stages, xs_cap;      // xs_cap: halves per batch row reserved for staged activations (max KP + pad)
  int relay_sites;                    // bit t: the barrier before phase type t parks stages (bit 5: the one before attention)
  int early_release;                  // hand a ring stage back before (1) or after (0) its MMA chain
  int max_inflight;                   // producer request throttle in stages (0 = the whole ring)
  int tm_relay;                       // park ring stages in TMEM while waiting in grid barriers (0 = off)
  int gate_mode;                      // producer waits for the attention K/V requests before: 1 the O weights, 2 the gate/up weights, 0 never
};

__device__ __forceinline__ void stamp(const MegaParams& p, int l, int k) {
  if (p.dbg != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[(static_cast<size_t>(blockIdx.x == 0 ? 0 : 1) * p.layers + l) * 32 + k] = t;
  }
}

__device__ __forceinline__ void stamp_val(const MegaParams& p, int l, int k, unsigned long long v) {
  if (p.dbg != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 77))
    p.dbg[(static_cast<size_t>(blockIdx.x == 0 ? 0 : 1) * p.layers + l) * 32 + k] = v;
}

// Per-step attention control block of this CTA's (fixed) work item: everything that would otherwise be a chain of
// dependent global loads inside every layer's attention phase (pos -> slot -> block table -> row address; RoPE row).
struct AttnCtl {
  int valid, b, kv, split, t, start, end, page_new;
  int pages[64];            // page id of key range page i (i = key/page_size - start/page_size)
  float cs[MG_HD / 2], sn[MG_HD / 2];
};
__device__ __forceinline__ void attn_ctl_init(const MegaParams& p, AttnCtl* c) {
  const int items = p.B * p.kvh * p.nsplit;
  const int it = blockIdx.x;
  if (threadIdx.x == 0) {
    c->valid = it < items;
    if (it < items) {
      c->split = it % p.nsplit; c->kv = (it / p.nsplit) % p.kvh; c->b = it / (p.nsplit * p.kvh);
      c->t = p.st.pos[c->b];
      int chunk = (c->t + p.nsplit - 1) / p.nsplit;
      chunk = (chunk + 31) & ~31;
      c->start = c->split * chunk;
      c->end = min(c->start + chunk, c->t);
    }
  }
  cons_sync();
  if (c->valid) {
    const int32_t* table = p.block_table + static_cast<size_t>(p.st.slot_ids[c->b]) * p.max_pages;
    const int p0 = c->start / p.page_size;
    const int np = (c->end > c->start) ? (c->end - 1) / p.page_size - p0 + 1 : 0;
    // unused entries point at page 0 (always mapped): predicated-off lanes still form addresses from pages[0]
    for (int i = threadIdx.x; i < 64; i += GV_CONS) c->pages[i] = (i < np) ? table[p0 + i] : 0;
    if (threadIdx.x == 64) c->page_new = table[c->t / p.page_size];
    if (threadIdx.x < MG_HD / 2) {
      c->cs[threadIdx.x] = p.cos_tab[static_cast<size_t>(c->t) * (MG_HD / 2) + threadIdx.x];
      c->sn[threadIdx.x] = p.sin_tab[static_cast<size_t>(c->t) * (MG_HD / 2) + threadIdx.x];
    }
  }
  cons_sync();
}

// ------------------------------------------------------------------------------------------------
// attention work item (b, kv, split) on the 8 consumer warps; scratch lives in the s_part/xs region
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void attn_item(const MegaParams& p, const MegaLayer& L, const AttnCtl& ctl, float* scratch, int dl,
                                          volatile int* gate) {
  // Warp-autonomous split-KV attention: every warp owns 16 keys per pass, keeps an online-softmax state (m, l, o) for
  // the 4 query heads of this kv head in registers and only meets the other warps once, for the CTA-level merge.
  //   scores : 2 lanes per key (64 dims each, 8 x 16-byte loads), pair-reduced with one shuffle
  //   PV     : lane <-> 4 dims, the 16 probabilities of the pass are broadcast with shuffles
  // All K and V rows of a pass are requested up front (one memory round trip per pass).
  const int b = ctl.b, kv = ctl.kv, split = ctl.split;
  float* qs = scratch;                                   // [4][128]  (pre-scaled, RoPE'd q)
  float* knew = qs + MG_G * MG_HD;                       // [128]
  float* vnew = knew + MG_HD;                            // [128]
  float* red_o = vnew + MG_HD;                           // [8 warps][4][128]
  float* red_ml = red_o + GV_WARPS * MG_G * MG_HD;       // [8 warps][4][2]
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int heads = p.heads, kvh = p.kvh, nsplit = p.nsplit, page_size = p.page_size;
  const int t = ctl.t;
  const int pg0 = ctl.start / page_size;
  const __half* row = p.qkv + static_cast<size_t>(b) * p.qkvw;
  const float scale = rsqrtf(static_cast<float>(MG_HD));
  const float* cs = ctl.cs;
  const float* sn = ctl.sn;
  // The cached K/V rows of the first pass do not depend on this step's q/k/v: request them BEFORE the q load + RoPE, so the
  // two memory round trips of the phase overlap instead of following each other.
  const int start = ctl.start;
  const int nloc = max(ctl.end - start, 0);
  const int hf = lane & 1;
  uint4 kreg[8];
  uint2 vreg[16];
  auto load_kv = [&](int base) {
    // K half row of key (warp*16 + lane/2), V 4-dim slices of the warp's 16 keys
    const int jk = base + warp * 16 + (lane >> 1);
    const int key = start + (jk < nloc ? jk : 0);
    const __half* kr = L.k_pool + ((static_cast<size_t>(ctl.pages[key / page_size - pg0]) * kvh + kv) * page_size + key % page_size) * MG_HD + hf * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) kreg[c] = *reinterpret_cast<const uint4*>(kr + c * 8);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = base + warp * 16 + u;
      const int keyv = start + (j < nloc ? j : 0);
      vreg[u] = *reinterpret_cast<const uint2*>(L.v_pool + ((static_cast<size_t>(ctl.pages[keyv / page_size - pg0]) * kvh + kv) * page_size + keyv % page_size) * MG_HD + lane * 4);
    }
  };
  load_kv(0);
  for (int i = tid; i < MG_G * (MG_HD / 2); i += GV_CONS) {
    const int h = i / (MG_HD / 2), d = i % (MG_HD / 2);
    const __half* qp = row + (kv * MG_G + h) * MG_HD;
    const float a = __half2float(__ldcg(qp + d)), c2 = __half2float(__ldcg(qp + d + MG_HD / 2));
    qs[h * MG_HD + d] = __half2float(__float2half_rn(a * cs[d] - c2 * sn[d])) * scale;
    qs[h * MG_HD + d + MG_HD / 2] = __half2float(__float2half_rn(c2 * cs[d] + a * sn[d])) * scale;
  }
  if (split == 0) {
    const int page = ctl.page_new, poff = t % page_size;
    const size_t dst = ((static_cast<size_t>(page) * kvh + kv) * page_size + poff) * MG_HD;
    if (tid < MG_HD / 2) {
      const __half* kp = row + (heads + kv) * MG_HD;
      const float a = __half2float(__ldcg(kp + tid)), c2 = __half2float(__ldcg(kp + tid + MG_HD / 2));
      const __half ra = __float2half_rn(a * cs[tid] - c2 * sn[tid]);
      const __half rb = __float2half_rn(c2 * cs[tid] + a * sn[tid]);
      L.k_pool[dst + tid] = ra;
      L.k_pool[dst + tid + MG_HD / 2] = rb;
      knew[tid] = __half2float(ra);
      knew[tid + MG_HD / 2] = __half2float(rb);
    } else if (tid < MG_HD) {
      const int d = (tid - MG_HD / 2) * 2;
      const __half* vp = row + (heads + kvh + kv) * MG_HD;
      const unsigned u = __ldcg(reinterpret_cast<const unsigned*>(vp + d));
      *reinterpret_cast<unsigned*>(L.v_pool + dst + d) = u;
      const float2 f = unpack_half2(u);
      vnew[d] = f.x;
      vnew[d + 1] = f.y;
    }
  }
  float m_w[MG_G], l_w[MG_G], o_w[MG_G][4];
#pragma unroll
  for (int h = 0; h < MG_G; ++h) { m_w[h] = -INFINITY; l_w[h] = 0.f; o_w[h][0] = o_w[h][1] = o_w[h][2] = o_w[h][3] = 0.f; }
  bool first = true;
  for (int base = 0; base < nloc || first; base += GV_WARPS * 16) {
    const int jk = base + warp * 16 + (lane >> 1);
    const bool kvalid = jk < nloc;
    if (base > 0) load_kv(base);              // pass 0 was requested before the q load
    // the LAST pass's latency-critical requests are out: the producer may start the O weights.  (With several passes per
    // CTA -- batch 4..8: 500-1000 keys per work item -- opening the gate after the first pass put every later pass behind
    // the weight stream's 3 us queue: 5.5 us per pass instead of ~1.5.)
    if (tid == 0 && base + GV_WARPS * 16 >= nloc) *gate = dl + 1;
    if (first) {
      cons_sync();                            // qs / knew / vnew visible
      stamp(p, dl, 20);
      first = false;
    }
    // ---- scores of the pair's key ----
    float acc[MG_G] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float kf[8];
      float2 p0 = unpack_half2(kreg[c].x), p1 = unpack_half2(kreg[c].y), p2 = unpack_half2(kreg[c].z), p3 = unpack_half2(kreg[c].w);
      kf[0] = p0.x; kf[1] = p0.y; kf[2] = p1.x; kf[3] = p1.y; kf[4] = p2.x; kf[5] = p2.y; kf[6] = p3.x; kf[7] = p3.y;
#pragma unroll
      for (int h = 0; h < MG_G; ++h) {
        const float4 q0 = *reinterpret_cast<const float4*>(qs + h * MG_HD + hf * 64 + c * 8);
        const float4 q1 = *reinterpret_cast<const float4*>(qs + h * MG_HD + hf * 64 + c * 8 + 4);
        acc[h] += q0.x * kf[0] + q0.y * kf[1] + q0.z * kf[2] + q0.w * kf[3] + q1.x * kf[4] + q1.y * kf[5] + q1.z * kf[6] + q1.w * kf[7];
      }
    }
#pragma unroll
    for (int h = 0; h < MG_G; ++h) {
      acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 1);
      if (!kvalid) acc[h] = -INFINITY;
      // online softmax update with the warp's 16 keys
      float mx = acc[h];
#pragma unroll
      for (int o = 2; o <= 16; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float mn = fmaxf(m_w[h], mx);
      const float pr = (mn == -INFINITY) ? 0.f : __expf(acc[h] - mn);       // both lanes of a pair hold the same value
      const float corr = (mn == -INFINITY) ? 1.f : __expf(m_w[h] - mn);
      float ps = hf == 0 ? pr : 0.f;
#pragma unroll
      for (int o = 1; o <= 16; o <<= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l_w[h] = l_w[h] * corr + ps;
      m_w[h] = mn;
      o_w[h][0] *= corr; o_w[h][1] *= corr; o_w[h][2] *= corr; o_w[h][3] *= corr;
      acc[h] = pr;
    }
    // ---- O += P V : lane <-> 4 dims, probabilities broadcast from the lane pair that owns the key ----
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float2 v0 = unpack_half2(vreg[u].x), v1 = unpack_half2(vreg[u].y);
#pragma unroll
      for (int h = 0; h < MG_G; ++h) {
        const float pr = __shfl_sync(0xffffffffu, acc[h], 2 * u);
        o_w[h][0] += pr * v0.x; o_w[h][1] += pr * v0.y; o_w[h][2] += pr * v1.x; o_w[h][3] += pr * v1.y;
      }
    }
  }
  if (split == 0 && warp == 0) {   // the token being decoded attends to itself: key/value from shared memory, not the cache
#pragma unroll
    for (int h = 0; h < MG_G; ++h) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) d += qs[h * MG_HD + lane * 4 + i] * knew[lane * 4 + i];
      d = warp_sum(d);
      const float mn = fmaxf(m_w[h], d);
      const float pr = __expf(d - mn), corr = __expf(m_w[h] - mn);
      l_w[h] = l_w[h] * corr + pr;
      m_w[h] = mn;
#pragma unroll
      for (int i = 0; i < 4; ++i) o_w[h][i] = o_w[h][i] * corr + pr * vnew[lane * 4 + i];
    }
  }
  stamp(p, dl, 21);
  // ---- CTA-level merge of the 8 warps ----
#pragma unroll
  for (int h = 0; h < MG_G; ++h) {
    *reinterpret_cast<float4*>(red_o + (warp * MG_G + h) * MG_HD + lane * 4) = make_float4(o_w[h][0], o_w[h][1], o_w[h][2], o_w[h][3]);
    if (lane == 0) { red_ml[(warp * MG_G + h) * 2] = m_w[h]; red_ml[(warp * MG_G + h) * 2 + 1] = l_w[h]; }
  }
  cons_sync();
  stamp(p, dl, 22);
  const size_t pbase = (static_cast<size_t>(b) * kvh + kv) * nsplit + split;
  if (tid < 128) {
    const int h = tid >> 5;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < GV_WARPS; ++w) M = fmaxf(M, red_ml[(w * MG_G + h) * 2]);
    float Ls = 0.f;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < GV_WARPS; ++w) {
      const float mw = red_ml[(w * MG_G + h) * 2], lw = red_ml[(w * MG_G + h) * 2 + 1];
      if (lw != 0.f) {
        const float wt = __expf(mw - M);
        Ls += wt * lw;
        const float4 x4 = *reinterpret_cast<const float4*>(red_o + (w * MG_G + h) * MG_HD + (tid & 31) * 4);
        a.x += wt * x4.x; a.y += wt * x4.y; a.z += wt * x4.z; a.w += wt * x4.w;
      }
    }
    __stcg(reinterpret_cast<float4*>(p.part_o + pbase * (MG_G * MG_HD) + tid * 4), a);
    if ((tid & 31) == 0) {
      __stcg(p.part_ml + pbase * (MG_G * 2) + h * 2, M);
      __stcg(p.part_ml + pbase * (MG_G * 2) + h * 2 + 1, Ls);
    }
  }
  stamp(p, dl, 23);
  cons_sync();
  if (tid == 0) {   // acq_rel: releases this CTA's partials (cumulative over the bar.sync) and acquires the others'
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + b * kvh + kv) : "memory");
    s_last = (old == nsplit - 1);
  }
  cons_sync();
  stamp(p, dl, 24);
  if (s_last) {     // the last split of this (row, kv head) combines the partials -> attn [B, heads*128] fp16
    if (tid < 128) {
      const size_t gb = (static_cast<size_t>(b) * kvh + kv) * nsplit;
      const int h = tid >> 5;
      float2 ml[20];
#pragma unroll
      for (int s = 0; s < 20; ++s)
        ml[s] = (s < nsplit) ? __ldcg(reinterpret_cast<const float2*>(p.part_ml + (gb + s) * (MG_G * 2) + h * 2)) : make_float2(-INFINITY, 0.f);
      float M = -INFINITY;
#pragma unroll
      for (int s = 0; s < 20; ++s) M = fmaxf(M, ml[s].x);
      float Lsum = 0.f;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half * 10 < nsplit) {
          float4 po[10];
#pragma unroll
          for (int s = 0; s < 10; ++s)
            po[s] = (half * 10 + s < nsplit) ? __ldcg(reinterpret_cast<const float4*>(p.part_o + (gb + half * 10 + s) * (MG_G * MG_HD) + tid * 4))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int s = 0; s < 10; ++s) {
            const float2 m2 = ml[half * 10 + s];
            if (half * 10 + s < nsplit && m2.y != 0.f) {
              const float w = __expf(m2.x - M);
              Lsum += w * m2.y;
              acc.x += w * po[s].x; acc.y += w * po[s].y; acc.z += w * po[s].z; acc.w += w * po[s].w;
            }
          }
        }
      }
      const float inv = 1.f / Lsum;
      uint2 pk;
      pk.x = pack_half2(acc.x * inv, acc.y * inv);
      pk.y = pack_half2(acc.z * inv, acc.w * inv);
      __stcg(reinterpret_cast<uint2*>(p.attn + static_cast<size_t>(b) * heads * MG_HD + (kv * MG_G) * MG_HD + tid * 4), pk);
    }
    if (tid == 0) p.counters[b * kvh + kv] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// finalize (CTA 0, 256 consumer threads)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void finalize_step(const MegaParams& p, int ncand) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_tok[GV_MAXB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const DecodeState& st = p.st;
  const int step = *st.step;
  for (int b = 0; b < p.B; ++b) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < ncand; i += GV_CONS) {
      const float v = __ldcg(p.cand_val + b * ncand + i);
      const int ix = __ldcg(p.cand_idx + b * ncand + i);
      if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    cons_sync();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (s_v[w] > bv || (s_v[w] == bv && s_i[w] < bi)) { bv = s_v[w]; bi = s_i[w]; }
      const int tok = st.finished[b] ? p.pad : bi;
      if (step < st.max_new) st.out_ids[static_cast<int64_t>(b) * st.max_new + step] = tok;
      if (!st.finished[b] && tok == p.eos) st.finished[b] = 1;
      if (st.heads[b] >= 0) {
        if (tok == p.vocab) st.heads[b] = 1;
        else if (tok == p.vocab + 1) st.heads[b] = 2;
        else if (tok == p.vocab + 14) st.heads[b] = 0;
      }
      st.pos[b] += 1;
      s_tok[b] = tok;
    }
    cons_sync();
  }
  if (tid == 0) {
    int all = 1;
    for (int b = 0; b < p.B; ++b) all &= st.finished[b];
    *st.step = step + 1;
    if (all && *st.done_step == 0) *st.done_step = step + 1;
  }
  for (int b = 0; b < p.B; ++b) {
    const int tok = s_tok[b];
    const __half* src;
    if (tok == p.vocab) src = p.sync_tab;
    else if (tok > p.vocab && tok < p.vocab + 14) src = p.time_tab + static_cast<size_t>(tok - p.vocab - 1) * p.H;
    else if (tok >= p.vocab + 14) src = p.score_tab + static_cast<size_t>(tok - p.vocab - 14) * p.H;
    else src = p.embed + static_cast<size_t>(tok % p.vocab) * p.H;
    for (int i = tid; i < p.H / 8; i += GV_CONS)
      reinterpret_cast<uint4*>(p.x + static_cast<size_t>(b) * p.H)[i] = reinterpret_cast<const uint4*>(src)[i];
  }
}

// Pull this CTA's K/V slice of layer L towards L2 while the other CTAs are still finishing the QKV phase: the cached
// rows do not depend on this step's q/k/v, so their HBM latency is taken off the attention critical path.
__device__ __forceinline__ void attn_prefetch(const MegaParams& p, const MegaLayer& L, const AttnCtl& c) {
  if (!c.valid) return;
  const int pg0 = c.start / p.page_size;
  // one 128-byte line per thread per pass: 2 lines per key row (256 B), K then V
  for (int i = threadIdx.x; i < (c.end - c.start) * 2; i += GV_CONS) {
    const int key = c.start + (i >> 1);
    const size_t off = ((static_cast<size_t>(c.pages[key / p.page_size - pg0]) * p.kvh + c.kv) * p.page_size + key % p.page_size) * MG_HD + (i & 1) * 64;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(L.k_pool + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(L.v_pool + off));
  }
}

// Phase geometry, indexed by phase type (0 qkv, 1 o, 2 gate/up, 3 down, 4 head).  The kernel body is ONE loop over the
// 4*layers+1 phases of the step with one copy of every building block: the step's code (~350 KB of SASS when every
// phase had its own inlined copy) is re-executed for each of the 32 layers, and anything that does not fit the
// instruction cache is re-fetched from L2 32 times per token in exactly the latency-bound sections between the
// streaming loops.
struct PhaseGeo { int N, K, KP, KS; };
__device__ __forceinline__ PhaseGeo phase_geo(const MegaParams& p, int type) {
  PhaseGeo g;
  switch (type) {
    case 0: g.N = p.qkvw; g.K = p.H; g.KP = p.kp_qkv; g.KS = p.ks_qkv; break;
    case 1: g.N = p.H; g.K = p.qd; g.KP = p.kp_o; g.KS = p.ks_o; break;
    case 2: g.N = 2 * p.ffn; g.K = p.H; g.KP = p.kp_gu; g.KS = p.ks_gu; break;
    case 3: g.N = p.H; g.K = p.ffn; g.KP = p.kp_dn; g.KS = p.ks_dn; break;
    default: g.N = p.head_rows; g.K = p.H; g.KP = p.kp_hd; g.KS = p.ks_hd; break;
  }
  return g;
}
__device__ __forceinline__ const __half* phase_weights(const MegaParams& p, const MegaLayer& L, int type) {
  return type == 0 ? L.qkv_w : type == 1 ? L.o_w : type == 2 ? L.gateup_w : type == 3 ? L.down_w : p.head_w;
}

// epilogue of a layer phase (256 consumer threads, after the partials are complete): epi 0 store, 1 residual add in
// place, 2 SwiGLU pairs
template <int NBC>
__device__ __forceinline__ void phase_epilogue(const GvSmem& m, int r0, int r1, int B, __half* out, int ldo, int epi) {
  const int nout = ((r1 - r0 + GV_TILE - 1) / GV_TILE) * (GV_TILE * NBC);
  // warp-uniform trip count (nout is only a multiple of 16*NBC): the SwiGLU partner exchange is a full-warp shuffle
  for (int base = 0; base < nout; base += GV_CONS) {
    const int idx = base + threadIdx.x;
    const bool in = idx < nout;
    const float v = in ? part_out<NBC>(m.s_part, idx) : 0.f;
    const int row = r0 + idx / NBC, n = idx % NBC;
    const float partner = __shfl_xor_sync(0xffffffffu, v, NBC);     // the next row of the same batch column
    if (!in) continue;
    if (epi == 2) {
      if ((row & 1) == 0 && n < B && row < r1) __stcg(out + static_cast<size_t>(n) * ldo + (row >> 1), __float2half_rn(silu_f(v) * partner));
    } else if (n < B && row < r1) {
      __half* o = out + static_cast<size_t>(n) * ldo + row;
      __stcg(o, __float2half_rn(epi == 1 ? v + __half2float(__ldcg(o)) : v));
    }
  }
}

// One copy of the barrier + relay code for all call sites.  `m` is passed BY VALUE: taking its address would force the
// kernel's copy into local memory and every shared-memory access derived from it (the whole consume loop) would be
// compiled as a generic LD/ST instead of LDS/STS.
__device__ __noinline__ void grid_sync_shared(unsigned* bar, unsigned& target, Stream& st, StreamTab& tab, GvSmem m, int stages, bool allow) {
  grid_sync(bar, target, st, tab, m, stages, allow);
}

template <int NB>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_step_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // carve with the largest stage (KS = 1024) and the reserved activation capacity
  GvSmem m;
  {
    size_t off = 0;
    m.ws = reinterpret_cast<__half*>(smem + off); off += static_cast<size_t>(p.stages) * GV_TILE * (1024 + WS_PAD) * 2;
    m.xs = reinterpret_cast<__half*>(smem + off); off += static_cast<size_t>(p.B) * p.xs_cap * 2;
    m.s_part = reinterpret_cast<float*>(smem + off); off += static_cast<size_t>(p.part_bytes);
    m.s_red = reinterpret_cast<float*>(smem + off); off += (GV_WARPS * GV_MAXB + GV_MAXB) * 4;
    m.full = reinterpret_cast<uint64_t*>(smem + off); off += GV_MAXSTAGES * 8;
    m.empty = reinterpret_cast<uint64_t*>(smem + off);
    m.ws_stride = GV_TILE * (1024 + WS_PAD);   // fixed stage stride: phases with a smaller KS must not alias other stages
  }
  __shared__ volatile int s_gate;     // layer index + 1 whose attention loads have been issued (consumers -> producer)
  __shared__ uint32_t s_tmem;
  __shared__ StreamTab s_tab;
  __shared__ int s_hlo[GV_MAXB], s_hhi[GV_MAXB];
  // the layer pointer table, read once: a dependent L2 round trip per phase (producer: before the first copy of the phase;
  // consumers: before the gamma prefetch) would sit on the critical path 129 times per token
  __shared__ MegaLayer s_layers[MG_MAX_LAYERS];
  for (int i = threadIdx.x; i < p.layers * static_cast<int>(sizeof(MegaLayer) / 8); i += MG_THREADS)
    reinterpret_cast<unsigned long long*>(s_layers)[i] = reinterpret_cast<const unsigned long long*>(p.L)[i];
  if (threadIdx.x == 0) { s_gate = 0; s_tab.arrived = 0; s_tab.released = 0; }
  if (p.tm_relay && threadIdx.x < 32) tmem_alloc(&s_tmem, TM_COLS);   // whole TMEM: one CTA per SM by shared-memory footprint
  // active-head row ranges: read ONCE, before anything of this step can rewrite heads[] (CTA 0's finalize does, and a
  // late producer warp must not see the new values or it would stream rows that nobody consumes)
  if (threadIdx.x < GV_MAXB) {
    int lo = 0, hi = 0;
    if (threadIdx.x < p.B) head_range(p.st.heads[threadIdx.x], p.vocab, lo, hi);
    s_hlo[threadIdx.x] = lo;
    s_hhi[threadIdx.x] = hi;
  }
  tc_fence_before();
  gv_init_barriers(m, p.stages);
  tc_fence_after();
  const int ncta = gridDim.x;
  int ulo = 0x7fffffff, uhi = 0;
  for (int b = 0; b < p.B; ++b) { ulo = min(ulo, s_hlo[b]); uhi = max(uhi, s_hhi[b]); }
  const int nseq = p.layers * 4 + 1;            // phases of the step: 4 per layer + head

  if (threadIdx.x >= GV_THREADS) {
    // ===================== poller warp: watches the grid barrier word while the consumers park stages =====================
    if (p.tm_relay && threadIdx.x == GV_THREADS) barrier_poller(p.gbar, s_tab);
    return;
  }
  if (threadIdx.x >= GV_CONS) {
    // ===================== producer warp: the whole step's weight stream =====================
    int s = 0;
    uint32_t ph = 0;
    for (int seq = 0; seq < nseq; ++seq) {
      const int type = seq == nseq - 1 ? 4 : (seq & 3), l = seq >> 2;
      const MegaLayer& L = s_layers[type == 4 ? 0 : l];
      if (type == p.gate_mode && type != 0) {       // gate_mode 1: before the O weights, 2: before the gate/up weights
        if ((threadIdx.x & 31) == 0)
          while (s_gate < l + 1) { }
        __syncwarp();
      }
      const PhaseGeo g = phase_geo(p, type);
      int r0, r1;
      cta_rows(g.N, r0, r1);
      // head: only the rows some batch row decodes from (heads[] was written by the previous step)
      const int vlo = type == 4 ? max(ulo, r0) : r0, vhi = type == 4 ? min(uhi, r1) : r1;
      if (vlo < vhi) gemv_producer(phase_weights(p, L, type), g.K, g.KP, g.KS, r0, r1, vlo, vhi, m, p.stages, s, ph, p.pf_ahead, p.max_inflight);
    }
    return;
  }

  // ===================== consumer warps =====================
  __shared__ AttnCtl s_ctl;
  AttnCtl* ctl = &s_ctl;
  attn_ctl_init(p, ctl);
  unsigned target = 0;
  // ---- weight-stream bookkeeping: items per phase type for THIS CTA (must mirror the producer's enumeration) ----
  if (threadIdx.x < 5) {
    const PhaseGeo g = phase_geo(p, threadIdx.x);
    int r0, r1;
    cta_rows(g.N, r0, r1);
    int nt = (r1 - r0 + GV_TILE - 1) / GV_TILE;
    if (threadIdx.x == 4) {
      const int vlo = max(ulo, r0), vhi = min(uhi, r1);
      nt = vlo < vhi ? (vhi - 1 - r0) / GV_TILE - (vlo - r0) / GV_TILE + 1 : 0;
    }
    s_tab.cnt[threadIdx.x] = nt * (g.K / g.KS);
    s_tab.ks[threadIdx.x] = g.KS;
  }
  cons_sync();
  Stream st;
  st.s = 0; st.ph = 0; st.tm_head = 0; st.tm_count = 0;
  st.seq_end = nseq;
  st.seq = 0; st.seq_left = s_tab.cnt[0];
  stream_skip_empty(st, s_tab);
  st.enabled = p.tm_relay;
  st.tm_base = p.tm_relay ? s_tmem + (static_cast<uint32_t>(32 * ((threadIdx.x >> 5) & 3)) << 16) + static_cast<uint32_t>((threadIdx.x >> 7) * (TM_SLOTS * 32)) : 0u;

  for (int seq = 0; seq < nseq; ++seq) {
    const int type = seq == nseq - 1 ? 4 : (seq & 3), l = type == 4 ? p.layers - 1 : (seq >> 2);
    const MegaLayer& L = s_layers[l];
    const PhaseGeo g = phase_geo(p, type);
    const bool pro = (type != 1 && type != 3);                                  // RMSNorm fused into the staging
    // weights needed right after the barrier: issued before it so that their latency hides behind the wait
    GammaRegs gm;
    if (pro) gm = load_gamma(type == 0 ? L.ln1 : type == 2 ? L.ln2 : p.final_norm, p.H);
    if (type == 0) attn_prefetch(p, L, *ctl);
    if (seq > 0) grid_sync_shared(p.gbar, target, st, s_tab, m, p.stages, (p.relay_sites >> type) & 1);   // the phase's input vector is complete
    stamp(p, l, type * 2 + (type > 0 ? 2 : 0));
    int r0, r1;
    cta_rows(g.N, r0, r1);
    const int vlo = type == 4 ? max(ulo, r0) : r0, vhi = type == 4 ? min(uhi, r1) : r1;
    const __half* xin = type == 1 ? p.attn : type == 3 ? p.act : p.x;
    if (vlo < vhi) {
      for (int k0 = 0; k0 < g.K; k0 += g.KP) {
        if (k0 > 0) cons_sync();
        if (pro) stage_x_rmsnorm_nb<NB>(m.xs, xin, g.K, p.B, g.K, gm, p.eps, m.s_red);
        else stage_x_plain(m.xs, xin, g.K, p.B, k0, g.KP);
        cons_sync();
        if (k0 == 0) { stamp(p, l, 10 + type * 2); stamp_val(p, l, 25 + type, static_cast<unsigned long long>(st.tm_count)); }
        consume_phase<NB>(st, s_tab, k0, g.KP, g.KS, r0, r1, vlo, vhi, m, p.B, p.stages, p.early_release);
      }
      cons_sync();
      stamp(p, l, 11 + type * 2);
    }
    if (type < 4) {
      __half* out = type == 0 ? p.qkv : type == 2 ? p.act : p.x;
      phase_epilogue<NB>(m, r0, r1, p.B, out, type == 0 ? p.qkvw : type == 2 ? p.ffn : p.H, type == 2 ? 2 : (type == 0 ? 0 : 1));
      stamp(p, l, type * 2 + (type > 0 ? 3 : 1));
      if (type == 0) {
        grid_sync_shared(p.gbar, target, st, s_tab, m, p.stages, (p.relay_sites >> 5) & 1);   // q, k, v complete
        stamp(p, l, 2);
        if (ctl->valid) attn_item(p, L, *ctl, m.s_part, l, &s_gate);
        else if (threadIdx.x == 0) s_gate = l + 1;
        stamp(p, l, 3);
      }
    } else {
      // ---- heads: logits of the active ranges (fp16-rounded like the reference's fp16 GEMM) + argmax candidates ----
      const int n = threadIdx.x % NB;          // 256 % NB == 0: the batch column of a thread is fixed
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      const int nout = ((r1 - r0 + GV_TILE - 1) / GV_TILE) * (GV_TILE * NB);
      __shared__ float s_cv[GV_WARPS][8];
      __shared__ int s_ci[GV_WARPS][8];
      if (vlo < vhi) {
        const int lo = s_hlo[n], hi = s_hhi[n];
        for (int idx = threadIdx.x; idx < nout; idx += GV_CONS) {
          const int r = r0 + idx / NB;
          float v = __half2float(__float2half_rn(part_out<NB>(m.s_part, idx)));
          const bool act = (n < p.B) && r < r1 && r >= lo && r < hi;
          if (!act) v = -INFINITY;
          if (p.logits_out != nullptr && n < p.B && r < r1 && r < p.n_valid) p.logits_out[static_cast<size_t>(n) * p.n_valid + r] = v;
          if (act && (v > bv || (v == bv && r < bi))) { bv = v; bi = r; }
        }
      } else if (p.logits_out != nullptr) {
        for (int idx = threadIdx.x; idx < nout; idx += GV_CONS) {
          const int r = r0 + idx / NB;
          if (n < p.B && r < r1 && r < p.n_valid) p.logits_out[static_cast<size_t>(n) * p.n_valid + r] = -INFINITY;
        }
      }
#pragma unroll
      for (int o = NB; o <= 16; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if ((threadIdx.x & 31) < NB) { s_cv[threadIdx.x >> 5][n] = bv; s_ci[threadIdx.x >> 5][n] = bi; }
      cons_sync();
      if (threadIdx.x < p.B) {
        float v = s_cv[0][threadIdx.x];
        int i = s_ci[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < GV_WARPS; ++w) {
          const float ov = s_cv[w][threadIdx.x];
          const int oi = s_ci[w][threadIdx.x];
          if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        __stcg(p.cand_val + threadIdx.x * ncta + blockIdx.x, v);
        __stcg(p.cand_idx + threadIdx.x * ncta + blockIdx.x, i);
      }
    }
  }
  grid_sync_shared(p.gbar, target, st, s_tab, m, p.stages, false);
  if (blockIdx.x == 0) finalize_step(p, ncta);
  // exit protocol: the last CTA to leave resets the barrier counters for the next launch
  cons_sync();
  if (threadIdx.x == 0) s_tab.arrived = 0xffffffffu;    // poller warp: quit
  if (p.tm_relay && threadIdx.x < 32) {
    __syncwarp();
    tmem_dealloc(s_tmem, TM_COLS);
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.gexit, 1u) == static_cast<unsigned>(ncta) - 1) {
      *p.gbar = 0;
      *p.gexit = 0;
      __threadfence();
    }
  }
}

}  // namespace

// host side ---------------------------------------------------------------------------------------------------
struct MegaPlan {
  MegaParams p;
  size_t smem;
  int grid;
};

static int pick_kp_cap(int K, int B, size_t cap_bytes) {
  // prefer k-phases made of full 1024-column stages (32 KB in flight per ring slot): at batch 8 the down projection used
  // to get KP = 3584 -> KS = 512, i.e. half-filled slots and a latency-bound stream (53 us instead of ~20)
  for (int kp = 16384; kp >= 1024; kp -= 1024)
    if (K % kp == 0 && kp <= K && static_cast<size_t>(B) * (kp + XS_PAD) * 2 <= cap_bytes) return kp;
  for (int kp = 16384; kp >= 256; kp -= 256)
    if (K % kp == 0 && kp <= K && static_cast<size_t>(B) * (kp + XS_PAD) * 2 <= cap_bytes) return kp;
  return 0;
}
static int pick_ks2(int KP) {
  for (int ks = 1024; ks >= 256; ks -= 256)
    if (KP % ks == 0) return ks;
  return 0;
}

int decode_mega_plan(int B, int H, int qkvw, int qd, int ffn, int head_rows, int max_ctx, int kvh, int num_sms, int* nsplit_out,
                     int* chunk_max_out, int* kp_ks /*[10]*/, int* max_tiles_out, int* stages_out, int* xs_cap_out, size_t* smem_out) {
  TB_CHECK(B >= 1 && B <= GV_MAXB, "decode batch must be in [1, 8]");
  const size_t xcap = 72 * 1024;   // B=8 x (4096 + pad) halves must fit (fused RMSNorm stages the whole hidden row)
  const int Ks[5] = {H, qd, H, ffn, H};
  int kpmax = 0;
  for (int i = 0; i < 5; ++i) {
    int kp = pick_kp_cap(Ks[i], B, xcap);
    TB_CHECK(kp > 0, "decode GEMV K must be a multiple of 256");
    if (i == 0 || i == 2 || i == 4) TB_CHECK(kp == Ks[i], "fused RMSNorm needs the whole hidden row staged");
    kp_ks[2 * i] = kp;
    kp_ks[2 * i + 1] = pick_ks2(kp);
    if (kp > kpmax) kpmax = kp;
  }
  const int Ns[5] = {qkvw, H, 2 * ffn, H, head_rows};
  int max_tiles = 0;
  for (int i = 0; i < 5; ++i) {
    const int rows = ((Ns[i] / 2 + num_sms - 1) / num_sms) * 2 + 2;
    const int t = (rows + GV_TILE - 1) / GV_TILE;
    if (t > max_tiles) max_tiles = t;
  }
  int nsplit = num_sms / (B * kvh);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 20) nsplit = 20;      // the merge keeps all split statistics in registers
  TB_CHECK(B * kvh * nsplit <= num_sms, "one attention work item per CTA");
  TB_CHECK((max_ctx + nsplit - 1) / nsplit <= 62 * 64, "attention chunk exceeds the per-CTA page table (64 pages)");
  const int chunk_max = (((max_ctx + nsplit - 1) / nsplit) + 31) & ~31;
  // attention scratch lives in the s_part region: qs + knew + vnew + red + sc
  const size_t attn_scratch = (static_cast<size_t>(MG_G) * MG_HD + 2 * MG_HD + GV_WARPS * MG_G * MG_HD + GV_WARPS * MG_G * 2) * 4;
  const int nbc = B == 1 ? 1 : B == 2 ? 2 : B <= 4 ? 4 : 8;     // batch bucket = kernel template argument
  size_t part_bytes = static_cast<size_t>(max_tiles) * GV_WARPS * GV_TILE * nbc * 4;
  if (part_bytes < attn_scratch) part_bytes = attn_scratch;
  part_bytes = (part_bytes + 127) & ~static_cast<size_t>(127);
  const int xs_cap = kpmax + XS_PAD;
  const size_t fixed = static_cast<size_t>(B) * xs_cap * 2 + part_bytes + (GV_WARPS * GV_MAXB + GV_MAXB) * 4 + 2 * GV_MAXSTAGES * 8 + 256;
  const size_t stage = static_cast<size_t>(GV_TILE) * (1024 + WS_PAD) * 2;
  const size_t budget = 220 * 1024;   // + ~5 KB static shared memory (layer table, control blocks) <= 227 KB
  TB_CHECK(fixed + 2 * stage <= budget, "decode step shared memory budget exceeded");
  int st = static_cast<int>((budget - fixed) / stage);
  if (st > GV_MAXSTAGES) st = GV_MAXSTAGES;
  *nsplit_out = nsplit;
  *chunk_max_out = chunk_max;
  *max_tiles_out = static_cast<int>(part_bytes);
  *stages_out = st;
  *xs_cap_out = xs_cap;
  *smem_out = fixed + st * stage;
  return 0;
}

int decode_step_fused(const DecodeMegaArgs& a, cudaStream_t stream) {
  MegaParams p{};
  p.B = a.B; p.H = a.H; p.qkvw = a.qkvw; p.qd = a.qd; p.ffn = a.ffn; p.layers = a.layers; p.heads = a.heads; p.kvh = a.kvh;
  p.vocab = a.vocab; p.head_rows = a.head_rows; p.n_valid = a.n_valid; p.eps = a.eps;
  p.L = reinterpret_cast<const MegaLayer*>(a.layer_table_dev);
  p.final_norm = a.final_norm; p.head_w = a.head_w; p.x = a.x; p.qkv = a.qkv; p.attn = a.attn; p.act = a.act;
  p.cos_tab = a.cos_tab; p.sin_tab = a.sin_tab; p.block_table = a.block_table; p.max_pages = a.max_pages; p.page_size = a.page_size;
  p.part_o = a.part_o; p.part_ml = a.part_ml; p.counters = a.counters; p.cand_val = a.cand_val; p.cand_idx = a.cand_idx;
  p.logits_out = a.logits_out; p.embed = a.embed; p.time_tab = a.time_tab; p.score_tab = a.score_tab; p.sync_tab = a.sync_tab;
  p.eos = a.eos; p.pad = a.pad; p.st = a.st; p.gbar = a.gbar; p.gexit = a.gbar + 1;
  p.pf_ahead = a.pf_ahead;
  p.tm_relay = a.tm_relay; p.gate_mode = a.gate_mode;
  p.max_inflight = a.max_inflight; p.early_release = a.early_release; p.relay_sites = a.relay_sites;
  p.dbg = a.dbg;
  int kpks[10];
  size_t smem = 0;
  if (decode_mega_plan(a.B, a.H, a.qkvw, a.qd, a.ffn, a.head_rows, a.max_ctx, a.kvh, a.num_sms, &p.nsplit, &p.chunk_max, kpks, &p.part_bytes,
                       &p.stages, &p.xs_cap, &smem))
    return 1;
  if (a.max_stages > 1 && p.stages > a.max_stages) { smem -= static_cast<size_t>(p.stages - a.max_stages) * GV_TILE * (1024 + WS_PAD) * 2; p.stages = a.max_stages; }
  p.kp_qkv = kpks[0]; p.ks_qkv = kpks[1]; p.kp_o = kpks[2]; p.ks_o = kpks[3]; p.kp_gu = kpks[4]; p.ks_gu = kpks[5];
  p.kp_dn = kpks[6]; p.ks_dn = kpks[7]; p.kp_hd = kpks[8]; p.ks_hd = kpks[9];
  TB_CHECK(a.layers <= MG_MAX_LAYERS, "fused decode step keeps the layer pointer table in shared memory (<= 48 layers)");
  TB_CHECK(a.heads == a.kvh * MG_G && a.qd == a.heads * MG_HD, "fused decode step is specialised for 128-dim heads, 4 q heads per kv head");
  // every CTA must be resident at once (grid barriers): one CTA per SM by shared-memory footprint, grid = #SMs
  auto launch = [&](auto kern) -> int {
    TB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<a.num_sms, MG_THREADS, smem, stream>>>(p);
    return 0;
  };
  int rc;
  if (a.B == 1) rc = launch(decode_step_kernel<1>);
  else if (a.B == 2) rc = launch(decode_step_kernel<2>);
  else if (a.B <= 4) rc = launch(decode_step_kernel<4>);
  else rc = launch(decode_step_kernel<8>);
  if (rc) return rc;
  TB_LAUNCH_OK();
  return 0;
}

int decode_mega_layer_table_bytes() { return static_cast<int>(sizeof(MegaLayer)); }

}  // namespace tb
