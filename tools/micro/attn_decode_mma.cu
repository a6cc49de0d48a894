// Round-2 prototype (standalone, self-checking): split-KV decode attention for GQA 32:8 x 128 on mma.sync tensor cores.
//
// Why: inside the fused decode step the attention of one (row, kv head, split) is computed with FMAs + warp shuffles,
// ~4 us per pass of 128 keys.  That is fine at batch 1 (one pass per CTA) but at batch 4..8 a work item covers 500-1000
// keys and the phase grows to 26-51 us per layer (profiles/r01_decode_ab.md).  Here the 4 query heads of a kv head are the
// (padded) M=16 rows of m16n8k16 tiles:  S[16 x 16 keys] = Q K^T (8 k-steps x 2 n-tiles), online softmax on the
// accumulator fragments, O[16 x 128] += P V (16 n-tiles) with V^T fragments from ldmatrix.trans.  K/V tiles of 64 keys are
// staged in shared memory with cp.async (double buffered), 4 warps x 16 keys per tile, warps merged once at the end.
//
//   ./attn_decode_mma [B=8] [T=2048] [nsplit=2]      prints max |err| against a naive fp32 kernel and the time per call
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/attn_decode_mma tools/micro/attn_decode_mma.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int HD = 128, G = 4, KVH = 8, HEADS = 32;
constexpr int TILE = 64;                 // keys per shared-memory tile (16 per warp)
constexpr int PITCH = HD + 8;            // halves; 272-byte rows: conflict-free 4-byte and ldmatrix accesses
constexpr int WARPS = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async_16(void* dst, const void* src, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// q [B][HEADS][HD] (RoPE applied), k/v [B][KVH][T][HD]; partial o [B][KVH][nsplit][G][HD] fp32 (unnormalised), ml [..][G][2]
__global__ void __launch_bounds__(WARPS * 32) attn_decode_mma(const __half* __restrict__ q, const __half* __restrict__ k,
                                                              const __half* __restrict__ v, int T, int nsplit, float scale_log2,
                                                              float* __restrict__ part_o, float* __restrict__ part_ml) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __half* sK = reinterpret_cast<__half*>(smem_raw);                    // [2][TILE][PITCH]
  __half* sV = sK + 2 * TILE * PITCH;                                  // [2][TILE][PITCH]
  float* s_red = reinterpret_cast<float*>(sV + 2 * TILE * PITCH);      // [WARPS][G][HD] + [WARPS][G][2]
  const int split = blockIdx.x, kv = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  int chunk = (T + nsplit - 1) / nsplit;
  chunk = (chunk + TILE - 1) / TILE * TILE;
  const int start = split * chunk, end = min(start + chunk, T);
  const int ntile = max(0, (end - start + TILE - 1) / TILE);
  const __half* kb = k + (static_cast<size_t>(b) * KVH + kv) * T * HD;
  const __half* vb = v + (static_cast<size_t>(b) * KVH + kv) * T * HD;

  auto load_tile = [&](int it, int buf) {
    const int key0 = start + it * TILE;
    for (int i = tid; i < TILE * (HD / 8); i += WARPS * 32) {          // 16-byte chunks: 64 keys x 16 chunks, K and V
      const int r = i / (HD / 8), c = i % (HD / 8);
      const bool ok = key0 + r < end;
      const size_t off = static_cast<size_t>(ok ? key0 + r : 0) * HD + c * 8;
      cp_async_16(sK + (buf * TILE + r) * PITCH + c * 8, kb + off, ok);
      cp_async_16(sV + (buf * TILE + r) * PITCH + c * 8, vb + off, ok);
    }
    cp_async_commit();
  };
  if (ntile > 0) load_tile(0, 0);

  // A fragments of Q: rows = the 4 query heads of this kv head (rows 4..15 are zero), all 8 k-steps kept in registers
  uint32_t qa[8][4];
  {
    const __half* qrow = q + (static_cast<size_t>(b) * HEADS + kv * G + g) * HD;   // only valid for g < G
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qa[ks][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t) : 0u;
      qa[ks][2] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t + 8) : 0u;
      qa[ks][1] = 0u;                                                              // rows g + 8: padding
      qa[ks][3] = 0u;
    }
  }
  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;                                            // row g (this thread's share of the sum)

  for (int it = 0; it < ntile; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntile) { load_tile(it + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const __half* Ks = sK + (buf * TILE + warp * 16) * PITCH;                      // this warp's 16 keys
    const __half* Vs = sV + (buf * TILE + warp * 16) * PITCH;
    const int key_w = start + it * TILE + warp * 16;
    // ---- S = Q K^T : two n-tiles of 8 keys ----
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t b00 = *reinterpret_cast<const uint32_t*>(Ks + g * PITCH + ks * 16 + 2 * t);
      const uint32_t b01 = *reinterpret_cast<const uint32_t*>(Ks + g * PITCH + ks * 16 + 2 * t + 8);
      const uint32_t b10 = *reinterpret_cast<const uint32_t*>(Ks + (g + 8) * PITCH + ks * 16 + 2 * t);
      const uint32_t b11 = *reinterpret_cast<const uint32_t*>(Ks + (g + 8) * PITCH + ks * 16 + 2 * t + 8);
      mma_16816(s0, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b00, b01);
      mma_16816(s1, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b10, b11);
    }
    // this thread: row g, keys key_w + {2t, 2t+1} (s0[0], s0[1]) and key_w + 8 + {2t, 2t+1} (s1[0], s1[1])
    float sv[4] = {s0[0], s0[1], s1[0], s1[1]};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int key = key_w + (i >> 1) * 8 + 2 * t + (i & 1);
      if (key >= end) sv[i] = -INFINITY;
    }
    float mx = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float ms = (m_new == -INFINITY) ? 0.f : m_new * scale_log2;
    const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run * scale_log2 - ms);
    float p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = exp2f(sv[i] * scale_log2 - ms);             // exp2(-inf) = 0 for masked keys
    l_run = l_run * corr + (p[0] + p[1] + p[2] + p[3]);
    m_run = m_new;
    // ---- O = O * corr + P V ----  A = P (rows g valid; rows g + 8 zero), k = the warp's 16 keys
    const uint32_t pa0 = pack_half2(p[0], p[1]), pa2 = pack_half2(p[2], p[3]);
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) {                                               // two d-tiles of 8 per ldmatrix.x4.trans
      o[2 * n2][0] *= corr; o[2 * n2][1] *= corr; o[2 * n2 + 1][0] *= corr; o[2 * n2 + 1][1] *= corr;
      uint32_t r0, r1, r2, r3;
      const int mtx = lane >> 3, rr = lane & 7;                                    // matrix index / row inside the matrix
      const __half* addr = Vs + ((mtx & 1) * 8 + rr) * PITCH + n2 * 16 + (mtx >> 1) * 8;
      ldmatrix_x4_trans(r0, r1, r2, r3, smem_u32(addr));
      mma_16816(o[2 * n2], pa0, 0u, pa2, 0u, r0, r1);
      mma_16816(o[2 * n2 + 1], pa0, 0u, pa2, 0u, r2, r3);
    }
    __syncthreads();                                                               // tile buffer may be overwritten
  }
  // ---- merge the 4 warps (each saw a different quarter of every tile) ----
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  float* red_o = s_red;                                 // [WARPS][G][HD]
  float* red_ml = s_red + WARPS * G * HD;               // [WARPS][G][2]
  if (g < G) {
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      red_o[(warp * G + g) * HD + n * 8 + 2 * t] = o[n][0];
      red_o[(warp * G + g) * HD + n * 8 + 2 * t + 1] = o[n][1];
    }
    if (t == 0) { red_ml[(warp * G + g) * 2] = m_run; red_ml[(warp * G + g) * 2 + 1] = l_run; }
  }
  __syncthreads();
  const size_t pbase = ((static_cast<size_t>(b) * KVH + kv) * nsplit + split);
  for (int i = tid; i < G * HD; i += WARPS * 32) {
    const int h = i / HD, d = i % HD;
    float M = -INFINITY;
    for (int w = 0; w < WARPS; ++w) M = fmaxf(M, red_ml[(w * G + h) * 2]);
    float acc = 0.f, L = 0.f;
    for (int w = 0; w < WARPS; ++w) {
      const float mw = red_ml[(w * G + h) * 2], lw = red_ml[(w * G + h) * 2 + 1];
      if (lw != 0.f) {
        const float wt = exp2f((mw - M) * scale_log2);
        acc += wt * red_o[(w * G + h) * HD + d];
        L += wt * lw;
      }
    }
    part_o[pbase * G * HD + i] = acc;
    if (d == 0) { part_ml[pbase * G * 2 + h * 2] = M; part_ml[pbase * G * 2 + h * 2 + 1] = L; }
  }
}

__global__ void merge_splits(const float* part_o, const float* part_ml, int nsplit, float scale_log2, float* out) {
  const int kv = blockIdx.x, b = blockIdx.y;   // out [B][HEADS][HD]
  for (int i = threadIdx.x; i < G * HD; i += blockDim.x) {
    const int h = i / HD;
    const size_t gb = (static_cast<size_t>(b) * KVH + kv) * nsplit;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_ml[(gb + s) * G * 2 + h * 2]);
    float acc = 0.f, L = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float ms = part_ml[(gb + s) * G * 2 + h * 2], ls = part_ml[(gb + s) * G * 2 + h * 2 + 1];
      if (ls != 0.f) {
        const float wt = exp2f((ms - M) * scale_log2);
        acc += wt * part_o[(gb + s) * G * HD + i];
        L += wt * ls;
      }
    }
    out[(static_cast<size_t>(b) * HEADS + kv * G + h) * HD + (i % HD)] = acc / L;
  }
}

// naive fp32 reference: one block per (b, head), thread per d
__global__ void attn_naive(const __half* q, const __half* k, const __half* v, int T, float scale, float* out) {
  const int head = blockIdx.x, b = blockIdx.y, kv = head / G, d = threadIdx.x;
  extern __shared__ float sc[];
  const __half* qr = q + (static_cast<size_t>(b) * HEADS + head) * HD;
  const __half* kb = k + (static_cast<size_t>(b) * KVH + kv) * T * HD;
  const __half* vb = v + (static_cast<size_t>(b) * KVH + kv) * T * HD;
  for (int j = d; j < T; j += blockDim.x) {
    float s = 0.f;
    for (int e = 0; e < HD; ++e) s += __half2float(qr[e]) * __half2float(kb[static_cast<size_t>(j) * HD + e]);
    sc[j] = s * scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = 0; j < T; ++j) mx = fmaxf(mx, sc[j]);
  float l = 0.f, acc = 0.f;
  for (int j = 0; j < T; ++j) {
    const float p = __half2float(__float2half_rn(expf(sc[j] - mx)));      // the fused path rounds P to fp16 before P V
    l += expf(sc[j] - mx);
    acc += p * __half2float(vb[static_cast<size_t>(j) * HD + d]);
  }
  out[(static_cast<size_t>(b) * HEADS + head) * HD + d] = acc / l;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 8, T = argc > 2 ? atoi(argv[2]) : 2048, nsplit = argc > 3 ? atoi(argv[3]) : 2;
  const size_t nq = static_cast<size_t>(B) * HEADS * HD, nkv = static_cast<size_t>(B) * KVH * T * HD;
  std::vector<__half> hq(nq), hk(nkv), hv(nkv);
  uint32_t seed = 12345;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (static_cast<int>(seed >> 9) % 2001 - 1000) / 1000.0f; };
  for (auto& x : hq) x = __float2half(rnd() * 2.0f);
  for (auto& x : hk) x = __float2half(rnd());
  for (auto& x : hv) x = __float2half(rnd());
  __half *dq, *dk, *dv; float *po, *pml, *out, *ref;
  cudaMalloc(&dq, nq * 2); cudaMalloc(&dk, nkv * 2); cudaMalloc(&dv, nkv * 2);
  cudaMalloc(&po, static_cast<size_t>(B) * KVH * nsplit * G * HD * 4); cudaMalloc(&pml, static_cast<size_t>(B) * KVH * nsplit * G * 2 * 4);
  cudaMalloc(&out, nq * 4); cudaMalloc(&ref, nq * 4);
  cudaMemcpy(dq, hq.data(), nq * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dk, hk.data(), nkv * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dv, hv.data(), nkv * 2, cudaMemcpyHostToDevice);
  const float scale = 1.0f / sqrtf(static_cast<float>(HD)), scale_log2 = scale * 1.4426950408889634f;
  const size_t smem = static_cast<size_t>(4) * TILE * PITCH * 2 + (WARPS * G * HD + WARPS * G * 2) * 4;
  cudaFuncSetAttribute(attn_decode_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaFuncSetAttribute(attn_naive, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 4);
  attn_naive<<<dim3(HEADS, B), HD, T * 4>>>(dq, dk, dv, T, scale, ref);
  float best = 1e9f;
  for (int rep = 0; rep < 10; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    attn_decode_mma<<<dim3(nsplit, KVH, B), WARPS * 32, smem>>>(dq, dk, dv, T, nsplit, scale_log2, po, pml);
    merge_splits<<<dim3(KVH, B), 128>>>(po, pml, nsplit, scale_log2, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  std::vector<float> ho(nq), hr(nq);
  cudaMemcpy(ho.data(), out, nq * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hr.data(), ref, nq * 4, cudaMemcpyDeviceToHost);
  double err = 0;
  for (size_t i = 0; i < nq; ++i) err = fmax(err, fabs(ho[i] - hr[i]));
  printf("B=%d T=%d nsplit=%d: max |mma - naive| = %.3g  (%s)  %.1f us per call, KV bytes %.1f MB -> %.0f GB/s\n", B, T, nsplit, err,
         cudaGetErrorString(cudaGetLastError()), best * 1e3, 2.0 * nkv * 2 / 1e6, 2.0 * nkv * 2 / (best * 1e-3) / 1e9);
  return err < 5e-3 ? 0 : 1;
}
