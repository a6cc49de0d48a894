// Host-callable launchers of the trace_b200 sm_100a kernels.  All pointers are device pointers, all
// launchers are asynchronous on `stream` and return 0 on success (tb::get_error() has the message).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

// ---- gemm_tc.cu : tcgen05/TMEM/TMA GEMM  out = epi(A[M,K] · W[N,K]^T) ------------------------------------
// act: 0 none, 1 quick_gelu.  swiglu: W rows are (gate,up) interleaved, out has N/2 columns.
int gemm_f16(const __half* A, int lda, const __half* W, int ldw, __half* out, int ldo, int M, int N, int K,
             const __half* bias, const __half* residual, int ldr, int act, int swiglu, int num_sms,
             cudaStream_t stream);

// ---- norm.cu ---------------------------------------------------------------------------------------------
int layernorm_f16(const __half* x, int ldx, const __half* w, const __half* b, float eps, __half* out, int ldo,
                  int rows, int D, cudaStream_t stream);
int rmsnorm_f16(const __half* x, int ldx, const __half* w, float eps, __half* out, int ldo, int rows, int D,
                cudaStream_t stream);

// ---- vit.cu ----------------------------------------------------------------------------------------------
// pixels [F,3,H,W] fp16 -> patches [F*P, Kp] (col = c*ps*ps + ky*ps + kx, zero padded to Kp)
int vit_im2col(const __half* pixels, __half* out, int F, int image, int ps, int Kp, cudaStream_t stream);
// h[f,0] = cls + pos[0]; h[f,1+p] = patch[f*P+p] + pos[1+p]; then pre-LayerNorm -> out [F*(P+1), D]
int vit_embed_ln(const __half* patch, const __half* cls, const __half* pos, const __half* ln_w, const __half* ln_b,
                 float eps, __half* out, int F, int P, int D, cudaStream_t stream);

// ---- attention_tc.cu : softmax(Q K^T * scale [+ causal mask]) V on tcgen05 (TMA-fed, S / PV accumulators in TMEM) ----
// qkv: fused activation [batch*seqlen, width]; Q/K/V heads at column offsets q_col0/k_col0/v_col0 (head h -> + h*HD).
int attention_tc_fwd(const __half* qkv, int width, __half* O, int seqlen, int o_stride, int64_t o_bstride, int q_col0, int k_col0,
                     int v_col0, int heads, int group, int batch, int head_dim, int causal, float scale, cudaStream_t stream);

// ---- projector.cu : fused SpatialSlotPool up to the slot-weighted sums --------------------------------------
// feats: rows of width D with pitch ld; frame f owns rows [f*frame_rows + row_off, +n).  res [F*8, D] fp16.
int slot_pool_fused(const __half* feats, int ld, int frame_rows, int row_off, int n, int D, int F,
                    const __half* ln_w, const __half* ln_b, float eps, const __half* slots /*[D,8]*/,
                    const float* cos_tab, const float* sin_tab /*[n, D/2]*/, float* scratch /*slot_pool_scratch_floats(F, n)*/,
                    __half* res, cudaStream_t stream);
int slot_pool_scratch_floats(int F, int n);

// ---- splice.cu : prefill activation assembly (prompt splice + time rows) -------------------------------------
int splice_prefill(const int64_t* ids /*[B,L]*/, int B, int L, int T, const __half* slots /*[B*T*8,H]*/,
                   const int32_t* time_ids /*[B,T,6]*/, const __half* embed, const __half* time_tab,
                   const __half* sync_tab, int vocab, int H, __half* out /*[B,S,H]*/, cudaStream_t stream);

// ---- rope_kv.cu : prefill RoPE (in place on the fused qkv buffer) + paged KV write ---------------------------
int rope_kv_prefill(__half* qkv /*[B*S, (heads+2*kvh)*hd]*/, int B, int S, int heads, int kvh, int hd,
                    const float* cos_tab, const float* sin_tab /*[max_pos, hd/2]*/, __half* kv_pool,
                    const int32_t* block_table /*[slots,max_pages]*/, const int32_t* slot_ids /*[B]*/, int max_pages,
                    int page_size, cudaStream_t stream);

// ---- preprocess.cu : expand2square + PIL bicubic resize (bit-exact integer passes) + crop + normalise + fp16 ----
// frames uint8 [F,H,W,3]; the (virtual) padded canvas is S_w x S_h with the frame pasted at (pad_x, pad_y); resized to
// OW x OH; centre crop `crop` at (crop_x, crop_y); tmp uint8 [F,rows,OW,3] holds canvas rows [row0, row0+rows).
int preprocess_frames(const uint8_t* frames, int F, int H, int W, int S, int pad_x, int pad_y, int OW, int OH, int crop, int crop_x,
                      int crop_y, int row0, int rows, const int32_t* hbounds, const int32_t* hcoef, int hk, const int32_t* vbounds,
                      const int32_t* vcoef, int vk, const float* mean, const float* stdv, const int* pad_rgb, uint8_t* tmp, __half* out,
                      cudaStream_t stream);

// ---- paged KV cache, streamed layout (rope_kv.cu writes it at prefill, decode_stream.cu appends and reads) ----
// one pool per layer: [page][kv_head][K|V][page_size positions][hd] fp16; element offset of a (page, kv head, K/V) tile:
__host__ __device__ inline int64_t kv_tile_offset(int page, int kvh, int kv, int is_v, int page_size, int hd) {
  return ((static_cast<int64_t>(page) * kvh + kv) * 2 + is_v) * page_size * hd;
}
// column of element d of the row at in-page position pos: 16-byte chunks (8 halves) are XOR-swizzled with pos & 7
__host__ __device__ inline int kv_chunk_swz(int d, int pos) { return (((d >> 3) ^ (pos & 7)) << 3) | (d & 7); }

// ---- decode_stream.cu : the decode step as one cooperative kernel, one HBM stream (weights + K/V pages), flag-in-data exchanges ----
struct DecodeStreamState {     // device-resident generation state; nothing in it is baked into a captured graph
  int32_t *pos, *heads, *finished, *step, *done_step;
  int32_t *row_step, *row_max; // [B] tokens emitted by the row so far / its budget (a row is done at EOS or at its budget)
  int64_t* out_ids;            // [B, ctrl[0]]
  int32_t* slot_ids;           // [B]
  int32_t* ctrl;               // [16]: 0 row stride of out_ids / forced, 1 eos, 2 pad, 3 do_sample, 4 temperature (float bits), 5 top_k,
                               //       6/7 seed lo/hi, 8 use_forced, 9 bit mask of the rows this launch finalizes
  const int64_t* forced;       // [B, ctrl[0]] teacher-forced ids (fed back instead of the chosen token when ctrl[8])
};
struct DecodeStreamArgs {
  int B, H, qkvw, qd, ffn, layers, heads, kvh, vocab, head_rows, n_valid, max_ctx, num_sms, page_size, max_pages;
  float eps;
  const void* layer_table_dev;   // [layers] x {ln1, qkv_w, o_w, ln2, gateup_w, down_w, kv_pool, -} device pointers
  const __half *final_norm, *head_w;
  __half* x;                     // [B,H] dense: input embedding of this step in, of the next step out
  void* exchange;                // zero-initialised, decode_stream_exchange_bytes() bytes
  const float *cos_tab, *sin_tab;
  const int32_t* block_table;
  float* logits_out;             // [B, n_valid] masked logits of this step (instrumented path) or null
  float* samp_logits;            // [B, n_valid] scratch for sampling
  const __half *embed, *time_tab, *score_tab, *sync_tab;
  DecodeStreamState st;
  unsigned* epoch_ctr;           // device launch counter (starts at 1); exchange epochs derive from it
  int* err;                      // device error word written by the kernel's watchdog before it traps
  int advance;                   // 0: head-only launch on the prefill's last hidden state (layers must be 0), 1: a decoder pass
  int cooperative, max_stages, poll_sleep_ns, watchdog_ms;
  unsigned long long* dbg;
};
int decode_stream_step(const DecodeStreamArgs& a, cudaStream_t stream);
int decode_stream_plan(const DecodeStreamArgs& a, int t, int32_t* out /*[8]: nsplit, ncomb, K/V pages per attention item at context t, kp_down, ks_down, ring stages, tiles per epilogue group, smem bytes*/);
size_t decode_stream_exchange_bytes(int H, int qkvw, int qd, int ffn, int num_sms, size_t* offs /*[16]*/);
int decode_stream_layer_bytes();

}  // namespace tb
