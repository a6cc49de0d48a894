/* trace_b200 -- C ABI of the B200-native TRACE inference hot path.
 *
 * The reference (gyxxyg/TRACE) is pure Python on top of `transformers`; it has no C/FFI/plugin seam
 * (SURVEY.md §8(b)).  The drop-in boundary of this repo is therefore the reference's Python surface
 * (trace_b200/model/…, mirroring trace/model/language_model/trace_mistral.py, trace/mm_utils.py,
 * trace/conversation.py) and THIS header is the thin C ABI that surface calls through ctypes.  Each entry
 * point states which reference call site it replaces.  See INTEGRATION.md for the binding a reference
 * maintainer would add.
 *
 * Conventions: every function returns 0 on success, non-zero on error; trace_b200_last_error() returns a
 * thread-local message.  All `dev` pointers are raw CUDA device pointers into caller-owned storage (e.g.
 * torch tensors); `stream` is a cudaStream_t passed as void*.  One context per device, one caller thread
 * per context (the reference is single-threaded, default stream).  No torch types appear here.
 */
#ifndef TRACE_B200_H_
#define TRACE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRACE_B200_ABI_VERSION 1
#define TRACE_B200_PAGE_SIZE 64          /* KV-cache positions per page */
#define TRACE_B200_MAX_DECODE_BATCH 8
#define TRACE_B200_NUM_SLOTS 8           /* SpatialSlotPool num_slots (multimodal_projector/builder.py:413) */

typedef struct trace_b200_ctx trace_b200_ctx;

/* Architecture = the keys of a TRACE checkpoint's config.json (SURVEY.md §5 "Config / flags") plus
 * workspace sizing. */
typedef struct trace_b200_config {
  /* CLIP vision tower (HF CLIPVisionConfig) */
  int32_t vit_width, vit_heads, vit_layers, vit_mlp, image_size, patch_size;
  float vit_ln_eps;
  /* Mistral decoder (HF MistralConfig) */
  int32_t hidden, layers, heads, kv_heads, head_dim, ffn, vocab;
  float rope_theta, rms_eps;
  /* workspaces */
  int32_t vit_chunk_frames;     /* frames encoded per ViT pass (activations stay L2-resident) */
  int32_t max_prefill_tokens;   /* B*S upper bound for one prefill call */
  int32_t max_ctx;              /* positions per sequence (prefill + generated) */
  int32_t kv_pages;             /* pages in the KV pool */
} trace_b200_config;

int trace_b200_abi_version(void);
const char* trace_b200_last_error(void);

/* Lifetime.  create() allocates every workspace and the KV pool; nothing is allocated afterwards. */
int trace_b200_create(const trace_b200_config* cfg, int device, trace_b200_ctx** out);
void trace_b200_destroy(trace_b200_ctx* ctx);

/* Weights: HF state-dict names exactly as the reference checkpoint stores them (SURVEY.md §4 state-dict
 * table; replaces TraceMistralForCausalLM.from_pretrained, trace/model/builder.py:114).  fp16 only
 * (builder.py:50 torch_dtype=float16).  The tensor is copied (and re-packed: fused QKV, interleaved
 * gate/up, padded patch kernel, stacked heads) into engine-owned memory. */
int trace_b200_load_weight(trace_b200_ctx* ctx, const char* hf_name, const void* dev_ptr, const int64_t* shape,
                           int ndim, void* stream);
int trace_b200_finalize_weights(trace_b200_ctx* ctx);   /* fails listing the first missing tensor */

/* encode_images_or_videos up to the projector (trace_arch.py:218-235): CLIPVisionTower.forward +
 * feature_select (clip_encoder.py:31-53, hidden_states[-2] without CLS) + SpatialSlotPool.forward
 * (multimodal_projector/builder.py:427-467).
 *   pixels  [F,3,H,W] fp16 (CLIP-normalised)       -> slots_out [F,8,hidden] fp16
 *   feats_out (optional, may be NULL) [F,patches,vit_width] fp16 receives the CLIP features. */
int trace_b200_encode_frames(trace_b200_ctx* ctx, const void* pixels, int F, void* slots_out, void* feats_out,
                             void* stream);

/* prepare_inputs_labels_for_multimodal, prefill branch (trace_arch.py:377-426) + the time rows of
 * encode_images_or_videos (trace_arch.py:237-258).  ids [B,L] int64 with exactly one -201 per row and -205
 * for <sync>; slots [B*T*8, hidden]; time_ids [B,T,6] int32 (host-formatted '0>6.1f' ids, time_encoder.py:58).
 * embeds_out [B, L-1+14T, hidden] fp16.
 * T == 0 (slots and time_ids may be NULL) is the text-only branch of generate (trace_mistral.py:306-307,
 * `embed_tokens(inputs)`): embeds_out [B, L, hidden]. */
int trace_b200_build_prefill(trace_b200_ctx* ctx, const int64_t* ids_dev, int B, int L, int T, const void* slots,
                             const int32_t* time_ids_dev, void* embeds_out, void* stream);

/* KV slots (replaces HF DynamicCache): a slot owns ceil(n_positions/64) pages until freed. */
int trace_b200_alloc_slot(trace_b200_ctx* ctx, int n_positions, int* slot_out);
int trace_b200_free_slot(trace_b200_ctx* ctx, int slot);

/* Step 0 of generate: MistralModel prefill over inputs_embeds (trace_mistral.py:178-188), all-ones
 * attention mask.  embeds [B,S,hidden] fp16; fills the slots' KV pages and keeps the last
 * position's hidden state for trace_b200_decode.  hidden_out (optional) [B,S,hidden] receives the final
 * (pre-norm) hidden states. */
int trace_b200_prefill(trace_b200_ctx* ctx, const int32_t* slots_host, int B, const void* embeds, int S,
                       void* hidden_out, void* stream);

/* The greedy loop: heads + -inf mask + argmax (trace_mistral.py:190-252 + HF greedy), head switch
 * (prepare_inputs_for_generation :336-344), decode-step embedding dispatch (trace_arch.py:345-375), decoder
 * layers with KV append -- entirely on the device.  heads_inout_host[b] in {0,1,2} or -1 for "no heads given"
 * (text head only, trace/__init__.py:infer); it is updated in place like the reference mutates `heads`.
 * ids_out_host [B,max_new] int64 receives the NEW tokens only; *n_out = tokens per row (stops when every row
 * hit EOS).  logits_out_dev (optional) [max_new,B,vocab+27] fp32 receives the masked logits of each step. */
int trace_b200_decode(trace_b200_ctx* ctx, const int32_t* slots_host, int B, int32_t* heads_inout_host, int max_new,
                      int eos, int pad, int64_t* ids_out_host, int* n_out, float* logits_out_dev, void* stream);

/* Continuous batching (replaces the serial batch-1 loops of evaluate.py:296-357 / videomme/evaluate.py:348-383).
 * A batch of `rows` decode rows stays open; requests are admitted into free rows and retired BETWEEN bursts of decode steps, so
 * rows of different clip length, prompt length, head and budget share every decode step and a finished row's slot in the MMA
 * is refilled instead of idling until the slowest row of a static batch ends.  Per-row state (position, head, emitted tokens,
 * budget) lives on the device; an inactive row points at a one-page dummy slot and an all-zero activation row.
 *   open   : rows in [1,8]; eos / pad as in trace_b200_decode; sampling as set by trace_b200_set_sampling at this moment.
 *   admit  : prefill of `embeds` [1,S,hidden] into `slot` (allocated with >= S + max_new positions) and the row's first token
 *            (heads on the prefill's last hidden state; the other rows are not touched).
 *   step   : n_steps decode steps for all rows (CUDA-graph replays); row_steps_out / finished_out [rows] = tokens emitted so
 *            far (can exceed the budget by the steps run after it was used up) and the EOS flags.
 *   read   : the new ids of a row (<= its budget) and its current head.     retire : frees the row (not the KV slot).
 *   close  : every row must have been retired. */
int trace_b200_batch_open(trace_b200_ctx* ctx, int rows, int eos, int pad);
int trace_b200_batch_admit(trace_b200_ctx* ctx, int row, int slot, const void* embeds, int S, int head, int max_new, void* stream);
int trace_b200_batch_step(trace_b200_ctx* ctx, int n_steps, int32_t* row_steps_out, int32_t* finished_out, void* stream);
int trace_b200_batch_read(trace_b200_ctx* ctx, int row, int64_t* ids_out, int capacity, int* n_out, int32_t* head_out);
int trace_b200_batch_retire(trace_b200_ctx* ctx, int row);
int trace_b200_batch_close(trace_b200_ctx* ctx);

/* Sampling of the following trace_b200_decode calls: HF GenerationMixin._sample with the default logits warpers as
 * scripts/inference/inference.py:62,71-73 drives it (do_sample=True, temperature=0.2; generation_config top_k=50, top_p=1):
 * logits / temperature, keep {z >= k-th largest} (top_k = 0: keep all), softmax, inverse-CDF draw over ascending token ids
 * with u = (splitmix64(seed ^ step*0x9E3779B97F4A7C15 ^ row*0xD1B54A32D192ED03) >> 40 + 0.5) / 2^24 -- a counter-based
 * generator, so a (seed, step, row) always draws the same token and the host can reproduce the draw.  do_sample = 0 restores
 * greedy decoding. */
int trace_b200_set_sampling(trace_b200_ctx* ctx, int do_sample, float temperature, int top_k, uint64_t seed);

/* Teacher forcing for the NEXT trace_b200_decode call only (B and n must equal its batch and max_new): the ids written to
 * ids_out are still the engine's own choices, but ids_host [B,n] are fed back instead (forward over a given continuation:
 * per-step logits against a reference trace stay comparable after an fp16 tie).  NULL clears it. */
int trace_b200_set_forced_ids(trace_b200_ctx* ctx, const int64_t* ids_host, int B, int n);

/* Launch plan of the fused decode step at batch B and context length context_len: out8 = {attention splits per (row, kv
 * head), combiner CTAs per (row, kv head), K/V page tiles per attention work item, k-phase width and stage width of the down
 * projection, ring stages, row tiles per epilogue group, dynamic shared memory bytes}.  Tests use it to assert that the
 * code paths of the full-size configurations (several K/V tiles per item, k-phased down projection) are the ones they run. */
int trace_b200_decode_plan(trace_b200_ctx* ctx, int B, int context_len, int32_t* out8);

/* Introspection used by bench.py: kernels launched by this context since creation, and bytes of weights. */
int64_t trace_b200_launch_count(trace_b200_ctx* ctx);
int64_t trace_b200_weight_bytes(trace_b200_ctx* ctx);
/* device time (CUDA events on the engine's decode stream) and decoder passes of the last trace_b200_decode */
int trace_b200_last_decode_timing(trace_b200_ctx* ctx, float* ms_out, int* steps_out);

/* debugging aid: %globaltimer stamps (and a few counters) of the last fused decode step, [2 CTAs][layers][32 slots]
 * (needs TRACE_B200_MEGA_TIMING=1 in the environment when the context is created; tools/profile_step.py decodes it).  A capacity
 * of 2*layers*32 + SMs*32 also returns the stamps of EVERY CTA for the middle layer (slot 31 = SM id). */
int trace_b200_debug_timeline(trace_b200_ctx* ctx, uint64_t* out_host, int capacity);

/* Frame pre-processing on the device: what process_video does per decoded frame on the host (trace/mm_utils.py:424-462):
 * expand2square with the CLIP mean colour (:259-270), CLIPImageProcessor.preprocess (PIL bicubic resize with antialiasing,
 * centre crop, 1/255, mean/std) and the driver's fp16 cast (evaluate.py:316).  Stateless (no ctx).
 *   frames uint8 [F,H,W,3] RGB (device)  ->  out fp16 [F,3,crop,crop] (device)
 * geom (host, 15 ints): canvas size, paste offset x/y, resized width/height, crop size, crop offset x/y, first canvas row
 * and number of canvas rows the vertical pass reads, horizontal / vertical taps per output pixel, pad colour r/g/b.
 * hbounds/vbounds [out,2] (first tap, taps) and hcoef/vcoef [out,taps] are Pillow's 22-bit fixed-point resampling tables
 * (device, produced by trace_b200/preprocess.py exactly like Resample.c precompute_coeffs / normalize_coeffs_8bpc);
 * mean/stdv: 3 floats (host); tmp: uint8 [F,rows,OW,3] scratch (device). */
int trace_b200_preprocess_frames(const uint8_t* frames, int F, int H, int W, const int32_t* geom, const int32_t* hbounds,
                                 const int32_t* hcoef, const int32_t* vbounds, const int32_t* vcoef, const float* mean,
                                 const float* stdv, uint8_t* tmp, void* out, void* stream);

/* ---- single-kernel entry points (unit-parity tests and micro-benchmarks) --------------------------- */
int trace_b200_op_gemm(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K,
                       const void* bias, const void* residual, int ldr, int act, int swiglu, void* stream);
int trace_b200_op_layernorm(const void* x, const void* w, const void* b, float eps, void* out, int rows, int D,
                            void* stream);
int trace_b200_op_rmsnorm(const void* x, const void* w, float eps, void* out, int rows, int D, void* stream);
/* tcgen05 / TMEM attention over a fused [batch*seqlen, (heads+2*kv_heads)*head_dim] activation */
int trace_b200_op_attention_tc(const void* qkv, int width, void* o, int seqlen, int heads, int kv_heads, int head_dim, int causal,
                               int batch, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRACE_B200_H_ */
