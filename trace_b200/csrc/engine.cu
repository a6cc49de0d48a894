// Engine: context, engine-owned packed weights, workspaces, paged KV pool, and the orchestration of
// encode_frames / build_prefill / prefill / decode behind the C ABI of include/trace_b200.h.
#include <math.h>
#include <string.h>

#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges show up in Nsight Systems / ncu --nvtx, free when no tool is attached

#include "../../include/trace_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace tb {

struct NvtxRange {            // one named range per stage of the path (tracing)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

namespace {

constexpr int kPage = TRACE_B200_PAGE_SIZE;
constexpr int kMaxSlots = 64;
constexpr int kMaxB = TRACE_B200_MAX_DECODE_BATCH;

struct VitLayer {
  __half *ln1_w, *ln1_b, *qkv_w, *qkv_b, *out_w, *out_b, *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};
struct LlmLayer {
  __half *ln1, *qkv_w, *o_w, *ln2, *gateup_w, *down_w;
};

}  // namespace
}  // namespace tb

using namespace tb;

struct trace_b200_ctx {
  trace_b200_config c;
  int device = 0;
  int num_sms = 148;
  int P = 0, Kp = 0, qkvw = 0, total_vocab = 0, head_rows = 0;
  int max_pages_seq = 0;
  int64_t launches = 0;
  int64_t weight_bytes = 0;
  std::vector<void*> allocs;
  std::set<std::string> loaded;
  std::vector<std::string> expected;
  bool finalized = false;

  // ViT weights
  __half *patch_w = nullptr, *cls = nullptr, *pos = nullptr, *pre_w = nullptr, *pre_b = nullptr;
  std::vector<VitLayer> vit;
  // projector / towers
  __half *pj_slots = nullptr, *pj_ln_w = nullptr, *pj_ln_b = nullptr, *pj_readout = nullptr;
  __half *time_tab = nullptr, *score_tab = nullptr, *sync_tab = nullptr;
  float *pj_cos = nullptr, *pj_sin = nullptr;
  // LLM weights
  __half *embed = nullptr, *final_norm = nullptr, *head_w = nullptr;
  std::vector<LlmLayer> llm;
  float *rope_cos = nullptr, *rope_sin = nullptr;

  // ViT workspaces (one chunk)
  __half *v_col = nullptr, *v_patch = nullptr, *v_h = nullptr, *v_y = nullptr, *v_qkv = nullptr, *v_attn = nullptr,
         *v_mlp = nullptr, *v_res = nullptr;
  float* pj_scratch = nullptr;    // slot pool: logits + LN statistics per (frame, patch)
  // prefill workspaces
  __half *p_h = nullptr, *p_y = nullptr, *p_qkv = nullptr, *p_attn = nullptr, *p_act = nullptr;
  // decode workspaces + state
  __half* d_x = nullptr;          // [8, hidden] input embedding of the next decode step (prefill's last row, then finalize's embedding)
  // KV pool
  int32_t* d_block_table = nullptr;
  std::vector<int32_t> h_block_table;
  std::vector<int> free_pages;
  std::vector<std::vector<int>> slot_pages;
  std::vector<int> slot_len;
  std::vector<bool> slot_used;
  // streamed KV layout (decode_stream.cu): one pool per layer, [page][kv head][K|V][64][128], chunk-swizzled rows
  std::vector<__half*> kv_pool;
  // decode graphs, keyed by batch size: nothing that varies between calls (max_new / eos / pad / sampling / forced ids) is a
  // kernel parameter -- it lives in device memory (sst.ctrl)
  std::map<int, cudaGraphExec_t> graphs;
  // prefill graphs keyed by (B, S): the 32 x 7 launches of the decoder layers only depend on the shapes (workspaces, weights,
  // KV pool and the slot-id array are context-owned), so a repeated shape replays one graph instead of 224 launches
  std::map<std::pair<int, int>, cudaGraphExec_t> prefill_graphs;
  std::map<std::pair<int, int>, int> prefill_seen;
  int prefill_graph = 1;          // TRACE_B200_PREFILL_GRAPH=0: plain launches
  // decode runs on an engine-owned stream (graph capture is illegal on the legacy default stream)
  cudaStream_t dstream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  float last_decode_ms = 0.f;
  int last_decode_steps = 0;
  // fused decode step (decode_mega.cu)
  void* d_layer_table = nullptr;
  unsigned long long* d_dbg = nullptr;
  int max_stages = 0;             // TRACE_B200_STAGES: cap on the ring stages of the decode step (0 = what fits)
  // fused decode step (decode_stream.cu)
  int poll_sleep_ns = 0, watchdog_ms = 4000;
  int cooperative = 1;
  void* d_exchange = nullptr;
  size_t exchange_bytes = 0;
  unsigned* d_epoch = nullptr;
  int64_t launches_since_reset = 0;
  int* h_err = nullptr;           // mapped pinned host word: the kernel's watchdog writes it before it traps
  int* d_err = nullptr;
  int32_t* d_ctrl = nullptr;
  int64_t* d_forced = nullptr;
  float* d_samp_logits = nullptr;
  DecodeStreamState sst{};
  // sampling / teacher forcing set through the ABI, consumed by the next decode call
  int do_sample = 0, top_k = 50;
  float temperature = 1.0f;
  uint64_t seed = 0;
  std::vector<int64_t> forced_host;
  int forced_B = 0, forced_n = 0;
  // continuous batching (trace_b200_batch_*): an open batch of `batch_rows` decode rows, rows admitted / retired between bursts
  int32_t* d_prefill_slots = nullptr;   // slot ids of the running prefill (its own array: a prefill may run while a batch is open)
  int batch_rows = 0;                   // 0 = no batch open
  int batch_eos = 2, batch_pad = 2;
  int batch_dummy_slot = -1;            // KV slot (1 page) the inactive rows point at
  std::vector<int> batch_slot;          // slot of every row, -1 = free
  std::vector<int> batch_budget;
};

namespace {

template <typename T>
int dalloc(trace_b200_ctx* ctx, T** p, size_t count, bool zero = false) {
  void* q = nullptr;
  TB_CUDA(cudaMalloc(&q, count * sizeof(T) + 256));
  if (zero) TB_CUDA(cudaMemset(q, 0, count * sizeof(T) + 256));
  ctx->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

int walloc(trace_b200_ctx* ctx, __half** p, size_t count, bool zero = false) {
  ctx->weight_bytes += static_cast<int64_t>(count) * 2;
  return dalloc(ctx, p, count, zero);
}

// cos/sin tables exactly as the reference builds them: inv_freq and the outer product in fp32, cos/sin in
// fp32, then cast to the activation dtype (fp16) -- stored here as fp32 holding the fp16-rounded value.
int build_rope_table(trace_b200_ctx* ctx, float** cos_d, float** sin_d, int npos, int dim, float base) {
  const int hd2 = dim / 2;
  std::vector<float> c(static_cast<size_t>(npos) * hd2), s(c.size());
  for (int j = 0; j < hd2; ++j) {
    const float inv = 1.0f / powf(base, static_cast<float>(2 * j) / static_cast<float>(dim));
    for (int p = 0; p < npos; ++p) {
      const float fr = static_cast<float>(p) * inv;
      c[static_cast<size_t>(p) * hd2 + j] = __half2float(__float2half_rn(static_cast<float>(cos(static_cast<double>(fr)))));
      s[static_cast<size_t>(p) * hd2 + j] = __half2float(__float2half_rn(static_cast<float>(sin(static_cast<double>(fr)))));
    }
  }
  if (dalloc(ctx, cos_d, c.size())) return 1;
  if (dalloc(ctx, sin_d, s.size())) return 1;
  TB_CUDA(cudaMemcpy(*cos_d, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemcpy(*sin_d, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

int allocate_all(trace_b200_ctx* ctx) {
  const trace_b200_config& c = ctx->c;
  const int D = c.vit_width, H = c.hidden;
  ctx->P = (c.image_size / c.patch_size) * (c.image_size / c.patch_size);
  const int kraw = 3 * c.patch_size * c.patch_size;
  ctx->Kp = (kraw + 7) / 8 * 8;
  ctx->qkvw = (c.heads + 2 * c.kv_heads) * c.head_dim;
  ctx->total_vocab = c.vocab + 27;
  ctx->head_rows = (ctx->total_vocab + 15) / 16 * 16;
  ctx->max_pages_seq = ceil_div(c.max_ctx, kPage);
  const int P = ctx->P, NT = P + 1;
  const std::string vt = "model.vision_tower.vision_tower.vision_model.";

  // ---- ViT ----
  if (walloc(ctx, &ctx->patch_w, static_cast<size_t>(D) * ctx->Kp, true)) return 1;
  if (walloc(ctx, &ctx->cls, D) || walloc(ctx, &ctx->pos, static_cast<size_t>(NT) * D) || walloc(ctx, &ctx->pre_w, D) ||
      walloc(ctx, &ctx->pre_b, D))
    return 1;
  ctx->expected = {vt + "embeddings.class_embedding", vt + "embeddings.patch_embedding.weight",
                   vt + "embeddings.position_embedding.weight", vt + "pre_layrnorm.weight", vt + "pre_layrnorm.bias"};
  const int used_layers = c.vit_layers - 1;   // hidden_states[-2]: the last block is never needed
  ctx->vit.resize(used_layers);
  for (int i = 0; i < used_layers; ++i) {
    VitLayer& l = ctx->vit[i];
    if (walloc(ctx, &l.ln1_w, D) || walloc(ctx, &l.ln1_b, D) || walloc(ctx, &l.qkv_w, static_cast<size_t>(3) * D * D) ||
        walloc(ctx, &l.qkv_b, 3 * D) || walloc(ctx, &l.out_w, static_cast<size_t>(D) * D) || walloc(ctx, &l.out_b, D) ||
        walloc(ctx, &l.ln2_w, D) || walloc(ctx, &l.ln2_b, D) || walloc(ctx, &l.fc1_w, static_cast<size_t>(c.vit_mlp) * D) ||
        walloc(ctx, &l.fc1_b, c.vit_mlp) || walloc(ctx, &l.fc2_w, static_cast<size_t>(D) * c.vit_mlp) ||
        walloc(ctx, &l.fc2_b, D))
      return 1;
    const std::string p = vt + "encoder.layers." + std::to_string(i) + ".";
    for (const char* n : {"self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "layer_norm1",
                          "layer_norm2", "mlp.fc1", "mlp.fc2"}) {
      ctx->expected.push_back(p + n + ".weight");
      ctx->expected.push_back(p + n + ".bias");
    }
  }
  // ---- projector + towers ----
  if (walloc(ctx, &ctx->pj_slots, static_cast<size_t>(D) * 8) || walloc(ctx, &ctx->pj_ln_w, D) ||
      walloc(ctx, &ctx->pj_ln_b, D) || walloc(ctx, &ctx->pj_readout, static_cast<size_t>(H) * D) ||
      walloc(ctx, &ctx->time_tab, static_cast<size_t>(13) * H) || walloc(ctx, &ctx->score_tab, static_cast<size_t>(13) * H) ||
      walloc(ctx, &ctx->sync_tab, H))
    return 1;
  for (const char* n : {"model.mm_projector.slots", "model.mm_projector.ln_vision.weight", "model.mm_projector.ln_vision.bias",
                        "model.mm_projector.readout.weight", "model.time_tower.embed_tokens.weight",
                        "model.score_tower.embed_tokens.weight", "model.sync_tower.embed_tokens.weight"})
    ctx->expected.push_back(n);
  if (build_rope_table(ctx, &ctx->pj_cos, &ctx->pj_sin, P, D, 10000.0f)) return 1;   // SlotRotaryEmbedding base 1e4
  // ---- LLM ----
  if (walloc(ctx, &ctx->embed, static_cast<size_t>(c.vocab) * H) || walloc(ctx, &ctx->final_norm, H) ||
      walloc(ctx, &ctx->head_w, static_cast<size_t>(ctx->head_rows) * H, true))
    return 1;
  for (const char* n : {"model.embed_tokens.weight", "model.norm.weight", "lm_head.weight", "sync_head.weight",
                        "time_head.weight", "score_head.weight"})
    ctx->expected.push_back(n);
  ctx->llm.resize(c.layers);
  const size_t qd = static_cast<size_t>(c.heads) * c.head_dim;
  for (int i = 0; i < c.layers; ++i) {
    LlmLayer& l = ctx->llm[i];
    if (walloc(ctx, &l.ln1, H) || walloc(ctx, &l.qkv_w, static_cast<size_t>(ctx->qkvw) * H) || walloc(ctx, &l.o_w, H * qd) ||
        walloc(ctx, &l.ln2, H) || walloc(ctx, &l.gateup_w, static_cast<size_t>(2) * c.ffn * H) ||
        walloc(ctx, &l.down_w, static_cast<size_t>(H) * c.ffn))
      return 1;
    const std::string p = "model.layers." + std::to_string(i) + ".";
    for (const char* n : {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                          "self_attn.o_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight",
                          "input_layernorm.weight", "post_attention_layernorm.weight"})
      ctx->expected.push_back(p + n);
  }
  if (build_rope_table(ctx, &ctx->rope_cos, &ctx->rope_sin, c.max_ctx, c.head_dim, c.rope_theta)) return 1;

  // ---- workspaces ----
  const size_t Fc = c.vit_chunk_frames;
  if (dalloc(ctx, &ctx->v_col, Fc * P * ctx->Kp) || dalloc(ctx, &ctx->v_patch, Fc * P * D) || dalloc(ctx, &ctx->v_h, Fc * NT * D) ||
      dalloc(ctx, &ctx->v_y, Fc * NT * D) || dalloc(ctx, &ctx->v_qkv, Fc * NT * 3 * D) || dalloc(ctx, &ctx->v_attn, Fc * NT * D) ||
      dalloc(ctx, &ctx->v_mlp, Fc * NT * c.vit_mlp) || dalloc(ctx, &ctx->v_res, Fc * 8 * D) ||
      dalloc(ctx, &ctx->pj_scratch, static_cast<size_t>(slot_pool_scratch_floats(static_cast<int>(Fc), static_cast<int>(P)))))
    return 1;
  const size_t Mp = c.max_prefill_tokens;
  if (dalloc(ctx, &ctx->p_h, Mp * H) || dalloc(ctx, &ctx->p_y, Mp * H) || dalloc(ctx, &ctx->p_qkv, Mp * ctx->qkvw) ||
      dalloc(ctx, &ctx->p_attn, Mp * qd) || dalloc(ctx, &ctx->p_act, Mp * c.ffn))
    return 1;
  if (dalloc(ctx, &ctx->d_x, static_cast<size_t>(kMaxB) * H)) return 1;
  if (dalloc(ctx, &ctx->sst.pos, kMaxB, true) || dalloc(ctx, &ctx->sst.heads, kMaxB, true) ||
      dalloc(ctx, &ctx->sst.finished, kMaxB, true) || dalloc(ctx, &ctx->sst.step, 1, true) ||
      dalloc(ctx, &ctx->sst.done_step, 1, true) || dalloc(ctx, &ctx->sst.out_ids, static_cast<size_t>(kMaxB) * c.max_ctx, true) ||
      dalloc(ctx, &ctx->sst.slot_ids, kMaxB, true) || dalloc(ctx, &ctx->sst.row_step, kMaxB, true) ||
      dalloc(ctx, &ctx->sst.row_max, kMaxB, true) || dalloc(ctx, &ctx->d_prefill_slots, kMaxB, true))
    return 1;
  if (getenv("TRACE_B200_MEGA_TIMING") != nullptr && dalloc(ctx, &ctx->d_dbg, static_cast<size_t>(2) * c.layers * 32 + static_cast<size_t>(ctx->num_sms) * 32, true)) return 1;
  {
    auto env_int = [](const char* name, int lo, int hi, int dflt) {
      const char* e = getenv(name);
      return (e != nullptr && atoi(e) >= lo && atoi(e) <= hi) ? atoi(e) : dflt;
    };
    ctx->cooperative = env_int("TRACE_B200_COOP", 0, 1, 1);
    ctx->prefill_graph = env_int("TRACE_B200_PREFILL_GRAPH", 0, 1, 1);
    ctx->max_stages = env_int("TRACE_B200_STAGES", 0, 8, 0);
    ctx->poll_sleep_ns = env_int("TRACE_B200_POLL_SLEEP", 0, 5000, 0);
    ctx->watchdog_ms = env_int("TRACE_B200_WATCHDOG_MS", 1, 3600000, 4000);    // compute-sanitizer runs need minutes
  }
  // ---- KV pool ----
  const size_t page_elems = static_cast<size_t>(c.kv_heads) * kPage * c.head_dim;
  ctx->kv_pool.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    // zero-filled: a partially filled page is streamed whole and its unwritten rows meet probability 0 in P.V
    if (dalloc(ctx, &ctx->kv_pool[i], 2 * page_elems * c.kv_pages, true)) return 1;
  }
  {
    size_t offs[16];
    ctx->exchange_bytes = decode_stream_exchange_bytes(c.hidden, ctx->qkvw, c.heads * c.head_dim, c.ffn, ctx->num_sms, offs);
    uint8_t* ex = nullptr;
    if (dalloc(ctx, &ex, ctx->exchange_bytes, true)) return 1;
    ctx->d_exchange = ex;
    if (dalloc(ctx, &ctx->d_epoch, 1) || dalloc(ctx, &ctx->d_ctrl, 16, true) ||
        dalloc(ctx, &ctx->d_forced, static_cast<size_t>(kMaxB) * c.max_ctx, true) ||
        dalloc(ctx, &ctx->d_samp_logits, static_cast<size_t>(kMaxB) * ctx->total_vocab, true))
      return 1;
    const unsigned one = 1;
    TB_CUDA(cudaMemcpy(ctx->d_epoch, &one, 4, cudaMemcpyHostToDevice));
    TB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_err), 4, cudaHostAllocMapped));
    *ctx->h_err = 0;
    TB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_err), ctx->h_err, 0));
    ctx->sst.ctrl = ctx->d_ctrl; ctx->sst.forced = ctx->d_forced;
  }
  if (dalloc(ctx, &ctx->d_block_table, static_cast<size_t>(kMaxSlots) * ctx->max_pages_seq, true)) return 1;
  ctx->h_block_table.assign(static_cast<size_t>(kMaxSlots) * ctx->max_pages_seq, 0);
  for (int i = c.kv_pages - 1; i >= 0; --i) ctx->free_pages.push_back(i);
  ctx->slot_pages.resize(kMaxSlots);
  ctx->slot_len.assign(kMaxSlots, 0);
  ctx->slot_used.assign(kMaxSlots, false);
  return 0;
}

bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

int copy_rows(void* dst, size_t dst_pitch_elems, const void* src, size_t src_pitch_elems, size_t width_elems, size_t rows,
              cudaStream_t s) {
  TB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch_elems * 2, src, src_pitch_elems * 2, width_elems * 2, rows, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int64_t numel(const int64_t* shape, int ndim) {
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  return n;
}

#define EXPECT_NUMEL(n)                                                                                         \
  TB_CHECK(numel(shape, ndim) == static_cast<int64_t>(n), "shape mismatch for " + name + ": got " +            \
                                                              std::to_string(numel(shape, ndim)) + " elements, want " + \
                                                              std::to_string(static_cast<int64_t>(n)))

int load_weight_impl(trace_b200_ctx* ctx, const std::string& name, const void* src, const int64_t* shape, int ndim,
                     cudaStream_t s) {
  const trace_b200_config& c = ctx->c;
  const size_t D = c.vit_width, H = c.hidden, hd = c.head_dim;
  const std::string vt = "model.vision_tower.vision_tower.vision_model.";
  auto plain = [&](__half* dst, size_t n) -> int {
    EXPECT_NUMEL(n);
    TB_CUDA(cudaMemcpyAsync(dst, src, n * 2, cudaMemcpyDeviceToDevice, s));
    return 0;
  };
  if (name.rfind(vt, 0) == 0) {
    const std::string r = name.substr(vt.size());
    if (r == "embeddings.class_embedding") return plain(ctx->cls, D);
    if (r == "embeddings.position_embedding.weight") return plain(ctx->pos, (ctx->P + 1) * D);
    if (r == "pre_layrnorm.weight") return plain(ctx->pre_w, D);
    if (r == "pre_layrnorm.bias") return plain(ctx->pre_b, D);
    if (r == "embeddings.patch_embedding.weight") {
      const size_t kraw = 3 * c.patch_size * c.patch_size;
      EXPECT_NUMEL(D * kraw);
      return copy_rows(ctx->patch_w, ctx->Kp, src, kraw, kraw, D, s);
    }
    if (r.rfind("post_layernorm.", 0) == 0) return 0;   // never used: features come from hidden_states[-2]
    if (r.rfind("encoder.layers.", 0) == 0) {
      const size_t dot = r.find('.', 15);
      const int li = atoi(r.substr(15, dot - 15).c_str());
      if (li >= static_cast<int>(ctx->vit.size())) return 0;   // the last block is never executed
      VitLayer& l = ctx->vit[li];
      const std::string t = r.substr(dot + 1);
      if (t == "layer_norm1.weight") return plain(l.ln1_w, D);
      if (t == "layer_norm1.bias") return plain(l.ln1_b, D);
      if (t == "layer_norm2.weight") return plain(l.ln2_w, D);
      if (t == "layer_norm2.bias") return plain(l.ln2_b, D);
      if (t == "self_attn.q_proj.weight") return plain(l.qkv_w, D * D);
      if (t == "self_attn.k_proj.weight") return plain(l.qkv_w + D * D, D * D);
      if (t == "self_attn.v_proj.weight") return plain(l.qkv_w + 2 * D * D, D * D);
      if (t == "self_attn.q_proj.bias") return plain(l.qkv_b, D);
      if (t == "self_attn.k_proj.bias") return plain(l.qkv_b + D, D);
      if (t == "self_attn.v_proj.bias") return plain(l.qkv_b + 2 * D, D);
      if (t == "self_attn.out_proj.weight") return plain(l.out_w, D * D);
      if (t == "self_attn.out_proj.bias") return plain(l.out_b, D);
      if (t == "mlp.fc1.weight") return plain(l.fc1_w, c.vit_mlp * D);
      if (t == "mlp.fc1.bias") return plain(l.fc1_b, c.vit_mlp);
      if (t == "mlp.fc2.weight") return plain(l.fc2_w, D * c.vit_mlp);
      if (t == "mlp.fc2.bias") return plain(l.fc2_b, D);
    }
    TB_CHECK(false, "unknown vision tower tensor " + name);
  }
  if (name == "model.mm_projector.slots") return plain(ctx->pj_slots, D * 8);
  if (name == "model.mm_projector.ln_vision.weight") return plain(ctx->pj_ln_w, D);
  if (name == "model.mm_projector.ln_vision.bias") return plain(ctx->pj_ln_b, D);
  if (name == "model.mm_projector.readout.weight") return plain(ctx->pj_readout, H * D);
  if (name == "model.time_tower.embed_tokens.weight") return plain(ctx->time_tab, 13 * H);
  if (name == "model.score_tower.embed_tokens.weight") return plain(ctx->score_tab, 13 * H);
  if (name == "model.sync_tower.embed_tokens.weight") return plain(ctx->sync_tab, H);
  if (name == "model.embed_tokens.weight") return plain(ctx->embed, static_cast<size_t>(c.vocab) * H);
  if (name == "model.norm.weight") return plain(ctx->final_norm, H);
  // stacked heads: rows are global token ids  [lm_head | sync_head | time_head | score_head]  (trace_mistral.py:248)
  if (name == "lm_head.weight") return plain(ctx->head_w, static_cast<size_t>(c.vocab) * H);
  if (name == "sync_head.weight") return plain(ctx->head_w + static_cast<size_t>(c.vocab) * H, H);
  if (name == "time_head.weight") return plain(ctx->head_w + static_cast<size_t>(c.vocab + 1) * H, 13 * H);
  if (name == "score_head.weight") return plain(ctx->head_w + static_cast<size_t>(c.vocab + 14) * H, 13 * H);
  if (name.rfind("model.layers.", 0) == 0) {
    const size_t dot = name.find('.', 13);
    const int li = atoi(name.substr(13, dot - 13).c_str());
    TB_CHECK(li >= 0 && li < c.layers, "layer index out of range in " + name);
    LlmLayer& l = ctx->llm[li];
    const std::string t = name.substr(dot + 1);
    const size_t qd = static_cast<size_t>(c.heads) * hd, kvd = static_cast<size_t>(c.kv_heads) * hd;
    if (t == "input_layernorm.weight") return plain(l.ln1, H);
    if (t == "post_attention_layernorm.weight") return plain(l.ln2, H);
    if (t == "self_attn.q_proj.weight") return plain(l.qkv_w, qd * H);
    if (t == "self_attn.k_proj.weight") return plain(l.qkv_w + qd * H, kvd * H);
    if (t == "self_attn.v_proj.weight") return plain(l.qkv_w + (qd + kvd) * H, kvd * H);
    if (t == "self_attn.o_proj.weight") return plain(l.o_w, H * qd);
    if (t == "mlp.down_proj.weight") return plain(l.down_w, H * c.ffn);
    // gate/up interleaved row-wise: row 2j = gate_j, row 2j+1 = up_j  (SwiGLU epilogues pair adjacent outputs)
    if (t == "mlp.gate_proj.weight") { EXPECT_NUMEL(static_cast<size_t>(c.ffn) * H); return copy_rows(l.gateup_w, 2 * H, src, H, H, c.ffn, s); }
    if (t == "mlp.up_proj.weight") { EXPECT_NUMEL(static_cast<size_t>(c.ffn) * H); return copy_rows(l.gateup_w + H, 2 * H, src, H, H, c.ffn, s); }
  }
  TB_CHECK(false, "unknown tensor name " + name);
}

// ------------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------------
#define RUN(x)            \
  do {                    \
    if ((x) != 0) return 1; \
    ctx->launches++;      \
  } while (0)

int vit_chunk(trace_b200_ctx* ctx, const __half* pixels, int F, __half* res_out, __half* feats_out, cudaStream_t s) {
  const trace_b200_config& c = ctx->c;
  const int D = c.vit_width, P = ctx->P, NT = P + 1, M = F * NT;
  const int sms = ctx->num_sms;
  RUN(vit_im2col(pixels, ctx->v_col, F, c.image_size, c.patch_size, ctx->Kp, s));
  RUN(gemm_f16(ctx->v_col, ctx->Kp, ctx->patch_w, ctx->Kp, ctx->v_patch, D, F * P, D, ctx->Kp, nullptr, nullptr, 0, 0, 0, sms, s));
  RUN(vit_embed_ln(ctx->v_patch, ctx->cls, ctx->pos, ctx->pre_w, ctx->pre_b, c.vit_ln_eps, ctx->v_h, F, P, D, s));
  const int hdv = D / c.vit_heads;
  for (const VitLayer& l : ctx->vit) {
    RUN(layernorm_f16(ctx->v_h, D, l.ln1_w, l.ln1_b, c.vit_ln_eps, ctx->v_y, D, M, D, s));
    RUN(gemm_f16(ctx->v_y, D, l.qkv_w, D, ctx->v_qkv, 3 * D, M, 3 * D, D, l.qkv_b, nullptr, 0, 0, 0, sms, s));
    RUN(attention_tc_fwd(ctx->v_qkv, 3 * D, ctx->v_attn, NT, D, static_cast<int64_t>(NT) * D, 0, D, 2 * D, c.vit_heads, 1, F, hdv, 0,
                         1.0f / sqrtf(static_cast<float>(hdv)), s));
    RUN(gemm_f16(ctx->v_attn, D, l.out_w, D, ctx->v_h, D, M, D, D, l.out_b, ctx->v_h, D, 0, 0, sms, s));
    RUN(layernorm_f16(ctx->v_h, D, l.ln2_w, l.ln2_b, c.vit_ln_eps, ctx->v_y, D, M, D, s));
    RUN(gemm_f16(ctx->v_y, D, l.fc1_w, D, ctx->v_mlp, c.vit_mlp, M, c.vit_mlp, D, l.fc1_b, nullptr, 0, 1, 0, sms, s));
    RUN(gemm_f16(ctx->v_mlp, c.vit_mlp, l.fc2_w, c.vit_mlp, ctx->v_h, D, M, D, c.vit_mlp, l.fc2_b, ctx->v_h, D, 0, 0, sms, s));
  }
  if (feats_out != nullptr) {   // hidden_states[-2][:, 1:]  (clip_encoder.py:32-35)
    TB_CUDA(cudaMemcpy2DAsync(feats_out, static_cast<size_t>(P) * D * 2, ctx->v_h + D, static_cast<size_t>(NT) * D * 2,
                              static_cast<size_t>(P) * D * 2, F, cudaMemcpyDeviceToDevice, s));
  }
  RUN(slot_pool_fused(ctx->v_h, D, NT, 1, P, D, F, ctx->pj_ln_w, ctx->pj_ln_b, 1e-6f, ctx->pj_slots, ctx->pj_cos,
                      ctx->pj_sin, ctx->pj_scratch, ctx->v_res, s));
  RUN(gemm_f16(ctx->v_res, D, ctx->pj_readout, D, res_out, c.hidden, F * 8, c.hidden, D, nullptr, nullptr, 0, 0, 0, sms, s));
  return 0;
}

int decoder_layers_prefill(trace_b200_ctx* ctx, int B, int S, const int32_t* slot_ids_dev, cudaStream_t s) {
  const trace_b200_config& c = ctx->c;
  const int H = c.hidden, M = B * S, qd = c.heads * c.head_dim, kvd = c.kv_heads * c.head_dim, sms = ctx->num_sms;
  for (int i = 0; i < c.layers; ++i) {
    const LlmLayer& l = ctx->llm[i];
    RUN(rmsnorm_f16(ctx->p_h, H, l.ln1, c.rms_eps, ctx->p_y, H, M, H, s));
    RUN(gemm_f16(ctx->p_y, H, l.qkv_w, H, ctx->p_qkv, ctx->qkvw, M, ctx->qkvw, H, nullptr, nullptr, 0, 0, 0, sms, s));
    RUN(rope_kv_prefill(ctx->p_qkv, B, S, c.heads, c.kv_heads, c.head_dim, ctx->rope_cos, ctx->rope_sin,
                        ctx->kv_pool[i], ctx->d_block_table, slot_ids_dev, ctx->max_pages_seq, kPage, s));
    RUN(attention_tc_fwd(ctx->p_qkv, ctx->qkvw, ctx->p_attn, S, qd, static_cast<int64_t>(S) * qd, 0, qd, qd + kvd, c.heads,
                         c.heads / c.kv_heads, B, c.head_dim, 1, 1.0f / sqrtf(static_cast<float>(c.head_dim)), s));
    RUN(gemm_f16(ctx->p_attn, qd, l.o_w, qd, ctx->p_h, H, M, H, qd, nullptr, ctx->p_h, H, 0, 0, sms, s));
    RUN(rmsnorm_f16(ctx->p_h, H, l.ln2, c.rms_eps, ctx->p_y, H, M, H, s));
    RUN(gemm_f16(ctx->p_y, H, l.gateup_w, H, ctx->p_act, c.ffn, M, 2 * c.ffn, H, nullptr, nullptr, 0, 0, 1, sms, s));
    RUN(gemm_f16(ctx->p_act, c.ffn, l.down_w, c.ffn, ctx->p_h, H, M, H, c.ffn, nullptr, ctx->p_h, H, 0, 0, sms, s));
  }
  return 0;
}

DecodeStreamArgs stream_args(trace_b200_ctx* ctx, int B, int layers_run, float* logits_out) {
  const trace_b200_config& c = ctx->c;
  DecodeStreamArgs a{};
  a.B = B; a.H = c.hidden; a.qkvw = ctx->qkvw; a.qd = c.heads * c.head_dim; a.ffn = c.ffn; a.layers = layers_run; a.heads = c.heads;
  a.kvh = c.kv_heads; a.vocab = c.vocab; a.head_rows = ctx->head_rows; a.n_valid = ctx->total_vocab; a.max_ctx = c.max_ctx;
  a.num_sms = ctx->num_sms; a.page_size = kPage; a.max_pages = ctx->max_pages_seq; a.eps = c.rms_eps;
  a.layer_table_dev = ctx->d_layer_table; a.final_norm = ctx->final_norm; a.head_w = ctx->head_w; a.x = ctx->d_x;
  a.exchange = ctx->d_exchange; a.cos_tab = ctx->rope_cos; a.sin_tab = ctx->rope_sin; a.block_table = ctx->d_block_table;
  a.logits_out = logits_out; a.samp_logits = ctx->d_samp_logits; a.embed = ctx->embed; a.time_tab = ctx->time_tab;
  a.score_tab = ctx->score_tab; a.sync_tab = ctx->sync_tab; a.st = ctx->sst; a.epoch_ctr = ctx->d_epoch; a.err = ctx->d_err;
  a.advance = layers_run > 0 ? 1 : 0; a.cooperative = ctx->cooperative; a.max_stages = ctx->max_stages; a.poll_sleep_ns = ctx->poll_sleep_ns; a.watchdog_ms = ctx->watchdog_ms; a.dbg = ctx->d_dbg;
  return a;
}

// one launch of the fused decode step (decode_stream.cu).  layers_run = 0: heads only, on the prefill's last hidden state
int decode_step(trace_b200_ctx* ctx, int B, int layers_run, float* logits_out, cudaStream_t s) {
  const DecodeStreamArgs a = stream_args(ctx, B, layers_run, logits_out);
  RUN(decode_stream_step(a, s));
  return 0;
}

// exchange epochs are 32-bit (512 per launch): long before they wrap, clear the buffers and restart the counter
int epoch_guard(trace_b200_ctx* ctx, cudaStream_t s) {
  if (ctx->launches_since_reset > (1 << 22)) {
    static const unsigned one = 1;
    TB_CUDA(cudaMemsetAsync(ctx->d_exchange, 0, ctx->exchange_bytes, s));
    TB_CUDA(cudaMemcpyAsync(ctx->d_epoch, &one, 4, cudaMemcpyHostToDevice, s));
    ctx->launches_since_reset = 0;
  }
  return 0;
}

int decode_loop(trace_b200_ctx* ctx, const int32_t* slots_host, int B, int32_t* heads_inout_host, int max_new, int eos, int pad,
              int64_t* ids_out_host, int* n_out, float* logits_out_dev, cudaStream_t s) {
  const trace_b200_config& c = ctx->c;
  TB_CHECK(ctx->batch_rows == 0, "a continuous batch is open on this context (trace_b200_batch_close first)");
  std::vector<int32_t> pos(B), zeros(kMaxB, 0), budget(kMaxB, max_new);
  for (int b = 0; b < B; ++b) pos[b] = ctx->slot_len[slots_host[b]];
  int32_t ctrl[16] = {0};
  ctrl[0] = max_new; ctrl[1] = eos; ctrl[2] = pad; ctrl[3] = ctx->do_sample; ctrl[9] = (1 << B) - 1;
  memcpy(&ctrl[4], &ctx->temperature, 4);
  ctrl[5] = ctx->top_k; ctrl[6] = static_cast<int32_t>(ctx->seed & 0xffffffffu); ctrl[7] = static_cast<int32_t>(ctx->seed >> 32);
  const bool forced = !ctx->forced_host.empty();
  if (forced) {
    TB_CHECK(ctx->forced_B == B && ctx->forced_n == max_new, "forced ids were set for another batch size / max_new");
    TB_CUDA(cudaMemcpyAsync(ctx->d_forced, ctx->forced_host.data(), ctx->forced_host.size() * 8, cudaMemcpyHostToDevice, s));
    ctrl[8] = 1;
  }
  if (epoch_guard(ctx, s)) return 1;
  TB_CUDA(cudaMemcpyAsync(ctx->d_ctrl, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.slot_ids, slots_host, B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.pos, pos.data(), B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.heads, heads_inout_host, B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.finished, zeros.data(), B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.row_step, zeros.data(), B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.row_max, budget.data(), B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.step, zeros.data(), 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.done_step, zeros.data(), 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaStreamSynchronize(s));     // ctrl / pos / forced_host are host stack or heap memory
  ctx->forced_host.clear();
  const size_t lstride = static_cast<size_t>(B) * ctx->total_vocab;

  auto fail = [&](cudaError_t e, const char* what) -> int {
    tb::set_error(std::string(what) + " failed: " + cudaGetErrorString(e) + " (decode watchdog word " + std::to_string(*ctx->h_err) +
                  ": code*1000 + CTA; 21/22 activation wait, 31-33 ring, 41-44 attention exchange, 51 candidates)");
    return 1;
  };
  int done = 0;
  auto poll_done = [&]() -> int {
    cudaError_t e = cudaMemcpyAsync(&done, ctx->sst.done_step, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(e, "decode step");
    return 0;
  };
  TB_CUDA(cudaEventRecord(ctx->ev_t0, s));
  // token 0: heads on the prefill's last hidden state, no decoder pass, positions do not advance
  if (decode_step(ctx, B, 0, logits_out_dev, s)) return 1;
  ++ctx->launches_since_reset;
  int steps_run = 1;
  if (logits_out_dev != nullptr || max_new <= 2) {
    // instrumented path: one eager step per token (parity tests read every step's logits)
    for (; steps_run < max_new; ++steps_run) {
      if (poll_done()) return 1;
      if (done) break;
      if (decode_step(ctx, B, c.layers, logits_out_dev ? logits_out_dev + lstride * steps_run : nullptr, s)) return 1;
      ++ctx->launches_since_reset;
    }
  } else {
    // production path: step 1 eager (also sets kernel attributes), then one captured step replayed.  Nothing that varies
    // between calls (max_new, eos, pad, sampling, forced ids) is a kernel parameter, so the graph is keyed by batch only.
    if (decode_step(ctx, B, c.layers, nullptr, s)) return 1;
    ++steps_run;
    cudaGraphExec_t exec = nullptr;
    auto it = ctx->graphs.find(B);
    if (it != ctx->graphs.end()) {
      exec = it->second;
    } else {
      cudaGraph_t graph = nullptr;
      TB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      const int64_t before = ctx->launches;
      const int rc = decode_step(ctx, B, c.layers, nullptr, s);
      ctx->launches = before;
      cudaError_t ce = cudaStreamEndCapture(s, &graph);
      TB_CHECK(rc == 0, std::string("decode step failed during graph capture: ") + tb::get_error());
      TB_CHECK(ce == cudaSuccess && graph != nullptr, std::string("graph capture failed: ") + cudaGetErrorString(ce));
      TB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      ctx->graphs[B] = exec;
    }
    while (steps_run < max_new) {
      const int burst = (max_new - steps_run < 16) ? (max_new - steps_run) : 16;
      for (int i = 0; i < burst; ++i) TB_CUDA(cudaGraphLaunch(exec, s));
      steps_run += burst;
      ctx->launches += burst;
      ctx->launches_since_reset += burst;
      if (poll_done()) return 1;
      if (done) break;
    }
  }
  TB_CUDA(cudaEventRecord(ctx->ev_t1, s));
  if (poll_done()) return 1;
  TB_CHECK(*ctx->h_err == 0, "decode watchdog fired: " + std::to_string(*ctx->h_err));
  TB_CUDA(cudaEventElapsedTime(&ctx->last_decode_ms, ctx->ev_t0, ctx->ev_t1));
  ctx->last_decode_steps = steps_run;
  const int n = done ? done : (steps_run < max_new ? steps_run : max_new);
  std::vector<int64_t> tmp(static_cast<size_t>(B) * max_new);
  TB_CUDA(cudaMemcpyAsync(tmp.data(), ctx->sst.out_ids, tmp.size() * 8, cudaMemcpyDeviceToHost, s));
  TB_CUDA(cudaMemcpyAsync(heads_inout_host, ctx->sst.heads, B * 4, cudaMemcpyDeviceToHost, s));
  TB_CUDA(cudaStreamSynchronize(s));
  for (int b = 0; b < B; ++b) {
    memcpy(ids_out_host + static_cast<size_t>(b) * max_new, tmp.data() + static_cast<size_t>(b) * max_new, static_cast<size_t>(n) * 8);
    ctx->slot_len[slots_host[b]] += n - 1;
  }
  *n_out = n;
  return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int trace_b200_abi_version(void) { return TRACE_B200_ABI_VERSION; }
const char* trace_b200_last_error(void) { return tb::get_error(); }

int trace_b200_create(const trace_b200_config* cfg, int device, trace_b200_ctx** out) {
  TB_CHECK(cfg != nullptr && out != nullptr, "null argument");
  TB_CHECK(cfg->vit_width % 64 == 0 && cfg->vit_width / cfg->vit_heads == 64, "ViT head_dim must be 64");
  TB_CHECK(cfg->head_dim == 128 && cfg->heads == 4 * cfg->kv_heads, "LLM attention must be 128-dim heads with 4 q per kv head");
  TB_CHECK(cfg->hidden % 256 == 0 && cfg->hidden <= 4096 && cfg->ffn % 256 == 0, "hidden/ffn must be multiples of 256, hidden <= 4096");
  TB_CHECK(cfg->vit_chunk_frames > 0 && cfg->max_prefill_tokens > 0 && cfg->max_ctx > 0 && cfg->kv_pages > 0, "workspace sizes must be positive");
  int ndev = 0;
  TB_CUDA(cudaGetDeviceCount(&ndev));
  TB_CHECK(device >= 0 && device < ndev, "no such CUDA device");
  TB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TB_CUDA(cudaGetDeviceProperties(&prop, device));
  TB_CHECK(prop.major == 10, std::string("trace_b200 kernels are sm_100a-only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
  trace_b200_ctx* ctx = new trace_b200_ctx();
  ctx->c = *cfg;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (allocate_all(ctx)) {
    trace_b200_destroy(ctx);
    return 1;
  }
  TB_CUDA(cudaStreamCreateWithFlags(&ctx->dstream, cudaStreamNonBlocking));
  TB_CUDA(cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming));
  TB_CUDA(cudaEventCreate(&ctx->ev_t0));
  TB_CUDA(cudaEventCreate(&ctx->ev_t1));
  *out = ctx;
  return 0;
}

void trace_b200_destroy(trace_b200_ctx* ctx) {
  if (ctx == nullptr) return;
  cudaSetDevice(ctx->device);
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  for (auto& kv : ctx->prefill_graphs) cudaGraphExecDestroy(kv.second);
  if (ctx->dstream) cudaStreamDestroy(ctx->dstream);
  if (ctx->ev_in) cudaEventDestroy(ctx->ev_in);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->h_err) cudaFreeHost(ctx->h_err);
  delete ctx;
}

int trace_b200_load_weight(trace_b200_ctx* ctx, const char* hf_name, const void* dev_ptr, const int64_t* shape, int ndim,
                           void* stream) {
  TB_CHECK(ctx != nullptr && hf_name != nullptr && dev_ptr != nullptr, "null argument");
  TB_CUDA(cudaSetDevice(ctx->device));
  const std::string name(hf_name);
  if (load_weight_impl(ctx, name, dev_ptr, shape, ndim, static_cast<cudaStream_t>(stream))) return 1;
  ctx->loaded.insert(name);
  return 0;
}

int trace_b200_finalize_weights(trace_b200_ctx* ctx) {
  TB_CHECK(ctx != nullptr, "null context");
  for (const std::string& n : ctx->expected) TB_CHECK(ctx->loaded.count(n) == 1, "missing weight " + n);
  {
    TB_CHECK(decode_stream_layer_bytes() == 8 * static_cast<int>(sizeof(void*)), "layer table layout mismatch");
    std::vector<const void*> tab;
    for (int i = 0; i < ctx->c.layers; ++i) {
      const LlmLayer& l = ctx->llm[i];
      for (const void* q : {static_cast<const void*>(l.ln1), static_cast<const void*>(l.qkv_w), static_cast<const void*>(l.o_w),
                            static_cast<const void*>(l.ln2), static_cast<const void*>(l.gateup_w), static_cast<const void*>(l.down_w),
                            static_cast<const void*>(ctx->kv_pool[i]), static_cast<const void*>(nullptr)})
        tab.push_back(q);
    }
    if (ctx->d_layer_table == nullptr) {
      const void** dt = nullptr;
      if (dalloc(ctx, &dt, tab.size())) return 1;
      ctx->d_layer_table = dt;
    }
    TB_CUDA(cudaMemcpy(ctx->d_layer_table, tab.data(), tab.size() * sizeof(void*), cudaMemcpyHostToDevice));
  }
  TB_CUDA(cudaDeviceSynchronize());
  ctx->finalized = true;
  return 0;
}

int trace_b200_encode_frames(trace_b200_ctx* ctx, const void* pixels, int F, void* slots_out, void* feats_out, void* stream) {
  TB_CHECK(ctx != nullptr && ctx->finalized, "context not finalized");
  TB_CHECK(pixels != nullptr && slots_out != nullptr && F >= 0, "bad arguments");
  NvtxRange nvtx("trace_b200.encode_frames");
  TB_CUDA(cudaSetDevice(ctx->device));
  const trace_b200_config& c = ctx->c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t frame_px = static_cast<size_t>(3) * c.image_size * c.image_size;
  for (int f0 = 0; f0 < F; f0 += c.vit_chunk_frames) {
    const int n = (F - f0 < c.vit_chunk_frames) ? F - f0 : c.vit_chunk_frames;
    __half* fo = feats_out ? static_cast<__half*>(feats_out) + static_cast<size_t>(f0) * ctx->P * c.vit_width : nullptr;
    if (vit_chunk(ctx, static_cast<const __half*>(pixels) + f0 * frame_px, n,
                  static_cast<__half*>(slots_out) + static_cast<size_t>(f0) * 8 * c.hidden, fo, s))
      return 1;
  }
  return 0;
}

int trace_b200_build_prefill(trace_b200_ctx* ctx, const int64_t* ids_dev, int B, int L, int T, const void* slots,
                             const int32_t* time_ids_dev, void* embeds_out, void* stream) {
  TB_CHECK(ctx != nullptr && ctx->finalized, "context not finalized");
  TB_CHECK(ids_dev && embeds_out && B > 0 && L > 0 && T >= 0, "bad arguments");
  TB_CHECK(T == 0 || (slots != nullptr && time_ids_dev != nullptr), "slots / time ids missing");
  NvtxRange nvtx("trace_b200.build_prefill");
  TB_CUDA(cudaSetDevice(ctx->device));
  RUN(splice_prefill(ids_dev, B, L, T, static_cast<const __half*>(slots), time_ids_dev, ctx->embed, ctx->time_tab, ctx->sync_tab,
                     ctx->c.vocab, ctx->c.hidden, static_cast<__half*>(embeds_out), static_cast<cudaStream_t>(stream)));
  return 0;
}

int trace_b200_alloc_slot(trace_b200_ctx* ctx, int n_positions, int* slot_out) {
  TB_CHECK(ctx != nullptr && slot_out != nullptr, "null argument");
  TB_CHECK(n_positions > 0 && n_positions <= ctx->c.max_ctx, "slot length exceeds max_ctx");
  const int need = ceil_div(n_positions, kPage);
  TB_CHECK(static_cast<int>(ctx->free_pages.size()) >= need, "KV pool exhausted");
  int slot = -1;
  for (int i = 0; i < kMaxSlots; ++i)
    if (!ctx->slot_used[i]) { slot = i; break; }
  TB_CHECK(slot >= 0, "no free KV slot");
  ctx->slot_used[slot] = true;
  ctx->slot_len[slot] = 0;
  ctx->slot_pages[slot].clear();
  for (int i = 0; i < need; ++i) {
    const int pg = ctx->free_pages.back();
    ctx->free_pages.pop_back();
    ctx->slot_pages[slot].push_back(pg);
    ctx->h_block_table[static_cast<size_t>(slot) * ctx->max_pages_seq + i] = pg;
  }
  TB_CUDA(cudaSetDevice(ctx->device));
  TB_CUDA(cudaMemcpy(ctx->d_block_table + static_cast<size_t>(slot) * ctx->max_pages_seq,
                     ctx->h_block_table.data() + static_cast<size_t>(slot) * ctx->max_pages_seq,
                     static_cast<size_t>(ctx->max_pages_seq) * 4, cudaMemcpyHostToDevice));
  *slot_out = slot;
  return 0;
}

int trace_b200_free_slot(trace_b200_ctx* ctx, int slot) {
  TB_CHECK(ctx != nullptr && slot >= 0 && slot < kMaxSlots && ctx->slot_used[slot], "bad slot");
  for (int pg : ctx->slot_pages[slot]) ctx->free_pages.push_back(pg);
  ctx->slot_pages[slot].clear();
  ctx->slot_used[slot] = false;
  ctx->slot_len[slot] = 0;
  return 0;
}

static int prefill_impl(trace_b200_ctx* ctx, const int32_t* slots_host, int B, const void* embeds, int S, void* hidden_out, int x_row0,
                        void* stream);
int trace_b200_prefill(trace_b200_ctx* ctx, const int32_t* slots_host, int B, const void* embeds, int S, void* hidden_out,
                       void* stream) {
  TB_CHECK(ctx != nullptr && ctx->batch_rows == 0, "a continuous batch is open on this context: rows are prefilled by trace_b200_batch_admit");
  return prefill_impl(ctx, slots_host, B, embeds, S, hidden_out, 0, stream);
}
// x_row0: first row of the decode activation buffer that receives the last hidden states (a row admitted into a running batch)
static int prefill_impl(trace_b200_ctx* ctx, const int32_t* slots_host, int B, const void* embeds, int S, void* hidden_out, int x_row0,
                        void* stream) {
  TB_CHECK(ctx != nullptr && ctx->finalized, "context not finalized");
  TB_CHECK(slots_host && embeds && B >= 1 && B <= kMaxB && S >= 1, "bad arguments");
  const trace_b200_config& c = ctx->c;
  TB_CHECK(static_cast<int64_t>(B) * S <= c.max_prefill_tokens, "prefill exceeds max_prefill_tokens");
  NvtxRange nvtx("trace_b200.prefill");
  TB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int b = 0; b < B; ++b) {
    const int sl = slots_host[b];
    TB_CHECK(sl >= 0 && sl < kMaxSlots && ctx->slot_used[sl], "prefill on an unallocated slot");
    TB_CHECK(static_cast<int>(ctx->slot_pages[sl].size()) * kPage >= S, "slot too small for the prompt");
    ctx->slot_len[sl] = S;
  }
  TB_CHECK(x_row0 >= 0 && x_row0 + B <= kMaxB, "bad decode row");
  TB_CUDA(cudaMemcpyAsync(ctx->d_prefill_slots, slots_host, B * 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->p_h, embeds, static_cast<size_t>(B) * S * c.hidden * 2, cudaMemcpyDeviceToDevice, s));
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);
  const bool graph_ok = ctx->prefill_graph && cap == cudaStreamCaptureStatusNone && s != nullptr && s != cudaStreamLegacy;   // capture is illegal on the legacy default stream
  const auto key = std::make_pair(B, S);
  auto it = ctx->prefill_graphs.find(key);
  if (graph_ok && it != ctx->prefill_graphs.end()) {
    TB_CUDA(cudaGraphLaunch(it->second, s));
    ctx->launches += static_cast<int64_t>(c.layers) * 7;
  } else if (graph_ok && ctx->prefill_seen[key]++ >= 1) {
    // the shape repeats (a batch of clips with one prompt, the bench): capture the launches once and replay them from now on.
    // A shape seen for the first time runs eagerly -- drivers with ever-changing prompt lengths never pay for a capture.
    cudaGraph_t graph = nullptr;
    TB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int64_t before = ctx->launches;
    const int rc = decoder_layers_prefill(ctx, B, S, ctx->d_prefill_slots, s);
    ctx->launches = before;
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    TB_CHECK(rc == 0, std::string("prefill failed during graph capture: ") + tb::get_error());
    TB_CHECK(ce == cudaSuccess && graph != nullptr, std::string("prefill graph capture failed: ") + cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    TB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    if (ctx->prefill_graphs.size() >= 16) {                                                 // bounded
      for (auto& kv : ctx->prefill_graphs) cudaGraphExecDestroy(kv.second);
      ctx->prefill_graphs.clear();
    }
    ctx->prefill_graphs[key] = exec;
    TB_CUDA(cudaGraphLaunch(exec, s));
    ctx->launches += static_cast<int64_t>(c.layers) * 7;
  } else if (decoder_layers_prefill(ctx, B, S, ctx->d_prefill_slots, s)) {
    return 1;
  }
  if (hidden_out != nullptr)
    TB_CUDA(cudaMemcpyAsync(hidden_out, ctx->p_h, static_cast<size_t>(B) * S * c.hidden * 2, cudaMemcpyDeviceToDevice, s));
  // last position of every row -> decode activation buffer
  TB_CUDA(cudaMemcpy2DAsync(ctx->d_x + static_cast<size_t>(x_row0) * c.hidden, static_cast<size_t>(c.hidden) * 2, ctx->p_h + static_cast<size_t>(S - 1) * c.hidden,
                            static_cast<size_t>(S) * c.hidden * 2, static_cast<size_t>(c.hidden) * 2, B, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int trace_b200_decode(trace_b200_ctx* ctx, const int32_t* slots_host, int B, int32_t* heads_inout_host, int max_new, int eos,
                      int pad, int64_t* ids_out_host, int* n_out, float* logits_out_dev, void* stream) {
  TB_CHECK(ctx != nullptr && ctx->finalized, "context not finalized");
  TB_CHECK(slots_host && heads_inout_host && ids_out_host && n_out && B >= 1 && B <= kMaxB && max_new >= 1, "bad arguments");
  const trace_b200_config& c = ctx->c;
  NvtxRange nvtx("trace_b200.decode");
  TB_CUDA(cudaSetDevice(ctx->device));
  // everything below runs on the engine's stream, ordered after the caller's stream (prefill was enqueued there)
  TB_CUDA(cudaEventRecord(ctx->ev_in, static_cast<cudaStream_t>(stream)));
  cudaStream_t s = ctx->dstream;
  TB_CUDA(cudaStreamWaitEvent(s, ctx->ev_in, 0));
  for (int b = 0; b < B; ++b) {
    const int sl = slots_host[b];
    TB_CHECK(sl >= 0 && sl < kMaxSlots && ctx->slot_used[sl] && ctx->slot_len[sl] > 0, "decode on a slot without prefill");
    TB_CHECK(ctx->slot_len[sl] + max_new - 1 <= static_cast<int>(ctx->slot_pages[sl].size()) * kPage, "slot too small for max_new tokens");
    TB_CHECK(ctx->slot_len[sl] + max_new - 1 <= c.max_ctx, "sequence exceeds max_ctx");
    TB_CHECK(heads_inout_host[b] >= -1 && heads_inout_host[b] <= 2, "head id must be -1, 0, 1 or 2");
  }
  TB_CHECK(max_new <= c.max_ctx, "max_new exceeds the output buffer");
  return decode_loop(ctx, slots_host, B, heads_inout_host, max_new, eos, pad, ids_out_host, n_out, logits_out_dev, s);
}

// ---- continuous batching -------------------------------------------------------------------------------------------------
namespace {
int batch_ctrl(trace_b200_ctx* ctx, int eos, int pad, int mask, cudaStream_t s) {
  int32_t ctrl[16] = {0};
  ctrl[0] = ctx->c.max_ctx; ctrl[1] = eos; ctrl[2] = pad; ctrl[3] = ctx->do_sample; ctrl[9] = mask;
  memcpy(&ctrl[4], &ctx->temperature, 4);
  ctrl[5] = ctx->top_k; ctrl[6] = static_cast<int32_t>(ctx->seed & 0xffffffffu); ctrl[7] = static_cast<int32_t>(ctx->seed >> 32);
  TB_CUDA(cudaMemcpyAsync(ctx->d_ctrl, ctrl, sizeof(ctrl), cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaStreamSynchronize(s));
  return 0;
}
int set_row(trace_b200_ctx* ctx, int row, int slot, int pos, int head, int finished, int budget, cudaStream_t s) {
  const int32_t v[6] = {slot, pos, head, finished, 0, budget};
  TB_CUDA(cudaMemcpyAsync(ctx->sst.slot_ids + row, &v[0], 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.pos + row, &v[1], 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.heads + row, &v[2], 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.finished + row, &v[3], 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.row_step + row, &v[4], 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.row_max + row, &v[5], 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaStreamSynchronize(s));
  return 0;
}
}  // namespace

int trace_b200_batch_open(trace_b200_ctx* ctx, int rows, int eos, int pad) {
  TB_CHECK(ctx != nullptr && ctx->finalized, "context not finalized");
  TB_CHECK(rows >= 1 && rows <= kMaxB, "a decode batch has 1..8 rows");
  TB_CHECK(ctx->batch_rows == 0, "a batch is already open");
  TB_CUDA(cudaSetDevice(ctx->device));
  int dummy = -1;
  if (trace_b200_alloc_slot(ctx, 1, &dummy)) return 1;
  cudaStream_t s = ctx->dstream;
  ctx->batch_dummy_slot = dummy;
  ctx->batch_slot.assign(rows, -1);
  ctx->batch_budget.assign(rows, 0);
  ctx->batch_eos = eos;
  ctx->batch_pad = pad;
  const int32_t zero = 0;
  TB_CUDA(cudaMemsetAsync(ctx->d_x, 0, static_cast<size_t>(kMaxB) * ctx->c.hidden * 2, s));   // inactive rows: x = 0 -> every projection of them is 0
  TB_CUDA(cudaMemcpyAsync(ctx->sst.step, &zero, 4, cudaMemcpyHostToDevice, s));
  TB_CUDA(cudaMemcpyAsync(ctx->sst.done_step, &zero, 4, cudaMemcpyHostToDevice, s));
  for (int r = 0; r < rows; ++r)
    if (set_row(ctx, r, dummy, 0, 0, 1, 0, s)) return 1;                                          // finished, budget 0: never advances, never emits
  ctx->batch_rows = rows;
  return 0;
}

int trace_b200_batch_admit(trace_b200_ctx* ctx, int row, int slot, const void* embeds, int S, int head, int max_new, void* stream) {
  TB_CHECK(ctx != nullptr && ctx->batch_rows > 0, "no batch open");
  TB_CHECK(row >= 0 && row < ctx->batch_rows && ctx->batch_slot[row] < 0, "row is not free");
  TB_CHECK(slot >= 0 && slot < kMaxSlots && ctx->slot_used[slot] && slot != ctx->batch_dummy_slot, "bad KV slot");
  TB_CHECK(max_new >= 1 && S + max_new <= static_cast<int>(ctx->slot_pages[slot].size()) * kPage && S + max_new <= ctx->c.max_ctx,
           "slot too small for prompt + max_new tokens");
  TB_CHECK(head >= -1 && head <= 2, "head id must be -1, 0, 1 or 2");
  NvtxRange nvtx("trace_b200.batch_admit");
  TB_CUDA(cudaSetDevice(ctx->device));
  const int32_t sl = slot;
  if (prefill_impl(ctx, &sl, 1, embeds, S, nullptr, row, stream)) return 1;
  TB_CUDA(cudaEventRecord(ctx->ev_in, static_cast<cudaStream_t>(stream)));
  cudaStream_t s = ctx->dstream;
  TB_CUDA(cudaStreamWaitEvent(s, ctx->ev_in, 0));
  if (set_row(ctx, row, slot, S, head, 0, max_new, s)) return 1;
  if (batch_ctrl(ctx, ctx->batch_eos, ctx->batch_pad, 1 << row, s)) return 1;
  // the row's first token: heads on the prefill's last hidden state; only this row is finalized
  if (decode_step(ctx, ctx->batch_rows, 0, nullptr, s)) return 1;
  ++ctx->launches_since_reset;
  TB_CUDA(cudaStreamSynchronize(s));
  ctx->batch_slot[row] = slot;
  ctx->batch_budget[row] = max_new;
  return 0;
}

int trace_b200_batch_step(trace_b200_ctx* ctx, int n_steps, int32_t* row_steps_out, int32_t* finished_out, void* stream) {
  TB_CHECK(ctx != nullptr && ctx->batch_rows > 0 && n_steps >= 1 && row_steps_out && finished_out, "bad arguments / no batch open");
  (void)stream;
  NvtxRange nvtx("trace_b200.batch_step");
  TB_CUDA(cudaSetDevice(ctx->device));
  const int B = ctx->batch_rows;
  cudaStream_t s = ctx->dstream;
  if (epoch_guard(ctx, s)) return 1;
  if (batch_ctrl(ctx, ctx->batch_eos, ctx->batch_pad, (1 << B) - 1, s)) return 1;
  cudaGraphExec_t exec = nullptr;
  auto it = ctx->graphs.find(B);
  int done_steps = 0;
  if (it != ctx->graphs.end()) {
    exec = it->second;
  } else {
    if (decode_step(ctx, B, ctx->c.layers, nullptr, s)) return 1;      // eager once (kernel attributes), then capture
    ++done_steps;
    cudaGraph_t graph = nullptr;
    TB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int64_t before = ctx->launches;
    const int rc = decode_step(ctx, B, ctx->c.layers, nullptr, s);
    ctx->launches = before;
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    TB_CHECK(rc == 0, std::string("decode step failed during graph capture: ") + tb::get_error());
    TB_CHECK(ce == cudaSuccess && graph != nullptr, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    TB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    ctx->graphs[B] = exec;
  }
  for (; done_steps < n_steps; ++done_steps) TB_CUDA(cudaGraphLaunch(exec, s));
  ctx->launches += n_steps;
  ctx->launches_since_reset += n_steps;
  TB_CUDA(cudaMemcpyAsync(row_steps_out, ctx->sst.row_step, B * 4, cudaMemcpyDeviceToHost, s));
  TB_CUDA(cudaMemcpyAsync(finished_out, ctx->sst.finished, B * 4, cudaMemcpyDeviceToHost, s));
  cudaError_t e = cudaStreamSynchronize(s);
  TB_CHECK(e == cudaSuccess && *ctx->h_err == 0, std::string("decode step failed: ") + cudaGetErrorString(e) + " (watchdog word " + std::to_string(*ctx->h_err) + ")");
  return 0;
}

int trace_b200_batch_read(trace_b200_ctx* ctx, int row, int64_t* ids_out, int capacity, int* n_out, int32_t* head_out) {
  TB_CHECK(ctx != nullptr && ctx->batch_rows > 0 && row >= 0 && row < ctx->batch_rows && ctx->batch_slot[row] >= 0, "row is not active");
  TB_CHECK(ids_out && n_out && head_out, "null argument");
  cudaStream_t s = ctx->dstream;
  int32_t st = 0;
  TB_CUDA(cudaMemcpyAsync(&st, ctx->sst.row_step + row, 4, cudaMemcpyDeviceToHost, s));
  TB_CUDA(cudaMemcpyAsync(head_out, ctx->sst.heads + row, 4, cudaMemcpyDeviceToHost, s));
  TB_CUDA(cudaStreamSynchronize(s));
  const int n = st < ctx->batch_budget[row] ? st : ctx->batch_budget[row];
  TB_CHECK(capacity >= n, "ids buffer too small");
  TB_CUDA(cudaMemcpyAsync(ids_out, ctx->sst.out_ids + static_cast<size_t>(row) * ctx->c.max_ctx, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, s));
  TB_CUDA(cudaStreamSynchronize(s));
  *n_out = n;
  return 0;
}

int trace_b200_batch_retire(trace_b200_ctx* ctx, int row) {
  TB_CHECK(ctx != nullptr && ctx->batch_rows > 0 && row >= 0 && row < ctx->batch_rows && ctx->batch_slot[row] >= 0, "row is not active");
  if (set_row(ctx, row, ctx->batch_dummy_slot, 0, 0, 1, 0, ctx->dstream)) return 1;
  TB_CUDA(cudaMemsetAsync(ctx->d_x + static_cast<size_t>(row) * ctx->c.hidden, 0, static_cast<size_t>(ctx->c.hidden) * 2, ctx->dstream));
  TB_CUDA(cudaStreamSynchronize(ctx->dstream));
  ctx->batch_slot[row] = -1;
  ctx->batch_budget[row] = 0;
  return 0;
}

int trace_b200_batch_close(trace_b200_ctx* ctx) {
  TB_CHECK(ctx != nullptr && ctx->batch_rows > 0, "no batch open");
  for (int r = 0; r < ctx->batch_rows; ++r) TB_CHECK(ctx->batch_slot[r] < 0, "retire every row before closing the batch");
  if (trace_b200_free_slot(ctx, ctx->batch_dummy_slot)) return 1;
  ctx->batch_dummy_slot = -1;
  ctx->batch_rows = 0;
  return 0;
}

int trace_b200_set_sampling(trace_b200_ctx* ctx, int do_sample, float temperature, int top_k, uint64_t seed) {
  TB_CHECK(ctx != nullptr, "null context");
  TB_CHECK(!do_sample || temperature > 0.f, "temperature must be positive");
  TB_CHECK(top_k >= 0, "top_k must be >= 0 (0 = off)");
  ctx->do_sample = do_sample ? 1 : 0;
  ctx->temperature = do_sample ? temperature : 1.0f;
  ctx->top_k = top_k;
  ctx->seed = seed;
  return 0;
}

int trace_b200_set_forced_ids(trace_b200_ctx* ctx, const int64_t* ids_host, int B, int n) {
  TB_CHECK(ctx != nullptr, "null context");
  if (ids_host == nullptr || B == 0) { ctx->forced_host.clear(); return 0; }
  TB_CHECK(B >= 1 && B <= kMaxB && n >= 1 && n <= ctx->c.max_ctx, "bad forced-id shape");
  ctx->forced_host.assign(ids_host, ids_host + static_cast<size_t>(B) * n);
  ctx->forced_B = B;
  ctx->forced_n = n;
  return 0;
}

int trace_b200_decode_plan(trace_b200_ctx* ctx, int B, int context_len, int32_t* out8) {
  TB_CHECK(ctx != nullptr && out8 != nullptr, "null argument");
  return decode_stream_plan(stream_args(ctx, B, ctx->c.layers, nullptr), context_len, out8);
}

int64_t trace_b200_launch_count(trace_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
int trace_b200_last_decode_timing(trace_b200_ctx* ctx, float* ms_out, int* steps_out) {
  TB_CHECK(ctx != nullptr && ms_out != nullptr && steps_out != nullptr, "null argument");
  *ms_out = ctx->last_decode_ms;
  *steps_out = ctx->last_decode_steps;
  return 0;
}
int64_t trace_b200_weight_bytes(trace_b200_ctx* ctx) { return ctx ? ctx->weight_bytes : 0; }
int trace_b200_debug_timeline(trace_b200_ctx* ctx, uint64_t* out_host, int capacity) {
  TB_CHECK(ctx != nullptr && out_host != nullptr, "null argument");
  TB_CHECK(ctx->d_dbg != nullptr, "timeline not enabled (set TRACE_B200_MEGA_TIMING=1 before creating the context)");
  const int n = 2 * ctx->c.layers * 32 + (capacity >= 2 * ctx->c.layers * 32 + ctx->num_sms * 32 ? ctx->num_sms * 32 : 0);
  TB_CHECK(capacity >= 2 * ctx->c.layers * 32, "timeline buffer too small");
  TB_CUDA(cudaMemcpy(out_host, ctx->d_dbg, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost));
  return 0;
}

// ---- single-kernel entry points -----------------------------------------------------------------
static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int trace_b200_preprocess_frames(const uint8_t* frames, int F, int H, int W, const int32_t* geom, const int32_t* hbounds,
                                 const int32_t* hcoef, const int32_t* vbounds, const int32_t* vcoef, const float* mean,
                                 const float* stdv, uint8_t* tmp, void* out, void* stream) {
  TB_CHECK(frames && geom && hbounds && hcoef && vbounds && vcoef && mean && stdv && tmp && out, "null argument");
  // geom: S, pad_x, pad_y, OW, OH, crop, crop_x, crop_y, row0, rows, hk, vk, pad_r, pad_g, pad_b
  return preprocess_frames(frames, F, H, W, geom[0], geom[1], geom[2], geom[3], geom[4], geom[5], geom[6], geom[7], geom[8], geom[9], hbounds,
                           hcoef, geom[10], vbounds, vcoef, geom[11], mean, stdv, geom + 12, tmp, static_cast<__half*>(out),
                           static_cast<cudaStream_t>(stream));
}

int trace_b200_op_gemm(const void* A, int lda, const void* W, int ldw, void* out, int ldo, int M, int N, int K, const void* bias,
                       const void* residual, int ldr, int act, int swiglu, void* stream) {
  return gemm_f16(static_cast<const __half*>(A), lda, static_cast<const __half*>(W), ldw, static_cast<__half*>(out), ldo, M, N, K,
                  static_cast<const __half*>(bias), static_cast<const __half*>(residual), ldr, act, swiglu, sm_count(),
                  static_cast<cudaStream_t>(stream));
}
int trace_b200_op_layernorm(const void* x, const void* w, const void* b, float eps, void* out, int rows, int D, void* stream) {
  return layernorm_f16(static_cast<const __half*>(x), D, static_cast<const __half*>(w), static_cast<const __half*>(b), eps,
                       static_cast<__half*>(out), D, rows, D, static_cast<cudaStream_t>(stream));
}
int trace_b200_op_rmsnorm(const void* x, const void* w, float eps, void* out, int rows, int D, void* stream) {
  return rmsnorm_f16(static_cast<const __half*>(x), D, static_cast<const __half*>(w), eps, static_cast<__half*>(out), D, rows, D,
                     static_cast<cudaStream_t>(stream));
}
int trace_b200_op_attention_tc(const void* qkv, int width, void* o, int seqlen, int heads, int kv_heads, int head_dim, int causal, int batch,
                               float scale, void* stream) {
  const int qd = heads * head_dim, kvd = kv_heads * head_dim;
  return attention_tc_fwd(static_cast<const __half*>(qkv), width, static_cast<__half*>(o), seqlen, qd, static_cast<int64_t>(seqlen) * qd, 0, qd,
                          qd + kvd, heads, heads / kv_heads, batch, head_dim, causal, scale, static_cast<cudaStream_t>(stream));
}
}  // extern "C"
