"""CPU oracle: a plain-torch fp32 restatement of the TRACE inference hot path.

TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product (``trace_b200/``) never does and fails
loudly when its CUDA library is missing.

Parity status: the reference has no tests or golden vectors of its own, and its
arithmetic lives in un-vendored ``transformers==4.40.1`` / ``timm==0.6.13``.  This restatement is
therefore pinned against *outputs of the reference itself run in the build container*
(``tests/golden/*.npz`` made by ``tests/golden/make_golden.py`` which imports /root/reference through
``oracle/ref_shim.py`` on top of the container's transformers 5.5.0) and against the known-answer table
of SURVEY.md §4.  ``tests/test_oracle_golden.py`` enforces both.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


# ------------------------------------------------------------------------------------------------
# time / score tokenisation  (host integer work)
# ------------------------------------------------------------------------------------------------
_SYM = {"<sync>": 0, "<sep>": 1, **{str(i): i + 2 for i in range(10)}, ".": 12}  # time_encoder.py:82-88
_INV = {v: k for k, v in _SYM.items()}


def encode_numbers(values: Sequence[float], fmt: str) -> List[int]:
    """TimeTower.encode (time_encoder.py:52-68, fmt '0>6.1f') / ScoreTower.encode
    (score_encoder.py:52-70, fmt '0>3.1f'): formatted numbers joined by <sep>, closed by <sync>."""
    ids: List[int] = []
    for i, v in enumerate(values):
        if i:
            ids.append(_SYM["<sep>"])
        ids.extend(_SYM[c] for c in format(v, fmt))
    ids.append(_SYM["<sync>"])
    return ids


def encode_time(values: Sequence[float]) -> List[int]:
    return encode_numbers(values, "0>6.1f")


def encode_score(values: Sequence[float]) -> List[int]:
    return encode_numbers(values, "0>3.1f")


def decode_symbol(i: int) -> str:
    """TimeTokenizer/ScoreTokenizer.decode of one id (time_encoder.py:97-98)."""
    return _INV[int(i)]


def frame_time_ids(video_timestamps: Sequence[Sequence[float]]) -> torch.Tensor:
    """Per-frame 6 time-token ids, trailing <sync> dropped (trace_arch.py:237-248, encode_time :271-288)."""
    rows = []
    for ts in video_timestamps:
        ids = encode_time(ts)[:-1]
        rows.append(ids)
    assert all(len(r) == len(rows[0]) for r in rows), "trace_arch.py:285 shape assert"
    return torch.tensor(rows, dtype=torch.long)


# ------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------
class TraceOracle:
    """fp32 CPU restatement.  `sd` maps HF parameter names to tensors (any float dtype)."""

    def __init__(self, dims, sd: Dict[str, torch.Tensor], dtype=torch.float32):
        self.d = dims
        self.dtype = dtype
        self.sd = {k: v.to(dtype) for k, v in sd.items()}
        self._vt = "model.vision_tower.vision_tower.vision_model."

    def p(self, name: str) -> torch.Tensor:
        return self.sd[name]

    # ---- CLIP vision tower: clip_encoder.py:41-53 -> HF CLIPVisionModel; feature_select :31-39 ----
    def vit_features(self, pixels: torch.Tensor) -> torch.Tensor:
        """[F,3,H,W] -> hidden_states[-2] without CLS: [F, patches, D]."""
        d, vt = self.d, self._vt
        x = pixels.to(self.dtype)
        Fn = x.shape[0]
        w = self.p(vt + "embeddings.patch_embedding.weight")
        pe = F.conv2d(x, w, stride=d.patch_size).flatten(2).transpose(1, 2)          # [F,P,D]
        cls = self.p(vt + "embeddings.class_embedding").expand(Fn, 1, -1)
        h = torch.cat([cls, pe], 1) + self.p(vt + "embeddings.position_embedding.weight")
        h = F.layer_norm(h, (d.vit_width,), self.p(vt + "pre_layrnorm.weight"),
                         self.p(vt + "pre_layrnorm.bias"), d.vit_ln_eps)
        nh, hd = d.vit_heads, d.vit_head_dim
        # hidden_states[-2] == output of block (L-1); block L and post_layernorm are never needed
        for i in range(d.vit_layers - 1):
            pre = f"{vt}encoder.layers.{i}."
            r = h
            y = F.layer_norm(h, (d.vit_width,), self.p(pre + "layer_norm1.weight"),
                             self.p(pre + "layer_norm1.bias"), d.vit_ln_eps)
            q = F.linear(y, self.p(pre + "self_attn.q_proj.weight"), self.p(pre + "self_attn.q_proj.bias"))
            k = F.linear(y, self.p(pre + "self_attn.k_proj.weight"), self.p(pre + "self_attn.k_proj.bias"))
            v = F.linear(y, self.p(pre + "self_attn.v_proj.weight"), self.p(pre + "self_attn.v_proj.bias"))
            q, k, v = (t.view(Fn, -1, nh, hd).transpose(1, 2) for t in (q, k, v))
            a = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1) @ v
            a = a.transpose(1, 2).reshape(Fn, -1, d.vit_width)
            h = r + F.linear(a, self.p(pre + "self_attn.out_proj.weight"), self.p(pre + "self_attn.out_proj.bias"))
            r = h
            y = F.layer_norm(h, (d.vit_width,), self.p(pre + "layer_norm2.weight"),
                             self.p(pre + "layer_norm2.bias"), d.vit_ln_eps)
            y = F.linear(y, self.p(pre + "mlp.fc1.weight"), self.p(pre + "mlp.fc1.bias"))
            y = y * torch.sigmoid(1.702 * y)                                         # quick_gelu
            h = r + F.linear(y, self.p(pre + "mlp.fc2.weight"), self.p(pre + "mlp.fc2.bias"))
        return h[:, 1:, :]

    # ---- SpatialSlotPool: multimodal_projector/builder.py:427-467 (+ rotary :289-359) ----
    def slot_pool(self, feats: torch.Tensor) -> torch.Tensor:
        """[F, n, D] -> [F, 8, hidden]."""
        D = feats.shape[-1]
        n = feats.shape[1]
        y = F.layer_norm(feats, (D,), self.p("model.mm_projector.ln_vision.weight"),
                         self.p("model.mm_projector.ln_vision.bias"), 1e-6)           # timm LayerNorm eps
        inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
        fr = torch.outer(torch.arange(n, dtype=torch.int64).float(), inv)
        emb = torch.cat([fr, fr], -1)
        cos, sin = emb.cos().to(y.dtype), emb.sin().to(y.dtype)
        rot = torch.cat([-y[..., D // 2:], y[..., : D // 2]], -1)                     # rotate_half :326-330
        y = y * cos + rot * sin
        logits = y @ self.p("model.mm_projector.slots")                               # [F,n,8]
        pm = torch.softmax(logits, dim=1)                                             # over the patch axis :458
        res = (y.transpose(1, 2) @ pm).transpose(1, 2)                                # [F,8,D]
        return F.linear(res, self.p("model.mm_projector.readout.weight"))

    # ---- encode_images_or_videos: trace_arch.py:218-266 ----
    def encode_clip(self, pixels: torch.Tensor, video_timestamps) -> torch.Tensor:
        """[T,3,H,W] + T×[float] -> [14T, hidden]; per frame 8 slot rows then 6 time rows."""
        slots = self.slot_pool(self.vit_features(pixels))                             # [T,8,H]
        tids = frame_time_ids(video_timestamps)                                       # [T,6]
        trow = self.p("model.time_tower.embed_tokens.weight")[tids]                   # [T,6,H]
        return torch.cat([slots, trow], 1).reshape(-1, self.d.hidden)

    # ---- prefill splice: trace_arch.py:377-426 ----
    def splice(self, ids: Sequence[int], vis: torch.Tensor) -> torch.Tensor:
        ids = list(int(i) for i in ids)
        vpos = [i for i, t in enumerate(ids) if t in (-201, -200)]
        assert len(vpos) == 1, "only have one video inputs! (trace_arch.py:408)"
        new_ids = ids[: vpos[0]] + [-201] * vis.shape[0] + ids[vpos[0] + 1:]
        t = torch.tensor(new_ids, dtype=torch.long)
        e = self.p("model.embed_tokens.weight")[t.clamp(min=0)].clone()
        e[t == -201] = vis
        e[t == -205] = self.p("model.sync_tower.embed_tokens.weight")[0]
        assert not ((t == -203) | (t == -204)).any(), "TIME/SCORE placeholders are training-only"
        return e

    # ---- decode-step embedding dispatch: trace_arch.py:345-375 ----
    def embed_token(self, tok: int) -> torch.Tensor:
        d = self.d
        if tok == d.vocab:
            return self.p("model.sync_tower.embed_tokens.weight")[0]
        if d.time_base <= tok < d.score_base:
            return self.p("model.time_tower.embed_tokens.weight")[tok - d.time_base]
        if tok >= d.score_base:
            return self.p("model.score_tower.embed_tokens.weight")[tok - d.score_base]
        return self.p("model.embed_tokens.weight")[tok % d.vocab]

    # ---- Mistral decoder (HF MistralModel via trace_mistral.py:178-188) ----
    def _rms(self, x, w):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return w * (x.float() * torch.rsqrt(v + self.d.rms_eps)).to(x.dtype)

    def _rope(self, x, pos):
        hd = self.d.head_dim
        inv = 1.0 / (self.d.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
        fr = torch.outer(pos.float(), inv)
        emb = torch.cat([fr, fr], -1)
        cos, sin = emb.cos().to(x.dtype)[None, :, :], emb.sin().to(x.dtype)[None, :, :]
        rot = torch.cat([-x[..., hd // 2:], x[..., : hd // 2]], -1)
        return x * cos + rot * sin

    def llm_forward(self, e: torch.Tensor, cache: Optional[list]) -> Tuple[torch.Tensor, list]:
        """e [q, hidden] appended after `cache` (list of (K,V) [kvh, t, hd]); returns final-normed
        hidden [q, hidden] and the new cache.  Mask is all-ones => positions = arange."""
        d = self.d
        qn = e.shape[0]
        t0 = 0 if cache is None else cache[0][0].shape[1]
        pos = torch.arange(t0, t0 + qn)
        new_cache = []
        h = e
        for i in range(d.layers):
            pre = f"model.layers.{i}."
            y = self._rms(h, self.p(pre + "input_layernorm.weight"))
            q = F.linear(y, self.p(pre + "self_attn.q_proj.weight")).view(qn, d.heads, d.head_dim).transpose(0, 1)
            k = F.linear(y, self.p(pre + "self_attn.k_proj.weight")).view(qn, d.kv_heads, d.head_dim).transpose(0, 1)
            v = F.linear(y, self.p(pre + "self_attn.v_proj.weight")).view(qn, d.kv_heads, d.head_dim).transpose(0, 1)
            q, k = self._rope(q, pos), self._rope(k, pos)
            if cache is not None:
                k = torch.cat([cache[i][0], k], 1)
                v = torch.cat([cache[i][1], v], 1)
            new_cache.append((k, v))
            g = d.heads // d.kv_heads
            kk, vv = k.repeat_interleave(g, 0), v.repeat_interleave(g, 0)
            s = (q @ kk.transpose(-1, -2)) * d.head_dim ** -0.5                      # [heads,q,t]
            tt = k.shape[1]
            causal = torch.arange(tt)[None, :] > (t0 + torch.arange(qn))[:, None]
            s = s.masked_fill(causal[None], NEG_INF)
            a = torch.softmax(s.float(), -1).to(q.dtype) @ vv
            a = a.transpose(0, 1).reshape(qn, d.heads * d.head_dim)
            h = h + F.linear(a, self.p(pre + "self_attn.o_proj.weight"))
            y = self._rms(h, self.p(pre + "post_attention_layernorm.weight"))
            gate = F.linear(y, self.p(pre + "mlp.gate_proj.weight"))
            up = F.linear(y, self.p(pre + "mlp.up_proj.weight"))
            h = h + F.linear(F.silu(gate) * up, self.p(pre + "mlp.down_proj.weight"))
        return self._rms(h, self.p("model.norm.weight")), new_cache

    # ---- heads + mask: trace_mistral.py:190-200, 244-252 ----
    def head_logits(self, h_last: torch.Tensor, head: Optional[int]) -> torch.Tensor:
        """[hidden] -> fp32 logits [V+27] with everything outside `head`'s range = -inf
        (or [V+1] text-only logits when head is None, trace_mistral.py:244)."""
        d = self.d
        text = torch.cat([F.linear(h_last, self.p("lm_head.weight")),
                          F.linear(h_last, self.p("sync_head.weight"))], -1).float()
        if head is None:
            return text
        z = torch.cat([text, F.linear(h_last, self.p("time_head.weight")).float(),
                       F.linear(h_last, self.p("score_head.weight")).float()], -1)
        lo, hi = [(0, d.time_base), (d.time_base, d.score_base), (d.score_base, d.total_vocab)][head]
        z[:lo] = NEG_INF
        z[hi:] = NEG_INF
        return z

    # ---- generate: trace_mistral.py:268-347 + HF greedy loop ----
    @torch.no_grad()
    def generate(self, ids: Sequence[int], pixels: torch.Tensor, video_timestamps, head: Optional[int] = 1,
                 max_new_tokens: int = 16, eos: int = 2, pad: int = 2, return_trace: bool = False):
        """One clip, greedy.  Returns new ids (list) [and per-step dict(logits of the active range,
        head) when return_trace].  `head` None reproduces trace/__init__.py:infer (text head only)."""
        d = self.d
        swap = {d.vocab: 1, d.time_base: 2, d.score_base: 0}                          # trace_mistral.py:86-88
        if pixels is None:                                                            # trace_mistral.py:306-307: plain text generate
            vis = None
            e = torch.stack([self.embed_token(int(i)) for i in ids], 0)
        else:
            vis = self.encode_clip(pixels, video_timestamps)
            e = self.splice(ids, vis)
        h, cache = self.llm_forward(e, None)
        out: List[int] = []
        trace = []
        finished = False
        for step in range(max_new_tokens):
            z = self.head_logits(h[-1], head)
            tok = int(torch.argmax(z))                                                # first max on ties
            if finished:
                tok = pad
            if return_trace:
                if head is None:
                    lo, hi = 0, d.time_base
                else:
                    lo, hi = [(0, d.time_base), (d.time_base, d.score_base), (d.score_base, d.total_vocab)][head]
                trace.append({"head": head, "logits": z[lo:hi].clone(), "lo": lo})
            out.append(tok)
            finished = finished or tok == eos
            if finished:
                break                                                                 # B=1: all rows done
            if head is not None and tok in swap:                                      # prepare_inputs_for_generation :336-344
                head = swap[tok]
            h, cache = self.llm_forward(self.embed_token(tok)[None, :], cache)
        if return_trace:
            return out, trace, {"vis": vis, "prefill_embeds": e}
        return out


# ------------------------------------------------------------------------------------------------
# output parser: trace/eval/evaluate.py:360-411
# ------------------------------------------------------------------------------------------------
def parse_events(output_ids: Sequence[int], decode_text, vocab: int = 32000):
    """ids -> {'timestamps','scores','captions'} exactly as the eval driver does (including its
    quirks: SURVEY §9-21).  `decode_text(list[int]) -> str`."""
    tb, sb = vocab + 1, vocab + 14
    outputs = {"timestamps": [], "scores": [], "captions": []}
    cur_ts, cur_t, cur_ss, cur_s, cur_cap = [], [], [], [], []
    for idx in output_ids:
        idx = int(idx)
        if idx <= vocab:
            if idx == vocab:
                outputs["captions"].append(decode_text(cur_cap))
                cur_cap = []
            else:
                cur_cap.append(idx)
        elif idx <= vocab + 13:
            if idx == tb:
                if cur_t:
                    cur_ts.append(float("".join(cur_t)))
                outputs["timestamps"].append(cur_ts)
                cur_ts, cur_t = [], []
            elif idx == tb + 1:
                if cur_t:
                    cur_ts.append(float("".join(cur_t)))
                cur_t = []
            else:
                cur_t.append(decode_symbol(idx - tb))
        else:
            if idx == sb:
                if cur_s:
                    cur_ss.append(float("".join(cur_s)))
                outputs["scores"].append(cur_ss)
                cur_ss, cur_s = [], []
            elif idx == sb + 1:
                if cur_s:
                    cur_ss.append(float("".join(cur_s)))
                cur_s = []
            else:
                cur_s.append(decode_symbol(idx - sb))
    if cur_cap:
        outputs["captions"].append(decode_text(cur_cap))
    return outputs
