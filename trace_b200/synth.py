"""Seeded synthetic TRACE checkpoints and a stand-in text tokenizer.

The released checkpoints (``Yongxin-Guo/trace*``), the CLIP weights and the Llama/Mistral
sentencepiece model are unreachable offline, so parity and throughput are
established on synthetic weights with the *exact* state-dict layout of the reference
(SURVEY.md §4 "state-dict extras"; trace_arch.py:31-40, trace_mistral.py:85-96,
multimodal_projector/builder.py:411-425 and the HF CLIP / Mistral module trees).

Every tensor is a pure function of ``(seed, parameter name, shape)`` so the golden generator
(which feeds the real reference), the CPU oracle and the CUDA engine see bit-identical
fp16-representable values on any machine.
"""
from __future__ import annotations

import dataclasses
import zlib
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch

TIME_VOCAB = 13   # time_encoder.py:82-88
SCORE_VOCAB = 13  # score_encoder.py:84-98
NUM_SLOTS = 8     # multimodal_projector/builder.py:413
TIME_TOKENS_PER_FRAME = 6  # format(t,'0>6.1f'), trailing <sync> dropped (trace_arch.py:243)


@dataclasses.dataclass
class TraceDims:
    """Architecture hyper-parameters (what `config.json` of a TRACE checkpoint carries)."""
    # --- CLIP vision tower (HF CLIPVisionConfig) ---
    vit_width: int = 1024
    vit_heads: int = 16
    vit_layers: int = 24          # checkpoint layers; only vit_layers-1 are needed ([-2], clip_encoder.py:32)
    vit_mlp: int = 4096
    image_size: int = 336
    patch_size: int = 14
    vit_ln_eps: float = 1e-5
    # --- Mistral decoder (HF MistralConfig) ---
    hidden: int = 4096            # trace_arch.py:38-40 hard-codes 4096 for the towers
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    head_dim: int = 128
    ffn: int = 14336
    vocab: int = 32000
    rope_theta: float = 1e6       # Mistral-7B-Instruct-v0.2; parametrised
    rms_eps: float = 1e-5
    # --- TRACE ---
    num_frames: int = 128

    @property
    def patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def vit_tokens(self) -> int:
        return self.patches + 1

    @property
    def vit_head_dim(self) -> int:
        return self.vit_width // self.vit_heads

    @property
    def tokens_per_frame(self) -> int:
        return NUM_SLOTS + TIME_TOKENS_PER_FRAME

    # logits layout: text [0,V+1) | time [V+1, V+14) | score [V+14, V+27)   (trace_mistral.py:248)
    @property
    def text_sync_id(self) -> int:
        return self.vocab

    @property
    def time_base(self) -> int:
        return self.vocab + 1

    @property
    def score_base(self) -> int:
        return self.vocab + 1 + TIME_VOCAB

    @property
    def total_vocab(self) -> int:
        return self.vocab + 1 + TIME_VOCAB + SCORE_VOCAB


def full_dims(num_frames: int = 128) -> TraceDims:
    """CLIP-ViT-L/14-336 + Mistral-7B (BASELINE.json configs 2-5)."""
    return TraceDims(num_frames=num_frames)


def tiny_dims(num_frames: int = 4) -> TraceDims:
    """Parity-test model the CPU oracle and the real reference finish in seconds.

    hidden stays 4096 (mandatory, trace_arch.py:38-40); head dims stay 64 (ViT) / 128 (LLM) so the
    same CUDA kernels are exercised; 56-px images -> 16 patches (+CLS = 17 tokens, a ragged tile).
    """
    return TraceDims(vit_width=128, vit_heads=2, vit_layers=3, vit_mlp=512, image_size=56,
                     hidden=4096, layers=2, heads=32, kv_heads=8, head_dim=128, ffn=256,
                     num_frames=num_frames)


def mid_dims(num_frames: int = 8) -> TraceDims:
    """A wider check (ragged 577-token ViT tile with the real image size, 3 heads, 3 LLM layers)."""
    return TraceDims(vit_width=192, vit_heads=3, vit_layers=3, vit_mlp=768, image_size=336,
                     hidden=4096, layers=3, heads=32, kv_heads=8, head_dim=128, ffn=512,
                     num_frames=num_frames)


def wide_dims(num_frames: int = 8) -> TraceDims:
    """Full widths of the BASELINE configs (CLIP-ViT-L/14-336 and Mistral-7B layer shapes, full vocabulary) at shallow
    depth (3 ViT blocks of which 2 are used, 2 decoder layers): the real reference finishes it on CPU in minutes, and
    every per-layer code path of the full model is taken."""
    return TraceDims(vit_layers=3, layers=2, num_frames=num_frames)


# ------------------------------------------------------------------------------------------------
# parameter inventory
# ------------------------------------------------------------------------------------------------
_VT = "model.vision_tower.vision_tower.vision_model."


def param_specs(d: TraceDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(hf_name, shape, kind)] for the full TRACE state dict.  kind selects the init law."""
    D, H = d.vit_width, d.hidden
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    s += [(_VT + "embeddings.class_embedding", (D,), "emb"),
          (_VT + "embeddings.patch_embedding.weight", (D, 3, d.patch_size, d.patch_size), "lin"),
          (_VT + "embeddings.position_embedding.weight", (d.vit_tokens, D), "emb"),
          (_VT + "pre_layrnorm.weight", (D,), "norm_w"), (_VT + "pre_layrnorm.bias", (D,), "norm_b")]
    for i in range(d.vit_layers):
        p = f"{_VT}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s += [(p + f"self_attn.{n}.weight", (D, D), "lin"), (p + f"self_attn.{n}.bias", (D,), "bias")]
        s += [(p + "layer_norm1.weight", (D,), "norm_w"), (p + "layer_norm1.bias", (D,), "norm_b"),
              (p + "mlp.fc1.weight", (d.vit_mlp, D), "lin"), (p + "mlp.fc1.bias", (d.vit_mlp,), "bias"),
              (p + "mlp.fc2.weight", (D, d.vit_mlp), "lin"), (p + "mlp.fc2.bias", (D,), "bias"),
              (p + "layer_norm2.weight", (D,), "norm_w"), (p + "layer_norm2.bias", (D,), "norm_b")]
    s += [(_VT + "post_layernorm.weight", (D,), "norm_w"), (_VT + "post_layernorm.bias", (D,), "norm_b")]
    # projector (multimodal_projector/builder.py:417-420)
    s += [("model.mm_projector.slots", (D, NUM_SLOTS), "slots"),
          ("model.mm_projector.ln_vision.weight", (D,), "norm_w"),
          ("model.mm_projector.ln_vision.bias", (D,), "norm_b"),
          ("model.mm_projector.readout.weight", (H, D), "lin")]
    # towers (trace_arch.py:38-40)
    s += [("model.time_tower.embed_tokens.weight", (TIME_VOCAB, H), "emb"),
          ("model.score_tower.embed_tokens.weight", (SCORE_VOCAB, H), "emb"),
          ("model.sync_tower.embed_tokens.weight", (1, H), "emb")]
    # Mistral
    s += [("model.embed_tokens.weight", (d.vocab, H), "emb")]
    q, kv = d.heads * d.head_dim, d.kv_heads * d.head_dim
    for i in range(d.layers):
        p = f"model.layers.{i}."
        s += [(p + "self_attn.q_proj.weight", (q, H), "lin"), (p + "self_attn.k_proj.weight", (kv, H), "lin"),
              (p + "self_attn.v_proj.weight", (kv, H), "lin"), (p + "self_attn.o_proj.weight", (H, q), "lin"),
              (p + "mlp.gate_proj.weight", (d.ffn, H), "lin"), (p + "mlp.up_proj.weight", (d.ffn, H), "lin"),
              (p + "mlp.down_proj.weight", (H, d.ffn), "lin"),
              (p + "input_layernorm.weight", (H,), "norm_w"),
              (p + "post_attention_layernorm.weight", (H,), "norm_w")]
    s += [("model.norm.weight", (H,), "norm_w"),
          ("lm_head.weight", (d.vocab, H), "head"), ("sync_head.weight", (1, H), "head"),
          ("time_head.weight", (TIME_VOCAB, H), "head"), ("score_head.weight", (SCORE_VOCAB, H), "head")]
    return s


def _gen(name: str, seed: int, device: str) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFF)
    return g


def synth_tensor(name: str, shape: Tuple[int, ...], kind: str, seed: int, device: str = "cpu") -> torch.Tensor:
    """One fp16 parameter.  Laws: linear/embedding N(0, 0.02) (HF init), heads N(0, 0.05) (wider
    logit margins than HF init so that greedy ids are not decided by fp16 noise), norm weights
    1+N(0,0.1), biases/norm biases N(0, 0.02), slots N(0, 0.05)."""
    g = _gen(name, seed, device)
    x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    if kind in ("lin", "emb", "bias", "norm_b"):
        x = x * 0.02
    elif kind == "head":
        x = x * 0.05
    elif kind == "slots":
        x = x * 0.05
    elif kind == "norm_w":
        x = 1.0 + 0.1 * x
    else:
        raise ValueError(kind)
    return x.to(torch.float16)


def iter_state_dict(d: TraceDims, seed: int = 0, device: str = "cpu") -> Iterator[Tuple[str, torch.Tensor]]:
    for name, shape, kind in param_specs(d):
        yield name, synth_tensor(name, shape, kind, seed, device)


def make_state_dict(d: TraceDims, seed: int = 0, device: str = "cpu") -> Dict[str, torch.Tensor]:
    return dict(iter_state_dict(d, seed, device))


# ------------------------------------------------------------------------------------------------
# structured variant: a checkpoint whose greedy decode walks TRACE's event grammar
# ------------------------------------------------------------------------------------------------
# Event grammar: time head "d d d d . d <sep> d d d d . d <sync>" -> score head
# "d . d <sync>" -> text head "caption tokens ... <sync>" -> back to the time head.  Random weights
# never close an event, so `events/s` would be undefined.  The structured variant keeps every
# layer's math and bytes (same shapes, dense non-zero weights) but damps the residual branches and
# ties each head row to the embedding of its predecessor token, which makes the greedy decode a
# first-order chain over tokens.  A first-order chain can visit each symbol once per head, so the
# synthetic event is:  time "0 1 2 . 3 <sep> 4 5 6 7 8 9 <sync>" (13) -> score "0 . 1 <sync>" (4)
# -> text: 22 distinct caption ids then <sync> (23)  = 40 tokens per event, the nominal DVC event.
EVENT_TIME_CHAIN = [2, 3, 4, 12, 5, 1, 6, 7, 8, 9, 10, 11, 0]   # time-vocab ids, ends with <sync>=0
EVENT_SCORE_CHAIN = [2, 12, 3, 0]                              # "0.1<sync>"
EVENT_CAPTION_LEN = 22
TOKENS_PER_EVENT = len(EVENT_TIME_CHAIN) + len(EVENT_SCORE_CHAIN) + EVENT_CAPTION_LEN + 1


def event_caption_ids(d: TraceDims) -> List[int]:
    return [1000 + 37 * i for i in range(EVENT_CAPTION_LEN)]


def make_structured_overrides(d: TraceDims, seed: int, get: Callable[[str], torch.Tensor],
                              eos_after_events: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Returns replacement tensors (same names/shapes/dtype) for the structured checkpoint.

    `get(name)` returns the plain random tensor.  Heads are rebuilt as  row[next] = unit(emb[prev])*G
    + small noise, with prev -> next the chain above; embeddings of chain tokens are amplified so they
    dominate the (damped) residual stream."""
    dev = get("model.norm.weight").device
    H = d.hidden
    out: Dict[str, torch.Tensor] = {}
    amp = 40.0   # chain-token embedding gain over N(0,0.02)

    def amplified(name: str, rows: List[int]) -> torch.Tensor:
        w = get(name).float().clone()
        w[rows] = w[rows] * amp
        out[name] = w.to(torch.float16)
        return out[name].float()

    cap = event_caption_ids(d)
    emb = amplified("model.embed_tokens.weight", cap)
    tt = amplified("model.time_tower.embed_tokens.weight", list(range(TIME_VOCAB)))
    st = amplified("model.score_tower.embed_tokens.weight", list(range(SCORE_VOCAB)))
    sy = amplified("model.sync_tower.embed_tokens.weight", [0])

    def unit(v: torch.Tensor) -> torch.Tensor:
        return v / v.norm() * (H ** 0.5)

    G = 0.25  # logit for the scripted successor ~= G * H / rms-normalised ... (>> random rows)

    def head(name: str, transitions: List[Tuple[torch.Tensor, int]]) -> None:
        w = get(name).float().clone() * 0.2
        for prev_vec, nxt in transitions:
            w[nxt] = w[nxt] + unit(prev_vec) * (G / (H ** 0.5))
        out[name] = w.to(torch.float16)

    # time head: entered after text <sync> (input embedding = sync_tower row, trace_arch.py:353-370)
    tr = [(sy[0], EVENT_TIME_CHAIN[0])]
    for a, b in zip(EVENT_TIME_CHAIN[:-1], EVENT_TIME_CHAIN[1:]):
        tr.append((tt[a], b))
    head("time_head.weight", tr)
    # score head: entered after time <sync> (input = time_tower row 0)
    tr = [(tt[0], EVENT_SCORE_CHAIN[0])]
    for a, b in zip(EVENT_SCORE_CHAIN[:-1], EVENT_SCORE_CHAIN[1:]):
        tr.append((st[a], b))
    head("score_head.weight", tr)
    # text head: entered after score <sync> (input = score_tower row 0); ends with sync_head firing
    tr = [(st[0], cap[0])]
    for a, b in zip(cap[:-1], cap[1:]):
        tr.append((emb[a], b))
    head("lm_head.weight", tr)
    w = get("sync_head.weight").float().clone() * 0.2
    w[0] = w[0] + unit(emb[cap[-1]]) * (G / (H ** 0.5))
    out["sync_head.weight"] = w.to(torch.float16)
    # damp the residual branches so the input-token embedding survives to the final norm
    for i in range(d.layers):
        for n in ("self_attn.o_proj.weight", "mlp.down_proj.weight"):
            name = f"model.layers.{i}.{n}"
            out[name] = (get(name).float() * 0.02).to(torch.float16)
    return out


# ------------------------------------------------------------------------------------------------
# stand-in text tokenizer (no sentencepiece model offline)
# ------------------------------------------------------------------------------------------------
class StandInTokenizer:
    """Deterministic word-hash tokenizer with the Llama special ids (bos=1, eos=2, unk=0, ids in
    [3, vocab)).  Only the properties the hot path observes are modelled: `tokenizer(text).input_ids`
    starts with BOS, `pad_token_id`, `eos_token_id`, `decode(ids, skip_special_tokens)`."""

    bos_token_id = 1
    eos_token_id = 2
    unk_token_id = 0
    pad_token_id = 0  # the Llama tokenizer of the checkpoints has pad = unk (evaluate.py:335)

    def __init__(self, vocab_size: int = 32000):
        self.vocab_size = vocab_size

    def _word_id(self, w: str) -> int:
        return 3 + zlib.crc32(w.encode()) % (self.vocab_size - 3)

    def __call__(self, text: str):
        import re
        import types
        ids = [self.bos_token_id] + [self._word_id(w) for w in re.findall(r"\w+|[^\w\s]", text)]
        return types.SimpleNamespace(input_ids=ids)

    def decode(self, ids, skip_special_tokens: bool = False) -> str:
        toks = []
        for i in ids:
            i = int(i)
            if skip_special_tokens and i in (0, 1, 2):
                continue
            toks.append(f"w{i}")
        return " ".join(toks)

    def batch_decode(self, seqs, skip_special_tokens: bool = False):
        return [self.decode(s, skip_special_tokens) for s in seqs]


def synthetic_clip(d: TraceDims, T: int, seed: int = 1234, duration_s: float = 120.0,
                   device: str = "cpu") -> Tuple[torch.Tensor, List[List[float]]]:
    """A CLIP-normalised synthetic clip [T,3,H,W] fp16 + per-frame timestamps."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    px = torch.randn((T, 3, d.image_size, d.image_size), generator=g, device=device).clamp_(-2.6, 2.6)
    ts = [[i * duration_s / max(T - 1, 1)] for i in range(T)]
    return px.to(torch.float16), ts
