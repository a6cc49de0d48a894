"""TraceMistralForCausalLM -- the HF-style wrapper surface of the reference
(trace/model/language_model/trace_mistral.py:62-347, trace/model/trace_arch.py:189-266) re-hosted on the
sm_100a engine.  Same method names, argument meaning and error behaviour:

* ``generate(inputs, images_or_videos=..., modal_list=..., video_timestamps=..., heads=..., attention_mask=...,
  do_sample, temperature, max_new_tokens, use_cache, pad_token_id, times, scores)`` returns ONLY the new
  token ids ``LongTensor[B, n]`` (HF ``inputs_embeds`` semantics), mutates ``heads`` in place
  (trace_mistral.py:338-344), raises ``NotImplementedError`` when ``inputs_embeds`` is passed (:282-283).
* ``get_model().time_tokenizer / .score_tokenizer`` decode ids for the drivers' parsers.
* ``config.image_aspect_ratio / .num_frames``, ``get_vision_tower().image_processor``, ``to(device)``.

All arithmetic runs in the CUDA library; this file only marshals arguments.
"""
from __future__ import annotations

import types
from typing import List, Optional

import torch

from .. import synth
from ..constants import MMODAL_TOKEN_INDEX, NUM_FRAMES
from ..engine import Engine, TraceB200Error
from .encoders import ScoreTokenizer, TimeTokenizer, frame_time_ids


class TraceMistralConfig:
    """The keys the reference reads back from config.json via getattr (SURVEY §5 'Config / flags')."""
    model_type = "trace_mistral"

    def __init__(self, dims: Optional[synth.TraceDims] = None, **kw):
        self.dims = dims or synth.full_dims()
        d = self.dims
        self.hidden_size, self.intermediate_size = d.hidden, d.ffn
        self.num_hidden_layers, self.num_attention_heads, self.num_key_value_heads = d.layers, d.heads, d.kv_heads
        self.vocab_size, self.rms_norm_eps, self.rope_theta = d.vocab, d.rms_eps, d.rope_theta
        self.mm_vision_select_layer, self.mm_vision_select_feature = -2, "patch"
        self.mm_projector_type, self.mm_hidden_size = "spatial_slot", d.vit_width
        self.time_vocab_size = self.score_vocab_size = 13
        self.image_aspect_ratio = "pad"
        self.num_frames = d.num_frames
        self.eos_token_id, self.bos_token_id, self.pad_token_id = 2, 1, None
        self.__dict__.update(kw)


class _ImageProcessor:
    """`processor.preprocess(images, return_tensors='pt')['pixel_values']` + `.image_mean`
    (mm_utils.py:458-459): CLIP preprocessing (bicubic shortest-edge resize, centre crop, /255, mean/std)."""
    image_mean = [0.48145466, 0.4578275, 0.40821073]
    image_std = [0.26862954, 0.26130258, 0.27577711]

    def __init__(self, size: int):
        self.size = size
        self.crop_size = {"height": size, "width": size}

    def preprocess(self, images, return_tensors="pt"):
        import numpy as np
        from PIL import Image
        out = []
        for im in images:
            if not isinstance(im, Image.Image):
                im = Image.fromarray(np.asarray(im))
            im = im.convert("RGB")
            w, h = im.size
            s = self.size / min(w, h)
            nw, nh = max(self.size, round(w * s)), max(self.size, round(h * s))
            im = im.resize((nw, nh), Image.BICUBIC)
            l, t = (nw - self.size) // 2, (nh - self.size) // 2
            im = im.crop((l, t, l + self.size, t + self.size))
            a = np.asarray(im, dtype=np.float32) / 255.0
            a = (a - np.array(self.image_mean, dtype=np.float32)) / np.array(self.image_std, dtype=np.float32)
            out.append(torch.from_numpy(a).permute(2, 0, 1))
        return {"pixel_values": torch.stack(out)}


class _VisionTower:
    def __init__(self, dims):
        self.image_processor = _ImageProcessor(dims.image_size)
        self.is_loaded = True
        self.hidden_size = dims.vit_width
        self.num_patches = dims.patches


class _Inner:
    """What `model.get_model()` returns in the reference (TraceMistralModel): tokenizers + towers."""

    def __init__(self, dims):
        self.time_tokenizer = TimeTokenizer()
        self.score_tokenizer = ScoreTokenizer()
        self.vision_tower = _VisionTower(dims)

    def get_vision_tower(self):
        return self.vision_tower


class TraceMistralForCausalLM:
    config_class = TraceMistralConfig

    _FORWARD_WEIGHTS = ("model.norm.weight", "lm_head.weight", "sync_head.weight", "time_head.weight", "score_head.weight")

    def __init__(self, config: TraceMistralConfig, device: int = 0, keep_head_weights: bool = False, **engine_kw):
        """`keep_head_weights`: also keep fp16 copies of the final norm and the four head matrices (262 MB for the full model)
        on the Python side, which `forward()` (full-sequence logits) needs; `generate()` never does."""
        self.keep_head_weights = keep_head_weights
        self._fw = {}
        self.config = config
        self.dims = config.dims
        self.vocab_size = config.vocab_size
        self.time_vocab_size = self.score_vocab_size = 13
        # trace_mistral.py:86-88
        self.swap_tokens = {self.vocab_size: 1, self.vocab_size + 1: 2, self.vocab_size + 14: 0}
        self.model = _Inner(self.dims)
        self.generation_config = types.SimpleNamespace(eos_token_id=config.eos_token_id, pad_token_id=config.pad_token_id)
        self.engine = Engine(self.dims, device=device, **engine_kw)
        self.device = torch.device(f"cuda:{device}")
        self.dtype = torch.float16
        self.dist_group = None   # set by trace_b200.parallel.attach_group for frame-sharded ViT encode

    # ---- reference surface ----------------------------------------------------------------------
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.model.get_vision_tower()

    def num_frames(self):
        return getattr(self.config, "num_frames", NUM_FRAMES)

    def to(self, *a, **k):
        return self   # weights already live on the engine's device

    def eval(self):
        return self

    def half(self):
        return self

    def load_state_dict(self, state_dict, strict: bool = True):
        items = list(state_dict.items() if hasattr(state_dict, "items") else state_dict)
        if self.keep_head_weights:
            for name, t in items:
                if name in self._FORWARD_WEIGHTS:
                    self._fw[name] = t.to(self.device, torch.float16).contiguous()
        self.engine.load_state_dict(items)

    # ---- forward (trace_mistral.py:114-264): full-sequence logits, prefill only ---------------------
    @torch.no_grad()
    def forward(self, input_ids: Optional[torch.Tensor] = None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                images_or_videos=None, return_dict=None, heads=None, video_timestamps=None, modal_list=None, **kwargs):
        """logits for EVERY position of the (spliced) prompt, fp32 [B, S, 32027] with everything outside `heads[b]`'s range
        set to -inf, or [B, S, 32001] when `heads` is None (trace_mistral.py:190-200, 244-252).  Inference only: `labels`
        (the training losses), `past_key_values` (HF drives incremental decoding through `generate`, which runs on the
        device-resident loop here) and `inputs_embeds` are rejected.  Without `images_or_videos` the ids are embedded as
        generate's text-only branch does (`embed_tokens(inputs)`, :306-307); the reference's own forward cannot take
        multi-token `input_ids` without images (trace_arch.py:359 applies Python `and` to tensors)."""
        if labels is not None or past_key_values is not None or inputs_embeds is not None:
            raise NotImplementedError("trace_b200.forward covers the inference prefill (input_ids [+ images_or_videos]) only")
        if not self._fw:
            raise TraceB200Error("forward() needs keep_head_weights=True at construction (generate() does not)")
        assert input_ids is not None and input_ids.dim() == 2
        B, L = input_ids.shape
        if heads is not None:
            assert len(heads) == B                                                               # trace_mistral.py:245
        ids = input_ids.to(self.device, torch.int64).contiguous()
        if images_or_videos is None:
            embeds = self.engine.build_prefill(ids, None, None)
        else:
            modal_list = modal_list or ["video"] * B
            nf = self.num_frames()
            clips = [x.unsqueeze(0).expand(nf, -1, -1, -1) if m == "image" else x for x, m in zip(images_or_videos, modal_list)]
            T = clips[0].shape[0]
            tid = torch.tensor([frame_time_ids(ts) for ts in video_timestamps], dtype=torch.int32, device=self.device)
            embeds = self.engine.build_prefill(ids, self._encode_slots(clips), tid)
        S = embeds.shape[1]
        from ..engine import op_gemm, op_rmsnorm
        d = self.dims
        V = d.vocab
        # the four heads stacked like the engine's own head matrix: row index = global token id (lm_head | sync | time | score)
        w = torch.cat([self._fw["lm_head.weight"], self._fw["sync_head.weight"], self._fw["time_head.weight"],
                       self._fw["score_head.weight"]], 0)
        pad = (-w.shape[0]) % 256                                                                # whole GEMM tiles
        if pad:
            w = torch.cat([w, torch.zeros(pad, w.shape[1], dtype=w.dtype, device=w.device)], 0)
        out = torch.empty((B, S, V + 27 if heads is not None else V + 1), dtype=torch.float32, device=self.device)
        for b in range(B):
            kv = self.engine.alloc_slot(S)
            try:
                hid = self.engine.prefill([kv], embeds[b:b + 1].contiguous(), return_hidden=True)
            finally:
                self.engine.free_slot(kv)
            hn = op_rmsnorm(hid[0].contiguous(), self._fw["model.norm.weight"], d.rms_eps)
            z = op_gemm(hn, w)[:, : V + 27].float()                                              # fp16 GEMM -> .float() (:191-200)
            if heads is None:
                out[b] = z[:, : V + 1]
            else:
                lo, hi = [(0, V + 1), (V + 1, V + 14), (V + 14, V + 27)][int(heads[b])]
                z[:, :lo] = float("-inf")
                z[:, hi:] = float("-inf")
                out[b] = z
        try:
            from transformers.modeling_outputs import CausalLMOutputWithPast
            return CausalLMOutputWithPast(loss=None, logits=out, past_key_values=None, hidden_states=None, attentions=None)
        except Exception:  # noqa: BLE001
            return types.SimpleNamespace(loss=None, logits=out, past_key_values=None, hidden_states=None, attentions=None)

    __call__ = forward

    # ---- encode_images_or_videos (trace_arch.py:218-266) ------------------------------------------
    def _encode_slots(self, clips: List[torch.Tensor], presliced_T: Optional[int] = None) -> torch.Tensor:
        """`presliced_T`: the clips hold only this rank's frame range of T-frame clips (trace_b200.parallel.generate_sharded)."""
        B, T = len(clips), clips[0].shape[0] if presliced_T is None else presliced_T
        frames = torch.cat([c.to(self.device, torch.float16, non_blocking=True) for c in clips], 0).contiguous()
        if self.dist_group is not None:
            from ..parallel import encode_frames_sharded
            slots = encode_frames_sharded(self.engine, frames, B, T, self.dist_group, presliced=presliced_T is not None)
        else:
            slots = self.engine.encode_frames(frames)
        return slots.view(B, T, synth.NUM_SLOTS, self.dims.hidden)

    def _generate_from_slots(self, ids: torch.Tensor, slots: torch.Tensor, tid: torch.Tensor, heads, max_new: int, eos: int,
                             pad: int) -> torch.Tensor:
        """Splice + prefill + greedy decode for rows whose visual tokens are already encoded.
        ids [B,L] int64, slots [B,T,8,H] fp16, tid [B,T,6] int32 (all on the engine's device); `heads` (list or None)
        is mutated in place.  Rows are decoded in groups of <= 8 (TRACE_B200_MAX_DECODE_BATCH)."""
        B, L = ids.shape
        T = 0 if tid is None else tid.shape[1]
        out_rows: List[torch.Tensor] = []
        hd_all = [(-1 if heads is None else int(h)) for h in (heads if heads is not None else [0] * B)]
        S = L if slots is None else L - 1 + self.dims.tokens_per_frame * T
        if S + max_new > self.engine.max_ctx:
            raise TraceB200Error(f"prompt ({S}) + max_new_tokens ({max_new}) exceeds the engine context ({self.engine.max_ctx})")
        if S > self.engine.max_prefill_tokens:
            raise TraceB200Error(f"prompt ({S}) exceeds the engine's prefill workspace ({self.engine.max_prefill_tokens} tokens)")
        # rows per prefill + decode group: the decode batch limit and what the prefill workspace holds
        group = max(1, min(8, self.engine.max_prefill_tokens // S))
        n_final = 0
        for b0 in range(0, B, group):
            b1 = min(B, b0 + group)
            if slots is None:
                embeds = self.engine.build_prefill(ids[b0:b1].contiguous(), None, None)
            else:
                embeds = self.engine.build_prefill(ids[b0:b1].contiguous(), slots[b0:b1].contiguous(), tid[b0:b1].contiguous())
            kv = [self.engine.alloc_slot(S + max_new) for _ in range(b1 - b0)]
            try:
                self.engine.prefill(kv, embeds)
                hd = hd_all[b0:b1]
                out = self.engine.decode(kv, hd, max_new, eos=eos, pad=pad)
                hd_all[b0:b1] = hd
            finally:
                for s in kv:
                    self.engine.free_slot(s)
            out_rows.append(out)
            n_final = max(n_final, out.shape[1])
        if heads is not None:
            for b in range(B):
                heads[b] = hd_all[b]
        res = torch.full((B, n_final), pad, dtype=torch.int64)
        r = 0
        for out in out_rows:
            res[r: r + out.shape[0], : out.shape[1]] = out
            r += out.shape[0]
        return res

    @torch.no_grad()
    def generate(self, inputs: Optional[torch.Tensor] = None, images_or_videos=None, times=None, scores=None,
                 video_timestamps=None, modal_list=None, **kwargs):
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")                       # trace_mistral.py:282-283
        attention_mask = kwargs.pop("attention_mask", None)
        kwargs.pop("position_ids", None)
        heads = kwargs.pop("heads", None)
        max_new = kwargs.pop("max_new_tokens", None) or 20
        do_sample = bool(kwargs.pop("do_sample", False))
        temperature = kwargs.pop("temperature", None)
        top_k = kwargs.pop("top_k", None)
        top_p = kwargs.pop("top_p", None)
        seed = kwargs.pop("seed", None)
        kwargs.pop("use_cache", None)
        kwargs.pop("stopping_criteria", None)   # constructed by every driver, only trace/__init__.py passes it
        pad = kwargs.pop("pad_token_id", None)
        eos = kwargs.pop("eos_token_id", None) or self.generation_config.eos_token_id
        if pad is None:
            pad = eos
        if times or scores:
            # training-time inputs: the reference embeds them at the -203 / -204 placeholders (trace_arch.py:423-426)
            raise NotImplementedError("`times` / `scores` (the <time>/<score> placeholders of training prompts) are not part of the inference path")
        if do_sample:
            # HF `_sample` with the default warpers, as scripts/inference/inference.py:62,71-73 drives it (T=0.2, top_k=50)
            if top_p is not None and float(top_p) < 1.0:
                raise NotImplementedError("top_p < 1 is not implemented (no TRACE driver sets it)")
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())        # follows torch.manual_seed like HF's multinomial
            self.engine.set_sampling(True, 1.0 if temperature is None else float(temperature),
                                     50 if top_k is None else int(top_k), seed)
        else:
            self.engine.set_sampling(False)
        assert inputs is not None and inputs.dim() == 2
        B, L = inputs.shape
        if attention_mask is not None and not bool(attention_mask.to(torch.bool).all()):
            raise NotImplementedError("left/right padded prompts are not supported (the reference cannot batch them either, SURVEY §9-14)")
        if heads is not None:
            assert len(heads) == B                                                               # trace_mistral.py:245
        if images_or_videos is None:
            # plain text generate: inputs_embeds = embed_tokens(inputs) (trace_mistral.py:306-307)
            ids = inputs.to(self.device, torch.int64).contiguous()
            if bool(((ids < 0) | (ids >= self.vocab_size + 27)).any()):
                raise ValueError("text-only generate: ids must be plain token ids (no multimodal placeholders, < vocab + 27)")
            return self._generate_from_slots(ids, None, None, heads, max_new, eos, pad).to(inputs.device)
        modal_list = modal_list or ["video"] * B
        nf = self.num_frames()
        clips = [x.unsqueeze(0).expand(nf, -1, -1, -1) if m == "image" else x for x, m in zip(images_or_videos, modal_list)]
        T = clips[0].shape[0]
        if any(c.shape != clips[0].shape for c in clips):
            raise RuntimeError("stack expects each tensor to be equal size (clips of one batch must share T)")  # trace_arch.py:222
        ids = inputs.to(self.device, torch.int64).contiguous()
        nvid = ((ids == MMODAL_TOKEN_INDEX["VIDEO"]) | (ids == MMODAL_TOKEN_INDEX["IMAGE"])).sum(1)
        assert bool((nvid == 1).all()), "only have one video inputs!"                            # trace_arch.py:408
        if bool(((ids == MMODAL_TOKEN_INDEX["TIME"]) | (ids == MMODAL_TOKEN_INDEX["SCORE"])).any()):
            raise NotImplementedError("<time>/<score> placeholders need `times` / `scores` (training prompts), not part of the inference path")
        if bool((ids >= self.vocab_size + 27).any()):
            raise ValueError("token id out of range")
        if video_timestamps is None:
            raise NotImplementedError("video_timestamps=None (no time rows) is not used by any TRACE driver")
        tid = torch.tensor([frame_time_ids(ts) for ts in video_timestamps], dtype=torch.int32, device=self.device)
        assert tid.shape == (B, T, 6), f"time ids shape {tuple(tid.shape)}"
        slots = self._encode_slots(clips)
        res = self._generate_from_slots(ids, slots, tid, heads, max_new, eos, pad)
        return res.to(inputs.device)
