"""load_pretrained_model (mirrors trace/model/builder.py:29-156, plain fp16 branch :112-154).

`model_path` is a directory holding `config.json` (the TRACE keys of SURVEY §5 + Mistral keys) and either
`*.safetensors` / `pytorch_model*.bin` weight shards with the HF names of SURVEY §4, or -- for the offline synthetic
checkpoints used by the tests and the bench -- a `synthetic.json` {"dims": {...}, "seed": n, "structured": bool}.
Returns (tokenizer, model, processor, context_len) exactly like the reference.  LoRA / 4-bit / 8-bit branches of
the reference are dead code (they reference a non-existent class) and are rejected here."""
from __future__ import annotations

import glob
import json
import os

import torch

from .. import synth
from .trace_mistral import TraceMistralConfig, TraceMistralForCausalLM


def _vision_tower_dir(cfg: dict, model_path: str):
    """The directory `mm_vision_tower` names when it is a local path (trace_arch.py:34-36 builds the tower from
    `config.mm_vision_tower`; clip_encoder.py:23-29 loads CLIPVisionModel + CLIPImageProcessor from it)."""
    vt = cfg.get("mm_vision_tower")
    if not vt:
        return None
    for cand in (vt, os.path.join(model_path, vt), os.path.join(os.path.dirname(os.path.abspath(model_path)), os.path.basename(str(vt)))):
        if os.path.isdir(cand):
            return cand
    return None


def _dims_from_hf_config(cfg: dict, tower_dir=None) -> synth.TraceDims:
    """Mistral keys from config.json.  The CLIP tower's geometry is not stored there in the reference (it loads
    `mm_vision_tower` by name, trace_arch.py:34-36): it is read from the tower directory's own config.json when that is a local
    directory, or from an optional `vision_config` block; otherwise the ViT-L/14-336 defaults apply."""
    v = dict(cfg.get("vision_config", {}) or {})
    if tower_dir is not None and os.path.exists(os.path.join(tower_dir, "config.json")):
        tc = json.load(open(os.path.join(tower_dir, "config.json")))
        v = {**tc.get("vision_config", tc), **v}
    extra = {k2: v[k1] for k1, k2 in (("hidden_size", "vit_width"), ("num_attention_heads", "vit_heads"), ("num_hidden_layers", "vit_layers"),
                                      ("intermediate_size", "vit_mlp"), ("image_size", "image_size"), ("patch_size", "patch_size")) if k1 in v}
    return synth.TraceDims(
        **extra,
        hidden=cfg.get("hidden_size", 4096), layers=cfg.get("num_hidden_layers", 32), heads=cfg.get("num_attention_heads", 32),
        kv_heads=cfg.get("num_key_value_heads", 8), head_dim=cfg.get("head_dim", 128), ffn=cfg.get("intermediate_size", 14336),
        vocab=cfg.get("vocab_size", 32000), rope_theta=float(cfg.get("rope_theta", 1e6)), rms_eps=float(cfg.get("rms_norm_eps", 1e-5)),
        num_frames=cfg.get("num_frames", 8))


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto", device="cuda",
                          use_flash_attn=False, **kwargs):
    if load_8bit or load_4bit or model_base is not None or "lora" in (model_name or "").lower():
        raise NotImplementedError("only the plain fp16 branch of load_pretrained_model (builder.py:112-114) is supported")
    dev = 0 if device == "cuda" else int(str(device).split(":")[-1])
    syn = os.path.join(model_path, "synthetic.json")
    if os.path.exists(syn):
        meta = json.load(open(syn))
        d = synth.TraceDims(**meta["dims"])
        model = TraceMistralForCausalLM(TraceMistralConfig(d), device=dev, **kwargs)
        sd = synth.make_state_dict(d, meta.get("seed", 0), device="cuda")
        if meta.get("structured"):
            sd.update(synth.make_structured_overrides(d, meta.get("seed", 0), lambda n: sd[n]))
        model.load_state_dict(sd)
        tokenizer = synth.StandInTokenizer(d.vocab)
    else:
        cfg = json.load(open(os.path.join(model_path, "config.json")))
        tower_dir = _vision_tower_dir(cfg, model_path)
        d = _dims_from_hf_config(cfg, tower_dir)
        model = TraceMistralForCausalLM(TraceMistralConfig(d, **{k: v for k, v in cfg.items() if k not in ("vision_config", "architectures")}), device=dev, **kwargs)
        shards = sorted(glob.glob(os.path.join(model_path, "*.safetensors"))) or sorted(glob.glob(os.path.join(model_path, "pytorch_model*.bin")))
        if not shards:
            raise FileNotFoundError(f"no weight shards under {model_path}")
        def read(sh):
            if sh.endswith(".safetensors"):
                from safetensors.torch import load_file
                return load_file(sh, device="cpu")
            return torch.load(sh, map_location="cpu")

        seen_tower = False
        for sh in shards:
            for name, t in read(sh).items():
                seen_tower = seen_tower or name.startswith("model.vision_tower.")
                model.engine.load_weight(name, t.half())
        if not seen_tower:
            # released checkpoints keep the (frozen) CLIP tower out of the LLM shards: it is loaded from `mm_vision_tower`
            # (trace_arch.py:34-36, clip_encoder.py:23-29) under the prefix the state dict would have had
            if tower_dir is None:
                raise FileNotFoundError(f"the checkpoint has no vision tower weights and config.mm_vision_tower={cfg.get('mm_vision_tower')!r} "
                                        "is not a local directory (there is no network)")
            tsh = sorted(glob.glob(os.path.join(tower_dir, "*.safetensors"))) or sorted(glob.glob(os.path.join(tower_dir, "pytorch_model*.bin")))
            if not tsh:
                raise FileNotFoundError(f"no weight shards under {tower_dir}")
            for sh in tsh:
                for name, t in read(sh).items():
                    if name.startswith("vision_model."):            # text tower / projections of a full CLIPModel are not used
                        model.engine.load_weight("model.vision_tower.vision_tower." + name, t.half())
        model.engine.load_state_dict([])
        if any(os.path.exists(os.path.join(model_path, f)) for f in ("tokenizer.model", "tokenizer.json", "tokenizer_config.json")):
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=False)                # builder.py:113
        else:   # weight-only export (offline tests): the deterministic stand-in used by the bench
            tokenizer = synth.StandInTokenizer(d.vocab)
    processor = model.get_vision_tower().image_processor                       # builder.py:149
    context_len = getattr(model.config, "max_sequence_length", 2048)            # builder.py:151-154
    return tokenizer, model, processor, context_len
