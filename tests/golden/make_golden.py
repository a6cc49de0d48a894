"""Generate golden fixtures by running the REAL reference (gyxxyg/TRACE) in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Needs /root/reference (absent on the GPU box, so the fixtures are committed).  The reference classes
are imported through oracle/ref_shim.py on the container's transformers 5.5.0 /
torch 2.11 in fp32 on CPU; weights and inputs come from trace_b200/synth.py so every consumer
(oracle, CUDA engine) can rebuild them bit-identically from (dims, seed).

What is driven: TraceMistralForCausalLM.prepare_inputs_labels_for_multimodal (prefill splice +
decode-step embedding dispatch), TraceMistralForCausalLM.forward (decoder, heads, head mask) and the
head-switch rule of prepare_inputs_for_generation (trace_mistral.py:336-344) in a hand-written greedy
loop, because transformers>=5 `generate()` passes input_ids=None on the first iteration and trips
trace_mistral.py:339.  The loop is the 4.40.1 `_greedy_search` contract: argmax,
finished rows emit pad, stop when all rows hit eos.
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from trace_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def build_reference_model(ref, d: synth.TraceDims, seed: int):
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel
    clip_dir = os.path.join(tempfile.mkdtemp(prefix="synth_"), "clip-synth")
    ccfg = CLIPVisionConfig(hidden_size=d.vit_width, intermediate_size=d.vit_mlp, num_hidden_layers=d.vit_layers,
                            num_attention_heads=d.vit_heads, image_size=d.image_size, patch_size=d.patch_size,
                            layer_norm_eps=d.vit_ln_eps, hidden_act="quick_gelu", projection_dim=64)
    CLIPVisionModel(ccfg).save_pretrained(clip_dir)
    CLIPImageProcessor(size={"shortest_edge": d.image_size},
                       crop_size={"height": d.image_size, "width": d.image_size}).save_pretrained(clip_dir)
    cfg = ref.TraceMistralConfig(
        hidden_size=d.hidden, intermediate_size=d.ffn, num_hidden_layers=d.layers, num_attention_heads=d.heads,
        num_key_value_heads=d.kv_heads, head_dim=d.head_dim, vocab_size=d.vocab, rms_norm_eps=d.rms_eps,
        rope_theta=d.rope_theta, sliding_window=None, max_position_embeddings=32768,
        mm_vision_tower=clip_dir, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
        mm_projector_type="spatial_slot", mm_hidden_size=d.vit_width, time_vocab_size=13, score_vocab_size=13,
        image_aspect_ratio="pad", num_frames=d.num_frames, attn_implementation="eager",
        pad_token_id=None, bos_token_id=1, eos_token_id=2)
    if getattr(cfg, "rope_parameters", None) is not None:
        cfg.rope_parameters["rope_theta"] = d.rope_theta
    model = ref.TraceMistralForCausalLM(cfg).eval().float()
    sd = {k: v.float() for k, v in synth.make_state_dict(d, seed).items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "position_ids" not in m and "inv_freq" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model


@torch.no_grad()
def reference_generate(model, d, ids_rows, clips, ts_rows, heads, max_new, eos=2, pad=2):
    """Hand greedy loop over the reference's own forward/prepare functions. Returns (ids [B,n], trace)."""
    B = len(ids_rows)
    ids = torch.tensor(ids_rows, dtype=torch.long)
    mask = torch.ones_like(ids)
    _, mask2, _, embeds, _, _, _ = model.prepare_inputs_labels_for_multimodal(
        ids, mask, None, None, [list(clips), ["video"] * B], [[]] * B, [[]] * B, video_timestamps=ts_rows)
    vis = model.encode_images_or_videos(list(clips), ["video"] * B, ts_rows)
    feats = model.get_model().get_vision_tower()(torch.stack(list(clips)).flatten(0, 1))
    heads = None if heads is None else list(heads)
    out = model(inputs_embeds=embeds, attention_mask=mask2, use_cache=True, heads=heads, return_dict=True,
                output_hidden_states=True)
    hid = out.hidden_states[-1]          # HF: last entry = after the final RMSNorm
    past = out.past_key_values
    finished = torch.zeros(B, dtype=torch.bool)
    gen, trace = [], []
    cur_mask = mask2
    for step in range(max_new):
        logits = out.logits[:, -1, :]
        tok = torch.argmax(logits, dim=-1)
        tok = torch.where(finished, torch.full_like(tok, pad), tok)
        trace.append({"heads": None if heads is None else list(heads), "logits": logits.clone()})
        gen.append(tok.clone())
        finished |= tok == eos
        if bool(finished.all()):
            break
        if heads is not None:  # trace_mistral.py:336-344
            for b, t in enumerate(tok):
                if int(t) in model.swap_tokens:
                    heads[b] = model.swap_tokens[int(t)]
        cur_mask = torch.cat([cur_mask, torch.ones(B, 1, dtype=cur_mask.dtype)], 1)
        out = model(input_ids=tok[:, None], attention_mask=cur_mask, past_key_values=past, use_cache=True,
                    heads=heads, return_dict=True)
        past = out.past_key_values
    return torch.stack(gen, 1), trace, {"embeds": embeds, "vis": vis, "feats": feats, "hid": hid}


def prompt_ids(d, L, seed):
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(3, d.vocab, (L,), generator=g).tolist()
    body[0] = 1            # BOS
    body[L // 3] = -201    # <video>
    body[-1] = -205        # trailing <sync> (evaluate.py:332)
    return body


def topk_pack(logit_row, lo, hi, k=8):
    seg = logit_row[lo:hi]
    k = min(k, seg.numel())
    v, i = torch.topk(seg, k)
    return v.numpy().astype(np.float32), (i + lo).numpy().astype(np.int64)


def make_case(ref, name, d, seed, B, T, L, heads0, max_new, strides=(3, 8, 32, 64)):
    """`heads0`: None (text-only logits), an int (every row starts on that head) or a list per row.
    `strides` = (feats patch stride, feats width stride, vis width stride, embeds / hidden width stride)."""
    model = build_reference_model(ref, d, seed)
    clips, ts_rows, ids_rows = [], [], []
    for b in range(B):
        px, ts = synth.synthetic_clip(d, T, seed=1234 + b, duration_s=120.0 + 7.3 * b)
        clips.append(px.float())
        ts_rows.append(ts)
        ids_rows.append(prompt_ids(d, L, 99 + b))
    heads = None if heads0 is None else (list(heads0) if isinstance(heads0, (list, tuple)) else [heads0] * B)
    ids, trace, mid = reference_generate(model, d, ids_rows, clips, ts_rows, heads, max_new)
    ranges = [(0, d.time_base), (d.time_base, d.score_base), (d.score_base, d.total_vocab)]
    tv, ti, hh = [], [], []
    for st in trace:
        rv, ri, rh = [], [], []
        for b in range(B):
            h = 0 if st["heads"] is None else st["heads"][b]
            lo, hi = ranges[h]
            v, i = topk_pack(st["logits"][b], lo, hi)
            rv.append(v), ri.append(i), rh.append(h)
        tv.append(np.stack(rv)), ti.append(np.stack(ri)), hh.append(rh)
    meta = {"name": name, "dims": dataclasses.asdict(d), "seed": seed, "B": B, "T": T, "L": L,
            "heads0": heads0, "max_new": max_new, "strides": list(strides), "clip_seeds": [1234 + b for b in range(B)],
            "durations": [120.0 + 7.3 * b for b in range(B)], "prompt_seeds": [99 + b for b in range(B)],
            "transformers": __import__("transformers").__version__, "torch": torch.__version__}
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        meta=json.dumps(meta), prompt_ids=np.array(ids_rows, dtype=np.int64),
        out_ids=ids.numpy().astype(np.int64),
        top_vals=np.stack(tv), top_idx=np.stack(ti), step_heads=np.array(hh, dtype=np.int64),
        # strided samples of the intermediate seams (full tensors would be MBs)
        feats_sample=mid["feats"][:, ::strides[0], ::strides[1]].numpy().astype(np.float32),   # CLIP hidden_states[-2][:,1:]
        vis_sample=mid["vis"][:, :, ::strides[2]].numpy().astype(np.float32),                  # [B,14T,hidden/stride]
        embeds_sample=mid["embeds"][:, :, ::strides[3]].numpy().astype(np.float32),            # [B,S,hidden/stride]
        hid_sample=mid["hid"][:, ::7, ::strides[3]].numpy().astype(np.float32),                # final-normed prefill hidden, every 7th position
        hid_last=mid["hid"][:, -1, :].numpy().astype(np.float32),                              # [B,hidden] the row the first token is decoded from
        prefill_len=np.array(mid["embeds"].shape[1]))
    print(name, "ids", ids.tolist(), "S", mid["embeds"].shape[1])


CASE_TABLE = {
    "tiny_b1": lambda ref: make_case(ref, "tiny_b1", synth.tiny_dims(4), seed=0, B=1, T=4, L=12, heads0=1, max_new=24),
    "tiny_b2": lambda ref: make_case(ref, "tiny_b2", synth.tiny_dims(4), seed=1, B=2, T=4, L=16, heads0=1, max_new=16),
    "tiny_text": lambda ref: make_case(ref, "tiny_text", synth.tiny_dims(4), seed=2, B=1, T=4, L=12, heads0=None, max_new=8),
    "mid_b1": lambda ref: make_case(ref, "mid_b1", synth.mid_dims(2), seed=3, B=1, T=2, L=20, heads0=1, max_new=12),
    # full WIDTHS of BASELINE configs 2-5 (ViT 1024 x 16 heads x 336 px, hidden 4096, ffn 14336, vocab 32000) at shallow
    # depth: every per-layer code path the full model takes (head partition over the whole vocabulary, k-phased down
    # projection at batch 4/8, several K/V tiles per attention work item) is compared with the reference itself
    "wide_b1": lambda ref: make_case(ref, "wide_b1", synth.wide_dims(8), seed=4, B=1, T=8, L=600, heads0=1, max_new=48,
                                     strides=(24, 32, 128, 256)),
    "wide_b4": lambda ref: make_case(ref, "wide_b4", synth.wide_dims(8), seed=5, B=4, T=8, L=600, heads0=[1, 0, 2, 0], max_new=32,
                                     strides=(24, 32, 128, 256)),
    "wide_b8": lambda ref: make_case(ref, "wide_b8", synth.wide_dims(8), seed=6, B=8, T=8, L=600,
                                     heads0=[1, 0, 2, 1, 0, 2, 0, 1], max_new=24, strides=(24, 32, 128, 256)),
    # context beyond 2304 positions (config 2 reaches 2527): 128 frames -> S = 2351
    "wide_long_b1": lambda ref: make_case(ref, "wide_long_b1", synth.wide_dims(128), seed=7, B=1, T=128, L=560, heads0=0,
                                          max_new=20, strides=(48, 64, 256, 512)),
}


def main():
    ref = load_reference()
    torch.manual_seed(0)
    names = sys.argv[1:] or list(CASE_TABLE)
    for n in names:
        import time
        t0 = time.time()
        CASE_TABLE[n](ref)
        print(f"{n}: {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
