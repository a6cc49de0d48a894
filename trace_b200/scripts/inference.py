"""Single-video inference script (mirrors scripts/inference/inference.py:15-128): load the model, pre-process one video (or
image), build the llama_2 prompt ending in `<sync>`, generate with `heads=[1]`, parse the ids into events and print them.

    python -m trace_b200.scripts.inference --model_path CKPT --video assets/sora.mp4 [--num_frames 64] [--question "..."]

The reference script samples (`do_sample=True, temperature=0.2`, :62,72); this engine implements the greedy path of the
evaluation drivers, so `--greedy` is the default and sampling is rejected loudly."""
from __future__ import annotations

import argparse

import torch

from ..constants import DEFAULT_MMODAL_TOKEN
from ..conversation import conv_templates
from ..events import parse_events
from ..mm_utils import get_model_name_from_path, process_image, process_video, tokenizer_MMODAL_token_all
from ..model.builder import load_pretrained_model

DEFAULT_QUESTION = ("Localize a series of activity events in the video, output the start and end timestamp for each event, "
                    "and describe each event with sentences.")


def inference(model_path, path, question=DEFAULT_QUESTION, num_frames=64, modal="video", max_new_tokens=1024, device_preprocess=False):
    tokenizer, model, processor, _ = load_pretrained_model(model_path, None, get_model_name_from_path(model_path))
    aspect = getattr(model.config, "image_aspect_ratio", "pad")
    if modal == "video":
        tensor, stamps = process_video(path, processor, aspect, num_frames=num_frames, device="cuda" if device_preprocess else None)
        tag = DEFAULT_MMODAL_TOKEN["VIDEO"]
    else:
        tensor = process_image(path, processor, aspect)[0]
        stamps = [[float(i)] for i in range(getattr(model.config, "num_frames", 8))]
        tag = DEFAULT_MMODAL_TOKEN["IMAGE"]
    conv = conv_templates["llama_2"].copy()
    conv.append_message(conv.roles[0], tag + "\n" + question)
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt() + "<sync>"
    input_ids = tokenizer_MMODAL_token_all(prompt, tokenizer, return_tensors="pt").unsqueeze(0)
    heads = [1]
    with torch.inference_mode():
        output_ids = model.generate(input_ids, attention_mask=torch.ones_like(input_ids), images_or_videos=[tensor.to(torch.float16)],
                                    modal_list=[modal], do_sample=False, temperature=0.0, max_new_tokens=max_new_tokens, use_cache=True,
                                    pad_token_id=tokenizer.eos_token_id, video_timestamps=[stamps], heads=heads)
    outputs = parse_events(output_ids[0].tolist(), lambda toks: tokenizer.decode(toks, skip_special_tokens=True),
                           getattr(model, "vocab_size", 32000))
    print(outputs)
    return outputs


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--video", default=None)
    ap.add_argument("--image", default=None)
    ap.add_argument("--question", default=DEFAULT_QUESTION)
    ap.add_argument("--num_frames", type=int, default=64)
    ap.add_argument("--max_new_tokens", type=int, default=1024)
    ap.add_argument("--device_preprocess", action="store_true")
    a = ap.parse_args(argv)
    assert (a.video is None) != (a.image is None), "pass exactly one of --video / --image"
    return inference(a.model_path, a.video or a.image, a.question, a.num_frames, "video" if a.video else "image", a.max_new_tokens,
                     a.device_preprocess)


if __name__ == "__main__":
    main()
