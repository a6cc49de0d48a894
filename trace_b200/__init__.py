"""trace_b200 -- B200-native (sm_100a) implementation of the TRACE video-temporal-grounding inference path.

Package API mirrors trace/__init__.py:13-86 (model_init / infer / x_infer)."""
from functools import partial

__all__ = ["model_init", "infer", "x_infer"]


def model_init(model_path=None, **kw):
    from .constants import NUM_FRAMES
    from .mm_utils import get_model_name_from_path, process_video
    from .model.builder import load_pretrained_model
    assert model_path is not None, "no network: pass a local checkpoint directory"
    tokenizer, model, processor, _ = load_pretrained_model(model_path, None, get_model_name_from_path(model_path), **kw)
    nf = getattr(model.config, "num_frames", NUM_FRAMES)
    return model, partial(process_video, aspect_ratio=None, processor=processor, num_frames=nf), tokenizer


def infer(model, video, instruct, tokenizer, do_sample=False, video_timestamps=None):
    """trace/__init__.py:23-75: text-head-only generate (no `heads`), returns the decoded string.

    `video` is what the `process_video` partial of `model_init` returns -- the `(frames, timestamps)` pair -- or the frame
    tensor alone together with `video_timestamps`.  The reference passes no timestamps here, which its own
    `encode_images_or_videos` cannot splice (trace_arch.py:236-258 leaves the features un-flattened and the 14-per-frame
    placeholder count of :414 no longer matches); the time rows are part of every TRACE prompt, so they are required."""
    import torch
    if isinstance(video, (tuple, list)) and len(video) == 2 and video_timestamps is None:
        video, video_timestamps = video
    if video_timestamps is None:
        raise ValueError("infer() needs the per-frame timestamps: pass the (frames, timestamps) pair returned by the process_video "
                         "partial of model_init, or video_timestamps=")
    from .constants import DEFAULT_MMODAL_TOKEN, MMODAL_TOKEN_INDEX
    from .conversation import conv_templates
    from .mm_utils import tokenizer_MMODAL_token
    conv = conv_templates["llama_2"].copy()
    conv.append_message(conv.roles[0], DEFAULT_MMODAL_TOKEN["VIDEO"] + "\n" + instruct)
    conv.append_message(conv.roles[1], None)
    ids = tokenizer_MMODAL_token(conv.get_prompt(), tokenizer, MMODAL_TOKEN_INDEX["VIDEO"], return_tensors="pt").unsqueeze(0).cuda()
    assert len(video_timestamps) == video.shape[0], "one [seconds] entry per frame"
    with torch.inference_mode():
        out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[video.half().cuda()], modal_list=["video"],
                             video_timestamps=[list(video_timestamps)], do_sample=do_sample, temperature=0.2 if do_sample else 0.0,
                             max_new_tokens=1024, use_cache=True,
                             pad_token_id=tokenizer.eos_token_id)
    return tokenizer.batch_decode(out, skip_special_tokens=True)[0].strip()


def x_infer(video, question, model, tokenizer, mode="vanilla", do_sample=False, video_timestamps=None):
    if mode == "mcqa":
        question = f"{question}\nAnswer with the option's letter from the given choices directly and only give the best option."
    elif mode == "openend":
        question = f"{question}\nAnswer the question using a single word or a short phrase with multiple words."
    return infer(model=model, tokenizer=tokenizer, video=video, instruct=question, do_sample=do_sample, video_timestamps=video_timestamps)
