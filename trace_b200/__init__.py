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


def infer(model, video, instruct, tokenizer, do_sample=False):
    """trace/__init__.py:23-75: text-head-only generate (no `heads`), returns the decoded string."""
    import torch
    from .constants import DEFAULT_MMODAL_TOKEN, MMODAL_TOKEN_INDEX
    from .conversation import conv_templates
    from .mm_utils import tokenizer_MMODAL_token
    conv = conv_templates["llama_2"].copy()
    conv.append_message(conv.roles[0], DEFAULT_MMODAL_TOKEN["VIDEO"] + "\n" + instruct)
    conv.append_message(conv.roles[1], None)
    ids = tokenizer_MMODAL_token(conv.get_prompt(), tokenizer, MMODAL_TOKEN_INDEX["VIDEO"], return_tensors="pt").unsqueeze(0).cuda()
    T = video.shape[0]
    with torch.inference_mode():
        out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[video.half().cuda()], modal_list=["video"],
                             video_timestamps=[[[float(i)] for i in range(T)]], do_sample=do_sample, max_new_tokens=1024, use_cache=True,
                             pad_token_id=tokenizer.eos_token_id)
    return tokenizer.batch_decode(out, skip_special_tokens=True)[0].strip()


def x_infer(video, question, model, tokenizer, mode="vanilla", do_sample=False):
    if mode == "mcqa":
        question = f"{question}\nAnswer with the option's letter from the given choices directly and only give the best option."
    elif mode == "openend":
        question = f"{question}\nAnswer the question using a single word or a short phrase with multiple words."
    return infer(model=model, tokenizer=tokenizer, video=video, instruct=question, do_sample=do_sample)
