"""Pre-processing utilities of the TRACE surface (mirrors trace/mm_utils.py: process_image :358-376, process_video :379-471,
expand2square :259-270, tokenizer_MMODAL_token(_all) :496-554, get_model_name_from_path :558-564,
KeywordsStoppingCriteria :567-600).  Host (CPU) work exactly as in the reference; frames are decoded with OpenCV
because decord is not available offline."""
from __future__ import annotations

import re
from typing import List

import numpy as np
import torch

from .constants import DEFAULT_MMODAL_TOKEN, IMAGE_TOKEN_INDEX, MAX_FRAMES, MMODAL_INDEX_TOKEN, MMODAL_TOKEN_INDEX, NUM_FRAMES


def expand2square(pil_img, background_color):
    """Pad to a square with the CLIP mean colour, image centred (mm_utils.py:259-270)."""
    from PIL import Image
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


def frame_sample(duration: int, num_frames: int) -> np.ndarray:
    """Uniform sampler: np.linspace(0, duration-1, num_frames, dtype=int) (mm_utils.py:382)."""
    return np.linspace(0, duration - 1, num_frames, dtype=int)


def _read_video(path: str):
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise FileNotFoundError(path)
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    fps = float(cap.get(cv2.CAP_PROP_FPS))
    return cap, n, fps


def process_image(image_path, processor, aspect_ratio="pad", num_frames=NUM_FRAMES, image_grid=False):
    """path -> Tensor[1,3,H,W] fp32: RGB, pad to square with the CLIP mean colour, CLIP preprocessing (mm_utils.py:358-376).
    The model repeats an `image` input `num_frames` times itself (trace_arch.py:221)."""
    from PIL import Image
    if image_grid:
        raise NotImplementedError("image_grid is not used by the inference drivers")
    image = Image.open(image_path).convert("RGB")
    if aspect_ratio == "pad":
        image = expand2square(image, tuple(int(x * 255) for x in processor.image_mean))
    return processor.preprocess([image], return_tensors="pt")["pixel_values"]


def process_video(video_path, processor, aspect_ratio="pad", num_frames=NUM_FRAMES, image_grid=False, sample_scheme="uniform",
                  device=None):
    """path -> (Tensor[T,3,H,W] fp32, [[seconds]]*T).  Raises ImportError for >9999 s / negative timestamps as the
    reference does (mm_utils.py:466-469).

    `device` (extension, default None = the reference's host path): a CUDA device; the decoded uint8 frames are uploaded
    once and padded / resized / normalised there (csrc/preprocess.cu, bit-exact integer resampling against PIL), and the
    returned tensor is the fp16 CUDA tensor the drivers build with `.to(dtype=torch.float16, device='cuda')`."""
    from PIL import Image
    import cv2
    if image_grid or sample_scheme != "uniform":
        raise NotImplementedError("only the uniform sampler used by the inference drivers is provided")
    cap, duration, fps = _read_video(video_path)
    ids = frame_sample(duration, num_frames)
    if len(ids) > MAX_FRAMES:
        ids = np.linspace(0, duration - 1, MAX_FRAMES, dtype=int)
    frames = []
    for i in ids:
        cap.set(cv2.CAP_PROP_POS_FRAMES, int(i))
        ok, bgr = cap.read()
        if not ok:
            raise RuntimeError(f"cannot decode frame {i} of {video_path}")
        frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    cap.release()
    stamps = [[int(i) / fps] for i in ids]
    if stamps[-1][0] > 9999:
        raise ImportError("The video is too long!")
    if stamps[0][0] < 0:
        raise ImportError("Timestamp can not be less than zero")
    if device is not None:
        from .preprocess import preprocess_frames_cuda
        size = getattr(processor, "size", 336)
        size = size if isinstance(size, int) else (size.get("shortest_edge") or size.get("height"))
        batch = torch.from_numpy(np.stack(frames)).to(device, non_blocking=True)
        video = preprocess_frames_cuda(batch.contiguous(), size, tuple(processor.image_mean), tuple(processor.image_std),
                                       "pad" if aspect_ratio == "pad" else "none")
        return video, stamps
    images = [Image.fromarray(f) for f in frames]
    if aspect_ratio == "pad":
        images = [expand2square(im, tuple(int(x * 255) for x in processor.image_mean)) for im in images]
    video = processor.preprocess(images, return_tensors="pt")["pixel_values"]
    if stamps[-1][0] > 9999:
        raise ImportError("The video is too long!")
    if stamps[0][0] < 0:
        raise ImportError("Timestamp can not be less than zero")
    return video, stamps


def tokenizer_MMODAL_token(prompt, tokenizer, MMODAL_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """Split on ONE modality tag and insert its sentinel id (mm_utils.py:496-516); BOS kept once."""
    tag = f"<{MMODAL_INDEX_TOKEN[MMODAL_token_index].lower()}>"
    chunks = [tokenizer(c).input_ids for c in prompt.split(tag)]
    ids: List[int] = []
    offset = 0
    if chunks and chunks[0] and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        ids.append(chunks[0][0])
    for i, c in enumerate(chunks):
        if i:
            ids.append(MMODAL_token_index)
        ids.extend(c[offset:])
    return torch.tensor(ids, dtype=torch.long) if return_tensors == "pt" else ids


def tokenizer_MMODAL_token_all(prompt, tokenizer, return_tensors=None):
    """Split on every modality tag (<video>, <sync>, ...) and insert the sentinel ids -201 / -205 / ...
    (mm_utils.py:519-554): first chunk keeps its BOS, later chunks drop theirs."""
    pattern = "|".join(map(re.escape, DEFAULT_MMODAL_TOKEN.values()))
    chunks = [tokenizer(c).input_ids for c in re.split(pattern, prompt)]
    seps = [MMODAL_TOKEN_INDEX[m[1:-1].upper()] for m in re.findall(pattern, prompt)]
    assert len(chunks) == len(seps) + 1 or len(chunks) == 0
    offset = 1 if chunks and chunks[0] and chunks[0][0] == tokenizer.bos_token_id else 0
    ids: List[int] = list(chunks[0])
    for c, s in zip(chunks[1:], seps):
        ids.append(s)
        ids.extend(c[offset:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return ids


def get_model_name_from_path(model_path: str) -> str:
    return "_".join(model_path.strip("/").split("/"))


class KeywordsStoppingCriteria:
    """Constructed by every driver; only trace/__init__.py:infer passes it to generate (mm_utils.py:567-600)."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.keyword_ids = []
        for k in keywords:
            ids = tokenizer(k).input_ids
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def __call__(self, output_ids, scores=None, **kw) -> bool:
        for row in output_ids:
            text = self.tokenizer.decode([int(t) for t in row[self.start_len:] if int(t) < self.tokenizer.vocab_size],
                                         skip_special_tokens=True)
            if not any(k in text for k in self.keywords):
                return False
        return True
