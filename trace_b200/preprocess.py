"""Device-side frame pre-processing: the per-frame host work of `process_video`
(trace/mm_utils.py:424-462: expand2square :259-270, CLIPImageProcessor.preprocess, fp16 cast at evaluate.py:316) as two
byte-exact integer resampling passes + normalisation on the GPU (csrc/preprocess.cu).

This module is the host half: it builds Pillow's resampling tables.  Pillow (third-party, Resample.c: `precompute_coeffs`,
`normalize_coeffs_8bpc`, `bicubic_filter`) resamples 8-bit images with per-output-pixel tap ranges and 22-bit fixed-point
coefficients computed in float64; the published algorithm is restated here so that the integer stage on the device is
bit-exact against PIL (pinned by tests/test_preprocess_cpu.py against PIL itself)."""
from __future__ import annotations

import ctypes
import math
from functools import lru_cache
from typing import Sequence, Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=64)
def pil_bicubic_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(bounds int32 [out,2] = (first tap, taps), coef int32 [out,ksize], ksize) of Pillow's BICUBIC resample in -> out."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coef = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            coef[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coef, ksize


def resize_geometry(h: int, w: int, size: int, pad: bool):
    """Canvas (after expand2square when `pad`), paste offset, resized size (CLIP shortest-edge rule) and centre crop."""
    if pad:
        side = max(h, w)
        cw, ch, px, py = side, side, (side - w) // 2, (side - h) // 2
    else:
        cw, ch, px, py = w, h, 0, 0
    short, long = (cw, ch) if cw <= ch else (ch, cw)
    new_long = int(size * long / short)
    ow, oh = (size, new_long) if cw <= ch else (new_long, size)
    return cw, ch, px, py, ow, oh, (ow - size) // 2, (oh - size) // 2


def preprocess_frames_cuda(frames: torch.Tensor, size: int = 336, image_mean: Sequence[float] = CLIP_MEAN,
                           image_std: Sequence[float] = CLIP_STD, aspect_ratio: str = "pad") -> torch.Tensor:
    """uint8 [F,H,W,3] RGB frames on the GPU -> fp16 [F,3,size,size] CLIP-normalised (what
    `process_video(...)[0].half().cuda()` holds in the reference)."""
    from .engine import _check, _stream, load_library
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3 and frames.is_contiguous()
    lib = load_library()
    F, H, W, _ = frames.shape
    cw, ch, px, py, ow, oh, cx, cy = resize_geometry(H, W, size, aspect_ratio == "pad")
    hb, hc, hk = pil_bicubic_tables(cw, ow)
    vb, vc, vk = pil_bicubic_tables(ch, oh)
    row0 = int(vb[0, 0])
    rows = int(vb[-1, 0] + vb[-1, 1]) - row0
    pad_rgb = [int(x * 255) for x in image_mean]                                   # mm_utils.py:455
    geom = (ctypes.c_int32 * 15)(cw, px, py, ow, oh, size, cx, cy, row0, rows, hk, vk, *pad_rgb)
    dev = frames.device
    t_hb, t_hc = torch.from_numpy(hb).to(dev), torch.from_numpy(hc).to(dev)
    t_vb, t_vc = torch.from_numpy(vb).to(dev), torch.from_numpy(vc).to(dev)
    tmp = torch.empty((F, rows, ow, 3), dtype=torch.uint8, device=dev)
    out = torch.empty((F, 3, size, size), dtype=torch.float16, device=dev)
    mean = (ctypes.c_float * 3)(*image_mean)
    std = (ctypes.c_float * 3)(*image_std)
    _check(lib.trace_b200_preprocess_frames(frames.data_ptr(), F, H, W, geom, t_hb.data_ptr(), t_hc.data_ptr(), t_vb.data_ptr(),
                                            t_vc.data_ptr(), mean, std, tmp.data_ptr(), out.data_ptr(), _stream()))
    return out
