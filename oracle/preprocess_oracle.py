"""TEST INFRASTRUCTURE ONLY (see oracle/trace_oracle.py): numpy restatement of the reference's per-frame pre-processing,
trace/mm_utils.py:424-462 -> expand2square (:259-270) + CLIPImageProcessor.preprocess.  The resize is Pillow's
(third-party, pinned here: Pillow 12.2.0 in this image; algorithm: Resample.c precompute_coeffs / normalize_coeffs_8bpc /
ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc), restated independently of trace_b200/preprocess.py and
pinned against PIL.Image.resize itself in tests/test_preprocess_cpu.py (bit-exact on uint8)."""
import math

import numpy as np

PREC = 22


def _cubic(x):
    x = abs(x)
    if x < 1.0:
        return (1.5 * x - 2.5) * x * x + 1.0
    if x < 2.0:
        return -0.5 * (((x - 5.0) * x + 8.0) * x - 4.0)
    return 0.0


def _tables(n_in, n_out):
    scale = n_in / n_out
    fs = scale if scale > 1.0 else 1.0
    sup = 2.0 * fs
    out = []
    for i in range(n_out):
        c = (i + 0.5) * scale
        lo = max(int(c - sup + 0.5), 0)
        hi = min(int(c + sup + 0.5), n_in)
        w = np.array([_cubic((x - c + 0.5) / fs) for x in range(lo, hi)], dtype=np.float64)
        tot = 0.0
        for v in w:
            tot += v
        if tot != 0.0:
            w = w / tot
        k = np.array([int(-0.5 + v * (1 << PREC)) if v < 0 else int(0.5 + v * (1 << PREC)) for v in w], dtype=np.int64)
        out.append((lo, k))
    return out


def _pass(img, n_out, axis):
    """img uint8 [..]; resample along `axis` (0 rows / 1 columns)."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    res = np.empty((n_out,) + img.shape[1:], dtype=np.uint8)
    for i, (lo, k) in enumerate(_tables(img.shape[0], n_out)):
        acc = (1 << (PREC - 1)) + np.tensordot(k, img[lo: lo + len(k)], axes=(0, 0))
        res[i] = np.clip(acc >> PREC, 0, 255).astype(np.uint8)
    return np.moveaxis(res, 0, axis)


def pil_resize_bicubic(img_u8, out_w, out_h):
    """uint8 [H,W,3] -> uint8 [out_h,out_w,3]; horizontal pass (only the rows the vertical pass needs), then vertical."""
    return _pass(_pass(img_u8, out_w, 1), out_h, 0)


def preprocess_frame(rgb_u8, size=336, mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711), pad=True):
    h, w, _ = rgb_u8.shape
    if pad and h != w:
        side = max(h, w)
        canvas = np.empty((side, side, 3), dtype=np.uint8)
        canvas[:] = np.array([int(x * 255) for x in mean], dtype=np.uint8)
        canvas[(side - h) // 2: (side - h) // 2 + h, (side - w) // 2: (side - w) // 2 + w] = rgb_u8
        rgb_u8 = canvas
        h = w = side
    short, long = (w, h) if w <= h else (h, w)
    new_long = int(size * long / short)
    ow, oh = (size, new_long) if w <= h else (new_long, size)
    r = pil_resize_bicubic(rgb_u8, ow, oh)
    top, left = (oh - size) // 2, (ow - size) // 2
    r = r[top: top + size, left: left + size]
    x = (r.astype(np.float64) * (1.0 / 255.0)).astype(np.float32)
    x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
