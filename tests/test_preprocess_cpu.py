"""Pins the resampling restatement (oracle/preprocess_oracle.py) and the product's coefficient tables
(trace_b200/preprocess.py) against Pillow itself and against the reference's processor chain (PIL + HF CLIPImageProcessor)."""
import numpy as np
import pytest
from PIL import Image

from oracle import preprocess_oracle as po
from trace_b200.preprocess import pil_bicubic_tables, resize_geometry


@pytest.mark.parametrize("h,w,ow,oh", [(480, 640, 336, 252), (90, 90, 56, 56), (336, 336, 336, 336), (57, 123, 200, 93), (720, 1280, 336, 336)])
def test_resize_restatement_is_bit_exact_against_pillow(h, w, ow, oh):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    got = po.pil_resize_bicubic(img, ow, oh)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("n_in,n_out", [(1920, 336), (336, 336), (90, 56), (57, 200), (1280, 597)])
def test_product_tables_equal_oracle_tables(n_in, n_out):
    bounds, coef, ksize = pil_bicubic_tables(n_in, n_out)
    tabs = po._tables(n_in, n_out)
    for i, (lo, k) in enumerate(tabs):
        assert bounds[i, 0] == lo and bounds[i, 1] == len(k) and len(k) <= ksize
        assert np.array_equal(coef[i, : len(k)], k) and not coef[i, len(k):].any()


@pytest.mark.parametrize("h,w,pad", [(360, 640, True), (640, 360, True), (336, 336, True), (480, 640, False)])
def test_frame_preprocess_matches_reference_chain(h, w, pad):
    """expand2square (mm_utils.py:259-270) + HF CLIPImageProcessor.preprocess as process_video calls them (:455-459)."""
    transformers = pytest.importorskip("transformers")
    from trace_b200.mm_utils import expand2square
    # the reference pins transformers 4.40.1, whose CLIPImageProcessor resizes with PIL; transformers 5.x keeps that class as
    # CLIPImageProcessorPil (its default CLIPImageProcessor is the torchvision backend: +-1 LSB on ~0.1 % of the pixels)
    cls = getattr(transformers, "CLIPImageProcessorPil", None) or transformers.CLIPImageProcessor
    proc = cls(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    pil = Image.fromarray(img)
    if pad:
        pil = expand2square(pil, tuple(int(x * 255) for x in proc.image_mean))
    want = proc.preprocess([pil], return_tensors="np")["pixel_values"][0]
    got = po.preprocess_frame(img, 336, tuple(proc.image_mean), tuple(proc.image_std), pad)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-6        # same uint8 pixels; float32 rescale/normalise order
    cw, ch, px, py, ow, oh, cx, cy = resize_geometry(h, w, 336, pad)
    assert (ow >= 336 and oh >= 336 and min(ow, oh) == 336)
