"""Shared helpers: rebuild a golden case's inputs from its meta (weights are seeded, not stored)."""
import dataclasses
import json
import os

import numpy as np
import torch

from trace_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tiny_b1", "tiny_b2", "tiny_text", "mid_b1"]
# full widths of the BASELINE configs at shallow depth (tests/golden/make_golden.py): outputs of the reference itself
WIDE_CASES = ["wide_b1", "wide_b4", "wide_b8", "wide_long_b1"]


def row_heads0(meta):
    """Initial head per row (None, an int for every row, or a list)."""
    h = meta["heads0"]
    if h is None:
        return None
    return list(h) if isinstance(h, (list, tuple)) else [h] * meta["B"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    meta = json.loads(str(z["meta"]))
    d = synth.TraceDims(**meta["dims"])
    clips, ts = [], []
    for s, dur in zip(meta["clip_seeds"], meta["durations"]):
        px, t = synth.synthetic_clip(d, meta["T"], seed=s, duration_s=dur)
        clips.append(px)
        ts.append(t)
    return {"z": z, "meta": meta, "dims": d, "clips": clips, "ts": ts,
            "prompt_ids": z["prompt_ids"].tolist()}


def ranges(d):
    return [(0, d.time_base), (d.time_base, d.score_base), (d.score_base, d.total_vocab)]


def compare_ids_with_margin(got, ref_ids, top_vals, tol, label=""):
    """Greedy ids must equal the reference's; a divergence is tolerated only at a step where the
    reference's own top-2 margin is below `tol` (an fp tie), after which the rows are incomparable.
    Returns (n_compared, n_tie_breaks)."""
    ties = 0
    n = 0
    for b in range(len(ref_ids)):
        for s in range(len(ref_ids[b])):
            if s >= len(got[b]):
                raise AssertionError(f"{label} row {b}: output shorter than reference at step {s}")
            if int(got[b][s]) != int(ref_ids[b][s]):
                margin = float(top_vals[s][b][0] - top_vals[s][b][1])
                assert margin < tol, (f"{label} row {b} step {s}: got {got[b][s]} want {ref_ids[b][s]} "
                                      f"with reference top-2 margin {margin:.4g} >= tol {tol}")
                ties += 1
                break
            n += 1
    return n, ties
