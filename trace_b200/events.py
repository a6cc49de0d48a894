"""ids -> events: the output parser every TRACE driver carries (trace/eval/evaluate.py:360-411; copies at
scripts/inference/inference.py:82-128, videomme/evaluate.py:388-436).  Works on a host list of ids (ONE
device->host copy per generate call instead of the reference's one `.item()` sync per token) and keeps the
reference's quirks: a caption is closed by <sync>=vocab or by the end of the sequence; `timestamps`/`scores`
entries are closed by their head's <sync>; digits are joined and parsed with float()."""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

_SYMS = ["<sync>", "<sep>"] + list("0123456789") + ["."]


def parse_events(output_ids: Sequence[int], decode_text: Callable[[List[int]], str], vocab: int = 32000) -> Dict[str, list]:
    tb, sb = vocab + 1, vocab + 14
    out = {"timestamps": [], "scores": [], "captions": []}
    cur_ts: List[float] = []
    cur_t: List[str] = []
    cur_ss: List[float] = []
    cur_s: List[str] = []
    cur_cap: List[int] = []
    for idx in output_ids:
        idx = int(idx)
        if idx <= vocab:
            if idx == vocab:
                out["captions"].append(decode_text(cur_cap))
                cur_cap = []
            else:
                cur_cap.append(idx)
        elif idx < sb:
            if idx == tb or idx == tb + 1:
                if cur_t:
                    cur_ts.append(float("".join(cur_t)))
                cur_t = []
                if idx == tb:
                    out["timestamps"].append(cur_ts)
                    cur_ts = []
            else:
                cur_t.append(_SYMS[idx - tb])
        else:
            if idx == sb or idx == sb + 1:
                if cur_s:
                    cur_ss.append(float("".join(cur_s)))
                cur_s = []
                if idx == sb:
                    out["scores"].append(cur_ss)
                    cur_ss = []
            else:
                cur_s.append(_SYMS[idx - sb])
    if cur_cap:
        out["captions"].append(decode_text(cur_cap))
    return out


def count_events(output_ids: Sequence[int], vocab: int = 32000) -> int:
    """Number of (timestamps, scores, caption) triples the parser emits = len(outputs['captions'])."""
    return len(parse_events(output_ids, lambda t: "", vocab)["captions"])
