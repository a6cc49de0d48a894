"""Host-side halves of the TRACE towers: the 13-symbol time / score tokenizers and number formatting
(time_encoder.py:52-108, score_encoder.py:52-116).  String formatting stays on the host exactly as in the
reference (`format(t, '0>6.1f')` is a correctly-rounded decimal of the binary float); the
embedding lookups themselves run inside the CUDA splice / decode kernels."""
from __future__ import annotations

from typing import List, Sequence

_VOCAB = {"<sync>": 0, "<sep>": 1, **{str(i): i + 2 for i in range(10)}, ".": 12}


class _NumberTokenizer:
    """decode(int) / get_vocab() subset of the reference's PreTrainedTokenizer subclasses
    (used by the drivers' output parsers: evaluate.py:395,408)."""

    number_format = "0>6.1f"

    def __init__(self):
        self.vocab = dict(_VOCAB)
        self.ids_to_tokens = {v: k for k, v in self.vocab.items()}

    def get_vocab(self):
        return self.vocab

    def get_vocab_size(self):
        return len(self.vocab)

    def decode(self, token_ids, **kwargs) -> str:
        if isinstance(token_ids, (list, tuple)):
            return "".join(self.ids_to_tokens[int(i)] for i in token_ids)
        return self.ids_to_tokens[int(token_ids)]

    def encode_numbers(self, values: Sequence[float]) -> List[int]:
        """TimeTower.encode / ScoreTower.encode: numbers joined by <sep>, closed by <sync>."""
        ids: List[int] = []
        for i, v in enumerate(values):
            if i:
                ids.append(self.vocab["<sep>"])
            ids.extend(self.vocab[ch] for ch in format(v, self.number_format))
        ids.append(self.vocab["<sync>"])
        return ids


class TimeTokenizer(_NumberTokenizer):
    number_format = "0>6.1f"   # time_encoder.py:58


class ScoreTokenizer(_NumberTokenizer):
    number_format = "0>3.1f"   # score_encoder.py:58


def frame_time_ids(video_timestamps: Sequence[Sequence[float]]) -> List[List[int]]:
    """Per frame: the 6 ids of format(t,'0>6.1f') (the trailing <sync> is dropped, trace_arch.py:243)."""
    tk = TimeTokenizer()
    rows = [tk.encode_numbers(ts)[:-1] for ts in video_timestamps]
    if any(len(r) != len(rows[0]) for r in rows):
        raise AssertionError(f"time token rows differ in length (timestamp >= 10000 s?): {[len(r) for r in rows]}")  # trace_arch.py:285
    return rows
