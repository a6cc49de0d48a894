"""Bucketed batch scheduler for the eval drivers (BASELINE configs 3 and 5).

The reference drivers are serial batch-1 (`evaluate.py:296-302`, `eval.sh:56`) and `generate` cannot batch clips of
different length at all (`torch.stack` at trace_arch.py:222; the unequal-prompt branch dereferences `labels=None`,
trace_arch.py:496).  On B200 one decode step streams the 14 GB of weights once for up to 8 rows, so batching is the
single biggest throughput lever.  This scheduler

  * buckets requests by (frames T, prompt length L, max_new_tokens, heads given?) -- rows of one bucket can share a
    prefill and a decode loop without padding, which the engine (like the reference) does not support;
  * encodes every distinct clip ONCE per `run()` (VideoMME / MVBench ask several questions about the same 128-frame
    clip and the reference re-runs CLIP for each, videomme/evaluate.py:348-383);
  * decodes each bucket in groups of <= 8 rows and returns the rows in request order, trimmed after the first EOS
    (rows of a batched `generate` are padded with `pad_token_id` up to the longest row of the group).

Everything numerical happens in `TraceMistralForCausalLM` / the CUDA engine; this file only reorders work."""
from __future__ import annotations

import dataclasses
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .events import parse_events
from .model.encoders import frame_time_ids


@dataclasses.dataclass
class ClipRequest:
    input_ids: torch.Tensor                     # LongTensor[L] with one -201 (<video>) and the trailing -205 (<sync>)
    clip: torch.Tensor                          # [T,3,H,W] CLIP-normalised frames (any float dtype, host or device)
    video_timestamps: Sequence[Sequence[float]]  # [[seconds]] * T, as returned by process_video
    head: Optional[int] = 1                     # initial head (evaluate.py:322); None = text head only
    max_new_tokens: int = 512
    tag: object = None                          # opaque, returned with the result


@dataclasses.dataclass
class ClipResult:
    ids: torch.Tensor                           # LongTensor[n] new tokens, trimmed after the first EOS
    head: Optional[int]                         # head after the last fed-back token (what the reference leaves in `heads`)
    tag: object = None

    def events(self, decode_text, vocab: int = 32000) -> Dict[str, list]:
        return parse_events(self.ids.tolist(), decode_text, vocab)


def bucket_key(r: ClipRequest) -> Tuple[int, int, int, bool]:
    return (int(r.clip.shape[0]), int(r.input_ids.numel()), int(r.max_new_tokens), r.head is None)


def plan_batches(requests: Sequence[ClipRequest], max_batch: int = 8) -> List[List[int]]:
    """Request indices grouped into decode batches: same bucket, arrival order kept inside a bucket, buckets in order
    of their first request.  Pure host logic (unit-tested on CPU)."""
    buckets: "OrderedDict[Tuple[int, int, int, bool], List[int]]" = OrderedDict()
    for i, r in enumerate(requests):
        buckets.setdefault(bucket_key(r), []).append(i)
    plan: List[List[int]] = []
    for idxs in buckets.values():
        for j in range(0, len(idxs), max_batch):
            plan.append(idxs[j: j + max_batch])
    return plan


def trim_after_eos(row: torch.Tensor, eos: int) -> torch.Tensor:
    hit = (row == eos).nonzero()
    return row[: int(hit[0]) + 1] if hit.numel() else row


class BatchScheduler:
    def __init__(self, model, max_batch: int = 8, eos_token_id: Optional[int] = None, pad_token_id: Optional[int] = None):
        assert 1 <= max_batch <= 8, "the decode step batches at most 8 rows (TRACE_B200_MAX_DECODE_BATCH)"
        self.model = model
        self.max_batch = max_batch
        self.eos = eos_token_id if eos_token_id is not None else model.generation_config.eos_token_id
        self.pad = pad_token_id if pad_token_id is not None else self.eos
        self.stats = {"clips_encoded": 0, "clip_encodes_saved": 0, "batches": 0, "rows": 0}

    @torch.no_grad()
    def run(self, requests: Sequence[ClipRequest]) -> List[ClipResult]:
        model = self.model
        dev = model.device
        results: List[Optional[ClipResult]] = [None] * len(requests)
        slot_cache: Dict[int, torch.Tensor] = {}                  # id(clip tensor) -> slots [T,8,H]
        for batch in plan_batches(requests, self.max_batch):
            rs = [requests[i] for i in batch]
            todo = [r for r in rs if id(r.clip) not in slot_cache]
            uniq: "OrderedDict[int, torch.Tensor]" = OrderedDict((id(r.clip), r.clip) for r in todo)
            if uniq:
                enc = model._encode_slots(list(uniq.values()))   # [U,T,8,H]; all clips of a bucket share T
                for k, key in enumerate(uniq):
                    slot_cache[key] = enc[k]
                self.stats["clips_encoded"] += len(uniq)
            self.stats["clip_encodes_saved"] += len(rs) - len(uniq)
            slots = torch.stack([slot_cache[id(r.clip)] for r in rs], 0).contiguous()
            ids = torch.stack([r.input_ids.to(dev, torch.int64) for r in rs], 0).contiguous()
            tid = torch.tensor([frame_time_ids(r.video_timestamps) for r in rs], dtype=torch.int32, device=dev)
            heads = None if rs[0].head is None else [int(r.head) for r in rs]
            out = model._generate_from_slots(ids, slots, tid, heads, int(rs[0].max_new_tokens), self.eos, self.pad)
            for k, i in enumerate(batch):
                results[i] = ClipResult(trim_after_eos(out[k], self.eos), None if heads is None else heads[k], requests[i].tag)
            self.stats["batches"] += 1
            self.stats["rows"] += len(rs)
        return results  # type: ignore[return-value]
