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


class ContinuousBatcher:
    """In-flight batching over the engine's `trace_b200_batch_*` calls.

    `BatchScheduler` above forms static buckets: rows of a bucket must share (T, L, max_new) and a bucket decodes until its
    slowest row ends.  Here ONE decode batch of `rows` rows stays open; a request is prefilled on its own and admitted into
    whatever row is free, rows of different clip length / prompt length / head / budget share every decode step (the decode
    kernel keeps position, head, emitted tokens and budget per row on the device), and a row that hits EOS or its budget is read
    out and refilled between bursts of `burst` steps.  Every distinct clip is encoded once.

    Overlapping the ViT of the next clip with the decode of the current one (the other half of SURVEY §8(f)-4) buys nothing on
    this design: the decode step is a cooperative kernel that owns all 148 SMs and 219 KB of each SM's shared memory, so kernels
    of another stream can only run between two steps -- time slicing, not overlap.  What does pay is keeping the 8 batch columns
    of every weight-streaming MMA occupied, which is what this class does."""

    def __init__(self, model, rows: int = 8, burst: int = 16, eos_token_id: Optional[int] = None, pad_token_id: Optional[int] = None):
        assert 1 <= rows <= 8 and burst >= 1
        self.model, self.rows, self.burst = model, rows, burst
        self.eos = eos_token_id if eos_token_id is not None else model.generation_config.eos_token_id
        self.pad = pad_token_id if pad_token_id is not None else self.eos
        self.stats = {"clips_encoded": 0, "admitted": 0, "bursts": 0, "row_steps": 0, "useful_row_steps": 0}

    @torch.no_grad()
    def run(self, requests: Sequence[ClipRequest]) -> List[ClipResult]:
        model, eng = self.model, self.model.engine
        dev = model.device
        results: List[Optional[ClipResult]] = [None] * len(requests)
        slot_cache: Dict[int, torch.Tensor] = {}
        pending = list(range(len(requests)))
        active: Dict[int, Tuple[int, int]] = {}                # row -> (request index, KV slot)
        rows = min(self.rows, max(1, len(requests)))
        eng.batch_open(rows, self.eos, self.pad)
        try:
            while pending or active:
                # ---- admit into free rows ----
                for row in range(rows):
                    if row in active or not pending:
                        continue
                    i = pending.pop(0)
                    r = requests[i]
                    if id(r.clip) not in slot_cache:
                        slot_cache[id(r.clip)] = model._encode_slots([r.clip])[0]
                        self.stats["clips_encoded"] += 1
                    ids = r.input_ids.to(dev, torch.int64).view(1, -1).contiguous()
                    tid = torch.tensor([frame_time_ids(r.video_timestamps)], dtype=torch.int32, device=dev)
                    emb = eng.build_prefill(ids, slot_cache[id(r.clip)][None].contiguous(), tid)
                    S = emb.shape[1]
                    if S + r.max_new_tokens > eng.max_ctx:
                        raise RuntimeError(f"prompt ({S}) + max_new_tokens ({r.max_new_tokens}) exceeds the engine context ({eng.max_ctx})")
                    slot = eng.alloc_slot(S + r.max_new_tokens)
                    try:
                        eng.batch_admit(row, slot, emb, -1 if r.head is None else int(r.head), int(r.max_new_tokens))
                    except Exception:
                        eng.free_slot(slot)
                        raise
                    active[row] = (i, slot)
                    self.stats["admitted"] += 1
                # ---- one burst of decode steps for every row ----
                steps, fin = eng.batch_step(self.burst)
                self.stats["bursts"] += 1
                self.stats["row_steps"] += self.burst * rows
                self.stats["useful_row_steps"] += self.burst * len(active)
                # ---- read out and retire rows that hit EOS or used up their budget ----
                for row in list(active):
                    i, slot = active[row]
                    if fin[row] or steps[row] >= requests[i].max_new_tokens:
                        ids_row, head = eng.batch_read(row, int(requests[i].max_new_tokens))
                        eng.batch_retire(row)
                        eng.free_slot(slot)
                        results[i] = ClipResult(trim_after_eos(ids_row, self.eos), None if requests[i].head is None else head, requests[i].tag)
                        del active[row]
        finally:
            for row, (i, slot) in list(active.items()):        # an exception mid-run: leave the engine usable
                eng.batch_retire(row)
                eng.free_slot(slot)
            eng.batch_close()
        return results  # type: ignore[return-value]
