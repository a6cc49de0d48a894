"""Multi-GPU: one process per GPU (torch.distributed, NCCL over NVLink 5 / NVSwitch).

The only exchange step of the path: frames are independent through CLIP + projector, so every
clip's T frames are split contiguously over the G ranks; each rank encodes its B*T/G frames and ONE
all-gather of the projected slot tokens ([B, T/G, 8, hidden] fp16 per rank = 1 MiB per clip per rank at
T/G=16) gives every rank the full [B, T, 8, hidden].  Gathering after the projector moves 18x fewer bytes than
gathering CLIP features.  The LLM (prefill + decode) is replicated; rank r decodes clips r, r+G, ...
No collective inside the decode loop."""
from __future__ import annotations

import torch
import torch.distributed as dist


def attach_group(model, group=None):
    """Make `model.generate` shard its ViT encode over `group` (default: the world group)."""
    model.dist_group = group if group is not None else dist.group.WORLD
    return model


def shard_bounds(T: int, world: int, rank: int):
    """Contiguous frame range of `rank` (rank-major == frame-major so the gathered buffer is in t order)."""
    assert T % world == 0, f"frames per clip ({T}) must divide evenly over {world} ranks"
    per = T // world
    return rank * per, (rank + 1) * per


def encode_frames_sharded(engine, frames: torch.Tensor, B: int, T: int, group, presliced: bool = False) -> torch.Tensor:
    """frames [B*T,3,H,W] (every rank holds all clips) -> slots [B*T, 8, hidden] on every rank.
    `presliced`: `frames` already is this rank's slice [B*(T/G),3,H,W] (frames [r*T/G, (r+1)*T/G) of every clip) -- a rank
    only ever needs its own frames, so a caller that loads clips from the host should upload just those."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(T, world, rank)
    if presliced:
        assert frames.shape[0] == B * (hi - lo), f"presliced frames: expected {B * (hi - lo)} frames, got {frames.shape[0]}"
        fr = frames.contiguous()
    else:
        fr = frames.view(B, T, *frames.shape[1:])[:, lo:hi].reshape(B * (hi - lo), *frames.shape[1:]).contiguous()
    local = engine.encode_frames(fr)                                   # [B*(T/G), 8, H]
    local = local.view(B, hi - lo, local.shape[1], local.shape[2])
    gathered = torch.empty((world * B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)   # concatenation along dim 0, rank-major
    gathered = gathered.view(world, B, *local.shape[1:])
    # [G, B, T/G, 8, H] -> [B, G, T/G, 8, H] -> [B*T, 8, H]
    return gathered.permute(1, 0, 2, 3, 4).reshape(B * T, local.shape[2], local.shape[3]).contiguous()


def clips_of_rank(n_clips: int, world: int, rank: int):
    """Config 4 of BASELINE.json: after the all-gather every rank holds every clip's slot tokens and decodes clips
    r, r+G, r+2G, ... with its full replica of the LLM."""
    return list(range(rank, n_clips, world))


@torch.no_grad()
def generate_sharded(model, input_ids: torch.Tensor, clips, video_timestamps, heads, max_new_tokens: int, group=None,
                     eos_token_id=None, pad_token_id=None, local_slices: bool = False):
    """Multi-GPU generate for a batch of B clips (same T, same prompt length):
      1. every rank encodes frames [r*T/G, (r+1)*T/G) of ALL clips, one all-gather of the projected slot tokens rebuilds
         [B,T,8,H] everywhere (`encode_frames_sharded`);
      2. rank r splices / prefills / decodes clips r, r+G, ... (<= 8 rows per decode loop) -- no collective in the loop;
      3. the new ids and the final heads are exchanged once (tiny object all-gather) so that every rank returns the full
         [B, n] tensor in clip order, padded with `pad_token_id` like a batched `generate`, and `heads` is updated in place.
    `local_slices`: `clips[i]` holds only this rank's frames of clip i ([T/G,3,H,W], frames shard_bounds(T, G, rank));
    T is then taken from `video_timestamps`.
    `model` needs `_encode_slots`, `_generate_from_slots`, `device`, `generation_config` (TraceMistralForCausalLM)."""
    from .model.encoders import frame_time_ids
    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = input_ids.shape[0]
    eos = eos_token_id if eos_token_id is not None else model.generation_config.eos_token_id
    pad = pad_token_id if pad_token_id is not None else eos
    prev = model.dist_group
    model.dist_group = group
    try:
        slots = model._encode_slots(list(clips), presliced_T=len(video_timestamps[0]) if local_slices else None)   # [B,T,8,H] on every rank
    finally:
        model.dist_group = prev
    mine = clips_of_rank(B, world, rank)
    local = None
    local_heads = None if heads is None else [int(heads[i]) for i in mine]
    if mine:
        idx = torch.tensor(mine, device=slots.device)
        ids = input_ids.to(model.device, torch.int64)[idx].contiguous()
        tid = torch.tensor([frame_time_ids(video_timestamps[i]) for i in mine], dtype=torch.int32, device=model.device)
        local = model._generate_from_slots(ids, slots[idx].contiguous(), tid, local_heads, max_new_tokens, eos, pad)
    parts = [None] * world
    dist.all_gather_object(parts, (mine, None if local is None else local.cpu(), local_heads), group=group)
    n = max((p[1].shape[1] for p in parts if p[1] is not None), default=0)
    out = torch.full((B, n), pad, dtype=torch.int64)
    for idxs, rows, hd in parts:
        for k, i in enumerate(idxs):
            out[i, : rows.shape[1]] = rows[k]
            if heads is not None:
                heads[i] = hd[k]
    return out
