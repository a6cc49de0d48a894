"""world_size-2 gloo test of the N>1 host logic (frame sharding + one all-gather, rank-major == frame-major)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEngine:
    """Stands in for the CUDA engine: slot token = f(frame id), so ordering bugs are visible."""

    def encode_frames(self, frames):
        ids = frames.reshape(frames.shape[0], -1)[:, 0]
        return (ids[:, None, None] * 10 + torch.arange(8)[None, :, None]).expand(-1, -1, 4).float().contiguous()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trace_b200.parallel import encode_frames_sharded, shard_bounds
    B, T = 3, 8
    frames = torch.arange(B * T, dtype=torch.float32).view(B * T, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()
    out = encode_frames_sharded(_FakeEngine(), frames, B, T, dist.group.WORLD)
    want = _FakeEngine().encode_frames(frames)
    ok = torch.equal(out, want) and shard_bounds(T, world, rank) == (rank * T // world, (rank + 1) * T // world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_encode_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_reject_ragged():
    from trace_b200.parallel import shard_bounds
    with pytest.raises(AssertionError):
        shard_bounds(10, 4, 0)


class _FakeModel:
    """Host-logic stand-in for TraceMistralForCausalLM: slots = sharded fake encode, 'decode' = a pure function of the row."""
    device = torch.device("cpu")
    dist_group = None

    def __init__(self):
        import types
        self.generation_config = types.SimpleNamespace(eos_token_id=2, pad_token_id=None)
        self.calls = []

    def _encode_slots(self, clips, presliced_T=None):
        from trace_b200.parallel import encode_frames_sharded
        B, T = len(clips), clips[0].shape[0] if presliced_T is None else presliced_T
        frames = torch.cat(clips, 0)
        out = (encode_frames_sharded(_FakeEngine(), frames, B, T, self.dist_group, presliced=presliced_T is not None)
               if self.dist_group is not None else _FakeEngine().encode_frames(frames))
        return out.view(B, T, 8, -1)

    def _generate_from_slots(self, ids, slots, tid, heads, max_new, eos, pad):
        self.calls.append(ids.shape[0])
        rows = []
        for b in range(ids.shape[0]):
            seed = int(slots[b].sum().item()) % 7 + 1                     # depends on ALL frames of the clip
            n = min(max_new, seed + 1)
            row = [int(ids[b, 0]) * 100 + seed + k for k in range(n - 1)] + [eos]
            rows.append(row + [pad] * (max_new - n))
            if heads is not None:
                heads[b] = (heads[b] + seed) % 3
        return torch.tensor(rows, dtype=torch.int64)


def _worker_generate(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trace_b200.parallel import clips_of_rank, generate_sharded
    B, T = 5, 8
    clips = [torch.full((T, 3, 2, 2), float(b)) + torch.arange(T).view(T, 1, 1, 1) * 0.125 for b in range(B)]
    ts = [[[float(i)] for i in range(T)] for _ in range(B)]
    ids = torch.arange(1, B + 1).view(B, 1).repeat(1, 4)
    heads = [1] * B
    m = _FakeModel()
    out = generate_sharded(m, ids, clips, ts, heads, 6, dist.group.WORLD, pad_token_id=9)
    # single-process reference: same fake model without a group
    ref_m = _FakeModel()
    ref_heads = [1] * B
    slots = ref_m._encode_slots(clips)
    ref = ref_m._generate_from_slots(ids, slots, None, ref_heads, 6, 2, 9)
    n = out.shape[1]
    ok = torch.equal(out, ref[:, :n]) and bool((ref[:, n:] == 9).all()) and heads == ref_heads
    ok = ok and m.calls == [len(clips_of_rank(B, world, rank))]
    # the same call with every rank holding only ITS frames of every clip (what a loader that reads from the host uploads)
    from trace_b200.parallel import shard_bounds
    lo, hi = shard_bounds(T, world, rank)
    heads2 = [1] * B
    out2 = generate_sharded(_FakeModel(), ids, [c[lo:hi] for c in clips], ts, heads2, 6, dist.group.WORLD, pad_token_id=9, local_slices=True)
    ok = ok and torch.equal(out2, out) and heads2 == heads
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_generate_sharded_matches_single_rank():
    """Config 4 host logic: frame-sharded encode + all-gather, clips r, r+G, ... decoded per rank, ids / heads exchanged once."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_generate, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
