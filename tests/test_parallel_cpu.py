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
