"""Host logic of the batch scheduler and the reformatters (no GPU)."""
import json
import os
import runpy
import sys

import pytest
import torch

from trace_b200.eval import reformat_dvc, reformat_tvg, reformat_vhd
from trace_b200.serving import ClipRequest, bucket_key, plan_batches, trim_after_eos

REF = "/root/reference/trace/eval"


def _req(T, L, max_new=16, head=1):
    return ClipRequest(torch.zeros(L, dtype=torch.long), torch.zeros(T, 3, 2, 2), [[0.0]] * T, head, max_new)


def test_plan_batches_buckets_and_caps():
    reqs = [_req(128 if i % 2 else 64, 40 + (i % 3 == 0)) for i in range(37)]
    plan = plan_batches(reqs, max_batch=8)
    seen = sorted(sum(plan, []))
    assert seen == list(range(37))
    for b in plan:
        assert 1 <= len(b) <= 8
        assert len({bucket_key(reqs[i]) for i in b}) == 1
        assert b == sorted(b)                      # arrival order inside a batch
    # text-head-only requests never share a batch with head-switching ones
    mixed = [_req(8, 10, head=None), _req(8, 10, head=1)]
    assert plan_batches(mixed) == [[0], [1]]


def test_trim_after_eos():
    assert trim_after_eos(torch.tensor([7, 2, 2, 2]), 2).tolist() == [7, 2]
    assert trim_after_eos(torch.tensor([7, 8]), 2).tolist() == [7, 8]
    assert trim_after_eos(torch.tensor([2]), 2).tolist() == [2]


PRED = [
    {"video": "a/b/v1.mp4", "id": 3, "timestamps": [[1.0, 5.5], [7.0], [8.0, 9.0]], "scores": [[4.5], [3.0], [2.0]],
     "captions": ["cut onion", "bad", "fry"]},
    {"video": "v2.mp4", "id": 4, "timestamps": [[3.0]], "scores": [], "captions": ["only"]},
    {"video": "c/v3.mp4", "id": 5, "timestamps": [[2.0], [6.1], [50.0], [7.9]], "scores": [[4.0], [2.0], [1.0], [0.0], [9.0]], "captions": ["q"]},
]
GT = {"annotations": [{"image_id": "v1.mp4", "duration": 12.0}, {"image_id": "v2.mp4", "duration": 9.0}, {"image_id": "v3.mp4", "duration": 10.0}]}


def test_reformatters_known_answers():
    dvc = reformat_dvc(PRED)
    assert dvc["v1.mp4"] == [{"caption": "cut onion", "timestamp": [1.0, 5.5]}, {"caption": "fry", "timestamp": [8.0, 9.0]}]
    assert dvc["v2.mp4"] == [] and dvc["v3.mp4"] == []
    tvg = reformat_tvg(PRED)
    assert tvg[3] == {"query": "cut onion", "timestamp": [1.0, 5.5], "vid": "v1.mp4"} and tvg[4] == {} and tvg[5] == {}
    vhd = reformat_vhd(PRED, GT["annotations"])
    assert vhd[0]["pred_saliency_scores"] == [4.5, 0.0, 3.0, 2.0, 0.0, 0.0]       # clip = max(0, int(t/2) - 1)
    assert vhd[1]["pred_saliency_scores"] == [0.0, 0.0, 0.0, 0.0]
    assert vhd[2]["pred_saliency_scores"] == [4.0, 0.0, 1.0, 0.0, 0.0]            # t=6.1 and t=7.9 -> clip 2: (2 + 0) / 2; t=50 > duration dropped; extra score cut


def test_vhd_short_scores_fail_like_the_reference():
    """reformat_vhd.py:73 pads missing scores with bare 0.0 and then calls len() on them (:33): a record with fewer
    score lists than timestamp lists raises TypeError in the reference; kept (same error behaviour)."""
    bad = [{"video": "v1.mp4", "id": 1, "timestamps": [[1.0], [3.0]], "scores": [[2.0]], "captions": ["q"]}]
    with pytest.raises(TypeError):
        reformat_vhd(bad, GT["annotations"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("task", ["dvc", "tvg", "vhd"])
def test_reformatters_equal_reference_scripts(task, tmp_path, capsys):
    """Run the reference's own CLI scripts on the same prediction file and compare the JSON they write."""
    pred, out, gt = tmp_path / "pred.json", tmp_path / "out.json", tmp_path / "gt.json"
    pred.write_text(json.dumps(PRED))
    gt.write_text(json.dumps(GT))
    argv = [f"reformat_{task}.py", "--pred_file", str(pred), "--out_file", str(out)] + (["--gtpath", str(gt)] if task == "vhd" else [])
    old = sys.argv
    try:
        sys.argv = argv
        runpy.run_path(os.path.join(REF, f"reformat_{task}.py"), run_name="__main__")
    finally:
        sys.argv = old
    capsys.readouterr()
    want = json.loads(out.read_text())
    got = {"dvc": reformat_dvc, "tvg": reformat_tvg}[task](PRED) if task != "vhd" else reformat_vhd(PRED, GT["annotations"])
    assert json.loads(json.dumps(got)) == want


def test_continuous_batcher_host_logic():
    """In-flight admission with a recorder engine: rows are refilled as soon as they end, every request is read out exactly
    once with its own budget, clips are encoded once, KV slots are returned, the batch is closed even when a request fails."""
    import torch
    from trace_b200.serving import ClipRequest, ContinuousBatcher

    class FakeEngine:
        max_ctx = 512

        def __init__(self):
            self.log, self.rows, self.free, self.next_slot = [], {}, [], 0
            self.open = False

        def build_prefill(self, ids, slots, tid):
            return torch.zeros(1, ids.shape[1] - 1 + 14 * tid.shape[1], 8, dtype=torch.float16)

        def alloc_slot(self, n):
            self.next_slot += 1
            self.log.append(("alloc", self.next_slot, n))
            return self.next_slot

        def free_slot(self, s):
            self.log.append(("free", s))

        def batch_open(self, rows, eos, pad):
            assert not self.open
            self.open, self.nrows = True, rows

        def batch_admit(self, row, slot, emb, head, max_new):
            assert row not in self.rows and 0 <= row < self.nrows
            if max_new == 13:
                raise RuntimeError("engine failure")
            self.rows[row] = {"slot": slot, "head": head, "budget": max_new, "emitted": 1, "eos_at": 5 if max_new % 2 else 10 ** 9, "S": emb.shape[1]}
            self.log.append(("admit", row, slot, emb.shape[1], head, max_new))

        def batch_step(self, n):
            st, fin = [0] * self.nrows, [1] * self.nrows
            for r, x in self.rows.items():
                x["emitted"] += n
                st[r], fin[r] = x["emitted"], int(x["emitted"] >= x["eos_at"])
            self.log.append(("step", n, sorted(self.rows)))
            return st, fin

        def batch_read(self, row, cap):
            x = self.rows[row]
            n = min(x["emitted"], x["budget"], cap)
            ids = [100 * x["slot"] + k for k in range(n)]
            if x["eos_at"] <= n:
                ids[x["eos_at"] - 1] = 2
            return torch.tensor(ids, dtype=torch.int64), (x["head"] + 1) % 3 if x["head"] >= 0 else -1

        def batch_retire(self, row):
            self.log.append(("retire", row))
            del self.rows[row]

        def batch_close(self):
            assert self.open and not self.rows
            self.open = False

    class FakeModel:
        device = torch.device("cpu")

        def __init__(self):
            import types
            self.engine = FakeEngine()
            self.generation_config = types.SimpleNamespace(eos_token_id=2)
            self.encodes = 0

        def _encode_slots(self, clips):
            self.encodes += len(clips)
            return torch.zeros(len(clips), clips[0].shape[0], 8, 8, dtype=torch.float16)

    clip_a, clip_b = torch.zeros(4, 3, 2, 2), torch.zeros(8, 3, 2, 2)
    ts4, ts8 = [[float(i)] for i in range(4)], [[float(i)] for i in range(8)]
    budgets = [8, 7, 20, 6, 9, 12]
    reqs = [ClipRequest(torch.tensor([1, 5, -201, 7 + k, -205]), clip_a if k % 2 == 0 else clip_b, ts4 if k % 2 == 0 else ts8, head=1 if k != 3 else None,
                        max_new_tokens=b, tag=f"r{k}") for k, b in enumerate(budgets)]
    m = FakeModel()
    sched = ContinuousBatcher(m, rows=2, burst=4)
    res = sched.run(reqs)
    eng = m.engine
    assert [r.tag for r in res] == [f"r{k}" for k in range(6)] and m.encodes == 2 and not eng.open
    for k, r in enumerate(res):
        assert 1 <= r.ids.numel() <= budgets[k]
        if budgets[k] % 2:                                   # odd budgets hit EOS at token 5: trimmed right after it
            assert r.ids.tolist()[-1] == 2 and r.ids.numel() == 5
        else:
            assert r.ids.numel() == budgets[k]
    assert res[3].head is None and res[0].head == 2
    admits = [e for e in eng.log if e[0] == "admit"]
    assert [a[3] for a in admits] == [4 + 14 * (4 if k % 2 == 0 else 8) for k in range(6)]        # each request prefilled on its own length
    assert sorted(e[1] for e in eng.log if e[0] == "free") == sorted(e[1] for e in eng.log if e[0] == "alloc")
    # never more rows than the batch has, and a freed row is refilled before the next burst
    for e in eng.log:
        if e[0] == "step":
            assert 1 <= len(e[2]) <= 2
    assert sched.stats["admitted"] == 6 and sched.stats["useful_row_steps"] <= sched.stats["row_steps"]
    # a failing admission closes the batch and returns the slots of the rows that were running
    m2 = FakeModel()
    bad = [reqs[0], ClipRequest(reqs[1].input_ids, clip_b, ts8, head=1, max_new_tokens=13)]
    with pytest.raises(RuntimeError):
        ContinuousBatcher(m2, rows=2, burst=4).run(bad)
    assert not m2.engine.open and not m2.engine.rows
    assert sorted(e[1] for e in m2.engine.log if e[0] == "free") == sorted(e[1] for e in m2.engine.log if e[0] == "alloc")
