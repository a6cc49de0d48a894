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
