"""Prediction-file reformatters (trace/eval/reformat_dvc.py:14-23, reformat_tvg.py:14-25, reformat_vhd.py:17-85).

The reference scripts are file-to-file CLIs; here each one is a pure function over the list of result records the
eval driver saves (`{'video','id','timestamps','scores','captions'}`, evaluate.py:360-411) plus the same CLI
(`python -m trace_b200.eval.reformat dvc|tvg|vhd --pred_file ... --out_file ...`).  Behaviours kept on purpose:
events whose timestamp list is not exactly [start, end] are dropped (dvc, tvg); tvg keeps only the FIRST valid
event of a record and yields {} when there is none; vhd maps (t, score) pairs to 2-second clips with
clip = max(0, int(t/2) - 1), averages duplicates, ignores t > duration and pads / truncates scores to the
number of timestamps."""
from __future__ import annotations

import argparse
import json
from typing import Dict, List, Sequence


def result_record(video: str, qid, events: Dict[str, list]) -> Dict[str, object]:
    """One row of the eval driver's result file (evaluate.py:360-372 + the parsed events)."""
    return {"video": video, "id": qid, "timestamps": events["timestamps"], "scores": events["scores"],
            "captions": events["captions"]}


def _video_name(item) -> str:
    return str(item["video"]).split("/")[-1]


def reformat_dvc(pred_data: Sequence[dict]) -> Dict[str, List[dict]]:
    out: Dict[str, List[dict]] = {}
    for item in pred_data:
        out[_video_name(item)] = [{"caption": cap, "timestamp": ts}
                                  for ts, cap in zip(item["timestamps"], item["captions"]) if len(ts) == 2]
    return out


def reformat_tvg(pred_data: Sequence[dict]) -> Dict[object, dict]:
    out: Dict[object, dict] = {}
    for item in pred_data:
        first = {}
        for ts, cap in zip(item["timestamps"], item["captions"]):
            if len(ts) == 2:
                first = {"query": cap, "timestamp": ts, "vid": _video_name(item)}
                break
        out[item["id"]] = first
    return out


def vhd_clip_scores(timestamps: Sequence[Sequence[float]], scores: Sequence[Sequence[float]], duration: float) -> List[float]:
    n = int(duration / 2)
    tot = [0.0] * n
    cnt = [0] * n
    for ts, sc in zip(timestamps, scores):
        if len(ts) == 0 or len(sc) == 0:
            continue
        t, s = ts[0], sc[0]
        if t > duration:
            continue
        cid = max(0, int(t / 2) - 1)
        tot[cid] += s
        cnt[cid] += 1
    return [tot[i] / cnt[i] if cnt[i] else 0.0 for i in range(n)]


def reformat_vhd(pred_data: Sequence[dict], gt_annotations: Sequence[dict]) -> List[dict]:
    by_vid = {a["image_id"]: a for a in gt_annotations}
    out = []
    for item in pred_data:
        vid = _video_name(item)
        ts = list(item["timestamps"]) if item["scores"] and item["timestamps"] else []
        sc = list(item["scores"]) if item["scores"] and item["timestamps"] else []
        sc = sc + [0.0] * (len(ts) - len(sc)) if len(sc) < len(ts) else sc[: len(ts)]
        out.append({"query": item["captions"][0], "vid": vid, "qid": item["id"],
                    "pred_saliency_scores": vhd_clip_scores(ts, sc, by_vid[vid]["duration"])})
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("task", choices=["dvc", "tvg", "vhd"])
    ap.add_argument("--pred_file", required=True)
    ap.add_argument("--out_file", required=True)
    ap.add_argument("--gtpath", default=None)
    a = ap.parse_args(argv)
    pred = json.load(open(a.pred_file))
    if a.task == "dvc":
        res = reformat_dvc(pred)
    elif a.task == "tvg":
        res = reformat_tvg(pred)
    else:
        res = reformat_vhd(pred, json.load(open(a.gtpath))["annotations"])
    json.dump(res, open(a.out_file, "w+"))


if __name__ == "__main__":
    main()
