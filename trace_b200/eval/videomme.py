"""Video-MME multiple-choice evaluation (mirrors trace/eval/videomme/evaluate.py:156-567) on the B200 engine -- BASELINE
config 5 (mixed frame counts, MCQ decode with max_new_tokens=128).

Same as the reference: annotation JSON = list of {"url": "...watch?v=ID", "duration", "question", "options": ["A. ...", ...],
"answer": "B", ...} with the video at `{data_dir}/data/ID.mp4`; segment-centre frame sampler (`get_index`, :214-227);
shortest-edge bicubic resize + centre crop + CLIP normalisation WITHOUT pad-to-square (:176-187); prompt = system text +
"Question: ...\\nOptions:\\n(A) ...\\n..." + "\\nOnly give the best option."; greedy, `heads=[1]`; the option letter is the
first match of `\\(*\\s*([a-z])\\s*[\\).]` in the first caption (:439-441); accuracy per duration category written to
`test.json` / `upload_leaderboard.json` (:537-558).

Different on purpose: questions are decoded in batches of <= 8 through `BatchScheduler`, and all questions of one video share
ONE clip encode (the reference re-runs CLIP per question); a failing sample is skipped with a message (the reference swallows
every exception with a bare `except: continue`, :533)."""
from __future__ import annotations

import argparse
import json
import os
import re
from pathlib import Path

import numpy as np
import torch

from ..constants import DEFAULT_MMODAL_TOKEN
from ..conversation import conv_templates
from ..mm_utils import get_model_name_from_path, tokenizer_MMODAL_token_all
from ..model.builder import load_pretrained_model
from ..serving import BatchScheduler, ClipRequest

SYSTEM = ("Carefully watch the video and pay attention to the cause and sequence of events, the detail and movement of objects, "
          "and the action and pose of persons. Based on your observations, select the best option that accurately addresses the "
          "question.\n")
QUESTION_PROMPT = "\nOnly give the best option."


def get_index(num_segments, fps, max_frame, bound=None, first_idx=0):
    """Centre of each of `num_segments` equal segments (videomme/evaluate.py:214-227)."""
    start, end = (bound[0], bound[1]) if bound else (-100000, 100000)
    start_idx = max(first_idx, round(start * fps))
    end_idx = min(round(end * fps), max_frame)
    seg_size = float(end_idx - start_idx) / num_segments
    idx = np.array([int(start_idx + (seg_size / 2) + np.round(seg_size * i)) for i in range(num_segments)])
    return idx, [[f / fps] for f in idx]


def qa_template(data):
    """(question text, "(X) answer text") -- videomme/evaluate.py:262-271."""
    q = f"Question: {data['question']}\n" + "Options:\n"
    ans = data["answer"]
    ans = f"({ans}) {data['options'][ord(ans) - ord('A')][3:]}"
    for c in data["options"]:
        q += f"({c[0]}) {c[3:]}\n"
    return q.rstrip(), ans


def extract_option(message: str) -> str:
    return re.findall(r"\(*\s*([a-z])\s*[\).]", message.lower())[0]


def read_video(path, num_segments, processor, device=None):
    """Frames at the segment centres, CLIP preprocessing without padding -> (Tensor[T,3,S,S], [[seconds]]*T)."""
    import cv2
    from PIL import Image
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise FileNotFoundError(path)
    max_frame = int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) - 1
    fps = float(cap.get(cv2.CAP_PROP_FPS))
    idx, msg = get_index(num_segments, fps, max_frame)
    frames = []
    for i in idx:
        cap.set(cv2.CAP_PROP_POS_FRAMES, int(i))
        ok, bgr = cap.read()
        if not ok:
            raise RuntimeError(f"cannot decode frame {i} of {path}")
        frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    cap.release()
    if device is not None:
        from ..preprocess import preprocess_frames_cuda
        size = processor.size if isinstance(processor.size, int) else (processor.size.get("shortest_edge") or processor.size.get("height"))
        batch = torch.from_numpy(np.stack(frames)).to(device)
        return preprocess_frames_cuda(batch.contiguous(), size, tuple(processor.image_mean), tuple(processor.image_std), "none"), msg
    return processor.preprocess([Image.fromarray(f) for f in frames], return_tensors="pt")["pixel_values"], msg


def build_prompt_ids(question: str, tokenizer) -> torch.Tensor:
    conv = conv_templates["llama_2"].copy()
    conv.append_message(conv.roles[0], DEFAULT_MMODAL_TOKEN["VIDEO"] + "\n" + SYSTEM + question + QUESTION_PROMPT)
    conv.append_message(conv.roles[1], None)
    return tokenizer_MMODAL_token_all(conv.get_prompt() + "<sync>", tokenizer, return_tensors="pt")


def main(args):
    tokenizer, model, processor, _ = load_pretrained_model(args.model_path, None, get_model_name_from_path(args.model_path))
    with open(args.anno_path, "r") as f:
        data_list = json.load(f)
    if args.debug:
        data_list = data_list[:10]
    Path(args.output_dir).mkdir(parents=True, exist_ok=True)
    sched = BatchScheduler(model, eos_token_id=tokenizer.eos_token_id, pad_token_id=tokenizer.eos_token_id)
    vocab = getattr(model, "vocab_size", 32000)
    dev = "cuda" if args.device_preprocess else None
    acc_dict, res_list = {}, []
    clips = {}                                                        # video path -> (tensor, timestamps): one decode + encode per video
    window = max(1, args.window)
    for w0 in range(0, len(data_list), window):
        reqs = []
        for k, data in enumerate(data_list[w0: w0 + window]):
            name = data["url"].split("watch?v=")[1]
            path = os.path.join(args.data_dir, "data", name + ".mp4")
            try:
                if path not in clips:
                    tensor, msg = read_video(path, args.num_frames, processor, dev)
                    clips[path] = (tensor.to(torch.float16), msg)
                tensor, msg = clips[path]
                question, answer = qa_template(data)
                reqs.append(ClipRequest(build_prompt_ids(question, tokenizer), tensor, msg, head=1, max_new_tokens=128,
                                        tag=(w0 + k, data["duration"], answer)))
            except Exception as ex:  # noqa: BLE001
                print(f"Error in {path}: {ex!r}")
        for r in sched.run(reqs):
            idx, cat, answer = r.tag
            ev = r.events(lambda toks: tokenizer.decode(toks, skip_special_tokens=True), vocab)
            acc = acc_dict.setdefault(cat, [0, 0])                     # correct, total
            acc[1] += 1
            try:
                pred = extract_option(ev["captions"][0])
            except Exception:  # noqa: BLE001  (no caption / no option letter: counted as wrong)
                pred = ""
            gt = answer[1].lower()
            acc[0] += int(pred == gt)
            res_list.append({"pred": [pred], "gt": [gt]})
            data_list[idx]["response"] = pred
            print(f"Pred: {pred}\nGT: {answer[1]}", flush=True)
        clips = {p: v for p, v in clips.items() if any(p.endswith(d["url"].split("watch?v=")[1] + ".mp4") for d in data_list[w0 + window: w0 + 2 * window])}
        with open(os.path.join(args.output_dir, "test.json"), "w") as f:
            json.dump({"acc_dict": acc_dict, "res_list": res_list}, f)
        final_res = {k: v[0] / v[1] * 100 for k, v in acc_dict.items()}
        tot = sum(v[1] for v in acc_dict.values())
        final_res["Avg"] = sum(v[0] for v in acc_dict.values()) / max(tot, 1) * 100
        print(final_res)
        with open(os.path.join(args.output_dir, "upload_leaderboard.json"), "w+") as f:
            json.dump(final_res, f)
    return acc_dict, res_list, sched.stats


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--anno_path", type=str, default="Video-MME/test.json")
    p.add_argument("--data_dir", type=str, default="Video-MME")
    p.add_argument("--output_dir", default="debug")
    p.add_argument("--num_frames", type=int, default=8)
    p.add_argument("--gpu_id", default="0")
    p.add_argument("--debug", action="store_true")
    p.add_argument("--model_path", default="")
    p.add_argument("--window", type=int, default=32, help="questions scheduled together (bucketed into decode batches of <= 8)")
    p.add_argument("--device_preprocess", action="store_true")
    for ignored in ("--cfg_path", "--video_path", "--model_type"):
        p.add_argument(ignored, type=str, default="")
    p.add_argument("--top_p", type=float, default=0.8)
    p.add_argument("--temperature", type=float, default=1)
    return p


if __name__ == "__main__":
    main(build_parser().parse_args())
