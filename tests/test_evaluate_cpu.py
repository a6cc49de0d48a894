"""The evaluation driver (trace_b200/eval/evaluate.py, mirroring trace/eval/evaluate.py) end to end on CPU with generated
videos and a recorder model: annotation loading, prompt construction, batching, event parsing, the result file the
reformat scripts read."""
import json
import os

import numpy as np
import torch

from trace_b200 import synth
from trace_b200.eval import evaluate as ev
from trace_b200.eval import reformat_dvc


def _write_video(path, n, fps=10.0, size=(64, 48)):
    import cv2
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), fps, size)
    for i in range(n):
        w.write(np.full((size[1], size[0], 3), (i * 5) % 255, np.uint8))
    w.release()


V = 32000


def _event_ids(k):
    """time head: 0012.5<sep>0030.0<sync>  score head: 4.5<sync>  text head: two caption tokens <sync>"""
    tb, sb = V + 1, V + 14
    t = [tb + 2, tb + 2, tb + 3, tb + 4, tb + 12, tb + 7, tb + 1, tb + 2, tb + 2, tb + 5, tb + 2, tb + 12, tb + 2, tb]
    s = [sb + 6, sb + 12, sb + 7, sb]
    return t + s + [100 + k, 200 + k, V]


class FakeModel:
    vocab_size = V
    device = torch.device("cpu")

    def __init__(self):
        import types
        self.config = types.SimpleNamespace(image_aspect_ratio="pad")
        self.generation_config = types.SimpleNamespace(eos_token_id=2, pad_token_id=None)
        self.batches = []

    def _encode_slots(self, clips):
        return torch.stack([c.float().mean().reshape(1, 1, 1).expand(c.shape[0], 8, 4) for c in clips], 0)

    def _generate_from_slots(self, ids, slots, tid, heads, max_new, eos, pad):
        self.batches.append(ids.shape[0])
        rows = []
        for b in range(ids.shape[0]):
            T = tid.shape[1]
            if T == 4:    # "short" videos: two events, the second caption closed by EOS
                row = _event_ids(1) + _event_ids(2)[:-1] + [eos]
            else:         # output that ends exactly on a <sync>: printed but not saved
                row = _event_ids(3)
            rows.append(row)
            heads[b] = 0
        n = max(len(r) for r in rows)                      # max_new_tokens reached right after the <sync> for the T=8 rows
        return torch.tensor([r + [pad] * (n - len(r)) for r in rows], dtype=torch.int64)


def test_evaluate_driver_end_to_end(tmp_path, monkeypatch, capsys):
    vids = tmp_path / "videos"
    vids.mkdir()
    anno = {"annotations": []}
    for k in range(5):
        _write_video(str(vids / f"v{k}.avi"), n=20 + k)
        anno["annotations"].append({"image_id": f"some/dir/v{k}.avi", "id": 100 + k, "caption": f" query {k}. ", "duration": 3.0})
    anno["annotations"].append({"image_id": "missing.avi", "id": 999, "caption": "x", "duration": 1.0})
    (tmp_path / "val.caption_coco_format.json").write_text(json.dumps(anno))
    (tmp_path / "prompt.txt").write_text("Find the moment described by: '{}'.\nsecond line is ignored\n")
    fake = FakeModel()
    from trace_b200.model.trace_mistral import _ImageProcessor
    tok = synth.StandInTokenizer()
    monkeypatch.setattr(ev, "load_pretrained_model", lambda *a, **k: (tok, fake, _ImageProcessor(56), 2048))
    args = ev.build_parser().parse_args(["--anno_path", str(tmp_path), "--video_path", str(vids), "--task", "tvg", "--dataset", "toy",
                                         "--output_dir", str(tmp_path / "out"), "--split", "val", "--num_frames", "4", "--batch_size", "3",
                                         "--max_new_tokens", "64", "--prompt_file", str(tmp_path / "prompt.txt"), "--model_path", "ckpt/toy"])
    out_path, results = ev.main(args)
    capsys.readouterr()
    assert os.path.basename(out_path) == "fmt_toy_val_f4_result.json" and json.load(open(out_path)) == results
    assert fake.batches == [3, 2]                       # 5 existing videos, --batch_size 3 (the reference would evaluate only 2 of them)
    assert [r["id"] for r in results] == [100, 101, 102, 103, 104]
    r0 = results[0]
    assert r0["video"].endswith("v0.avi") and r0["timestamps"] == [[12.5, 30.0], [12.5, 30.0]] and r0["scores"] == [[4.5], [4.5]]
    assert len(r0["captions"]) == 2
    dvc = reformat_dvc(results)
    assert dvc["v0.avi"][0]["timestamp"] == [12.5, 30.0]
    # prompt: first line of the file only, formatted with the stripped query, llama_2 template, <video> sentinel, trailing <sync>
    ids = ev.build_prompt_ids(ev.read_txt(str(tmp_path / "prompt.txt")).format("query 0."), tok)
    assert ids[0] == 1 and int((ids == -201).sum()) == 1 and ids[-1] == -205


def test_output_ending_on_sync_is_not_saved(tmp_path, monkeypatch, capsys):
    vids = tmp_path / "videos"
    vids.mkdir()
    _write_video(str(vids / "a.avi"), n=30)
    (tmp_path / "val.caption_coco_format.json").write_text(json.dumps({"annotations": [{"image_id": "a.avi", "id": 1, "caption": "c", "duration": 3.0}]}))
    (tmp_path / "prompt.txt").write_text("Describe.\n")
    fake = FakeModel()
    from trace_b200.model.trace_mistral import _ImageProcessor
    monkeypatch.setattr(ev, "load_pretrained_model", lambda *a, **k: (synth.StandInTokenizer(), fake, _ImageProcessor(56), 2048))
    args = ev.build_parser().parse_args(["--anno_path", str(tmp_path), "--video_path", str(vids), "--task", "dvc", "--dataset", "toy",
                                         "--output_dir", str(tmp_path / "out"), "--num_frames", "8", "--batch_size", "1",
                                         "--prompt_file", str(tmp_path / "prompt.txt"), "--model_path", "ckpt/toy"])
    _, results = ev.main(args)
    out = capsys.readouterr().out
    assert results == [] and "'timestamps': [[12.5, 30.0]]" in out      # printed, not saved (evaluate.py:409-411)


def test_inference_script_video_and_image(tmp_path, monkeypatch, capsys):
    """scripts/inference flow (inference.py:15-128) with a recorder model: video and image inputs, prompt, parsed events."""
    from PIL import Image
    from trace_b200.model.trace_mistral import _ImageProcessor
    from trace_b200.scripts import inference as inf

    class GenModel:
        vocab_size = V

        def __init__(self):
            import types
            self.config = types.SimpleNamespace(image_aspect_ratio="pad", num_frames=4)
            self.calls = []

        def generate(self, input_ids, **kw):
            self.calls.append((tuple(input_ids.shape), kw["modal_list"], tuple(kw["images_or_videos"][0].shape), len(kw["video_timestamps"][0]),
                               kw["heads"], kw["do_sample"], kw["max_new_tokens"]))
            return torch.tensor([_event_ids(1) + [7, 8, 2]])

    m = GenModel()
    monkeypatch.setattr(inf, "load_pretrained_model", lambda *a, **k: (synth.StandInTokenizer(), m, _ImageProcessor(56), 2048))
    vid = str(tmp_path / "c.avi")
    _write_video(vid, n=24)
    out = inf.main(["--model_path", "ckpt/toy", "--video", vid, "--num_frames", "6", "--max_new_tokens", "32"])
    assert out["timestamps"] == [[12.5, 30.0]] and out["scores"] == [[4.5]] and len(out["captions"]) == 2
    img = str(tmp_path / "i.png")
    Image.fromarray(np.full((40, 70, 3), 90, np.uint8)).save(img)
    inf.main(["--model_path", "ckpt/toy", "--image", img])
    capsys.readouterr()
    (s0, modal0, shape0, nts0, heads0, samp0, mx0), (s1, modal1, shape1, nts1, *_rest) = m.calls
    assert modal0 == ["video"] and shape0 == (6, 3, 56, 56) and nts0 == 6 and heads0 == [1] and samp0 is False and mx0 == 32
    assert modal1 == ["image"] and shape1 == (3, 56, 56) and nts1 == 4


def test_videomme_driver(tmp_path, monkeypatch, capsys):
    """Video-MME MCQ flow (videomme/evaluate.py): segment-centre sampler, no-pad CLIP transform, option extraction, per-category
    accuracy files; several questions about one video share a single clip encode (config 5: mixed 4 / 8 frame buckets)."""
    from trace_b200.eval import videomme as vm
    from trace_b200.model.trace_mistral import _ImageProcessor
    idx, msg = vm.get_index(4, 10.0, 39)
    assert idx.tolist() == [4, 14, 24, 33] and msg[1] == [1.4]                      # int(4.875 + np.round(9.75 * i)): 4, 14, 24 (19.5 -> 20), 33
    q, a = vm.qa_template({"question": "What?", "options": ["A. cat", "B. dog", "C. emu", "D. fox"], "answer": "C"})
    assert q == "Question: What?\nOptions:\n(A) cat\n(B) dog\n(C) emu\n(D) fox" and a == "(C) emu"
    assert vm.extract_option("(B) dog") == "b" and vm.extract_option("Best option: c.") == "c"
    (tmp_path / "data").mkdir()
    for name, n in (("vidA", 40), ("vidB", 30)):
        _write_video(str(tmp_path / "data" / f"{name}.mp4"), n=n)
    anno = []
    for k, (name, ans) in enumerate([("vidA", "A"), ("vidA", "B"), ("vidB", "A"), ("vidA", "C"), ("nope", "A")]):
        anno.append({"url": f"https://www.youtube.com/watch?v={name}", "duration": "short" if name == "vidA" else "long",
                     "question": f"q{k}", "options": ["A. a", "B. b", "C. c", "D. d"], "answer": ans})
    (tmp_path / "test.json").write_text(json.dumps(anno))

    class McqModel(FakeModel):
        def _generate_from_slots(self, ids, slots, tid, heads, max_new, eos, pad):
            self.batches.append(ids.shape[0])
            assert max_new == 128
            tb = V + 1
            # timestamps "0001.0<sync>", score "1.0<sync>", caption tokens spelling "( a )" in the stand-in vocabulary, EOS
            row = [tb + 2, tb + 2, tb + 2, tb + 3, tb + 12, tb + 2, tb] + [V + 14 + 3, V + 14 + 12, V + 14 + 2, V + 14] + [300, eos]
            return torch.tensor([row] * ids.shape[0], dtype=torch.int64)

    fake = McqModel()
    tok = synth.StandInTokenizer()
    monkeypatch.setattr(tok, "decode", lambda toks, **k: "(A) a", raising=False)
    monkeypatch.setattr(vm, "load_pretrained_model", lambda *a, **k: (tok, fake, _ImageProcessor(56), 2048))
    args = vm.build_parser().parse_args(["--anno_path", str(tmp_path / "test.json"), "--data_dir", str(tmp_path), "--output_dir",
                                         str(tmp_path / "out"), "--num_frames", "4", "--model_path", "ckpt/toy"])
    acc, res, stats = vm.main(args)
    capsys.readouterr()
    assert acc == {"short": [1, 3], "long": [1, 1]}                                  # every prediction is "a"
    assert stats["clips_encoded"] == 2 and stats["clip_encodes_saved"] == 2          # vidA encoded once for its 3 questions
    board = json.load(open(tmp_path / "out" / "upload_leaderboard.json"))
    assert abs(board["short"] - 100 / 3) < 1e-9 and board["long"] == 100.0 and board["Avg"] == 50.0
    assert json.load(open(tmp_path / "out" / "test.json"))["res_list"][0] == {"pred": ["a"], "gt": ["a"]}


def test_videomme_helpers_equal_reference_source():
    """get_index / qa_template executed from the reference's own source text (the module itself needs decord / accelerate)."""
    import ast
    import types
    import pytest
    ref = "/root/reference/trace/eval/videomme/evaluate.py"
    if not os.path.exists(ref):
        pytest.skip("/root/reference not mounted (GPU box)")
    from trace_b200.eval import videomme as vm
    src = open(ref).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MME_dataset")
    ns = {"np": np}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("get_index", "qa_template"):
            exec(compile(ast.Module(body=[fn], type_ignores=[]), ref, "exec"), ns)
    rng = np.random.default_rng(0)
    for _ in range(200):
        T = int(rng.integers(1, 130))
        fps = float(rng.uniform(5, 60))
        max_frame = int(rng.integers(T, 100000))
        me = types.SimpleNamespace(num_segments=T)
        want_idx, want_msg = ns["get_index"](me, None, fps, max_frame, first_idx=0)
        got_idx, got_msg = vm.get_index(T, fps, max_frame)
        assert got_idx.tolist() == want_idx.tolist() and got_msg == want_msg
    d = {"question": "Why?", "options": ["A. one", "B. two words", "C. 3", "D. four"], "answer": "B"}
    assert vm.qa_template(d) == ns["qa_template"](None, d)
