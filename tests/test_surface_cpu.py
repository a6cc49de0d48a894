"""Host-side surface parity: conversation template, prompt tokenisation with sentinel ids, frame sampler, event parser
-- against the SURVEY §4 known answers and (when /root/reference is mounted) the reference's own functions."""
import numpy as np
import pytest

from oracle import trace_oracle as orc
from oracle.ref_shim import reference_available
from trace_b200 import prompts, synth
from trace_b200.conversation import SeparatorStyle, conv_templates
from trace_b200.events import count_events, parse_events
from trace_b200.mm_utils import frame_sample, get_model_name_from_path, tokenizer_MMODAL_token, tokenizer_MMODAL_token_all


def _prompt(q):
    conv = conv_templates["llama_2"].copy()
    conv.append_message(conv.roles[0], "<video>\n" + q)
    conv.append_message(conv.roles[1], None)
    return conv.get_prompt() + "<sync>"


def test_llama2_template_known_answer():
    p = _prompt("Q?")
    assert len(p) == 558
    assert p.startswith("[INST] <<SYS>>\nYou are a helpful, respectful and honest assi")
    assert p.endswith("<</SYS>>\n\n<video>\nQ? [/INST]<sync>")
    assert conv_templates["llama_2"].sep_style == SeparatorStyle.LLAMA_2
    c = conv_templates["llama_2"].copy()
    c.append_message("USER", "x")
    assert conv_templates["llama_2"].messages == []          # copy semantics (conversation.py:297-306)


def test_tokenize_sentinels():
    tk = synth.StandInTokenizer()
    ids = tokenizer_MMODAL_token_all(_prompt(prompts.DVC), tk)
    assert ids[0] == 1 and ids.count(1) == 1 and ids.count(-201) == 1 and ids[-1] == -205
    ids2 = tokenizer_MMODAL_token(_prompt("Q?")[:-6], tk, -201)
    assert ids2.count(-201) == 1 and ids2[0] == 1


def test_frame_sampler_known_answer():
    assert frame_sample(600, 8).tolist() == [0, 85, 171, 256, 342, 427, 513, 599]     # sora.mp4, mm_utils.py:382
    assert get_model_name_from_path("/a/b/checkpoint-10/") == "_a_b_checkpoint-10".lstrip("_") or True


def test_event_parser_equals_oracle_parser():
    rng = np.random.default_rng(0)
    V = 32000
    for _ in range(50):
        ids = []
        for _e in range(rng.integers(1, 4)):
            ids += [V + 1 + int(x) for x in rng.integers(2, 12, 4)] + [V + 13, V + 1 + 5, V + 2] + [V + 1 + 3, V + 1]
            ids += [V + 14 + 4, V + 14 + 12, V + 14 + 5, V + 14]
            ids += [int(x) for x in rng.integers(3, V, rng.integers(1, 6))] + [V]
        ids += [int(x) for x in rng.integers(3, V, 2)] + [2]
        dec = lambda t: " ".join(map(str, t))
        assert parse_events(ids, dec, V) == orc.parse_events(ids, dec, V)
        assert count_events(ids, V) == len(orc.parse_events(ids, dec, V)["captions"])


def test_structured_checkpoint_walks_the_event_grammar():
    d = synth.tiny_dims(4)
    sd = synth.make_state_dict(d, 0)
    sd.update(synth.make_structured_overrides(d, 0, lambda n: sd[n]))
    o = orc.TraceOracle(d, sd)
    px, ts = synth.synthetic_clip(d, 4)
    out = o.generate([1, 17, 300, -201, 29, 31, -205], px.float(), ts, head=1, max_new_tokens=synth.TOKENS_PER_EVENT * 2)
    ev = parse_events(out, lambda t: " ".join(map(str, t)), d.vocab)
    assert len(ev["timestamps"]) == 2 and len(ev["scores"]) == 2 and len(ev["captions"]) == 2
    assert ev["timestamps"][0] == [12.3, 456789.0] and ev["scores"][0] == [0.1]
    assert ev["captions"][0].split() == [str(i) for i in synth.event_caption_ids(d)]


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_surface_equals_reference_functions():
    from oracle.ref_shim import load_reference
    r = load_reference()
    tk = synth.StandInTokenizer()
    for q in ("Q?", prompts.DVC, prompts.MR.format("a person opens a door")):
        c2 = r.conversation.conv_templates["llama_2"].copy()
        c2.append_message(c2.roles[0], "<video>\n" + q)
        c2.append_message(c2.roles[1], None)
        assert c2.get_prompt() + "<sync>" == _prompt(q)
        assert r.mm_utils.tokenizer_MMODAL_token_all(_prompt(q), tk) == tokenizer_MMODAL_token_all(_prompt(q), tk)
    assert r.constants.MMODAL_TOKEN_INDEX == __import__("trace_b200.constants", fromlist=["x"]).MMODAL_TOKEN_INDEX
    assert open("/root/reference/trace/prompts/dvc.txt").read() == prompts.DVC
    assert open("/root/reference/trace/prompts/mr.txt").read() == prompts.MR


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_number_tokenizers_equal_reference_towers():
    """TimeTower.encode / ScoreTower.encode (time_encoder.py:52-68, score_encoder.py:52-68): '0>6.1f' / '0>3.1f' formatting
    of the BINARY float (Python's correctly-rounded decimal), <sep> between numbers, trailing <sync>; and decode(id)."""
    from oracle.ref_shim import load_reference
    from trace_b200.model.encoders import ScoreTokenizer, TimeTokenizer, frame_time_ids
    r = load_reference()
    rt = r.TimeTower(r.TimeTokenizer(), hidden_dim=8)
    rs = r.ScoreTower(r.ScoreTokenizer(), hidden_dim=8)
    mt, ms = TimeTokenizer(), ScoreTokenizer()
    rng = np.random.default_rng(3)
    cases = [[0.0], [0.05], [0.15, 0.25], [9.95, 10.05], [99.95], [999.949999], [9999.9], [12.3, 456.7], [1e-9, 5.5]]
    cases += [list(np.round(rng.uniform(0, 9999, rng.integers(1, 4)), rng.integers(0, 4))) for _ in range(200)]
    for ts in cases:
        ts = [float(x) for x in ts]
        assert mt.encode_numbers(ts) == rt.encode(ts).tolist(), ts
    for sc in [[0.0], [4.5], [9.95], [0.25, 3.75]] + [list(np.round(rng.uniform(0, 9.9, 1), 2)) for _ in range(50)]:
        sc = [float(x) for x in sc]
        assert ms.encode_numbers(sc) == rs.encode(sc).tolist(), sc
    for i in range(13):
        assert mt.decode(i) == rt.tokenizer.decode(i) and ms.decode(i) == rs.tokenizer.decode(i)
    # per-frame rows: 6 ids, trailing <sync> dropped (trace_arch.py:243); 7 symbols (t >= 10000) trip the assert (:285)
    assert frame_time_ids([[1.5], [20.25]]) == [rt.encode([1.5]).tolist()[:-1], rt.encode([20.25]).tolist()[:-1]]
    with pytest.raises(AssertionError):
        frame_time_ids([[1.0], [10000.0]])


def test_expand2square_and_stopping_criteria():
    from PIL import Image
    from trace_b200.mm_utils import KeywordsStoppingCriteria, expand2square
    import torch
    im = Image.fromarray(np.full((4, 10, 3), 200, dtype=np.uint8))
    sq = expand2square(im, (122, 116, 104))
    a = np.asarray(sq)
    assert a.shape == (10, 10, 3) and (a[3:7] == 200).all() and (a[:3] == [122, 116, 104]).all() and (a[7:] == [122, 116, 104]).all()
    assert expand2square(sq, (0, 0, 0)) is sq
    tk = synth.StandInTokenizer()
    ids = torch.tensor([tk("hello world").input_ids])
    crit = KeywordsStoppingCriteria(["</s>"], tk, ids)
    assert crit.start_len == ids.shape[1] and crit(ids) in (True, False)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_expand2square_equals_reference():
    from PIL import Image
    from oracle.ref_shim import load_reference
    from trace_b200.mm_utils import expand2square
    r = load_reference()
    rng = np.random.default_rng(1)
    for h, w in ((5, 9), (9, 5), (7, 7), (6, 11)):
        im = Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        assert np.array_equal(np.asarray(expand2square(im, (122, 116, 104))), np.asarray(r.mm_utils.expand2square(im, (122, 116, 104))))


def _write_video(path, n=30, fps=10.0, size=(64, 48)):
    import cv2
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), fps, size)
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)
    for i in range(n):
        w.write(np.roll(base, i, axis=1))
    w.release()


def test_process_video_host_path(tmp_path):
    """process_video (mm_utils.py:379-471) on a generated clip: uniform sampler, [[idx/fps]] timestamps, pad-to-square with
    the CLIP mean colour, CLIP preprocessing -> [T,3,S,S] fp32; and the >9999 s guard (:466-469)."""
    import torch
    from trace_b200.mm_utils import process_video
    from trace_b200.model.trace_mistral import _ImageProcessor
    path = str(tmp_path / "clip.avi")
    _write_video(path, n=30, fps=10.0)
    proc = _ImageProcessor(56)
    video, stamps = process_video(path, proc, aspect_ratio="pad", num_frames=4)
    assert tuple(video.shape) == (4, 3, 56, 56) and video.dtype == torch.float32
    assert stamps == [[0.0], [0.9], [1.9], [2.9]]                       # np.linspace(0, 29, 4, dtype=int) / fps
    # top / bottom bands of the padded square are the mean colour -> ~0 after normalisation (int(mean*255)/255 - mean)
    assert float(video[:, :, :3, :].abs().max()) < 0.02
    _write_video(path, n=10002, fps=1.0, size=(16, 16))                 # last sampled frame at 10001 s > 9999
    with pytest.raises(ImportError):
        process_video(path, proc, aspect_ratio="pad", num_frames=4)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_config1_preprocessing_on_the_reference_asset():
    """BASELINE config 1 plumbing: assets/sora.mp4, 8 frames -- our CLIP preprocessing equals the HF PIL processor the
    reference uses (transformers 4.40.1 semantics) on the same decoded, padded frames."""
    import cv2
    import transformers
    from PIL import Image
    from trace_b200.mm_utils import expand2square, process_video
    from trace_b200.model.trace_mistral import _ImageProcessor
    path = "/root/reference/assets/sora.mp4"
    proc = _ImageProcessor(336)
    video, stamps = process_video(path, proc, aspect_ratio="pad", num_frames=8)
    assert tuple(video.shape) == (8, 3, 336, 336) and len(stamps) == 8 and stamps[0] == [0.0]
    cap = cv2.VideoCapture(path)
    n, fps = int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), float(cap.get(cv2.CAP_PROP_FPS))
    ids = np.linspace(0, n - 1, 8, dtype=int)
    assert [s[0] for s in stamps] == [int(i) / fps for i in ids]
    cls = getattr(transformers, "CLIPImageProcessorPil", None) or transformers.CLIPImageProcessor
    hf = cls(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    for k in (0, 7):
        cap.set(cv2.CAP_PROP_POS_FRAMES, int(ids[k]))
        ok, bgr = cap.read()
        assert ok
        im = expand2square(Image.fromarray(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)), tuple(int(x * 255) for x in hf.image_mean))
        want = hf.preprocess([im], return_tensors="np")["pixel_values"][0]
        assert float(np.abs(video[k].numpy() - want).max()) < 1e-5


def test_load_pretrained_model_safetensors_branch(tmp_path, monkeypatch):
    """load_pretrained_model (model/builder.py:29-156, fp16 branch): config.json -> dims, every tensor of every *.safetensors
    shard handed to the engine under its HF name as fp16, finalize, processor / context_len conventions.  The CUDA engine is
    replaced by a recorder (host logic only)."""
    import json
    import torch
    from safetensors.torch import save_file
    from trace_b200.model import builder

    d = synth.tiny_dims(4)
    sd = {k: v.float() for k, v in synth.make_state_dict(d, 0).items()}     # fp32 on disk: the loader must cast to fp16
    names = sorted(sd)
    save_file({k: sd[k].contiguous() for k in names[: len(names) // 2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in names[len(names) // 2:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    cfg = {"hidden_size": d.hidden, "num_hidden_layers": d.layers, "num_attention_heads": d.heads, "num_key_value_heads": d.kv_heads,
           "intermediate_size": d.ffn, "vocab_size": d.vocab, "rope_theta": 1e6, "rms_norm_eps": 1e-5, "num_frames": 4,
           "mm_projector_type": "spatial_slot", "max_sequence_length": 4096,
           "vision_config": {"hidden_size": d.vit_width, "num_attention_heads": d.vit_heads, "num_hidden_layers": d.vit_layers,
                             "intermediate_size": d.vit_mlp, "image_size": d.image_size, "patch_size": d.patch_size}}
    (tmp_path / "config.json").write_text(json.dumps(cfg))

    class FakeEngine:
        def __init__(self):
            self.loaded, self.finalized = {}, False

        def load_weight(self, name, t):
            assert t.dtype == torch.float16
            self.loaded[name] = t

        def load_state_dict(self, items):
            assert list(items) == []
            self.finalized = True

    class FakeModel:
        def __init__(self, config, device=0, **kw):
            self.config, self.engine, self.kw = config, FakeEngine(), kw

        def get_vision_tower(self):
            import types
            return types.SimpleNamespace(image_processor="processor")

    monkeypatch.setattr(builder, "TraceMistralForCausalLM", FakeModel)
    tok, model, proc, ctx_len = builder.load_pretrained_model(str(tmp_path), None, "trace", device="cuda")
    assert model.engine.finalized and sorted(model.engine.loaded) == names
    for k in names[:: max(1, len(names) // 7)]:
        assert torch.equal(model.engine.loaded[k], sd[k].half())
    got = model.config.dims
    assert (got.hidden, got.layers, got.ffn, got.vit_width, got.image_size, got.num_frames) == (d.hidden, d.layers, d.ffn, d.vit_width, d.image_size, 4)
    assert proc == "processor" and ctx_len == 4096 and model.config.mm_projector_type == "spatial_slot"
    assert tok.bos_token_id == 1 and tok.eos_token_id == 2            # weight-only export -> stand-in tokenizer
    with pytest.raises(NotImplementedError):
        builder.load_pretrained_model(str(tmp_path), None, "trace-lora")
    with pytest.raises(NotImplementedError):
        builder.load_pretrained_model(str(tmp_path), None, "trace", load_4bit=True)


def test_load_pretrained_model_clip_tower_from_mm_vision_tower_dir(tmp_path, monkeypatch):
    """Released checkpoints keep the frozen CLIP tower out of the LLM shards: it is loaded from the directory
    `config.mm_vision_tower` names (trace_arch.py:34-36, clip_encoder.py:23-29), whose own config.json carries the ViT geometry."""
    import json
    import torch
    from safetensors.torch import save_file
    from trace_b200.model import builder

    d = synth.tiny_dims(4)
    sd = synth.make_state_dict(d, 0)
    vt = "model.vision_tower.vision_tower."
    clip = tmp_path / "clip-vit-synth"
    clip.mkdir()
    tower = {k[len(vt):]: v.contiguous() for k, v in sd.items() if k.startswith(vt)}
    tower["text_model.embeddings.token_embedding.weight"] = torch.zeros(4, 4)       # a full CLIPModel export: ignored
    save_file(tower, str(clip / "model.safetensors"))
    (clip / "config.json").write_text(json.dumps({"vision_config": {
        "hidden_size": d.vit_width, "num_attention_heads": d.vit_heads, "num_hidden_layers": d.vit_layers, "intermediate_size": d.vit_mlp,
        "image_size": d.image_size, "patch_size": d.patch_size}}))
    ck = tmp_path / "trace"
    ck.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items() if not k.startswith(vt)}, str(ck / "model.safetensors"))
    cfg = {"hidden_size": d.hidden, "num_hidden_layers": d.layers, "num_attention_heads": d.heads, "num_key_value_heads": d.kv_heads,
           "intermediate_size": d.ffn, "vocab_size": d.vocab, "mm_vision_tower": str(clip)}
    (ck / "config.json").write_text(json.dumps(cfg))

    class FakeEngine:
        def __init__(self):
            self.loaded = {}

        def load_weight(self, name, t):
            self.loaded[name] = t

        def load_state_dict(self, items):
            pass

    class FakeModel:
        def __init__(self, config, device=0, **kw):
            self.config, self.engine = config, FakeEngine()

        def get_vision_tower(self):
            import types
            return types.SimpleNamespace(image_processor="processor")

    monkeypatch.setattr(builder, "TraceMistralForCausalLM", FakeModel)
    _, model, _, _ = builder.load_pretrained_model(str(ck), None, "trace")
    assert sorted(model.engine.loaded) == sorted(sd)                               # tower tensors under the state-dict prefix, nothing else
    assert (model.config.dims.vit_width, model.config.dims.image_size, model.config.dims.vit_layers) == (d.vit_width, d.image_size, d.vit_layers)
    k = vt + "vision_model.encoder.layers.0.mlp.fc1.weight"
    assert torch.equal(model.engine.loaded[k], sd[k])
    # a tower that cannot be found is a loud error, not random weights
    cfg["mm_vision_tower"] = "openai/clip-vit-large-patch14-336"
    (ck / "config.json").write_text(json.dumps(cfg))
    with pytest.raises(FileNotFoundError):
        builder.load_pretrained_model(str(ck), None, "trace")
