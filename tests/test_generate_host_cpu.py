"""Argument handling and error behaviour of TraceMistralForCausalLM.generate / forward (trace_mistral.py:268-314) with the
CUDA engine replaced by a recorder: what the wrapper accepts, rejects, mutates and returns -- no arithmetic."""
import pytest
import torch

from trace_b200 import synth
from trace_b200.model import trace_mistral as tm


class FakeEngine:
    max_ctx = 256
    max_prefill_tokens = 4096

    def __init__(self, dims, device=0, **kw):
        self.dims, self.kw, self.log = dims, kw, []
        self._slot = 0

    def set_sampling(self, do_sample, temperature=1.0, top_k=50, seed=0):
        self.log.append(("sampling", bool(do_sample), float(temperature), int(top_k), int(seed)))

    def encode_frames(self, frames):
        self.log.append(("encode", tuple(frames.shape)))
        return torch.zeros(frames.shape[0], 8, self.dims.hidden, dtype=torch.float16)

    def build_prefill(self, ids, slots, tid):
        B, L = ids.shape
        S = L if slots is None else L - 1 + 14 * tid.shape[1]
        self.log.append(("splice", B, L, None if tid is None else tid.shape[1]))
        return torch.zeros(B, S, 8, dtype=torch.float16)

    def alloc_slot(self, n):
        self._slot += 1
        return self._slot

    def free_slot(self, s):
        self.log.append(("free", s))

    def prefill(self, kv, embeds, return_hidden=False):
        self.log.append(("prefill", len(kv), embeds.shape[1]))

    def decode(self, kv, heads, max_new, eos=2, pad=2, return_logits=False):
        self.log.append(("decode", len(kv), list(heads), max_new, eos, pad))
        B = len(kv)
        n = min(max_new, 3 + B % 3)
        for b in range(B):
            if heads[b] >= 0:
                heads[b] = (heads[b] + 1) % 3
        return torch.arange(B * n, dtype=torch.int64).view(B, n) + 10

    def load_state_dict(self, items):
        pass


@pytest.fixture()
def model(monkeypatch):
    monkeypatch.setattr(tm, "Engine", FakeEngine)
    monkeypatch.setattr(tm.TraceMistralForCausalLM, "device", torch.device("cpu"), raising=False)
    d = synth.tiny_dims(4)
    m = tm.TraceMistralForCausalLM(tm.TraceMistralConfig(d), device=0)
    m.device = torch.device("cpu")
    return m


def _inputs(B, T=4, L=7, d=None):
    d = d or synth.tiny_dims(4)
    ids = torch.tensor([[1, 5, -201, 6, 7, 8, -205][:L]] * B)
    clips = [torch.zeros(T, 3, d.image_size, d.image_size) for _ in range(B)]
    ts = [[[float(i)] for i in range(T)] for _ in range(B)]
    return ids, clips, ts


def test_generate_returns_new_tokens_and_mutates_heads(model):
    ids, clips, ts = _inputs(3)
    heads = [1, 1, 1]
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=clips, modal_list=["video"] * 3, video_timestamps=ts,
                         heads=heads, do_sample=False, temperature=0.0, max_new_tokens=5, use_cache=True, pad_token_id=2,
                         stopping_criteria=["ignored"])
    assert out.dtype == torch.int64 and out.shape[0] == 3 and out.shape[1] <= 5
    assert heads == [2, 2, 2]                                         # updated in place (trace_mistral.py:338-344)
    kinds = [e[0] for e in model.engine.log]
    assert kinds.count("encode") == 1 and kinds.count("decode") == 1 and kinds.count("free") == 3
    assert ("splice", 3, 7, 4) in model.engine.log and ("prefill", 3, 7 - 1 + 14 * 4) in model.engine.log


def test_generate_splits_more_than_eight_rows_and_pads(model):
    ids, clips, ts = _inputs(11)
    out = model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1] * 11, max_new_tokens=6, pad_token_id=9)
    dec = [e for e in model.engine.log if e[0] == "decode"]
    assert [e[1] for e in dec] == [8, 3]                              # TRACE_B200_MAX_DECODE_BATCH rows per decode loop
    assert out.shape == (11, 5)                                       # longest group: 3 + 8 % 3 = 5; the other (3 + 3 % 3 = 3) padded
    assert bool((out[8:, 3:] == 9).all())


def test_generate_without_heads_is_text_head_only(model):
    ids, clips, ts = _inputs(2)
    model.generate(ids, images_or_videos=clips, video_timestamps=ts, max_new_tokens=4)
    dec = [e for e in model.engine.log if e[0] == "decode"][0]
    assert dec[2] == [-1, -1] and dec[4] == 2 and dec[5] == 2         # eos from generation_config, pad defaults to eos


def test_generate_text_only_and_image_modality(model):
    ids = torch.tensor([[1, 5, 6, 7]])
    model.generate(ids, max_new_tokens=3)
    assert ("splice", 1, 4, None) in model.engine.log and not any(e[0] == "encode" for e in model.engine.log)
    d = synth.tiny_dims(4)
    ids2, _, ts = _inputs(1)
    img = torch.zeros(3, d.image_size, d.image_size)
    model.generate(ids2, images_or_videos=[img], modal_list=["image"], video_timestamps=ts, heads=[1], max_new_tokens=3)
    assert ("encode", (4, 3, d.image_size, d.image_size)) in model.engine.log      # repeated num_frames times (trace_arch.py:221)


def test_generate_error_behaviour(model):
    ids, clips, ts = _inputs(2)
    with pytest.raises(NotImplementedError):
        model.generate(ids, images_or_videos=clips, video_timestamps=ts, inputs_embeds=torch.zeros(1))      # trace_mistral.py:282-283
    with pytest.raises(NotImplementedError):                                                                # training-only inputs
        model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1, 1], times=[[1.0]], scores=[[2.0]])
    bad = ids.clone()
    bad[0, 3] = -203
    with pytest.raises(NotImplementedError):                                                                # <time> placeholder without `times`
        model.generate(bad, images_or_videos=clips, video_timestamps=ts, heads=[1, 1])
    with pytest.raises(NotImplementedError):
        model.generate(ids, images_or_videos=clips, video_timestamps=ts, do_sample=True, top_p=0.9)
    with pytest.raises(ValueError):                                                                         # text-only: placeholders are not ids
        model.generate(ids, max_new_tokens=3)
    with pytest.raises(RuntimeError):                                                                       # torch.stack, trace_arch.py:222
        model.generate(ids, images_or_videos=[clips[0], clips[1][:2]], video_timestamps=ts, heads=[1, 1])
    with pytest.raises(AssertionError):                                                                     # trace_mistral.py:245
        model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1])
    two = ids.clone()
    two[0, 1] = -201
    with pytest.raises(AssertionError):                                                                     # trace_arch.py:408
        model.generate(two, images_or_videos=clips, video_timestamps=ts, heads=[1, 1])
    mask = torch.ones_like(ids)
    mask[0, 0] = 0
    with pytest.raises(NotImplementedError):
        model.generate(ids, attention_mask=mask, images_or_videos=clips, video_timestamps=ts, heads=[1, 1])
    with pytest.raises(tm.TraceB200Error):                                                                  # context overflow is loud
        model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1, 1], max_new_tokens=10_000)
    with pytest.raises(tm.TraceB200Error):                                                                  # forward needs the head weights
        model.forward(ids)


def test_generate_sampling_arguments_reach_the_engine(model):
    """scripts/inference/inference.py:62,71-73: do_sample=True, temperature=0.2 (generation_config: top_k=50)."""
    ids, clips, ts = _inputs(1)
    model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1], do_sample=True, temperature=0.2, max_new_tokens=3, seed=7)
    assert ("sampling", True, 0.2, 50, 7) in [tuple(round(x, 6) if isinstance(x, float) else x for x in e) for e in model.engine.log]
    model.engine.log.clear()
    model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1], do_sample=False, temperature=0.0, max_new_tokens=3)
    assert model.engine.log[0][:2] == ("sampling", False)
    # without an explicit seed the draw follows torch's global generator (like HF's multinomial)
    torch.manual_seed(3)
    model.engine.log.clear()
    model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1], do_sample=True, temperature=0.2, max_new_tokens=3)
    s1 = model.engine.log[0][4]
    torch.manual_seed(3)
    model.engine.log.clear()
    model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1], do_sample=True, temperature=0.2, max_new_tokens=3)
    assert model.engine.log[0][4] == s1


def test_prefill_groups_respect_the_workspace(model):
    """Rows are prefilled in groups that fit max_prefill_tokens (a 128-frame batch of 8 does not fit the default 4096-token
    workspace in one call): here S = 6 + 14*4 = 62, workspace 256 tokens -> groups of 4."""
    ids, clips, ts = _inputs(6)
    model.engine.max_prefill_tokens = 256
    model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1] * 6, max_new_tokens=3)
    assert [e[1] for e in model.engine.log if e[0] == "prefill"] == [4, 2]
    model.engine.max_prefill_tokens = 40
    with pytest.raises(tm.TraceB200Error):
        model.generate(ids, images_or_videos=clips, video_timestamps=ts, heads=[1] * 6, max_new_tokens=3)


def test_package_infer_passes_real_timestamps(model, monkeypatch):
    """trace_b200.infer / x_infer (trace/__init__.py:23-86): text-head-only generate.  The per-frame timestamps that
    process_video returned must reach generate (they become the time rows of the prompt); fabricating them is an error."""
    import trace_b200
    seen = {}

    def fake_generate(ids, **kw):
        seen.update(kw)
        seen["ids"] = ids
        return torch.tensor([[5, 6, 2]])

    monkeypatch.setattr(model, "generate", fake_generate)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    tok = synth.StandInTokenizer()
    d = synth.tiny_dims(4)
    video = torch.zeros(4, 3, d.image_size, d.image_size)
    ts = [[0.0], [1.7], [3.4], [5.1]]
    out = trace_b200.infer(model, (video, ts), "What happens?", tok)
    assert isinstance(out, str) and seen["video_timestamps"] == [ts] and "heads" not in seen
    assert seen["do_sample"] is False and seen["max_new_tokens"] == 1024 and int((seen["ids"] == -201).sum()) == 1
    trace_b200.x_infer(video, "Which one?", model, tok, mode="mcqa", do_sample=True, video_timestamps=ts)
    assert seen["do_sample"] is True and abs(seen["temperature"] - 0.2) < 1e-9 and seen["video_timestamps"] == [ts]
    with pytest.raises(ValueError):
        trace_b200.infer(model, video, "What happens?", tok)            # no timestamps: refuse instead of inventing them
