"""End-to-end parity of the CUDA path (through the C ABI / the reference-shaped Python surface) against
the CPU oracle and the committed golden fixtures (reference outputs).  fp16 engine vs fp32 oracle:
tolerances are stated inline; greedy ids must match except at reference top-2 margins below the logit
tolerance (an fp16 tie), see tests/util.compare_ids_with_margin."""
import numpy as np
import pytest
import torch

from oracle import trace_oracle as orc
from trace_b200 import synth
from tests.util import CASES, load_case, ranges, compare_ids_with_margin

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.06      # |logit| ~ 5..10 in these synthetic models: fp16 activations + fp16-rounded logits
FEAT_TOL = 0.03


def _engine(d, seed, **kw):
    from trace_b200.engine import Engine
    e = Engine(d, device=0, vit_chunk_frames=4, max_prefill_tokens=1024, max_ctx=512, **kw)
    e.load_state_dict(synth.iter_state_dict(d, seed))
    return e


def _rel(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return float((got - want).abs().max() / (want.abs().max() + 1e-6))


@pytest.mark.parametrize("name", ["tiny_b1", "mid_b1"])
def test_stages_against_oracle(name):
    c = load_case(name)
    d, meta, z = c["dims"], c["meta"], c["z"]
    eng = _engine(d, meta["seed"])
    o = orc.TraceOracle(d, synth.make_state_dict(d, meta["seed"]))
    px = c["clips"][0].cuda()
    slots, feats = eng.encode_frames(px, return_feats=True)
    torch.cuda.synchronize()
    of = o.vit_features(c["clips"][0].float())
    assert _rel(feats, of) < FEAT_TOL, f"CLIP features rel err {_rel(feats, of)}"
    np.testing.assert_allclose(feats.float().cpu()[:, ::3, ::8].numpy(), z["feats_sample"][: meta["T"]], atol=FEAT_TOL * float(of.abs().max()))
    os_ = o.slot_pool(of)
    assert _rel(slots, os_) < FEAT_TOL, f"slot pool rel err {_rel(slots, os_)}"
    # splice
    from trace_b200.model.encoders import frame_time_ids
    ids = torch.tensor([c["prompt_ids"][0]], dtype=torch.int64, device="cuda")
    tid = torch.tensor([frame_time_ids(c["ts"][0])], dtype=torch.int32, device="cuda")
    emb = eng.build_prefill(ids, slots[None].contiguous(), tid)
    oe = o.splice(c["prompt_ids"][0], o.encode_clip(c["clips"][0].float(), c["ts"][0]))
    assert emb.shape[1] == oe.shape[0] == int(z["prefill_len"])
    assert _rel(emb[0], oe) < FEAT_TOL
    # text rows of the splice are pure gathers: bit-exact against the fp16 table
    L = len(c["prompt_ids"][0])
    vpos = c["prompt_ids"][0].index(-201)
    tab = synth.synth_tensor("model.embed_tokens.weight", (d.vocab, d.hidden), "emb", meta["seed"])
    for p in range(vpos):
        assert torch.equal(emb[0, p].cpu(), tab[c["prompt_ids"][0][p]]), f"gather mismatch at row {p}"
    # prefill hidden states
    kv = eng.alloc_slot(emb.shape[1] + 8)
    hid = eng.prefill([kv], emb, return_hidden=True)
    h_or, _ = o.llm_forward(oe, None)
    # engine returns pre-final-norm hidden; compare after the oracle's final norm applied on the engine output
    hn = o._rms(hid[0].float().cpu(), o.p("model.norm.weight"))
    assert _rel(hn, h_or) < FEAT_TOL, f"prefill hidden rel err {_rel(hn, h_or)}"
    eng.free_slot(kv)
    eng.close()


@pytest.mark.parametrize("name", CASES)
def test_generate_matches_reference(name):
    c = load_case(name)
    d, meta, z = c["dims"], c["meta"], c["z"]
    from trace_b200.model import TraceMistralConfig, TraceMistralForCausalLM
    model = TraceMistralForCausalLM(TraceMistralConfig(d), device=0, vit_chunk_frames=4, max_prefill_tokens=1024, max_ctx=512)
    model.load_state_dict(dict(synth.iter_state_dict(d, meta["seed"])))
    B = meta["B"]
    ids = torch.tensor(c["prompt_ids"], dtype=torch.int64, device="cuda")
    heads = None if meta["heads0"] is None else [meta["heads0"]] * B
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[x.cuda() for x in c["clips"]],
                         modal_list=["video"] * B, video_timestamps=c["ts"], heads=heads, do_sample=False, temperature=0.0,
                         max_new_tokens=meta["max_new"], use_cache=True, pad_token_id=2)
    assert out.dtype == torch.int64 and out.shape[0] == B
    n, ties = compare_ids_with_margin(out.cpu().tolist(), z["out_ids"].tolist(), z["top_vals"], LOGIT_TOL, name)
    print(f"{name}: {n} ids compared bit-exact, {ties} fp16 tie-breaks")
    if heads is not None and ties == 0:
        # `heads` mutated in place exactly like the reference (trace_mistral.py:338-344): the head that produced
        # the last step, updated once more by the last emitted token (only when that token was fed back)
        assert all(h in (0, 1, 2) for h in heads)


@pytest.mark.parametrize("name", ["tiny_b1", "tiny_b2"])
def test_logits_against_oracle(name):
    c = load_case(name)
    d, meta, z = c["dims"], c["meta"], c["z"]
    eng = _engine(d, meta["seed"])
    o = orc.TraceOracle(d, synth.make_state_dict(d, meta["seed"]))
    from trace_b200.model.encoders import frame_time_ids
    B, T = meta["B"], meta["T"]
    px = torch.cat(c["clips"], 0).cuda()
    slots = eng.encode_frames(px).view(B, T, 8, d.hidden)
    ids = torch.tensor(c["prompt_ids"], dtype=torch.int64, device="cuda")
    tid = torch.tensor([frame_time_ids(t) for t in c["ts"]], dtype=torch.int32, device="cuda")
    emb = eng.build_prefill(ids, slots.contiguous(), tid)
    kv = [eng.alloc_slot(emb.shape[1] + meta["max_new"]) for _ in range(B)]
    eng.prefill(kv, emb)
    heads = [meta["heads0"]] * B
    out, logits = eng.decode(kv, heads, meta["max_new"], return_logits=True)
    rg = ranges(d)
    worst = 0.0
    for b in range(B):
        ref_ids, trace, _ = o.generate(c["prompt_ids"][b], c["clips"][b].float(), c["ts"][b], head=meta["heads0"],
                                       max_new_tokens=meta["max_new"], return_trace=True)
        for s, st in enumerate(trace):
            if s >= out.shape[1] or (s > 0 and out[b, s - 1].item() != ref_ids[s - 1]):
                break   # after a tie-break the rows are no longer comparable
            lo = st["lo"]
            got = logits[s, b, lo: lo + st["logits"].numel()].cpu()
            err = float((got - st["logits"]).abs().max())
            worst = max(worst, err)
            assert err < LOGIT_TOL, f"{name} row {b} step {s}: active-head logits differ by {err}"
            # inactive ranges are exactly -inf (trace_mistral.py:244-252)
            full = logits[s, b].cpu()
            mask = torch.ones_like(full, dtype=torch.bool)
            mask[lo: lo + st["logits"].numel()] = False
            assert bool(torch.isinf(full[mask]).all()) and bool((full[mask] < 0).all())
    print(f"{name}: worst active-head logit error {worst:.4g} (tolerance {LOGIT_TOL})")
    for s in kv:
        eng.free_slot(s)
    eng.close()


def test_graph_path_equals_eager_path():
    """The CUDA-graph production loop and the instrumented eager loop must emit identical ids."""
    c = load_case("tiny_b2")
    d, meta = c["dims"], c["meta"]
    eng = _engine(d, meta["seed"])
    from trace_b200.model.encoders import frame_time_ids
    B, T = meta["B"], meta["T"]
    slots = eng.encode_frames(torch.cat(c["clips"], 0).cuda()).view(B, T, 8, d.hidden)
    ids = torch.tensor(c["prompt_ids"], dtype=torch.int64, device="cuda")
    tid = torch.tensor([frame_time_ids(t) for t in c["ts"]], dtype=torch.int32, device="cuda")
    emb = eng.build_prefill(ids, slots.contiguous(), tid)
    outs = []
    for want_logits in (True, False, False):
        kv = [eng.alloc_slot(emb.shape[1] + 40) for _ in range(B)]
        eng.prefill(kv, emb)
        r = eng.decode(kv, [1] * B, 40, return_logits=want_logits)
        outs.append((r[0] if want_logits else r).tolist())
        for s in kv:
            eng.free_slot(s)
    assert outs[0] == outs[1] == outs[2]
    eng.close()
