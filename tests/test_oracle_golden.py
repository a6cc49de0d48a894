"""Pin the CPU oracle (oracle/trace_oracle.py) against the reference's own outputs (tests/golden,
made by make_golden.py from /root/reference) and the SURVEY §4 known-answer table."""
import numpy as np
import pytest
import torch

from oracle import trace_oracle as orc
from trace_b200 import synth
from tests.util import CASES, WIDE_CASES, load_case, ranges, row_heads0


def test_known_answers_time_score():
    assert orc.encode_time([12.3, 45.6]) == [2, 2, 3, 4, 12, 5, 1, 2, 2, 6, 7, 12, 8, 0]
    assert orc.encode_time([0.0]) == [2, 2, 2, 2, 12, 2, 0]
    assert orc.encode_time([9999.9]) == [11, 11, 11, 11, 12, 11, 0]
    assert orc.encode_time([]) == [0]
    assert orc.encode_time([85 / 30.050377423088836]) == [2, 2, 2, 4, 12, 10, 0]
    assert orc.encode_score([4.5]) == [6, 12, 7, 0]
    assert orc.encode_score([]) == [0]
    assert [orc.decode_symbol(i) for i in range(13)] == ["<sync>", "<sep>"] + list("0123456789") + ["."]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    c = load_case(name)
    d, z, meta = c["dims"], c["z"], c["meta"]
    o = orc.TraceOracle(d, synth.make_state_dict(d, meta["seed"]))
    rg = ranges(d)
    for b in range(meta["B"]):
        ids, trace, mid = o.generate(c["prompt_ids"][b], c["clips"][b].float(), c["ts"][b], head=meta["heads0"],
                                     max_new_tokens=meta["max_new"], return_trace=True)
        feats = o.vit_features(c["clips"][b].float())
        T = meta["T"]
        np.testing.assert_allclose(feats[:, ::3, ::8].numpy(), z["feats_sample"][b * T:(b + 1) * T], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(mid["vis"][:, ::32].numpy(), z["vis_sample"][b], atol=2e-4, rtol=1e-4)
        np.testing.assert_allclose(mid["prefill_embeds"][:, ::64].numpy(), z["embeds_sample"][b], atol=2e-4, rtol=1e-4)
        assert mid["prefill_embeds"].shape[0] == int(z["prefill_len"])
        ref_ids = z["out_ids"][b].tolist()
        assert ids == ref_ids[: len(ids)], (ids, ref_ids)
        for s, st in enumerate(trace):
            h = int(z["step_heads"][s][b])
            assert (0 if st["head"] is None else st["head"]) == h
            idx = z["top_idx"][s][b]
            full = torch.full((d.total_vocab,), float("nan"))
            full[st["lo"]: st["lo"] + st["logits"].numel()] = st["logits"]
            np.testing.assert_allclose(full[idx].numpy(), z["top_vals"][s][b], atol=5e-4, rtol=1e-4)


@pytest.mark.parametrize("name,rows,steps", [("wide_b1", [0], 16), ("wide_b4", [1, 2], 6)])
def test_oracle_matches_reference_full_width(name, rows, steps):
    """The oracle at the full widths of the BASELINE configs (shallow depth), pinned to the reference's own outputs: stage
    samples, prefill hidden state, and the first decode steps of a few rows (the whole suite has to stay within minutes)."""
    c = load_case(name)
    d, z, meta = c["dims"], c["z"], c["meta"]
    s0, s1, s2, s3 = meta["strides"]
    o = orc.TraceOracle(d, synth.make_state_dict(d, meta["seed"]))
    heads0 = row_heads0(meta)
    T = meta["T"]
    for b in rows:
        ids, trace, mid = o.generate(c["prompt_ids"][b], c["clips"][b].float(), c["ts"][b], head=heads0[b], max_new_tokens=steps, return_trace=True)
        feats = o.vit_features(c["clips"][b].float())
        np.testing.assert_allclose(feats[:, ::s0, ::s1].numpy(), z["feats_sample"][b * T:(b + 1) * T], atol=5e-4, rtol=1e-4)
        np.testing.assert_allclose(mid["vis"][:, ::s2].numpy(), z["vis_sample"][b], atol=5e-4, rtol=1e-4)
        np.testing.assert_allclose(mid["prefill_embeds"][:, ::s3].numpy(), z["embeds_sample"][b], atol=5e-4, rtol=1e-4)
        h, _ = o.llm_forward(mid["prefill_embeds"], None)
        np.testing.assert_allclose(h[::7, ::s3].numpy(), z["hid_sample"][b], atol=2e-3, rtol=1e-3)
        np.testing.assert_allclose(h[-1].numpy(), z["hid_last"][b], atol=2e-3, rtol=1e-3)
        assert ids == z["out_ids"][b].tolist()[: len(ids)]
        for s, st in enumerate(trace):
            assert st["head"] == int(z["step_heads"][s][b])
            full = torch.full((d.total_vocab,), float("nan"))
            full[st["lo"]: st["lo"] + st["logits"].numel()] = st["logits"]
            np.testing.assert_allclose(full[z["top_idx"][s][b]].numpy(), z["top_vals"][s][b], atol=2e-3, rtol=1e-4)


def test_prefill_length_known_answer():
    # 12 prompt ids incl. one -201, T=4  ->  67 = 11 + 4*(8+6)
    c = load_case("tiny_b1")
    assert int(c["z"]["prefill_len"]) == 67


def test_parse_events_roundtrip():
    V = 32000
    ids = [V + 1 + 2, V + 1 + 3, V + 1 + 12, V + 1 + 4, V + 2, V + 1 + 5, V + 1,   # "01.2<sep>3<sync>"
           V + 14 + 6, V + 14 + 12, V + 14 + 7, V + 14,                            # "4.5<sync>"
           10, 11, 12, V,                                                          # caption <sync>
           V + 1 + 9, V + 1, V + 14 + 2, V + 14, 13, 2]                            # 2nd event closed by EOS
    ev = orc.parse_events(ids, lambda t: " ".join(map(str, t)), V)
    assert ev["timestamps"] == [[1.2, 3.0], [7.0]]
    assert ev["scores"] == [[4.5], [0.0]]
    assert ev["captions"] == ["10 11 12", "13 2"]


def test_oracle_forward_and_text_only_equal_live_reference():
    """Run the REAL reference here (when /root/reference is mounted) on the branches the golden files do not cover:
    `forward` logits at EVERY prompt position (with the `heads` mask and without) and the text-only path
    (`images_or_videos=None`, trace_mistral.py:306-307 / forward with input_ids only), and pin the oracle to them."""
    from oracle.ref_shim import load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = load_reference()
    d = synth.tiny_dims(4)
    model = mg.build_reference_model(ref, d, 6)
    o = orc.TraceOracle(d, synth.make_state_dict(d, 6))
    g = torch.Generator().manual_seed(3)
    rows = [[1] + torch.randint(3, d.vocab, (9,), generator=g).tolist() for _ in range(2)]
    ids = torch.tensor(rows, dtype=torch.long)
    with torch.no_grad():
        # the text-only prefill of generate: inputs_embeds = embed_tokens(inputs) (trace_mistral.py:306-307).  (forward with
        # multi-token input_ids and no images is NOT a valid reference call: the decode branch of
        # prepare_inputs_labels_for_multimodal uses Python `and` on tensors, trace_arch.py:359.)
        emb = model.get_model().embed_tokens(ids)
        for heads in ([1, 2], [0, 0], None):
            out = model(inputs_embeds=emb, attention_mask=torch.ones_like(ids), heads=None if heads is None else list(heads), return_dict=True)
            assert out.logits.dtype == torch.float32
            assert out.logits.shape[-1] == (d.total_vocab if heads is not None else d.vocab + 1)
            for b in range(2):
                e = torch.stack([o.embed_token(t) for t in rows[b]], 0)
                h, _ = o.llm_forward(e, None)
                for pos in (0, 4, 9):
                    want = out.logits[b, pos]
                    got = o.head_logits(h[pos], None if heads is None else heads[b])
                    fin = torch.isfinite(want)
                    assert torch.equal(torch.isfinite(got), fin)
                    np.testing.assert_allclose(got[fin].numpy(), want[fin].numpy(), atol=5e-4, rtol=1e-4)
        # text-only greedy: prefill through inputs_embeds, then one token at a time with the KV cache (HF contract)
        out = model(inputs_embeds=emb[:1], attention_mask=torch.ones(1, ids.shape[1], dtype=torch.long), use_cache=True, return_dict=True)
        want_ids = []
        mask = torch.ones(1, ids.shape[1], dtype=torch.long)
        for _ in range(5):
            tok = int(torch.argmax(out.logits[0, -1]))
            want_ids.append(tok)
            mask = torch.cat([mask, torch.ones(1, 1, dtype=torch.long)], 1)
            out = model(input_ids=torch.tensor([[tok]]), attention_mask=mask, past_key_values=out.past_key_values, use_cache=True, return_dict=True)
    assert o.generate(rows[0], None, None, head=None, max_new_tokens=5) == want_ids
