"""Pin the CPU oracle (oracle/trace_oracle.py) against the reference's own outputs (tests/golden,
made by make_golden.py from /root/reference) and the SURVEY §4 known-answer table."""
import numpy as np
import pytest
import torch

from oracle import trace_oracle as orc
from trace_b200 import synth
from tests.util import CASES, load_case, ranges


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
