"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/trace_b200.h declares (no compute without a GPU), and the host logic of the reference-shaped surface."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from trace_b200 import engine
    lib = engine.load_library()
    hdr = open(os.path.join(ROOT, "include", "trace_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(trace_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/trace_b200.h but not exported"
    assert sorted(engine.EXPORTS) == declared
    assert lib.trace_b200_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from trace_b200 import synth
    from trace_b200.engine import Engine, TraceB200Error
    with pytest.raises(TraceB200Error):
        Engine(synth.tiny_dims())


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "trace_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_time_score_tokenizers_known_answers():
    from trace_b200.model.encoders import ScoreTokenizer, TimeTokenizer, frame_time_ids
    t = TimeTokenizer()
    assert t.encode_numbers([12.3, 45.6]) == [2, 2, 3, 4, 12, 5, 1, 2, 2, 6, 7, 12, 8, 0]
    assert t.encode_numbers([9999.9]) == [11, 11, 11, 11, 12, 11, 0]
    assert t.encode_numbers([]) == [0]
    assert ScoreTokenizer().encode_numbers([4.5]) == [6, 12, 7, 0]
    assert [t.decode(i) for i in range(13)] == ["<sync>", "<sep>"] + list("0123456789") + ["."]
    assert frame_time_ids([[85 / 30.050377423088836]]) == [[2, 2, 2, 4, 12, 10]]
    with pytest.raises(AssertionError):
        frame_time_ids([[1.0], [10000.0]])
