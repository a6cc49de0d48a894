"""Import the *real* reference (gyxxyg/TRACE, /root/reference) under a shim.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) generate the golden fixtures in
``tests/golden`` (``tests/golden/make_golden.py``) and (b) let the CPU test-suite re-validate the
oracle restatement (``oracle/trace_oracle.py``) against the reference's own classes whenever
``/root/reference`` is present (it is absent on the GPU box).  Nothing in ``trace_b200/`` imports it.

Recipe = SURVEY.md §11: the repo is exposed as package ``Trace`` (drivers import ``Trace.trace.*``),
``transformers`` is imported first, then stub modules are injected for the optional third-party
imports the reference performs at module import time (timm, decord, imageio, moviepy, scenedetect).
The timm ``LayerNorm`` stub is ``nn.LayerNorm(eps=1e-6)`` (timm 0.6.13 default) -- see
multimodal_projector/builder.py:22,419.
"""
from __future__ import annotations

import functools
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("TRACE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trace", "model"))


_loaded = None


def load_reference():
    """Returns a namespace with the reference's classes (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    import torch
    import torch.nn as nn
    import transformers  # noqa: F401  (must precede the stubs: transformers probes `timm`)
    from transformers import CLIPVisionModel, MistralForCausalLM  # noqa: F401

    scratch = tempfile.mkdtemp(prefix="trace_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(scratch, "Trace"))
    sys.path.insert(0, scratch)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:  # placeholder for classes that are only referenced, never called here
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")

    stub("timm")
    stub("timm.models")
    stub("timm.models.regnet", RegStage=_Dummy)
    stub("timm.models.layers",
         LayerNorm=functools.partial(nn.LayerNorm, eps=1e-6),
         LayerNorm2d=_Dummy)
    bridge = types.SimpleNamespace(set_bridge=lambda *a, **k: None)
    stub("decord", VideoReader=_Dummy, cpu=lambda *a, **k: None, bridge=bridge)
    stub("imageio")
    stub("moviepy")
    stub("moviepy.editor", VideoFileClip=_Dummy)
    stub("scenedetect", open_video=_Dummy, SceneManager=_Dummy)
    stub("scenedetect.detectors", ContentDetector=_Dummy)
    stub("scenedetect.stats_manager", StatsManager=_Dummy)
    # removed in transformers>=5 but imported by multimodal_projector/builder.py:23.  NB: importing
    # model classes re-binds sys.modules['transformers'], so patch the live module object.
    tf_live = sys.modules["transformers"]
    if "TRANSFORMERS_CACHE" not in tf_live.__dict__:
        tf_live.TRANSFORMERS_CACHE = os.path.join(scratch, "hf_cache")

    from Trace.trace.model.language_model.trace_mistral import (  # type: ignore
        TraceMistralConfig, TraceMistralForCausalLM)
    from Trace.trace.model.multimodal_projector.builder import SpatialSlotPool  # type: ignore
    from Trace.trace.model.multimodal_encoder.time_encoder import TimeTower, TimeTokenizer  # type: ignore
    from Trace.trace.model.multimodal_encoder.score_encoder import ScoreTower, ScoreTokenizer  # type: ignore
    from Trace.trace import conversation as ref_conversation  # type: ignore
    from Trace.trace import mm_utils as ref_mm_utils  # type: ignore
    from Trace.trace import constants as ref_constants  # type: ignore

    _loaded = types.SimpleNamespace(
        torch=torch, scratch=scratch,
        TraceMistralConfig=TraceMistralConfig, TraceMistralForCausalLM=TraceMistralForCausalLM,
        SpatialSlotPool=SpatialSlotPool, TimeTower=TimeTower, TimeTokenizer=TimeTokenizer,
        ScoreTower=ScoreTower, ScoreTokenizer=ScoreTokenizer,
        conversation=ref_conversation, mm_utils=ref_mm_utils, constants=ref_constants)
    return _loaded
