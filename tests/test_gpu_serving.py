"""Batched configurations of BASELINE.json on the GPU (through the reference-shaped surface / C ABI) against the CPU
oracle: config 3 (batch of 8 clips, one decode loop), config 5 (mixed frame counts -> bucketed batches, one clip asked
several questions), and the text-only branch of generate (trace_mistral.py:306-307)."""
import pytest
import torch

from oracle import trace_oracle as orc
from trace_b200 import synth

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.06


def _model(d, seed, **kw):
    from trace_b200.model import TraceMistralConfig, TraceMistralForCausalLM
    m = TraceMistralForCausalLM(TraceMistralConfig(d), device=0, vit_chunk_frames=8, max_prefill_tokens=2048, max_ctx=512, **kw)
    m.load_state_dict(dict(synth.iter_state_dict(d, seed)))
    return m


def _check_row(got, oracle, ids, px, ts, head, max_new, label):
    """Greedy ids must equal the oracle's; a divergence is tolerated only where the oracle's own top-2 margin is below
    the fp16 logit tolerance (after which the row is no longer comparable)."""
    ref, trace, _ = oracle.generate(ids, None if px is None else px.float(), ts, head=head, max_new_tokens=max_new, return_trace=True)
    got = [int(x) for x in got]
    for s, (g, r) in enumerate(zip(got, ref)):
        if g != r:
            top2 = torch.topk(trace[s]["logits"], 2).values
            assert float(top2[0] - top2[1]) < LOGIT_TOL, f"{label}: step {s} got {g} want {r} (margin {float(top2[0] - top2[1]):.3g})"
            return s, 1
    assert len(got) >= len(ref), f"{label}: output shorter than the oracle's"
    return len(ref), 0


def _prompt(n_text, seed, with_video=True):
    g = torch.Generator().manual_seed(seed)
    body = torch.randint(3, 30000, (n_text,), generator=g).tolist()
    return ([1] + body[: n_text // 2] + [-201] + body[n_text // 2:] + [-205]) if with_video else ([1] + body)


def test_batch_of_eight_clips_matches_oracle_rowwise():
    """Config 3 shape at parity size: 8 clips x T frames in ONE prefill + decode loop (batch slot of the decode MMA full)."""
    d = synth.tiny_dims(4)
    model = _model(d, 3)
    o = orc.TraceOracle(d, synth.make_state_dict(d, 3))
    B, T, max_new = 8, 4, 20
    clips, stamps, prompts = [], [], []
    for b in range(B):
        px, ts = synth.synthetic_clip(d, T, seed=100 + b, duration_s=30.0 + 7 * b)
        clips.append(px)
        stamps.append(ts)
        prompts.append(_prompt(10, 50 + b))
    ids = torch.tensor(prompts, dtype=torch.int64, device="cuda")
    heads = [1] * B
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[c.cuda() for c in clips], modal_list=["video"] * B,
                         video_timestamps=stamps, heads=heads, do_sample=False, max_new_tokens=max_new, pad_token_id=2)
    assert out.shape[0] == B
    n = ties = 0
    for b in range(B):
        a, t = _check_row(out[b].tolist(), o, prompts[b], clips[b], stamps[b], 1, max_new, f"row {b}")
        n += a
        ties += t
    print(f"batch-8: {n} ids compared, {ties} fp16 tie-breaks")
    assert ties <= 2


def test_text_only_generate_matches_oracle():
    d = synth.tiny_dims(4)
    model = _model(d, 5)
    o = orc.TraceOracle(d, synth.make_state_dict(d, 5))
    prompts = [_prompt(14, 7, with_video=False), _prompt(14, 8, with_video=False)]
    ids = torch.tensor(prompts, dtype=torch.int64, device="cuda")
    out = model.generate(ids, attention_mask=torch.ones_like(ids), do_sample=False, max_new_tokens=12, pad_token_id=2)
    for b in range(2):
        _check_row(out[b].tolist(), o, prompts[b], None, None, None, 12, f"text row {b}")
    # with `heads` the same branch decodes from the masked 32027-wide logits
    heads = [0, 0]
    out2 = model.generate(ids, attention_mask=torch.ones_like(ids), do_sample=False, max_new_tokens=12, pad_token_id=2, heads=heads)
    for b in range(2):
        _check_row(out2[b].tolist(), o, prompts[b], None, None, 0, 12, f"text row {b} (heads)")


def test_scheduler_buckets_mixed_clips_and_reuses_encodes():
    """Config 5 shape: clips with different frame counts cannot share a batch (the reference cannot batch them at all);
    the scheduler buckets by (T, L), encodes every distinct clip once, and every request must come out exactly as a
    batch-1 generate of that request (the oracle is the arbiter for both)."""
    from trace_b200.serving import BatchScheduler, ClipRequest, plan_batches
    d = synth.tiny_dims(8)
    model = _model(d, 9)
    o = orc.TraceOracle(d, synth.make_state_dict(d, 9))
    clipA, tsA = synth.synthetic_clip(d, 8, seed=1, duration_s=64.0)
    clipB, tsB = synth.synthetic_clip(d, 4, seed=2, duration_s=20.0)
    clipC, tsC = synth.synthetic_clip(d, 8, seed=3, duration_s=90.0)
    devA, devB, devC = clipA.cuda(), clipB.cuda(), clipC.cuda()
    reqs = []
    spec = [(devA, clipA, tsA, 11), (devB, clipB, tsB, 12), (devA, clipA, tsA, 13), (devC, clipC, tsC, 14), (devA, clipA, tsA, 15),
            (devB, clipB, tsB, 16)]
    for k, (dev, _, ts, ps) in enumerate(spec):
        reqs.append(ClipRequest(torch.tensor(_prompt(10, ps)), dev, ts, head=1, max_new_tokens=16, tag=k))
    plan = plan_batches(reqs)
    assert sorted(sum(plan, [])) == list(range(len(reqs))) and len(plan) == 2
    sched = BatchScheduler(model)
    res = sched.run(reqs)
    assert [r.tag for r in res] == list(range(len(reqs)))
    assert sched.stats["clips_encoded"] == 3 and sched.stats["clip_encodes_saved"] == 3
    for k, (dev, host, ts, ps) in enumerate(spec):
        _check_row(res[k].ids.tolist(), o, _prompt(10, ps), host, ts, 1, 16, f"request {k}")
        assert res[k].head in (0, 1, 2)


def test_early_eos_rows_emit_pad_and_loop_stops():
    """HF greedy semantics: a row that emitted EOS keeps emitting pad_token_id while the other rows continue;
    the loop stops when every row has finished.  EOS is chosen from the oracle's own output so that row 0 stops early."""
    d = synth.tiny_dims(4)
    model = _model(d, 3)
    o = orc.TraceOracle(d, synth.make_state_dict(d, 3))
    clips, stamps, prompts, free_run = [], [], [], []
    for b in range(2):
        px, ts = synth.synthetic_clip(d, 4, seed=200 + b, duration_s=40.0)
        clips.append(px)
        stamps.append(ts)
        prompts.append(_prompt(10, 70 + b))
        free_run.append(o.generate(prompts[b], px.float(), ts, head=1, max_new_tokens=12))
    # an id that row 0 emits at step >= 2 and that row 1 does not emit before that step
    eos = step0 = None
    for s in range(2, 12):
        if free_run[0][s] not in free_run[1][: s + 1]:
            eos, step0 = free_run[0][s], s
            break
    if eos is None:
        pytest.skip("the two rows emit the same ids (degenerate synthetic weights)")
    ids = torch.tensor(prompts, dtype=torch.int64, device="cuda")
    heads = [1, 1]
    pad = 7
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[c.cuda() for c in clips], modal_list=["video"] * 2,
                         video_timestamps=stamps, heads=heads, do_sample=False, max_new_tokens=12, pad_token_id=pad, eos_token_id=eos)
    want1 = o.generate(prompts[1], clips[1].float(), stamps[1], head=1, max_new_tokens=12, eos=eos, pad=pad)
    row0, row1 = out[0].tolist(), out[1].tolist()
    first = row0.index(eos)
    assert first <= step0 and row0[:first + 1] == free_run[0][:first + 1]
    assert all(t == pad for t in row0[first + 1:]), row0
    assert row1[: len(want1)] == want1
    assert out.shape[1] == len(want1) or out.shape[1] == 12


def test_engine_errors_are_loud():
    from trace_b200.engine import Engine, TraceB200Error
    d = synth.tiny_dims(4)
    eng = Engine(d, device=0, vit_chunk_frames=4, max_prefill_tokens=256, max_ctx=128, kv_pages=3)
    eng.load_state_dict(synth.iter_state_dict(d, 0))
    with pytest.raises(TraceB200Error):
        eng.alloc_slot(129)                      # longer than max_ctx
    a = eng.alloc_slot(128)                      # 2 of the 3 pages
    with pytest.raises(TraceB200Error):
        eng.alloc_slot(128)                      # KV pool exhausted
    eng.free_slot(a)
    b = eng.alloc_slot(128)
    eng.free_slot(b)
    with pytest.raises(TraceB200Error):
        eng.prefill([99], torch.zeros(1, 4, d.hidden, dtype=torch.float16, device="cuda"))   # unallocated slot
    eng.close()
    e2 = Engine(d, device=0, vit_chunk_frames=4, max_prefill_tokens=256, max_ctx=128)
    with pytest.raises(TraceB200Error):          # weights missing: finalize must name the first missing tensor
        e2.load_state_dict([])
    e2.close()


@pytest.mark.parametrize("use_heads", [True, False])
def test_forward_full_sequence_logits(use_heads):
    """model.forward (trace_mistral.py:114-264): fp32 logits for every position of the spliced prompt, masked by `heads`
    (32027 wide) or text-only (32001 wide), against the oracle's head_logits at each position."""
    from trace_b200.model import TraceMistralConfig, TraceMistralForCausalLM
    d = synth.tiny_dims(4)
    model = TraceMistralForCausalLM(TraceMistralConfig(d), device=0, keep_head_weights=True, vit_chunk_frames=8, max_prefill_tokens=2048, max_ctx=512)
    model.load_state_dict(dict(synth.iter_state_dict(d, 4)))
    o = orc.TraceOracle(d, synth.make_state_dict(d, 4))
    px, ts = synth.synthetic_clip(d, 4, seed=31, duration_s=25.0)
    prompt = _prompt(10, 21)
    ids = torch.tensor([prompt, prompt], dtype=torch.int64, device="cuda")
    heads = [1, 2] if use_heads else None
    out = model.forward(ids, attention_mask=torch.ones_like(ids), images_or_videos=[px.cuda(), px.cuda()], modal_list=["video"] * 2,
                        video_timestamps=[ts, ts], heads=heads)
    logits = out.logits
    e = o.splice(prompt, o.encode_clip(px.float(), ts))
    h, _ = o.llm_forward(e, None)
    S = e.shape[0]
    assert logits.dtype == torch.float32 and tuple(logits.shape) == (2, S, d.total_vocab if use_heads else d.vocab + 1)
    for b in range(2):
        head = heads[b] if use_heads else None
        for pos in (0, 3, S // 2, S - 1):
            want = o.head_logits(h[pos], head)
            got = logits[b, pos].cpu()
            fin = torch.isfinite(want)
            assert torch.equal(torch.isfinite(got), fin), f"-inf mask differs at row {b} pos {pos}"
            assert float((got[fin] - want[fin]).abs().max()) < LOGIT_TOL, f"row {b} pos {pos}: {float((got[fin] - want[fin]).abs().max())}"
    with pytest.raises(NotImplementedError):
        model.forward(ids, labels=ids)


def test_safetensors_checkpoint_loads_and_generates(tmp_path):
    """load_pretrained_model on a real checkpoint directory (config.json + sharded *.safetensors with the HF names, the CLIP
    tower in its own `mm_vision_tower` directory as the released checkpoints have it, trace_arch.py:34-36) -> packed engine
    layout (fused QKV, interleaved gate/up, stacked heads) -> generate: ids must equal the reference's golden ids."""
    import json
    from safetensors.torch import save_file
    from tests.util import load_case, compare_ids_with_margin
    from trace_b200.model.builder import load_pretrained_model
    c = load_case("tiny_b1")
    d, meta, z = c["dims"], c["meta"], c["z"]
    sd = synth.make_state_dict(d, meta["seed"])
    vt = "model.vision_tower.vision_tower."
    clip_dir = tmp_path / "clip-vit-synth"
    clip_dir.mkdir()
    # the tower's own checkpoint: HF CLIPVisionModel names ("vision_model....")
    save_file({k[len(vt):]: v.contiguous() for k, v in sd.items() if k.startswith(vt)}, str(clip_dir / "model.safetensors"))
    json.dump({"hidden_size": d.vit_width, "num_attention_heads": d.vit_heads, "num_hidden_layers": d.vit_layers, "intermediate_size": d.vit_mlp,
               "image_size": d.image_size, "patch_size": d.patch_size, "layer_norm_eps": d.vit_ln_eps}, open(clip_dir / "config.json", "w"))
    ck = tmp_path / "trace-synth"
    ck.mkdir()
    rest = {k: v.contiguous() for k, v in sd.items() if not k.startswith(vt)}
    names = sorted(rest)
    half = len(names) // 2
    save_file({k: rest[k] for k in names[:half]}, str(ck / "model-00001-of-00002.safetensors"))
    save_file({k: rest[k] for k in names[half:]}, str(ck / "model-00002-of-00002.safetensors"))
    json.dump({"architectures": ["TraceMistralForCausalLM"], "model_type": "trace_mistral", "hidden_size": d.hidden, "intermediate_size": d.ffn,
               "num_hidden_layers": d.layers, "num_attention_heads": d.heads, "num_key_value_heads": d.kv_heads, "vocab_size": d.vocab,
               "rms_norm_eps": d.rms_eps, "rope_theta": d.rope_theta, "mm_vision_tower": str(clip_dir), "mm_projector_type": "spatial_slot",
               "num_frames": meta["T"], "image_aspect_ratio": "pad"}, open(ck / "config.json", "w"))
    tokenizer, model, processor, context_len = load_pretrained_model(str(ck), None, "trace-synth", vit_chunk_frames=4, max_prefill_tokens=1024,
                                                                     max_ctx=512)
    assert context_len == 2048 and processor is model.get_vision_tower().image_processor
    ids = torch.tensor(c["prompt_ids"], dtype=torch.int64, device="cuda")
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[c["clips"][0].cuda()], modal_list=["video"],
                         video_timestamps=c["ts"], heads=[1], do_sample=False, max_new_tokens=meta["max_new"], pad_token_id=2)
    n, ties = compare_ids_with_margin(out.cpu().tolist(), z["out_ids"].tolist(), z["top_vals"], LOGIT_TOL, "safetensors checkpoint")
    assert n > 0
    # a second generate on the SAME model with another budget and sampling switched on and off again (decode graph reuse)
    out2 = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[c["clips"][0].cuda()], modal_list=["video"],
                          video_timestamps=c["ts"], heads=[1], do_sample=False, max_new_tokens=7, pad_token_id=2)
    assert torch.equal(out2, out[:, :7])


def test_continuous_batching_mixed_requests_match_oracle():
    """Config-5 shape at parity size: more requests than decode rows, clips of 4 and 8 frames, prompts of different length,
    different heads and budgets, one clip asked twice -- all through ONE open decode batch with in-flight admission, every
    request against the oracle."""
    from trace_b200.serving import ClipRequest, ContinuousBatcher
    d = synth.tiny_dims(8)
    model = _model(d, 4)
    o = orc.TraceOracle(d, synth.make_state_dict(d, 4))
    clips = {}
    for k, T in enumerate([4, 8, 4, 8, 8]):
        clips[k] = synth.synthetic_clip(d, T, seed=300 + k, duration_s=20.0 + 11 * k)
    spec = [(0, 8, 1, 14), (1, 12, 1, 22), (2, 6, 0, 9), (3, 10, 2, 18), (1, 7, 1, 12), (4, 9, 1, 25), (0, 11, None, 10), (2, 8, 1, 16),
            (3, 6, 1, 7), (4, 13, 0, 20), (1, 9, 2, 11)]
    reqs, refs = [], []
    for n, (k, ntext, head, max_new) in enumerate(spec):
        px, ts = clips[k]
        ids = _prompt(ntext, 700 + n)
        reqs.append(ClipRequest(torch.tensor(ids), px, ts, head=head, max_new_tokens=max_new, tag=n))
    sched = ContinuousBatcher(model, rows=4, burst=5)
    res = sched.run(reqs)
    assert [r.tag for r in res] == list(range(len(spec)))
    assert sched.stats["clips_encoded"] == 5 and sched.stats["admitted"] == len(spec)
    total = ties = 0
    for n, (k, ntext, head, max_new) in enumerate(spec):
        px, ts = clips[k]
        got = res[n].ids.tolist()
        assert 1 <= len(got) <= max_new
        cnt, t = _check_row(got, o, reqs[n].input_ids.tolist(), px, ts, head, max_new, f"request {n}")
        total += cnt
        ties += t
        if t == 0 and 2 not in got:
            assert len(got) == max_new
    print(f"continuous batching: {total} ids compared, {ties} tie-breaks, stats {sched.stats}")
    # the engine is back in its normal state: a plain generate still works
    ids = torch.tensor([_prompt(8, 1)], dtype=torch.int64, device="cuda")
    px, ts = clips[0]
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=[px.cuda()], modal_list=["video"], video_timestamps=[ts],
                         heads=[1], do_sample=False, max_new_tokens=6, pad_token_id=2)
    _check_row(out[0].tolist(), o, ids[0].tolist(), px, ts, 1, 6, "generate after a continuous batch")
