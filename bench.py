#!/usr/bin/env python
"""bench.py -- events/sec end-to-end on synthetic 128-frame 336x336 clips (BASELINE.json metric, config 2).

One "step" = the whole hot path over one clip per GPU: CLIP-ViT-L/14-336 encode of 128 frames -> spatial-slot
projector -> prompt splice -> Mistral-7B prefill (S = 14*128 + 224 - 1 = 2015) -> greedy decode of 512 new tokens
with the three interleaved heads, through the reference-shaped surface `TraceMistralForCausalLM.generate`.

  value : whole-job events/s with the clip already resident in HBM when the timed region starts
  e2e   : the same metric with the clip in pinned HOST memory (H2D copy + D2H of the ids inside the timed region)
  roofline : the decode step (32 x {QKV, attention, O, gate/up, down} + head + finalize = one CUDA-graph replay),
             HBM-bound: algorithmic bytes per step / mean device time per step measured with CUDA events on the
             engine's decode stream inside the timed region, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the CPU oracle (port of the reference path) timed on the host cores on a bounded sample.

N > 1 (torchrun, one rank per GPU, NCCL): every rank owns one clip per step (weak scaling); each clip's 128
frames are split 128/N per rank through CLIP + projector and ONE all-gather of the projected slot tokens
rebuilds the clip on every rank; prefill + decode are replicated per clip.

`--impl reference` times the CPU oracle port (oracle/trace_oracle.py, a restatement of the reference's PyTorch
path; /root/reference itself does not exist on the GPU box) with all host threads on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

TOKENS_PER_EVENT = 40     # synth.TOKENS_PER_EVENT: 13 time + 4 score + 22 caption + 1 <sync>
L_TEXT = 224              # fixed prompt length
MAX_NEW = 512             # evaluate.py:449
# dram__bytes_read.sum + dram__bytes_write.sum of ONE decode_step_kernel launch (ncu, B=1, context ~2020), see profiles/
NCU_DRAM_BYTES_PER_LAUNCH = 14245.12e6 + 118.06e6
NCU_DRAM_SOURCE = "profiles/r01_ncu_full_decode_step_final.csv (ncu --set full, one decode_step_kernel<1> launch, context 2017)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--max-new", type=int, default=MAX_NEW)
    ap.add_argument("--weights", default="structured", choices=["structured", "random"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0))), "measured"
    return 6650.0, 1400.0, "fallback"


def bench_prompt_ids(T):
    """llama_2 template + <video>\\n + trace/prompts/dvc.txt + <sync>, stand-in tokenizer, padded to L_TEXT ids."""
    from trace_b200 import prompts, synth
    from trace_b200.conversation import conv_templates
    from trace_b200.mm_utils import tokenizer_MMODAL_token_all
    conv = conv_templates["llama_2"].copy()
    conv.append_message(conv.roles[0], "<video>\n" + prompts.DVC)
    conv.append_message(conv.roles[1], None)
    ids = tokenizer_MMODAL_token_all(conv.get_prompt() + "<sync>", synth.StandInTokenizer(), return_tensors="pt").tolist()
    filler = [3 + (7919 * i) % 31000 for i in range(max(0, L_TEXT - len(ids)))]
    ids = ids[:1] + filler + ids[1:]
    return ids[:L_TEXT - 1] + [ids[-1]] if len(ids) > L_TEXT else ids


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.q}", "--format=csv,noheader,nounits", "-i", str(self.gpu),
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_sample(T, max_new, threads):
    """Bounded sample of the same workload on the host cores with the oracle port, scaled to events/s.

    Sample: 2 frames through 1 CLIP block (x23 blocks x T/2), 1 decoder layer prefill at S=2015 (x32), 4 decode tokens
    through 1 decoder layer at t~2015 (x32 layers x max_new tokens) + 1 text-head evaluation (x text-head steps).
    Full widths, fp32, depth scaled linearly -- stated in the JSON as `sample`."""
    from oracle import trace_oracle as orc
    from trace_b200 import synth
    import dataclasses
    torch.set_num_threads(threads)
    full = synth.full_dims(T)
    d = dataclasses.replace(full, vit_layers=2, layers=1)          # 1 executed CLIP block, 1 decoder layer
    sd = {}
    for name, shape, kind in synth.param_specs(d):
        if name in ("model.embed_tokens.weight",):
            sd[name] = torch.zeros(shape)                          # gathers only; no need to randomise 131M values
        else:
            sd[name] = synth.synth_tensor(name, shape, kind, 0).float()
    o = orc.TraceOracle(d, sd)
    px, _ = synth.synthetic_clip(d, 2)
    # thread count: all host threads the path can USE -- torch's intra-op pool degrades badly when oversubscribed on
    # many-core boxes (128 threads on the GEMV-shaped decode step), so calibrate on one decode token and keep the best
    e1 = torch.randn(1, d.hidden) * 0.02
    _, c1 = o.llm_forward(torch.randn(64, d.hidden) * 0.02, None)
    best = (1e9, threads)
    for nt in sorted({n for n in (8, 16, 32, 64, threads) if n <= threads}):
        torch.set_num_threads(nt)
        o.llm_forward(e1, c1)
        t0 = time.perf_counter()
        for _ in range(3):
            o.llm_forward(e1, c1)
        dt = (time.perf_counter() - t0) / 3
        if dt < best[0]:
            best = (dt, nt)
    threads = best[1]
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    feats = o.vit_features(px.float())
    t_vit_block_2f = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.slot_pool(feats)
    t_proj_2f = time.perf_counter() - t0
    S = 14 * T + L_TEXT - 1
    e = torch.randn(S, d.hidden) * 0.02
    t0 = time.perf_counter()
    h, cache = o.llm_forward(e, None)
    t_prefill_layer = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(4):
        h1, cache = o.llm_forward(e[:1], cache)
    t_decode_layer_tok = (time.perf_counter() - t0) / 4
    t0 = time.perf_counter()
    o.head_logits(h1[-1], 0)
    t_head = time.perf_counter() - t0
    vit = t_vit_block_2f * (full.vit_layers - 1) * (T / 2)
    proj = t_proj_2f * (T / 2)
    prefill = t_prefill_layer * full.layers
    decode = (t_decode_layer_tok * full.layers) * max_new + t_head * max_new * (23.0 / TOKENS_PER_EVENT)
    total = vit + proj + prefill + decode
    events = max_new / TOKENS_PER_EVENT
    return {"value": events / total, "unit": "events/s", "cores": threads, "kind": "port",
            "sample": (f"oracle port fp32, full widths: 2 frames x 1 CLIP block ({t_vit_block_2f:.2f}s) x{full.vit_layers - 1} blocks x{T // 2}; "
                       f"1 decoder layer prefill S={S} ({t_prefill_layer:.2f}s) x{full.layers}; 4 decode tokens x 1 layer "
                       f"({t_decode_layer_tok * 1e3:.1f} ms/token/layer) x{full.layers} x{max_new}; text head {t_head * 1e3:.0f} ms; "
                       f"=> {total:.0f} s per {T}-frame clip ({vit:.0f} ViT + {prefill:.0f} prefill + {decode:.0f} decode)"),
            "seconds_per_clip": total}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(args.frames, args.max_new, threads)
        if i >= args.warmup:
            vals.append(r)
        if i == 0 and r["seconds_per_clip"] > 0 and args.warmup + args.steps > 2:
            pass
    v = sum(x["value"] for x in vals) / len(vals)
    ms = 1e3 * sum(x["seconds_per_clip"] for x in vals) / len(vals)
    line = {"impl": "reference", "metric": "events/sec end-to-end, 128-frame clip", "value": v, "unit": "events/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config 2: one {args.frames}-frame 336px clip, DVC prompt L={L_TEXT}, greedy {args.max_new} new tokens "
                                   f"(reference is serial batch-1: N GPUs do not apply)", "tokens_per_event": TOKENS_PER_EVENT},
            "cpu_baseline": {"value": v, "unit": "events/s", "cores": threads, "kind": "port", "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 leg
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from trace_b200 import synth
    from trace_b200.events import count_events
    from trace_b200.model import TraceMistralConfig, TraceMistralForCausalLM
    from trace_b200.parallel import attach_group

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    T, max_new = args.frames, args.max_new
    d = synth.full_dims(T)
    S = 14 * T + L_TEXT - 1
    model = TraceMistralForCausalLM(TraceMistralConfig(d), device=local, vit_chunk_frames=32, max_prefill_tokens=S,
                                    max_ctx=S + max_new + 64, kv_pages=2 * ((S + max_new + 64 + 63) // 64))
    t_load = time.perf_counter()
    cache = {}

    def get(name):
        if name not in cache:
            spec = next(s for s in synth.param_specs(d) if s[0] == name)
            cache[name] = synth.synth_tensor(spec[0], spec[1], spec[2], 0, device="cuda")
        return cache[name]

    over = synth.make_structured_overrides(d, 0, get) if args.weights == "structured" else {}
    cache.clear()
    for name, shape, kind in synth.param_specs(d):
        t = over.pop(name) if name in over else synth.synth_tensor(name, shape, kind, 0, device="cuda")
        model.engine.load_weight(name, t)
        del t
    model.engine.load_state_dict([])      # finalize (all tensors already loaded)
    torch.cuda.synchronize()
    t_load = time.perf_counter() - t_load
    if world > 1:
        attach_group(model)

    ids_row = bench_prompt_ids(T)
    assert len(ids_row) == L_TEXT, len(ids_row)
    ids = torch.tensor([ids_row], dtype=torch.int64, device="cuda")
    mask = torch.ones_like(ids)
    # one clip per rank per step; under sharding every rank needs every clip of the step (all ranks encode a slice of each)
    clips_host, stamps = [], []
    for r in range(world):
        px, ts = synth.synthetic_clip(d, T, seed=1234 + r)
        clips_host.append(px.pin_memory())
        stamps.append(ts)
    clips_dev = [c.cuda() for c in clips_host]
    h2d_bytes = sum(c.numel() * 2 for c in clips_host) + ids.numel() * 8
    eng = model.engine

    def step(resident: bool):
        """One clip per rank.  Sharded ViT: every rank encodes its frame range of all `world` clips, all-gather, then
        decodes its own clip."""
        srcs = clips_dev if resident else [c.cuda(non_blocking=True) for c in clips_host]
        heads = [1]
        if world == 1:
            out = model.generate(ids, attention_mask=mask, images_or_videos=[srcs[0]], modal_list=["video"], video_timestamps=[stamps[0]],
                                 heads=heads, do_sample=False, temperature=0.0, max_new_tokens=max_new, use_cache=True, pad_token_id=2)
        else:
            from trace_b200.parallel import encode_frames_sharded
            from trace_b200.model.encoders import frame_time_ids
            frames = torch.cat(srcs, 0)
            slots = encode_frames_sharded(eng, frames, world, T, model.dist_group).view(world, T, 8, d.hidden)
            tid = torch.tensor([frame_time_ids(stamps[rank])], dtype=torch.int32, device="cuda")
            emb = eng.build_prefill(ids, slots[rank:rank + 1].contiguous(), tid)
            kv = [eng.alloc_slot(S + max_new)]
            eng.prefill(kv, emb)
            out = eng.decode(kv, heads, max_new)
            eng.free_slot(kv[0])
        ms, st = eng.last_decode_timing()
        return out.cpu(), ms, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out, _, _ = step(True)
    launches0 = eng.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- timed region A: inputs resident in HBM ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dec_ms, dec_steps, n_tok, n_ev = 0.0, 0, 0, 0
    for _ in range(args.steps):
        out, ms, st = step(True)
        dec_ms += ms
        dec_steps += st
        n_tok += out.shape[1]
        n_ev += count_events(out[0].tolist(), d.vocab)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    # ---- timed region B: end to end from pinned host memory ----
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    n_ev_e2e = 0
    d2h = 0
    for _ in range(args.steps):
        out, _, _ = step(False)
        n_ev_e2e += count_events(out[0].tolist(), d.vocab)
        d2h += out.numel() * 8
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.finish() if sampler else None

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_ev, n_ev_e2e, n_tok], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    ev_total, ev_e2e, tok_total = float(cnt[0]), float(cnt[1]), float(cnt[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-stage device times outside the timed regions (explain the headline; not part of `value`) ----
    def timed(fn, reps=3):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    from trace_b200.model.encoders import frame_time_ids
    ms_vit = timed(lambda: eng.encode_frames(clips_dev[0]))
    slots0 = eng.encode_frames(clips_dev[0]).view(1, T, 8, d.hidden)
    tid0 = torch.tensor([frame_time_ids(stamps[0])], dtype=torch.int32, device="cuda")

    def prefill_once():
        emb = eng.build_prefill(ids, slots0, tid0)
        kv = [eng.alloc_slot(S + 8)]
        eng.prefill(kv, emb)
        eng.free_slot(kv[0])

    ms_prefill = timed(prefill_once)
    hbm_peak, tf_peak, peak_src = peaks()
    vit_flop = 366.0e9 * T                                          # 23 of 24 blocks + patch embed
    prefill_flop = 2 * 6.979e9 * S + 262144.0 * S * S
    # algorithmic bytes of one decode step at B=1: fp16 layer weights + head actually read + KV
    layer_w = d.layers * ((d.heads + 2 * d.kv_heads) * d.head_dim * d.hidden + d.heads * d.head_dim * d.hidden + 3 * d.ffn * d.hidden) * 2
    text_frac = 23.0 / TOKENS_PER_EVENT if args.weights == "structured" else 1.0
    head_b = (d.vocab + 1) * d.hidden * 2 * text_frac + 26 * d.hidden * 2 * (1 - text_frac)
    t_avg = S + max_new / 2
    kv_b = d.layers * 2 * d.kv_heads * d.head_dim * 2 * (t_avg + 1)
    bytes_step = layer_w + head_b + kv_b
    step_ms = dec_ms / max(dec_steps, 1)
    achieved = bytes_step / (step_ms * 1e-3) / 1e9
    value = ev_total / (ms_total * 1e-3)
    line = {
        "metric": "events/sec end-to-end, 128-frame clip", "value": value, "unit": "events/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": (f"config 2 per GPU: one {T}-frame 336px clip, CLIP-ViT-L/14-336 (23 blocks) + spatial-slot projector + "
                                f"Mistral-7B, llama_2+dvc prompt L={L_TEXT} (S={S}), greedy {max_new} new tokens, heads=[1]"),
                   "clips_per_step": world, "frames": T, "prompt_len": L_TEXT, "prefill_len": S, "max_new_tokens": max_new,
                   "weights": args.weights + " synthetic fp16 (seed 0)", "tokens_per_event": TOKENS_PER_EVENT,
                   "decode_path": "per-op kernels + CUDA graph" if os.environ.get("TRACE_B200_DECODE_PEROP") == "1" else "fused persistent decode-step kernel, 1 launch/token (CUDA graph)",
                   "parallelism": "1 GPU" if world == 1 else f"frame-sharded ViT x{world} + 1 all-gather, replicated LLM, 1 clip/GPU",
                   "l2_policy": "per-step working set (14.5 GB weights) >> 126 MB L2, no flush needed",
                   "weight_init_s": round(t_load, 1)},
        "events_parsed": ev_total, "tokens_generated": tok_total, "tokens_per_s": tok_total / (ms_total * 1e-3),
        "nominal_events_per_s": tok_total / TOKENS_PER_EVENT / (ms_total * 1e-3),
        "gpu_launches": int(launches),
        "e2e": {"value": ev_e2e / (ms_e2e * 1e-3), "unit": "events/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h // args.steps,
                "ms_per_step": ms_e2e / args.steps},
        "roofline": {"bound": "hbm", "kernel": "decode_step_kernel (one launch per generated token: 32 layers + heads + argmax/head-switch)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_DRAM_SOURCE, "bytes_per_step": bytes_step, "ms_per_decode_step": step_ms,
                     "decode_share_of_step": dec_ms / ms_total},
        "stages": {
            "vit_encode": {"ms": ms_vit, "tflops": vit_flop / (ms_vit * 1e-3) / 1e12, "frac_of_tensor_peak": vit_flop / (ms_vit * 1e-3) / 1e12 / tf_peak,
                           "tensor_peak_tflops": tf_peak, "flop": vit_flop, "note": "CLIP-ViT-L/14-336 x%d frames + slot pool, tcgen05 GEMMs; tensor-pipe %% per kernel from ncu in profiles/" % T},
            "prefill": {"ms": ms_prefill, "tflops": prefill_flop / (ms_prefill * 1e-3) / 1e12,
                        "frac_of_tensor_peak": prefill_flop / (ms_prefill * 1e-3) / 1e12 / tf_peak, "flop": prefill_flop, "tokens": S},
            "decode": {"ms_per_token": step_ms, "tokens": max_new},
        },
        "clocks": clocks,
    }
    if not args.skip_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_sample(T, max_new, os.cpu_count() or 1).items() if k != "seconds_per_clip"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
