"""Run one generate() of the bench workload (for ncu launch lists / --set full captures).
    python tools/profile_step.py --frames 128 --max-new 4 [--warm 1]
Weights are random-synthetic (grammar irrelevant for profiling)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import L_TEXT, bench_prompt_ids  # noqa: E402
from trace_b200 import synth  # noqa: E402
from trace_b200.model import TraceMistralConfig, TraceMistralForCausalLM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--max-new", type=int, default=4)
ap.add_argument("--warm", type=int, default=0)
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
def box_hbm_gbs():
    x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del x, y
    return best


BOX_GBS = box_hbm_gbs()
d = synth.full_dims(a.frames)
S = 14 * a.frames + L_TEXT - 1
model = TraceMistralForCausalLM(TraceMistralConfig(d), device=0, vit_chunk_frames=32, max_prefill_tokens=S * a.batch,
                                max_ctx=S + a.max_new + 64, kv_pages=(a.batch + 1) * ((S + a.max_new + 127) // 64))
for name, shape, kind in synth.param_specs(d):
    model.engine.load_weight(name, synth.synth_tensor(name, shape, kind, 0, device="cuda"))
model.engine.load_state_dict([])
ids = torch.tensor([bench_prompt_ids(a.frames)] * a.batch, dtype=torch.int64, device="cuda")
clips, stamps = [], []
for b in range(a.batch):
    px, ts = synth.synthetic_clip(d, a.frames, seed=1234 + b)
    clips.append(px.cuda())
    stamps.append(ts)
for i in range(a.warm + 1):
    torch.cuda.synchronize()
    out = model.generate(ids, attention_mask=torch.ones_like(ids), images_or_videos=clips, modal_list=["video"] * a.batch,
                         video_timestamps=stamps, heads=[1] * a.batch, do_sample=False, max_new_tokens=a.max_new, pad_token_id=2)
torch.cuda.synchronize()
ms, st = model.engine.last_decode_timing()
import hashlib
print("ids", out.tolist()[0][:16], "sha", hashlib.sha1(str(out.tolist()).encode()).hexdigest()[:12])
if os.environ.get("TRACE_B200_MEGA_TIMING"):
    tl = model.engine.debug_timeline().double() / 1e3   # us
    names = ["qkv", "sync1", "attn", "sync2", "o", "sync3", "gateup", "sync4", "down"]
    for c in range(2):
        x = tl[c, 1:-1]
        d = x[:, 1:10] - x[:, 0:9]
        nxt = tl[c, 2:, 0] - tl[c, 1:-1, 9]
        print(f"cta{c}: " + " ".join(f"{n}={float(d[:, i].mean()):.2f}" for i, n in enumerate(names)) + f" sync0={float(nxt.mean()):.2f} layer={float((tl[c, 2:, 0] - tl[c, 1:-1, 0]).mean()):.2f} us")
        for nm, st0, k in (("qkv", 0, 10), ("o", 4, 12), ("gateup", 6, 14), ("down", 8, 16)):
            print(f"   {nm}: stage_x={float((x[:, k] - x[:, st0]).mean()):.2f} consume={float((x[:, k + 1] - x[:, k]).mean()):.2f} epilogue={float((x[:, st0 + 1] - x[:, k + 1]).mean()):.2f}")
if os.environ.get("TRACE_B200_MEGA_TIMING"):
    for c in range(2):
        x = tl[c, 1:-1]
        pts = [x[:, 2], x[:, 20], x[:, 21], x[:, 22], x[:, 23], x[:, 24], x[:, 3]]
        nm = ["q+rope", "K/V+scores", "softmax", "PV+reduce", "partials+atomic", "merge(last)"]
        print(f"   cta{c} attn: " + " ".join(f"{n}={float((pts[i + 1] - pts[i]).mean()):.2f}" for i, n in enumerate(nm)))
if os.environ.get("TRACE_B200_MEGA_TIMING"):
    raw = model.engine.debug_timeline()
    for c in range(2):
        print(f"   cta{c} stages parked in TMEM at phase start: " + " ".join(
            f"{n}={float(raw[c, 1:-1, 25 + i].double().mean()):.2f}" for i, n in enumerate(["qkv", "o", "gateup", "down"])))
print(f"DECODE_TIMING ms={ms:.2f} passes={st} ms_per_step={ms / max(st - 1, 1):.4f} box_copy_GBs={BOX_GBS:.0f} floor_ms_at_box_bw={14.0e9 / BOX_GBS / 1e6:.3f}")
