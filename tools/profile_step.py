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
model = TraceMistralForCausalLM(TraceMistralConfig(d), device=0, vit_chunk_frames=64, max_prefill_tokens=S * a.batch,
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
    # decode_stream.cu stamps per layer: 10+2*type staged, 24+type consumed, 11+2*type epilogue written (type 0 qkv, 1 o, 2 gate/up,
    # 3 down); attention: 20 q ready, 21 K/V tiles done, 22 partials written, 23 combine written, 3 phase left
    tl = model.engine.debug_timeline().double() / 1e3   # us
    L = tl.shape[1]
    # whole-step accounting of the last step, CTA 0: entry -> prologue -> layer 0 .. L-1 -> head phase -> candidates -> finalize
    e = tl[0]
    t_entry, t_pro = float(e[0, 0]), float(e[0, 1])
    print(f"step CTA0: previous launch's end -> entry {float(e[0, 2]):.2f} us (launch gap + drain)")
    print(f"step CTA0: entry->prologue {t_pro - t_entry:.2f} us | prologue->layer0 qkv staged {float(e[0, 10]) - t_pro:.2f} | layer0 {float(e[0, 17] - e[0, 10]):.2f} | "
          f"layers 1..{L - 1} {float(e[L - 1, 17] - e[0, 17]):.2f} ({float(e[L - 1, 17] - e[0, 17]) / (L - 1):.2f} each) | head: wait x {float(e[L - 1, 18] - e[L - 1, 17]):.2f} "
          f"stream {float(e[L - 1, 28] - e[L - 1, 18]):.2f} epilogue {float(e[L - 1, 19] - e[L - 1, 28]):.2f} candidates {float(e[L - 1, 29] - e[L - 1, 19]):.2f} "
          f"finalize {float(e[L - 1, 30] - e[L - 1, 29]):.2f} | entry->end {float(e[L - 1, 30]) - t_entry:.2f} us; CTA77 entry->candidates {float(tl[1, L - 1, 29] - tl[1, 0, 0]):.2f}")
    for c in range(2):
        x = tl[c, 1:-1]            # layers 1 .. L-2 (layer 0 follows the prologue, the last one is overwritten by the head phase)
        prev_dn = tl[c, 0:-2, 17]
        seg = [("qkv.wait_x", prev_dn, x[:, 10]), ("qkv.stream", x[:, 10], x[:, 24]), ("qkv.epi", x[:, 24], x[:, 11]),
               ("attn.wait_q", x[:, 11], x[:, 20]), ("attn.tiles", x[:, 20], x[:, 21]), ("attn.merge", x[:, 21], x[:, 22]),
               ("attn.combine", x[:, 22], x[:, 23]), ("o.wait_attn", x[:, 23], x[:, 12]), ("o.stream", x[:, 12], x[:, 25]),
               ("o.epi", x[:, 25], x[:, 13]), ("gu.wait_x", x[:, 13], x[:, 14]), ("gu.stream", x[:, 14], x[:, 26]),
               ("gu.epi", x[:, 26], x[:, 15]), ("dn.wait_act", x[:, 15], x[:, 16]), ("dn.stream", x[:, 16], x[:, 27]),
               ("dn.epi", x[:, 27], x[:, 17])]
        per_layer = (x[:, 17] - prev_dn)
        print(f"   cta{c} per-layer us: " + " ".join(f"{float(v):.0f}" for v in per_layer) + f" | o.wait_attn per layer: " + " ".join(f"{float(v):.0f}" for v in (x[:, 12] - x[:, 23])))
        tot = float((x[:, 17] - prev_dn).mean())
        print(f"cta{c}: layer={tot:.2f} us | " + " ".join(f"{n}={float((b - a_).mean()):.2f}" for n, a_, b in seg))
        stream = sum(float((b - a_).mean()) for n, a_, b in seg if n.endswith(".stream"))
        wait = sum(float((b - a_).mean()) for n, a_, b in seg if ".wait" in n)
        print(f"   cta{c}: streaming {stream:.2f} us, exchange waits {wait:.2f} us, attention compute {float((x[:, 23] - x[:, 20]).mean()):.2f} us, "
              f"epilogues {sum(float((b - a_).mean()) for n, a_, b in seg if n.endswith('.epi')):.2f} us")
if os.environ.get("TRACE_B200_MEGA_TIMING"):
    _, pc = model.engine.debug_timeline(per_cta=True)
    pc = pc.double()
    import json as _json
    rec = {"smid": pc[:, 31].long().tolist()}
    dur = [round(float(v) / 1e3, 1) for v in pc[:, 4]][:a.max_new]
    gap = [round(float(v) / 1e3, 1) for v in pc[:, 5]][:a.max_new]
    print("per-step entry->end us:", dur)
    print("per-step gap before us:", gap)
    t0 = float(pc[:, 10].min())
    for nm, k in (("qkv_staged", 10), ("qkv_done", 24), ("o_staged", 12), ("o_done", 25), ("gu_staged", 14), ("gu_done", 26),
                  ("dn_staged", 16), ("dn_done", 27)):
        rec[nm] = [round((float(v) - t0) / 1e3, 3) for v in pc[:, k]]
    for a_, b_ in (("gu_staged", "gu_done"), ("dn_staged", "dn_done")):
        dur = [y - x for x, y in zip(rec[a_], rec[b_])]
        srt = sorted(dur)
        print(f"per-CTA {a_}->{b_}: min {srt[0]:.2f} p10 {srt[len(srt) // 10]:.2f} median {srt[len(srt) // 2]:.2f} p90 {srt[9 * len(srt) // 10]:.2f} max {srt[-1]:.2f} us")
    for nm in ("gu_done", "dn_done", "qkv_done", "o_done"):
        srt = sorted(rec[nm])
        print(f"per-CTA {nm} finish spread: first {srt[0]:.2f} median {srt[len(srt) // 2]:.2f} p90 {srt[9 * len(srt) // 10]:.2f} last {srt[-1]:.2f} us")
    os.makedirs("gpurun_out", exist_ok=True)
    _json.dump(rec, open(f"gpurun_out/per_cta_{os.environ.get('AB_NAME', 'run')}.json", "w"))
print(f"DECODE_TIMING ms={ms:.2f} passes={st} ms_per_step={ms / max(st - 1, 1):.4f} box_copy_GBs={BOX_GBS:.0f} floor_ms_at_box_bw={14.0e9 / BOX_GBS / 1e6:.3f}")
