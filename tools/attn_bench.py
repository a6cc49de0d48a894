"""Attention micro-benchmark: mma.sync vs tcgen05 at the bench shapes (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trace_b200 import engine as E

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for name, B, S, heads, kvh, hd, causal in (("vit", 32, 577, 16, 16, 64, False), ("prefill", 1, 2015, 32, 8, 128, True), ("prefill8", 8, 1119, 32, 8, 128, True)):
    qkv = torch.randn(B, S, (heads + 2 * kvh) * hd, device="cuda").half()
    fl = 4.0 * B * heads * S * S * hd * (0.5 if causal else 1.0)
    outs = {}
    for impl, fn in (("tcgen05", E.op_attention_tc),):
        outs[impl] = fn(qkv, heads, kvh, hd, causal)
        ms = timeit(lambda: fn(qkv, heads, kvh, hd, causal))
        print(f"attn {name:9s} {impl:9s} B={B} S={S} h={heads}/{kvh} d={hd}  {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")
    print(f"   max |tcgen05 - mma.sync| = {float((outs['tcgen05'].float() - outs['mma.sync'].float()).abs().max()):.4g}")
