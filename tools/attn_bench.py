"""Attention micro-benchmark: the tcgen05 kernel at the bench shapes (CUDA events), checked against fp32 torch."""
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
    out = E.op_attention_tc(qkv, heads, kvh, hd, causal)
    ms = timeit(lambda: E.op_attention_tc(qkv, heads, kvh, hd, causal), n=20)
    print(f"attn {name:9s} tcgen05 B={B} S={S} h={heads}/{kvh} d={hd}  {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")
    # fp32 torch reference on a slice of the batch
    nb = min(B, 2)
    q, k, v = qkv[:nb].float().split([heads * hd, kvh * hd, kvh * hd], dim=-1)
    q = q.view(nb, S, heads, hd).transpose(1, 2)
    k = k.view(nb, S, kvh, hd).transpose(1, 2).repeat_interleave(heads // kvh, dim=1)
    v = v.view(nb, S, kvh, hd).transpose(1, 2).repeat_interleave(heads // kvh, dim=1)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(nb, S, heads * hd)
    print(f"   max |tcgen05 - fp32 torch| = {float((out[:nb].float() - ref).abs().max()):.4g}  (max |ref| {float(ref.abs().max()):.3g})")
