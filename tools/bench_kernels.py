"""Micro-benchmarks of the tensor-core kernels at the bench workload's shapes (CUDA events, L2 flushed by cycling
through distinct buffers).  Prints TFLOP/s and fraction of MEASURED_PEAKS bf16_tflops."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from trace_b200 import engine as E  # noqa: E402

peak = 1702.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["bf16_tflops"]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def gemm(name, M, N, K, bias=False, res=False, act=0, swiglu=False):
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * 0.02).half()
    b = torch.randn(N, device="cuda").half() if bias else None
    r = torch.randn(M, N, device="cuda").half() if res else None
    ms = timeit(lambda: E.op_gemm(a, w, bias=b, residual=r, act=act, swiglu=swiglu))
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"gemm {name:28s} M={M:6d} N={N:6d} K={K:6d}  {ms * 1e3:9.1f} us  {tf:7.1f} TFLOP/s  {tf / peak:5.2f} of measured cuBLAS peak")


def attn(name, B, S, heads, kvh, hd, causal):
    qkv = torch.randn(B, S, (heads + 2 * kvh) * hd, device="cuda").half()
    fl = 4.0 * B * heads * S * S * hd * (0.5 if causal else 1.0)
    for impl, fn in (("tcgen05", E.op_attention_tc),):
        try:
            ms = timeit(lambda: fn(qkv, heads, kvh, hd, causal))
            print(f"attn {name:10s} {impl:9s} B={B} S={S} h={heads}/{kvh} d={hd}  {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")
        except Exception as ex:  # noqa: BLE001
            print(f"attn {name} {impl} failed: {ex}")


M = 32 * 577
gemm("vit qkv (+bias)", M, 3072, 1024, bias=True)
gemm("vit out (+bias+res)", M, 1024, 1024, bias=True, res=True)
gemm("vit fc1 (+bias+gelu)", M, 4096, 1024, bias=True, act=1)
gemm("vit fc2 (+bias+res)", M, 1024, 4096, bias=True, res=True)
gemm("vit patch embed", 32 * 576, 1024, 592)
S = 2015
gemm("prefill qkv", S, 6144, 4096)
gemm("prefill o (+res)", S, 4096, 4096, res=True)
gemm("prefill gate/up (swiglu)", S, 28672, 4096, swiglu=True)
gemm("prefill down (+res)", S, 4096, 14336, res=True)
attn("vit", 32, 577, 16, 16, 64, False)
attn("prefill", 1, 2015, 32, 8, 128, True)
