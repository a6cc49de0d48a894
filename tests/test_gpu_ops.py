"""Single-kernel parity (through the C ABI) against plain fp32 PyTorch restatements of the same op.
Floating-point kernels: the tolerance is stated per test (fp16 inputs/outputs, fp32 accumulation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).half()


def _close(got, want, rtol, atol, what):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    bad = err > bound
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; max abs err "
                                 f"{float(err.max()):.4g}, max |want| {float(want.abs().max()):.4g}, "
                                 f"first bad index {bad.nonzero()[0].tolist()}")


@pytest.mark.parametrize("M,N,K", [(300, 384, 128), (67, 6144, 512), (1000, 1024, 592), (5000, 1024, 1024),
                                   (129, 512, 4096), (16, 4096, 128), (2015, 1024, 4096)])
def test_gemm_plain(M, N, K):
    from trace_b200 import engine as E
    a, w = _rand((M, K), 1.0, 1), _rand((N, K), 0.05, 2)
    out = E.op_gemm(a, w)
    torch.cuda.synchronize()
    want = a.float() @ w.float().t()
    _close(out, want, 2e-3, 2e-3 * float(want.abs().max()), f"gemm {M}x{N}x{K}")


def test_gemm_epilogues():
    from trace_b200 import engine as E
    M, N, K = 777, 1024, 256
    a, w, bias, res = _rand((M, K), 1.0, 3), _rand((N, K), 0.05, 4), _rand((N,), 0.5, 5), _rand((M, N), 1.0, 6)
    base = a.float() @ w.float().t() + bias.float()
    _close(E.op_gemm(a, w, bias=bias), base, 2e-3, 4e-3, "bias")
    _close(E.op_gemm(a, w, bias=bias, act=1), base * torch.sigmoid(1.702 * base), 2e-3, 4e-3, "quick_gelu")
    _close(E.op_gemm(a, w, bias=bias, residual=res), base + res.float(), 2e-3, 4e-3, "residual")
    out = res.clone()
    # in-place residual (out aliases residual) as the engine uses it
    import ctypes
    lib = E.load_library()
    rc = lib.trace_b200_op_gemm(a.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, bias.data_ptr(), out.data_ptr(), N, 0, 0,
                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.trace_b200_last_error()
    _close(out, base + res.float(), 2e-3, 4e-3, "in-place residual")
    g = a.float() @ w.float().t()
    want = torch.nn.functional.silu(g[:, 0::2]) * g[:, 1::2]
    _close(E.op_gemm(a, w, swiglu=True), want, 2e-3, 4e-3, "swiglu")


def test_gemm_wide_tiles_all_epilogues():
    """Shapes that take the 256x256 two-CTA (cta_group::2) path: ragged M, every epilogue."""
    from trace_b200 import engine as E
    M, N, K = 5003, 2048, 320
    a, w, bias, res = _rand((M, K), 1.0, 31), _rand((N, K), 0.05, 32), _rand((N,), 0.5, 33), _rand((M, N), 1.0, 34)
    base = a.float() @ w.float().t()
    _close(E.op_gemm(a, w), base, 2e-3, 4e-3, "wide plain")
    _close(E.op_gemm(a, w, bias=bias, act=1), (base + bias.float()) * torch.sigmoid(1.702 * (base + bias.float())), 2e-3, 4e-3, "wide gelu")
    _close(E.op_gemm(a, w, bias=bias, residual=res), base + bias.float() + res.float(), 2e-3, 6e-3, "wide residual")
    _close(E.op_gemm(a, w, swiglu=True), torch.nn.functional.silu(base[:, 0::2]) * base[:, 1::2], 2e-3, 4e-3, "wide swiglu")
    M, N, K = 18464, 1024, 1024
    a, w = _rand((M, K), 1.0, 35), _rand((N, K), 0.03, 36)
    _close(E.op_gemm(a, w), a.float() @ w.float().t(), 2e-3, 6e-3, "vit-shaped")


def test_norms():
    from trace_b200 import engine as E
    for D in (128, 192, 1024, 4096):     # 128 / 192: one-CTA-per-row kernel; 1024 / 4096: warp-per-row kernel
        x, w, b = _rand((333, D), 2.0, 7), (1 + 0.1 * _rand((D,), 1.0, 8).float()).half(), _rand((D,), 0.1, 9)
        if D <= 1024:
            want = torch.nn.functional.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5)
            _close(E.op_layernorm(x, w, b, 1e-5), want, 2e-3, 2e-3, f"layernorm D={D}")
        xf = x.float()
        want = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).half().float()
        _close(E.op_rmsnorm(x, w, 1e-5), want, 1e-3, 1e-3, f"rmsnorm D={D}")


@pytest.mark.parametrize("B,S,heads,kvh,hd,causal", [(3, 577, 2, 2, 64, False), (2, 17, 2, 2, 64, False),
                                                     (2, 200, 32, 8, 128, True), (1, 67, 32, 8, 128, True),
                                                     (1, 1000, 8, 2, 128, True)])
def test_attention(B, S, heads, kvh, hd, causal):
    from trace_b200 import engine as E
    qkv = _rand((B, S, (heads + 2 * kvh) * hd), 1.0, 11)
    out = E.op_attention_tc(qkv, heads, kvh, hd, causal)
    torch.cuda.synchronize()
    q = qkv[..., : heads * hd].float().view(B, S, heads, hd).transpose(1, 2)
    k = qkv[..., heads * hd: (heads + kvh) * hd].float().view(B, S, kvh, hd).transpose(1, 2)
    v = qkv[..., (heads + kvh) * hd:].float().view(B, S, kvh, hd).transpose(1, 2)
    k, v = k.repeat_interleave(heads // kvh, 1), v.repeat_interleave(heads // kvh, 1)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(S, S, device="cuda", dtype=torch.bool), 1), float("-inf"))
    want = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, heads * hd)
    _close(out, want, 5e-3, 5e-3, f"attention B{B} S{S} h{heads}/{kvh} d{hd} causal={causal}")


@pytest.mark.parametrize("S,hd,causal", [(577, 64, False), (900, 128, True)])
def test_attention_tcgen05_lazy_rescale(S, hd, causal):
    """Peaky scores (|s| up to ~40): the running maximum of a row grows by more than 2^8 between key tiles, so the
    TMEM-resident O accumulator and the row sum are rescaled in place (tcgen05.ld -> mul -> tcgen05.st); rows whose
    maximum moves by less keep their stale reference.  Both paths must agree with fp32 softmax."""
    from trace_b200 import engine as E
    heads, kvh, B = 4, 4, 2
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn((B, S, (heads + 2 * kvh) * hd), generator=g, device="cuda")
    qkv[..., : heads * hd] *= 3.0                       # q
    qkv[..., heads * hd: (heads + kvh) * hd] *= 3.0     # k
    # make late keys win for half of the rows: the maximum keeps growing along the key axis
    qkv[:, S // 2:, heads * hd: (heads + kvh) * hd] *= 1.5
    qkv = qkv.half()
    out = E.op_attention_tc(qkv, heads, kvh, hd, causal)
    torch.cuda.synchronize()
    q = qkv[..., : heads * hd].float().view(B, S, heads, hd).transpose(1, 2)
    k = qkv[..., heads * hd: (heads + kvh) * hd].float().view(B, S, kvh, hd).transpose(1, 2)
    v = qkv[..., (heads + kvh) * hd:].float().view(B, S, kvh, hd).transpose(1, 2)
    s_ = (q @ k.transpose(-1, -2)) * hd ** -0.5
    if causal:
        s_ = s_.masked_fill(torch.triu(torch.ones(S, S, device="cuda", dtype=torch.bool), 1), float("-inf"))
    want = (torch.softmax(s_, -1) @ v).transpose(1, 2).reshape(B, S, heads * hd)
    _close(out, want, 1e-2, 1e-2, f"attention tcgen05 lazy rescale S{S} d{hd}")


@pytest.mark.parametrize("h,w,pad", [(360, 640, True), (480, 270, True), (336, 336, True), (240, 426, False), (1080, 1920, True)])
def test_device_preprocess_matches_pil_chain(h, w, pad):
    """expand2square + PIL bicubic resize + crop + normalise + fp16 on the device against the numpy restatement that is
    pinned to Pillow / the HF PIL processor on CPU (tests/test_preprocess_cpu.py).  Integer resampling is bit-exact and
    the float32 normalisation uses the same operation order, so the fp16 tensors must be identical."""
    from oracle import preprocess_oracle as po
    from trace_b200.preprocess import preprocess_frames_cuda
    rng = np.random.default_rng(h + w)
    frames = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    frames[1, : h // 2] = 255                                    # saturated region: bicubic overshoot must clip at 255 / 0
    frames[1, h // 2:] = 0
    got = preprocess_frames_cuda(torch.from_numpy(frames).cuda(), 336, aspect_ratio="pad" if pad else "none").cpu()
    assert got.shape == (3, 3, 336, 336) and got.dtype == torch.float16
    for f in range(3):
        want = torch.from_numpy(po.preprocess_frame(frames[f], 336, pad=pad)).half()
        assert torch.equal(got[f], want), f"frame {f}: {(got[f] != want).sum().item()} of {want.numel()} fp16 values differ"
