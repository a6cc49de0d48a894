"""ctypes binding of ``libtrace_b200.so`` (the C ABI declared in ``include/trace_b200.h``).

PyTorch is used for device memory and streams only: every call passes raw ``tensor.data_ptr()`` device
pointers plus explicit shapes, and the current CUDA stream handle.  There is no CPU or PyTorch fallback:
if the library is missing or the device is not sm_100 the import / constructor raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import synth

# TRACE_B200_LIB selects another in-tree build of the same ABI (A/B runs of two kernel versions on one box)
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ.get("TRACE_B200_LIB", "libtrace_b200.so"))
_lib = None

EXPORTS = [
    "trace_b200_abi_version", "trace_b200_last_error", "trace_b200_create", "trace_b200_destroy",
    "trace_b200_load_weight", "trace_b200_finalize_weights", "trace_b200_encode_frames",
    "trace_b200_build_prefill", "trace_b200_alloc_slot", "trace_b200_free_slot", "trace_b200_prefill",
    "trace_b200_decode", "trace_b200_batch_open", "trace_b200_batch_admit", "trace_b200_batch_step", "trace_b200_batch_read",
    "trace_b200_batch_retire", "trace_b200_batch_close", "trace_b200_set_sampling", "trace_b200_set_forced_ids", "trace_b200_decode_plan", "trace_b200_launch_count", "trace_b200_last_decode_timing", "trace_b200_weight_bytes", "trace_b200_debug_timeline", "trace_b200_op_gemm",
    "trace_b200_preprocess_frames", "trace_b200_op_layernorm", "trace_b200_op_rmsnorm", "trace_b200_op_attention_tc",
]


class TraceB200Error(RuntimeError):
    pass


class _Config(ctypes.Structure):
    _fields_ = [
        ("vit_width", ctypes.c_int32), ("vit_heads", ctypes.c_int32), ("vit_layers", ctypes.c_int32),
        ("vit_mlp", ctypes.c_int32), ("image_size", ctypes.c_int32), ("patch_size", ctypes.c_int32),
        ("vit_ln_eps", ctypes.c_float),
        ("hidden", ctypes.c_int32), ("layers", ctypes.c_int32), ("heads", ctypes.c_int32),
        ("kv_heads", ctypes.c_int32), ("head_dim", ctypes.c_int32), ("ffn", ctypes.c_int32),
        ("vocab", ctypes.c_int32), ("rope_theta", ctypes.c_float), ("rms_eps", ctypes.c_float),
        ("vit_chunk_frames", ctypes.c_int32), ("max_prefill_tokens", ctypes.c_int32),
        ("max_ctx", ctypes.c_int32), ("kv_pages", ctypes.c_int32),
    ]


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree library; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise TraceB200Error(f"{_LIB_PATH} not built -- run __graft_entry__.build() (there is no CPU fallback)")
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    lib.trace_b200_abi_version.restype = i32
    lib.trace_b200_last_error.restype = ctypes.c_char_p
    lib.trace_b200_create.argtypes = [ctypes.POINTER(_Config), i32, ctypes.POINTER(vp)]
    lib.trace_b200_destroy.argtypes = [vp]
    lib.trace_b200_destroy.restype = None
    lib.trace_b200_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32, vp]
    lib.trace_b200_finalize_weights.argtypes = [vp]
    lib.trace_b200_encode_frames.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.trace_b200_build_prefill.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.trace_b200_alloc_slot.argtypes = [vp, i32, ctypes.POINTER(i32)]
    lib.trace_b200_free_slot.argtypes = [vp, i32]
    lib.trace_b200_prefill.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), i32, vp, i32, vp, vp]
    lib.trace_b200_decode.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), i32, ctypes.POINTER(ctypes.c_int32), i32, i32,
                                      i32, ctypes.POINTER(i64), ctypes.POINTER(i32), vp, vp]
    lib.trace_b200_batch_open.argtypes = [vp, i32, i32, i32]
    lib.trace_b200_batch_admit.argtypes = [vp, i32, i32, vp, i32, i32, i32, vp]
    lib.trace_b200_batch_step.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), vp]
    lib.trace_b200_batch_read.argtypes = [vp, i32, ctypes.POINTER(i64), i32, ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_int32)]
    lib.trace_b200_batch_retire.argtypes = [vp, i32]
    lib.trace_b200_batch_close.argtypes = [vp]
    lib.trace_b200_set_sampling.argtypes = [vp, i32, f32, i32, ctypes.c_uint64]
    lib.trace_b200_set_forced_ids.argtypes = [vp, ctypes.POINTER(i64), i32, i32]
    lib.trace_b200_decode_plan.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_int32)]
    lib.trace_b200_launch_count.argtypes = [vp]
    lib.trace_b200_launch_count.restype = i64
    lib.trace_b200_last_decode_timing.argtypes = [vp, ctypes.POINTER(f32), ctypes.POINTER(i32)]
    lib.trace_b200_debug_timeline.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64), i32]
    lib.trace_b200_weight_bytes.argtypes = [vp]
    lib.trace_b200_weight_bytes.restype = i64
    lib.trace_b200_op_gemm.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp]
    lib.trace_b200_preprocess_frames.argtypes = [vp, i32, i32, i32, ctypes.POINTER(ctypes.c_int32), vp, vp, vp, vp,
                                                 ctypes.POINTER(f32), ctypes.POINTER(f32), vp, vp, vp]
    lib.trace_b200_op_layernorm.argtypes = [vp, vp, vp, f32, vp, i32, i32, vp]
    lib.trace_b200_op_rmsnorm.argtypes = [vp, vp, f32, vp, i32, i32, vp]
    lib.trace_b200_op_attention_tc.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp]
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise TraceB200Error(load_library().trace_b200_last_error().decode("utf-8", "replace"))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f16c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.float16 and t.is_contiguous(), "expected a contiguous CUDA fp16 tensor"
    return t


class Engine:
    """One context on one device.  Mirrors the life-cycle of the reference model object:
    construct (from_pretrained) -> encode / prefill / decode per generate() call."""

    def __init__(self, dims: synth.TraceDims, device: int = 0, vit_chunk_frames: int = 64,
                 max_prefill_tokens: int = 4096, max_ctx: int = 4096, kv_pages: Optional[int] = None):
        if not torch.cuda.is_available():
            raise TraceB200Error("trace_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.dims = dims
        self.device = device
        max_ctx = (max_ctx + 63) // 64 * 64
        if kv_pages is None:
            kv_pages = 8 * (max_ctx // 64)
        self.cfg = _Config(dims.vit_width, dims.vit_heads, dims.vit_layers, dims.vit_mlp, dims.image_size, dims.patch_size,
                           dims.vit_ln_eps, dims.hidden, dims.layers, dims.heads, dims.kv_heads, dims.head_dim, dims.ffn,
                           dims.vocab, dims.rope_theta, dims.rms_eps, vit_chunk_frames, max_prefill_tokens, max_ctx, kv_pages)
        self.max_ctx = max_ctx
        self.max_prefill_tokens = max_prefill_tokens
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _check(self.lib.trace_b200_create(ctypes.byref(self.cfg), device, ctypes.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.trace_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------
    def load_weight(self, name: str, tensor: torch.Tensor) -> None:
        t = _f16c(tensor.to(device=f"cuda:{self.device}", dtype=torch.float16).contiguous())
        shape = (ctypes.c_int64 * t.dim())(*t.shape)
        _check(self.lib.trace_b200_load_weight(self._h, name.encode(), t.data_ptr(), shape, t.dim(), _stream()))
        torch.cuda.current_stream().synchronize()   # `t` may be a temporary

    def load_state_dict(self, items: Iterable[Tuple[str, torch.Tensor]]) -> None:
        for name, t in items:
            self.load_weight(name, t)
        _check(self.lib.trace_b200_finalize_weights(self._h))

    # ---- stages --------------------------------------------------------------------------------
    def encode_frames(self, pixels: torch.Tensor, return_feats: bool = False):
        """[F,3,H,W] fp16 -> slots [F,8,hidden] fp16 (and CLIP features [F,P,D] when asked)."""
        px = _f16c(pixels)
        F = px.shape[0]
        d = self.dims
        assert tuple(px.shape[1:]) == (3, d.image_size, d.image_size), f"bad frame shape {tuple(px.shape)}"
        slots = torch.empty((F, synth.NUM_SLOTS, d.hidden), dtype=torch.float16, device=px.device)
        feats = torch.empty((F, d.patches, d.vit_width), dtype=torch.float16, device=px.device) if return_feats else None
        _check(self.lib.trace_b200_encode_frames(self._h, px.data_ptr(), F, slots.data_ptr(), _ptr(feats), _stream()))
        return (slots, feats) if return_feats else slots

    def build_prefill(self, ids: torch.Tensor, slots: Optional[torch.Tensor], time_ids: Optional[torch.Tensor]) -> torch.Tensor:
        """ids [B,L] int64 (one -201 per row), slots [B,T,8,H] fp16, time_ids [B,T,6] int32 -> embeds [B,S,H].
        slots=None / time_ids=None: text-only prompt, plain `embed_tokens(ids)` -> [B,L,H]."""
        assert ids.is_cuda and ids.dtype == torch.int64 and ids.is_contiguous()
        B, L = ids.shape
        if slots is None:
            out = torch.empty((B, L, self.dims.hidden), dtype=torch.float16, device=ids.device)
            _check(self.lib.trace_b200_build_prefill(self._h, ids.data_ptr(), B, L, 0, None, None, out.data_ptr(), _stream()))
            return out
        assert time_ids.is_cuda and time_ids.dtype == torch.int32 and time_ids.is_contiguous()
        T = time_ids.shape[1]
        S = L - 1 + self.dims.tokens_per_frame * T
        sl = _f16c(slots)
        out = torch.empty((B, S, self.dims.hidden), dtype=torch.float16, device=ids.device)
        _check(self.lib.trace_b200_build_prefill(self._h, ids.data_ptr(), B, L, T, sl.data_ptr(), time_ids.data_ptr(),
                                                 out.data_ptr(), _stream()))
        return out

    def alloc_slot(self, n_positions: int) -> int:
        s = ctypes.c_int(-1)
        _check(self.lib.trace_b200_alloc_slot(self._h, n_positions, ctypes.byref(s)))
        return s.value

    def free_slot(self, slot: int) -> None:
        _check(self.lib.trace_b200_free_slot(self._h, slot))

    def prefill(self, slots: Sequence[int], embeds: torch.Tensor, return_hidden: bool = False):
        e = _f16c(embeds)
        B, S, _ = e.shape
        arr = (ctypes.c_int32 * B)(*slots)
        hid = torch.empty_like(e) if return_hidden else None
        _check(self.lib.trace_b200_prefill(self._h, arr, B, e.data_ptr(), S, _ptr(hid), _stream()))
        return hid

    def set_sampling(self, do_sample: bool, temperature: float = 1.0, top_k: int = 50, seed: int = 0) -> None:
        """HF `_sample` with the default warpers (temperature, top-k) for the following decode() calls; see trace_b200.h."""
        _check(self.lib.trace_b200_set_sampling(self._h, int(bool(do_sample)), float(temperature), int(top_k), int(seed) & (2 ** 64 - 1)))

    def decode(self, slots: Sequence[int], heads: List[int], max_new: int, eos: int = 2, pad: int = 2,
               return_logits: bool = False, forced_ids=None):
        """Greedy (or sampled, see set_sampling) loop.  `heads` is mutated in place (as the reference mutates its `heads`
        list).  `forced_ids` [B, max_new]: teacher forcing -- the returned ids are still the engine's own choices.
        Returns ids [B, n] int64 (CPU) and optionally the masked logits [n, B, vocab+27] (CUDA fp32)."""
        B = len(slots)
        if forced_ids is not None:
            f = torch.as_tensor(forced_ids, dtype=torch.int64).contiguous()
            assert tuple(f.shape) == (B, max_new), f"forced ids must be [B, max_new], got {tuple(f.shape)}"
            arr = (ctypes.c_int64 * (B * max_new))(*f.flatten().tolist())
            _check(self.lib.trace_b200_set_forced_ids(self._h, arr, B, max_new))
        sl = (ctypes.c_int32 * B)(*slots)
        hd = (ctypes.c_int32 * B)(*heads)
        ids = (ctypes.c_int64 * (B * max_new))()
        n = ctypes.c_int(0)
        logits = None
        if return_logits:
            logits = torch.full((max_new, B, self.dims.total_vocab), float("nan"), dtype=torch.float32,
                                device=f"cuda:{self.device}")
        _check(self.lib.trace_b200_decode(self._h, sl, B, hd, max_new, eos, pad, ids, ctypes.byref(n), _ptr(logits), _stream()))
        for b in range(B):
            heads[b] = hd[b]
        out = torch.tensor(list(ids), dtype=torch.int64).view(B, max_new)[:, : n.value].contiguous()
        if return_logits:
            return out, logits[: n.value]
        return out

    # ---- continuous batching (include/trace_b200.h: trace_b200_batch_*) ------------------------------------
    def batch_open(self, rows: int, eos: int = 2, pad: int = 2) -> None:
        _check(self.lib.trace_b200_batch_open(self._h, rows, eos, pad))
        self._batch_rows = rows

    def batch_admit(self, row: int, slot: int, embeds: torch.Tensor, head: int, max_new: int) -> None:
        e = _f16c(embeds)
        assert e.dim() == 3 and e.shape[0] == 1, "one request per admission: embeds [1,S,hidden]"
        _check(self.lib.trace_b200_batch_admit(self._h, row, slot, e.data_ptr(), e.shape[1], head, max_new, _stream()))

    def batch_step(self, n_steps: int):
        """n decode steps for every row; returns (tokens emitted per row, EOS flag per row)."""
        B = self._batch_rows
        st, fin = (ctypes.c_int32 * B)(), (ctypes.c_int32 * B)()
        _check(self.lib.trace_b200_batch_step(self._h, n_steps, st, fin, _stream()))
        return list(st), list(fin)

    def batch_read(self, row: int, capacity: int):
        ids = (ctypes.c_int64 * capacity)()
        n, head = ctypes.c_int(0), ctypes.c_int32(0)
        _check(self.lib.trace_b200_batch_read(self._h, row, ids, capacity, ctypes.byref(n), ctypes.byref(head)))
        return torch.tensor(list(ids[: n.value]), dtype=torch.int64), head.value

    def batch_retire(self, row: int) -> None:
        _check(self.lib.trace_b200_batch_retire(self._h, row))

    def batch_close(self) -> None:
        _check(self.lib.trace_b200_batch_close(self._h))
        self._batch_rows = 0

    def decode_plan(self, B: int, context_len: int) -> dict:
        out = (ctypes.c_int32 * 8)()
        _check(self.lib.trace_b200_decode_plan(self._h, B, context_len, out))
        keys = ("nsplit", "ncomb", "kv_tiles_per_item", "kp_down", "ks_down", "stages", "group_tiles", "smem_bytes")
        return dict(zip(keys, list(out)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.trace_b200_launch_count(self._h))

    def last_decode_timing(self):
        """(milliseconds on the device, decoder passes) of the most recent decode() call."""
        ms, st = ctypes.c_float(0), ctypes.c_int(0)
        _check(self.lib.trace_b200_last_decode_timing(self._h, ctypes.byref(ms), ctypes.byref(st)))
        return ms.value, st.value

    def debug_timeline(self, per_cta: bool = False):
        """[2 CTAs][layers][32] %globaltimer stamps of the last decode step; with per_cta also [SMs][32] stamps of every CTA
        for the middle layer (slot 31 = %smid)."""
        n0 = 2 * self.dims.layers * 32
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        n = n0 + (sms * 32 if per_cta else 0)
        buf = (ctypes.c_uint64 * n)()
        _check(self.lib.trace_b200_debug_timeline(self._h, buf, n))
        t = torch.tensor(list(buf), dtype=torch.int64)
        if per_cta:
            return t[:n0].view(2, self.dims.layers, 32), t[n0:].view(sms, 32)
        return t.view(2, self.dims.layers, 32)

    @property
    def weight_bytes(self) -> int:
        return int(self.lib.trace_b200_weight_bytes(self._h))


# ---- single-kernel wrappers (tests / micro-benchmarks) ------------------------------------------------
def op_gemm(a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, act: int = 0, swiglu: bool = False) -> torch.Tensor:
    lib = load_library()
    a, w = _f16c(a), _f16c(w)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N // 2 if swiglu else N), dtype=torch.float16, device=a.device)
    _check(lib.trace_b200_op_gemm(a.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), out.shape[1], M, N, K, _ptr(bias),
                                  _ptr(residual), 0 if residual is None else residual.shape[1], act, int(swiglu), _stream()))
    return out


def op_layernorm(x, w, b, eps):
    lib = load_library()
    out = torch.empty_like(x)
    _check(lib.trace_b200_op_layernorm(_f16c(x).data_ptr(), w.data_ptr(), b.data_ptr(), eps, out.data_ptr(), x.shape[0], x.shape[1], _stream()))
    return out


def op_rmsnorm(x, w, eps):
    lib = load_library()
    out = torch.empty_like(x)
    _check(lib.trace_b200_op_rmsnorm(_f16c(x).data_ptr(), w.data_ptr(), eps, out.data_ptr(), x.shape[0], x.shape[1], _stream()))
    return out


def op_attention_tc(qkv: torch.Tensor, heads: int, kv_heads: int, head_dim: int, causal: bool) -> torch.Tensor:
    """tcgen05/TMEM attention: qkv [B, S, (heads + 2*kv_heads)*head_dim] fused -> out [B, S, heads*head_dim]."""
    lib = load_library()
    qkv = _f16c(qkv)
    B, S, W = qkv.shape
    out = torch.empty((B, S, heads * head_dim), dtype=torch.float16, device=qkv.device)
    _check(lib.trace_b200_op_attention_tc(qkv.data_ptr(), W, out.data_ptr(), S, heads, kv_heads, head_dim, int(causal), B,
                                          head_dim ** -0.5, _stream()))
    return out
