"""Tensor-level operators over the C ABI (include/fpq_b200.h).

Host-side plumbing only: argument checks, output allocation, the current CUDA stream.  All
arithmetic happens in the sm_100a kernels of fpqvar_b200/csrc.  CPU tensors are rejected --
there is no fallback path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib as L

_DT = {torch.float32: L.FPQ_F32, torch.float16: L.FPQ_F16}

_EXT = None


def _ext():
    """The torch extension over the C ABI (fpqvar_b200/dropin/quant_cuda*.so): the four hot operators do their argument
    checks, output allocation and launch in C++ (a few microseconds of host time per call).  No fallback: if it is not
    built, the hot operators raise."""
    global _EXT
    if _EXT is None:
        try:
            from .dropin import quant_cuda as ext
        except ImportError as e:
            raise L.FpqError("the torch extension fpqvar_b200/dropin/quant_cuda is not built (bash fpqvar_b200/csrc/build_torch_ext.sh, "
                             f"or __graft_entry__.build()): {e}") from None
        L.lib()                         # same library instance for the ctypes entries (launch counter, tunables)
        _EXT = ext
    return _EXT


def _call(fn, *args):
    try:
        return fn(*args)
    except RuntimeError as e:           # std::runtime_error / c10::Error from the extension
        if isinstance(e, L.FpqError):
            raise
        raise L.FpqError(str(e).split("\n")[0]) from None


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise L.FpqError(f"{what}: expected a CUDA tensor (fpqvar_b200 has no CPU fallback); got "
                         f"{getattr(x, 'device', type(x))}")


def _dt(x: torch.Tensor, what: str) -> int:
    try:
        return _DT[x.dtype]
    except KeyError:
        raise L.FpqError(f"{what}: dtype {x.dtype} is not supported (float16 / float32 only)") from None


try:                                    # raw handle of torch's current stream without building a Stream object (~0.3 us vs ~2 us)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                  # pragma: no cover - older torch
    _raw_stream = None


def _stream(device_index: Optional[int] = None) -> int:
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device() if device_index is None else device_index)
    return torch.cuda.current_stream(device_index).cuda_stream


class _on_device:
    """Device guard that costs nothing when the tensor already lives on the current device (the common case); the CUDA
    runtime launches on the calling thread's current device."""
    __slots__ = ("idx", "prev")

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self.idx

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _rows(x: torch.Tensor, row_len: Optional[int]) -> tuple[int, int]:
    if row_len is None:
        row_len = x.shape[-1] if x.dim() > 0 else 1
    n = x.numel()
    if row_len <= 0 or n % row_len != 0:
        raise L.FpqError(f"numel {n} is not a multiple of the group/row length {row_len}")
    return n // row_len, row_len


def _smooth_f32(smooth: Optional[torch.Tensor], n_cols: int, what: str) -> Optional[torch.Tensor]:
    """The GALT factor as the kernels take it: a float32 contiguous CUDA vector of `n_cols` entries.  Anything else is
    converted on the current stream (the kernels read it only after their programmatic-dependency wait, so a cast
    kernel launched right in front of them is ordered like any other producer)."""
    if smooth is None:
        return None
    _require_cuda(smooth, f"{what}(smooth)")
    if smooth.numel() != n_cols:
        raise L.FpqError(f"{what}: smooth has {smooth.numel()} entries, expected {n_cols}")
    smooth = smooth.detach()
    if smooth.dtype != torch.float32 or not smooth.is_contiguous():
        smooth = smooth.to(torch.float32).contiguous()
    return smooth.reshape(-1)


def fake_quant(x: torch.Tensor, fmt: str, row_len: Optional[int] = 128, tie: str = "kernel", clamp3: bool = False,
               out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Symmetric fake-quant (fpq_fake_quant). ``row_len=None``: the last dim shares one scale
    (per_token / per_channel).  Output dtype: the input's for tie="kernel" (the ``*_cuda``
    reference functions), float32 for tie="argmin" (quant_utils.py:308 promotes)."""
    if not isinstance(x, torch.Tensor):
        raise L.FpqError(f"fake_quant: expected a CUDA tensor; got {type(x)}")
    od = -1 if out_dtype is None else _DT.get(out_dtype)
    if od is None:
        raise L.FpqError(f"fake_quant: out dtype {out_dtype} is not supported (float16 / float32 only)")
    return _call(_ext().fake_quant, x, L.FMT[fmt], -1 if row_len is None else row_len, L.TIE[tie], L.FLAG_CLAMP3 if clamp3 else 0, od)


def fake_quant_segments_(buf: torch.Tensor, lo: int, hi: int, fmt: str, row_len: int) -> None:
    """In-place symmetric fake-quant (kernel tie rule) of ``buf[:, lo:hi]`` for a contiguous fp16 ``buf`` of shape
    [B, L, ...]: B segments of (hi - lo) * prod(shape[2:]) halves, L * prod(shape[2:]) apart, quantized where they lie in ONE
    launch (fpq_fake_quant_segments) -- no gather copy, no scatter copy.  ``row_len`` (64 | 128) halves share a scale."""
    _require_cuda(buf, "fake_quant_segments_")
    if buf.dtype != torch.float16 or not buf.is_contiguous() or buf.dim() < 2:
        raise L.FpqError("fake_quant_segments_: expected a contiguous float16 tensor [B, L, ...]")
    B, Lmax = buf.shape[0], buf.shape[1]
    if not (0 <= lo <= hi <= Lmax):
        raise L.FpqError(f"fake_quant_segments_: bad range [{lo}, {hi}) for length {Lmax}")
    inner = buf[0, 0].numel()
    seg = (hi - lo) * inner
    if seg % row_len:
        raise L.FpqError(f"fake_quant_segments_: {seg} elements per segment is not a multiple of the row length {row_len}")
    if seg == 0 or B == 0:
        return
    ptr = buf.data_ptr() + lo * inner * 2
    with _on_device(buf) as _di:
        rc = L.lib().fpq_fake_quant_segments(ptr, ptr, B, seg // row_len, row_len, Lmax * inner, Lmax * inner, L.FMT[fmt], _stream(_di))
    L.check(rc, "fpq_fake_quant_segments")


_WS = {}


def _clip_workspace(device: torch.device) -> torch.Tensor:
    """{flag, ticket} of FPQ_FLAG_GLOBAL_CLIP, one per (device, stream): the kernels leave it zeroed."""
    key = (device.index, _stream(device.index))
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.zeros(2, dtype=torch.int32, device=device)
    return ws


def fake_quant_signsplit(x: torch.Tensor, split_fmt: str, row_len: Optional[int] = 128, tie: str = "kernel",
                         global_clip: bool = False, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Sign-split fake-quant (fpq_fake_quant_signsplit).  ``global_clip=True`` reproduces the
    reference's whole-tensor ``clamp(x, -|x|max, |x|max)`` (quant_utils.py:421-422), which only
    matters when the tensor holds a NaN (then the whole output is zero)."""
    if not isinstance(x, torch.Tensor):
        raise L.FpqError(f"fake_quant_signsplit: expected a CUDA tensor; got {type(x)}")
    od = -1 if out_dtype is None else _DT.get(out_dtype)
    if od is None:
        raise L.FpqError(f"fake_quant_signsplit: out dtype {out_dtype} is not supported (float16 / float32 only)")
    ws = _clip_workspace(x.device) if (global_clip and x.is_cuda) else None
    return _call(_ext().fake_quant_signsplit, x, L.SPLIT[split_fmt], -1 if row_len is None else row_len, L.TIE[tie],
                 L.FLAG_GLOBAL_CLIP if global_clip else 0, ws, od)


def gelu_fake_quant_signsplit(x: torch.Tensor, split_fmt: str, global_clip: bool = False) -> torch.Tensor:
    """``fake_quant_signsplit(gelu(x, approximate="tanh"), split_fmt, 128, "kernel")`` in one pass over an fp16 tensor
    (fpq_gelu_fake_quant_signsplit): the reference's ``fc2.act_quant(self.act(self.fc1(x)))``, basic_var.py:108,120."""
    _require_cuda(x, "gelu_fake_quant_signsplit")
    if x.dtype != torch.float16:
        raise L.FpqError("gelu_fake_quant_signsplit: x must be float16 (the autocast output of fc1)")
    x = x.contiguous()
    n_rows, _ = _rows(x, 128)
    out = torch.empty_like(x)
    ws = _clip_workspace(x.device) if global_clip else None
    with _on_device(x) as _di:
        rc = L.lib().fpq_gelu_fake_quant_signsplit(x.data_ptr(), out.data_ptr(), n_rows, L.SPLIT[split_fmt], L.FLAG_GLOBAL_CLIP if global_clip else 0,
                                                   ws.data_ptr() if ws is not None else None, _stream(_di))
    L.check(rc, "fpq_gelu_fake_quant_signsplit")
    return out


def gelu_table() -> torch.Tensor:
    """fp16 [65536]: the library's GELU(tanh) of every fp16 bit pattern (fpq_selftest_gelu)."""
    t = torch.empty(65536, dtype=torch.float16, device="cuda")
    L.check(L.lib().fpq_selftest_gelu(t.data_ptr(), _stream()), "fpq_selftest_gelu")
    return t


def quant_grid(x: torch.Tensor, grid: torch.Tensor, tie: str = "kernel") -> torch.Tensor:
    """Nearest grid value, reference scan semantics (fpq_quant_grid). fp32 in, fp32 out."""
    _require_cuda(x, "quant_grid")
    _require_cuda(grid, "quant_grid(grid)")
    if x.dtype != torch.float32 or grid.dtype != torch.float32:
        raise L.FpqError("quant_grid: x and grid must be float32")
    x = x.contiguous()
    grid = grid.contiguous()
    z = torch.empty_like(x)
    with _on_device(x) as _di:
        rc = L.lib().fpq_quant_grid(x.data_ptr(), grid.data_ptr(), grid.numel(), x.numel(), z.data_ptr(), L.TIE[tie], _stream(_di))
    L.check(rc, "fpq_quant_grid")
    return z


def pack_sign_bits(signs) -> "ctypes.Array":
    """+-1 vector of length 128 -> the 4 x uint32 mask the kernels take (bit set = +1)."""
    vals = [float(v) for v in (signs.tolist() if hasattr(signs, "tolist") else signs)]
    if len(vals) != 128 or any(v not in (1.0, -1.0) for v in vals):
        raise L.FpqError("sign vector must hold 128 entries of +-1")
    words = [0, 0, 0, 0]
    for i, v in enumerate(vals):
        if v > 0:
            words[i >> 5] |= 1 << (i & 31)
    return (ctypes.c_uint32 * 4)(*words)


def transform_rotate_quant(x: torch.Tensor, smooth: Optional[torch.Tensor], sign_bits, fmt: Optional[str],
                           return_rotated: bool = False):
    """Fused ``(x * smooth) @ Q_block128`` -> fp16 -> per-group fake quant (fpq_transform_rotate_quant).
    x: fp32 [..., C]; returns fp16 of the same shape (and the pre-quantization rotated fp16
    tensor when ``return_rotated``).  ``fmt=None`` skips the quantizer."""
    if not isinstance(x, torch.Tensor):
        raise L.FpqError(f"transform_rotate_quant: expected a CUDA tensor; got {type(x)}")
    r = _call(_ext().transform_rotate_quant, x, smooth, list(sign_bits), -1 if fmt is None else L.FMT[fmt], return_rotated)
    return (r[0], r[1]) if return_rotated else r[0]


def modulate_transform_rotate_quant(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, smooth: Optional[torch.Tensor], sign_bits,
                                    fmt: Optional[str], return_rotated: bool = False):
    """Fused ``((x * (scale + 1) + shift) * smooth) @ Q_block128`` -> fp16 -> per-group fake quant
    (fpq_modulate_transform_rotate_quant).  x: fp32 [B, L, C] (LayerNorm output); scale, shift:
    broadcastable [B, 1, C] (the adaLN tensors of basic_var.py:258), fp32 or fp16.

    fp16 scale / shift is what the call site sees under the reference's fp16 autocast
    (evaluate_fp_quant_transform_rotate.py:195: ada_lin is an autocast Linear): there `scale.add(1)` is an
    fp16 add and only its rounded result meets the fp32 LayerNorm output.  That add is done here with the
    same ATen op on the tiny [B, 1, C] tensor and handed to the kernel as a gain (FPQ_MOD_GAIN)."""
    if not isinstance(x, torch.Tensor):
        raise L.FpqError(f"modulate_transform_rotate_quant: expected a CUDA tensor; got {type(x)}")
    r = _call(_ext().modulate_transform_rotate_quant, x, scale, shift, smooth, list(sign_bits), -1 if fmt is None else L.FMT[fmt], return_rotated)
    return (r[0], r[1]) if return_rotated else r[0]


def transform_rotate_weight(w: torch.Tensor, smooth: Optional[torch.Tensor], sign_bits, inplace: bool = False) -> torch.Tensor:
    """``(W / smooth).double() @ Q_block128`` -> fp32 (fpq_transform_rotate_weight)."""
    _require_cuda(w, "transform_rotate_weight")
    if w.dtype != torch.float32 or w.dim() != 2:
        raise L.FpqError("transform_rotate_weight: w must be a 2-D float32 tensor")
    if not w.is_contiguous():
        if inplace:
            raise L.FpqError("transform_rotate_weight: in-place needs a contiguous tensor")
        w = w.contiguous()
    smooth = _smooth_f32(smooth, w.shape[1], "transform_rotate_weight")
    out = w if inplace else torch.empty_like(w)
    with _on_device(w) as _di:
        rc = L.lib().fpq_transform_rotate_weight(w.data_ptr(), smooth.data_ptr() if smooth is not None else None, sign_bits,
                                                 out.data_ptr(), w.shape[0], w.shape[1], _stream(_di))
    L.check(rc, "fpq_transform_rotate_weight")
    return out


def _fmt_code(name: str) -> int:
    return L.FMT[name] if name in L.FMT else 16 + L.SPLIT[name]


def score_formats(x: torch.Tensor, formats: Sequence[str], tie: str = "kernel", sse: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum of squared quantization errors per candidate format, x read once (fpq_score_formats).
    Returns / accumulates into a float64 tensor of len(formats)."""
    _require_cuda(x, "score_formats")
    x = x.contiguous()
    if sse is None:
        sse = torch.zeros(len(formats), dtype=torch.float64, device=x.device)
    elif not (isinstance(sse, torch.Tensor) and sse.is_cuda and sse.device == x.device and sse.dtype == torch.float64
              and sse.is_contiguous() and sse.numel() >= len(formats)):
        raise L.FpqError(f"score_formats: sse must be a contiguous float64 tensor of >= {len(formats)} entries on {x.device}")
    codes = (ctypes.c_int * len(formats))(*[_fmt_code(f) for f in formats])
    n_rows, rl = _rows(x, 128)
    with _on_device(x) as _di:
        rc = L.lib().fpq_score_formats(x.data_ptr(), n_rows, rl, _dt(x, "score_formats"), codes, len(formats), L.TIE[tie],
                                       sse.data_ptr(), _stream(_di))
    L.check(rc, "fpq_score_formats")
    return sse


def sse_rows(a: torch.Tensor, b: torch.Tensor, row_weight: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out += sum_r w[r] * sum_c (a[r, c] - b[r, c])^2 in one read of both matrices (fpq_sse_rows): the output-level loss of
    the format search without a difference tensor.  a, b: [rows, C] of the same dtype (fp16 | fp32); row_weight: float64
    [rows] or None; returns / accumulates into a float64 scalar tensor."""
    _require_cuda(a, "sse_rows")
    _require_cuda(b, "sse_rows")
    if a.shape != b.shape or a.dtype != b.dtype or a.dim() != 2 or a.device != b.device:
        raise L.FpqError("sse_rows: a and b must be 2-D tensors of the same shape, dtype and device")
    a, b = a.contiguous(), b.contiguous()
    if row_weight is not None:
        if not (row_weight.is_cuda and row_weight.device == a.device and row_weight.dtype == torch.float64 and row_weight.is_contiguous()
                and row_weight.numel() == a.shape[0]):
            raise L.FpqError(f"sse_rows: row_weight must be a contiguous float64 tensor of {a.shape[0]} entries on {a.device}")
    if out is None:
        out = torch.zeros((), dtype=torch.float64, device=a.device)
    elif not (out.is_cuda and out.device == a.device and out.dtype == torch.float64 and out.numel() == 1):
        raise L.FpqError("sse_rows: out must be one float64 element on the input's device")
    with _on_device(a) as _di:
        rc = L.lib().fpq_sse_rows(a.data_ptr(), b.data_ptr(), a.shape[0], a.shape[1], _dt(a, "sse_rows"),
                                  row_weight.data_ptr() if row_weight is not None else None, out.data_ptr(), _stream(_di))
    L.check(rc, "fpq_sse_rows")
    return out


def selftest_rounding(fmt_code: int, tie: str) -> tuple[int, int]:
    """(mismatch count, first mismatching fp32 bit pattern) of closed form vs scan over all 2^32 inputs."""
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = L.lib().fpq_selftest_rounding(fmt_code, L.TIE[tie], res.data_ptr(), _stream())
    L.check(rc, "fpq_selftest_rounding")
    r = res.cpu()
    return int(r[0]), int(r[1]) & 0xFFFFFFFF


def selftest_f16_flow(fmt_code: int) -> tuple[int, int]:
    """(mismatch count, first mismatching (scale, x) index) of the packed fp16 element function vs the
    literal reference sequence over every possible (x, scale) fp16 pair."""
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    rc = L.lib().fpq_selftest_f16_flow(fmt_code, res.data_ptr(), _stream())
    L.check(rc, "fpq_selftest_f16_flow")
    r = res.cpu()
    return int(r[0]), int(r[1])


def launch_count() -> int:
    return int(L.lib().fpq_launch_count())
