"""ORACLE (test infrastructure only; nothing under fpqvar_b200/ may import this).

CPU statement of the packed low-bit operand format and of the low-bit GEMM of include/fpq_b200.h (fpq_pack_codes,
fpq_unpack_codes, fpq_codes_to_nibbles, fpq_gemm_codes).  The reference has no such operator -- QuantizedLinear.forward
(models_fp_quant_transform_rotate/quant_utils.py:764-769) multiplies fp16 tensors whose values are scale * grid value -- so
the statement is anchored on the reference's quantizer: `quantize_codes` returns exactly the (grid value, scale) pair that
oracle.fake_quant (fp_quant_*_per_group_cuda, qu.py:265-378, 537-574) multiplies in its last step, and
`dequantize(quantize_codes(x)) == fake_quant(x)` bit for bit is the first test (tests/test_oracle_gemm_codes.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
GROUP = 128
TILE_ROWS = 128


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_gemm.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_gemm.so"])
        lib = ctypes.CDLL(so)
        lib.gemm_codes_ref.restype = None
        lib.gemm_codes_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def rows_padded(rows: int) -> int:
    return (rows + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS


def quantize_codes(x: np.ndarray, fmt: str, scale_group=GROUP):
    """(q, s): q fp32 [rows, k] grid values, s fp32 [rows, k/scale_group] scales, as fp_quant_*_per_group_cuda computes them
    (qu.py:320-326: scale = absmax / max|grid| in the input dtype, x / scale in the input dtype, quant_cuda.quant);
    scale_group=None: one scale per row, the per_token / per_channel functions (qu.py:503-534)."""
    x = np.asarray(x)
    assert x.ndim == 2 and x.shape[1] % GROUP == 0 and x.dtype in (np.float16, np.float32)
    dt = x.dtype
    grid = O.GRIDS[fmt]
    scale_group = x.shape[1] if scale_group is None else scale_group
    g = O._as_groups(x, scale_group)
    a = O._absmax_lastdim(g)
    s = O._div(a, np.full_like(a, O.grid_absmax(fmt), dtype=np.float32), dt)
    v = O._div(g, s, dt)
    q = O.scan_quant(v.astype(np.float32), grid).reshape(x.shape)
    return q.astype(np.float32), s.astype(np.float32).reshape(x.shape[0], x.shape[1] // scale_group)


def dequantize(q: np.ndarray, s: np.ndarray, dtype) -> np.ndarray:
    """qu.py:328-329: quantized * scale in fp32, cast to the output dtype."""
    with np.errstate(invalid="ignore", over="ignore", under="ignore"):
        out = q.reshape(q.shape[0], s.shape[1], -1) * s[:, :, None]
        return out.reshape(q.shape).astype(dtype)


def e4m3_encode(q: np.ndarray) -> np.ndarray:
    """e4m3 (bias 7, 3 mantissa bits, no infinities) byte of every value; asserts that the value is representable --
    true for every grid of the reference (qu.py:233-235, 458-486)."""
    q = np.asarray(q, dtype=np.float32)
    sign = (np.signbit(q)).astype(np.uint8) << 7
    a = np.abs(q).astype(np.float64)
    out = np.zeros(q.shape, dtype=np.uint8)
    nz = a > 0
    m, e = np.frexp(a[nz])                      # a = m * 2^e, m in [0.5, 1)
    e = e - 1                                   # a = (2m) * 2^e, 2m in [1, 2)
    normal = e >= -6
    mant = np.where(normal, (2 * m - 1) * 8, a[nz] * 512)        # subnormals: multiples of 2^-9
    assert np.all(mant == np.round(mant)) and np.all(e <= 8), "value is not an e4m3 number"
    byte = np.where(normal, ((e + 7).astype(np.int64) << 3) | mant.astype(np.int64), mant.astype(np.int64))
    out[nz] = byte.astype(np.uint8)
    return out | sign


def e4m3_decode(b: np.ndarray) -> np.ndarray:
    b = np.asarray(b, dtype=np.uint8)
    e = ((b >> 3) & 0xF).astype(np.int32)
    m = (b & 7).astype(np.float32)
    mag = np.where(e == 0, m * np.float32(2.0 ** -9), (1 + m / 8) * np.exp2((e - 7).astype(np.float32))).astype(np.float32)
    return np.where(b >> 7 == 1, -mag, mag).astype(np.float32)


def to_blocked(codes2d: np.ndarray) -> np.ndarray:
    """[rows, k] bytes -> the flat array of include/fpq_b200.h:
    [k/128 slabs][rows_pad/8][8 chunks of 16 K][8 rows][16 bytes], padding rows = 0."""
    rows, k = codes2d.shape
    rp = rows_padded(rows)
    full = np.zeros((rp, k), dtype=np.uint8)
    full[:rows] = codes2d
    t = full.reshape(rp // 8, 8, k // GROUP, 8, 16)            # [rb, r, slab, chunk, byte]
    return np.ascontiguousarray(t.transpose(2, 0, 3, 1, 4)).reshape(-1)


def from_blocked(flat: np.ndarray, rows: int, k: int) -> np.ndarray:
    rp = rows_padded(rows)
    t = np.asarray(flat, dtype=np.uint8).reshape(k // GROUP, rp // 8, 8, 8, 16)       # [slab, rb, chunk, r, byte]
    return np.ascontiguousarray(t.transpose(1, 3, 0, 2, 4)).reshape(rp, k)[:rows]


def scales_layout(s: np.ndarray) -> np.ndarray:
    """[rows, k/128] -> [k/128, rows_pad] with zeros for the padding rows."""
    rows, slabs = s.shape
    out = np.zeros((slabs, rows_padded(rows)), dtype=np.float32)
    out[:, :rows] = s.T
    return out


def pack_codes(x: np.ndarray, fmt: str, scale_group=GROUP):
    """What fpq_pack_codes writes: (flat code bytes, scales [k/scale_group, rows_pad])."""
    q, s = quantize_codes(x, fmt, scale_group)
    return to_blocked(e4m3_encode(q)), scales_layout(s)


_HALF = {"e2m1": [0, .5, 1, 1.5, 2, 3, 4, 6], "e1m2": [0, .25, .5, .75, 1, 1.25, 1.5, 1.75], "e3m0": [0, .25, .5, 1, 2, 4, 8, 16]}


def codes_to_nibbles(codes: np.ndarray, fmt: str) -> np.ndarray:
    """nibble = sign << 3 | index of |q| in the ascending half grid; byte i = codes 2i (low nibble), 2i+1."""
    v = e4m3_decode(codes)
    half = np.asarray(_HALF[fmt], dtype=np.float32)
    idx = np.searchsorted(half, np.abs(v))
    assert np.all(half[idx] == np.abs(v))
    nib = (idx | ((np.asarray(codes) >> 7).astype(np.int64) << 3)).astype(np.uint8)
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)


def nibbles_to_codes(nib: np.ndarray, fmt: str) -> np.ndarray:
    half = np.asarray(_HALF[fmt], dtype=np.float32)
    n = np.empty(nib.size * 2, dtype=np.uint8)
    n[0::2] = nib & 0xF
    n[1::2] = nib >> 4
    v = half[n & 7] * np.where((n >> 3) == 1, np.float32(-1), np.float32(1))
    return e4m3_encode(np.where(v == 0, np.float32(0), v))


def gemm_codes(qa: np.ndarray, sa: np.ndarray, qw: np.ndarray, sw: np.ndarray, bias=None) -> np.ndarray:
    """fp32 [m, n] in the fixed operation order of oracle/gemm_codes.c.  qa [m, k], sa [m, groups], qw [n, k], sw [n, groups]."""
    m, k = qa.shape
    assert sa.shape[1] == sw.shape[1] and k % sa.shape[1] == 0
    n = qw.shape[0]
    qa = np.ascontiguousarray(qa, dtype=np.float32)
    qw = np.ascontiguousarray(qw, dtype=np.float32)
    sat = np.ascontiguousarray(sa.T, dtype=np.float32)          # [slabs, m]
    swt = np.ascontiguousarray(sw.T, dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    c = np.empty((m, n), dtype=np.float32)
    _lib().gemm_codes_ref(qa.ctypes.data, sat.ctypes.data, m, qw.ctypes.data, swt.ctypes.data, n, k, k // sa.shape[1],
                          None if b is None else b.ctypes.data, c.ctypes.data)
    return c


def linear_f64(qa, sa, qw, sw, bias=None) -> np.ndarray:
    """The reference's expression F.linear(q_x * s_x, W_q * s_w, b) evaluated in float64 on the exact operand values."""
    a = qa.astype(np.float64).reshape(qa.shape[0], sa.shape[1], -1) * sa.astype(np.float64)[:, :, None]
    w = qw.astype(np.float64).reshape(qw.shape[0], sw.shape[1], -1) * sw.astype(np.float64)[:, :, None]
    c = a.reshape(qa.shape) @ w.reshape(qw.shape).T
    return c if bias is None else c + np.asarray(bias, dtype=np.float64)
