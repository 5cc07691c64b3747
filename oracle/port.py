"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

ctypes front end of oracle/fakequant_port.c: the multi-threaded C port of the reference's fake-quant
functions that bench.py times as the CPU baseline (``cpu_baseline.kind == "port"``) and that
``bench.py --impl reference`` runs.  tests/test_oracle_port.py checks it bit for bit against the
numpy oracle.  Nothing under fpqvar_b200/ may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_c = ctypes


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_port.so")
        src = os.path.join(_HERE, "fakequant_port.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
            subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_port.so"])
        lib = ctypes.CDLL(so)
        sym = [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_void_p, _c.c_int, _c.c_float, _c.c_int, _c.c_int]
        for n in ("f32_f32", "f32_f16", "f16_f16", "f16_f32"):
            fn = getattr(lib, f"port_fake_quant_{n}")
            fn.restype, fn.argtypes = None, sym
        spl = [_c.c_void_p, _c.c_void_p, _c.c_size_t, _c.c_size_t, _c.c_void_p, _c.c_int, _c.c_float, _c.c_void_p, _c.c_int,
               _c.c_float, _c.c_int]
        for n in ("f32_f32", "f16_f16", "f16_f32"):
            fn = getattr(lib, f"port_signsplit_{n}")
            fn.restype, fn.argtypes = None, spl
        lib.port_transform_rotate_quant.restype = None
        lib.port_transform_rotate_quant.argtypes = [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t,
                                                    _c.c_size_t, _c.c_void_p, _c.c_int, _c.c_float]
        lib.port_num_threads.restype = _c.c_int
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().port_num_threads())


_TAG = {np.dtype(np.float32): "f32", np.dtype(np.float16): "f16"}
_TIE = {"kernel": 0, "argmin": 1}


def _rows(x, group_size):
    rl = x.shape[-1] if group_size is None else group_size
    assert x.size % rl == 0
    return x.size // rl, rl


def fake_quant(x: np.ndarray, fmt: str, group_size=128, tie="kernel", clamp3=False, out_dtype=None, out=None) -> np.ndarray:
    x = np.ascontiguousarray(x)
    if out_dtype is None:
        out_dtype = x.dtype if tie == "kernel" else np.float32
    if out is None:
        out = np.empty(x.shape, dtype=out_dtype)
    grid = O.GRIDS[fmt]
    n_rows, rl = _rows(x, group_size)
    fn = getattr(_lib(), f"port_fake_quant_{_TAG[x.dtype]}_{_TAG[np.dtype(out_dtype)]}")
    fn(x.ctypes.data, out.ctypes.data, n_rows, rl, grid.ctypes.data, grid.size, float(O.grid_absmax(fmt)), _TIE[tie], int(clamp3))
    return out


def fake_quant_signsplit(x: np.ndarray, fmt: str, group_size=128, tie="kernel", out=None) -> np.ndarray:
    x = np.ascontiguousarray(x)
    out_dtype = x.dtype if tie == "kernel" else np.float32
    if out is None:
        out = np.empty(x.shape, dtype=out_dtype)
    gneg, gpos = (O.GRIDS[n] for n in O.SPLIT[fmt])
    n_rows, rl = _rows(x, group_size)
    fn = getattr(_lib(), f"port_signsplit_{_TAG[x.dtype]}_{_TAG[np.dtype(out_dtype)]}")
    fn(x.ctypes.data, out.ctypes.data, n_rows, rl, gneg.ctypes.data, gneg.size, float(np.max(np.abs(gneg))),
       gpos.ctypes.data, gpos.size, float(np.max(np.abs(gpos))), _TIE[tie])
    return out


_Q128 = None


def q128_f32() -> np.ndarray:
    global _Q128
    if _Q128 is None:
        _Q128 = np.ascontiguousarray(O.random_hadamard_matrix(O.sign_vector()).astype(np.float32))
    return _Q128


def transform_rotate_quant(x: np.ndarray, smooth, fmt="e2m1", out=None, return_rotated=False):
    x = np.ascontiguousarray(x, dtype=np.float32)
    c = x.shape[-1]
    assert c % 128 == 0
    if out is None:
        out = np.empty(x.shape, dtype=np.float16)
    rot = np.empty(x.shape, dtype=np.float16) if return_rotated else None
    s = None if smooth is None else np.ascontiguousarray(smooth, dtype=np.float32)
    q = q128_f32()
    grid = None if fmt is None else O.GRIDS[fmt]
    _lib().port_transform_rotate_quant(x.ctypes.data, None if s is None else s.ctypes.data, q.ctypes.data, out.ctypes.data,
                                       None if rot is None else rot.ctypes.data, x.size // c, c,
                                       None if grid is None else grid.ctypes.data, 0 if grid is None else grid.size,
                                       0.0 if grid is None else float(O.grid_absmax(fmt)))
    return (out, rot) if return_rotated else out
