"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + the plain-C rounding rules in scan_quant.c) of the reference's
floating-point fake-quantization path.  Nothing under ``fpqvar_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do,
and only as the checker / the timed CPU baseline.

Parity status: the reference has no tests and no golden vectors, so this
oracle is pinned by
  * tests/golden/*.npz -- outputs of the reference's own Python functions imported from
    /root/reference in the build container (generator: tests/golden/make_golden.py), and
  * the unmodified reference CUDA extension built into oracle/_ref (oracle/build_ref.sh),
    run against ``scan_quant`` on the GPU box (tests/test_gpu_ref_ext.py).

All file:line citations are relative to /root/reference/.  "qu.py" abbreviates
models_fp_quant_transform_rotate/quant_utils.py and "qu0.py" models_fp_quant/quant_utils.py.

Every function states, operation by operation, the dtype each reference step runs in:
torch computes an fp16 ``a / b`` as ``half(float(a) / float(b))`` and a 0-dim divisor does not
promote an fp16 tensor, so with fp16 input the scale and the normalised tensor are rounded to
fp16 before the grid rounding, which always happens in fp32 (qu.py:323, quant/quant_kernel.cu:28).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """Load (building on first use) the plain-C rounding rules."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_scan.so")
        src = os.path.join(_HERE, "scan_quant.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_scan.so"])
        lib = ctypes.CDLL(so)
        for name in ("oracle_scan_quant_f32", "oracle_argmin_quant_f32"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
        _LIB = lib
    return _LIB


# --------------------------------------------------------------------------------------
# Grids, exactly as the reference spells them (including the duplicated 0 in the FP6 tables)
# --------------------------------------------------------------------------------------
def _sym(pos):
    pos = list(pos)
    return np.array([-v for v in reversed(pos)] + [0.0] + pos, dtype=np.float32)


def _sym_double_zero(pos):
    pos = list(pos)
    return np.array([-v for v in reversed(pos)] + [0.0, 0.0] + pos, dtype=np.float32)


_E2M1_POS = [0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]
_E1M2_POS = [0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75]
_E3M0_POS = [0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 16.0]
_E2M3_POS = ([0.125 * i for i in range(1, 16)] + [2.0 + 0.25 * i for i in range(8)]
             + [4.0 + 0.5 * i for i in range(8)])
_E3M2_POS = ([0.0625 * i for i in range(1, 8)] + [0.5 + 0.125 * i for i in range(4)]
             + [1.0 + 0.25 * i for i in range(4)] + [2.0 + 0.5 * i for i in range(4)]
             + [4.0 + 1.0 * i for i in range(4)] + [8.0 + 2.0 * i for i in range(4)]
             + [16.0 + 4.0 * i for i in range(4)])

GRIDS = {
    # qu.py:233-235
    "e3m0": _sym(_E3M0_POS),
    "e2m1": _sym(_E2M1_POS),
    "e1m2": _sym(_E1M2_POS),
    # qu.py:458-486
    "e2m3": _sym_double_zero(_E2M3_POS),
    "e3m2": _sym_double_zero(_E3M2_POS),
    # qu.py:488-493, 495-500
    "int_neg": np.array([float(-i) for i in range(32, -1, -1)], dtype=np.float32),
    "e2m3_pos": np.array([0.0] + _E2M3_POS, dtype=np.float32),
    # qu.py:418-419
    "e1m2_neg": np.array([-v for v in reversed(_E1M2_POS)] + [0.0], dtype=np.float32),
    "e2m1_pos": np.array([0.0] + _E2M1_POS, dtype=np.float32),
    # qu0.py:501 (the AFPQ negative half is the mirrored e2m1 half, despite its variable name)
    "e2m1_neg": np.array([-v for v in reversed(_E2M1_POS)] + [0.0], dtype=np.float32),
}

# sign-split formats: name -> (negative-half grid, positive-half grid)
SPLIT = {
    "e1m2_neg_e2m1_pos": ("e1m2_neg", "e2m1_pos"),      # qu.py:415-452
    "int_neg_e2m3_pos": ("int_neg", "e2m3_pos"),        # qu.py:577-646
    "afpq_e2m1": ("e2m1_neg", "e2m1_pos"),              # qu0.py:498-535
}


def grid_absmax(name: str) -> np.float32:
    """``quant_grid.abs().max()`` (qu.py:320)."""
    return np.float32(np.max(np.abs(GRIDS[name])))


# --------------------------------------------------------------------------------------
# Element-wise rounding rules
# --------------------------------------------------------------------------------------
def _round_with(fn_name: str, x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    g32 = np.ascontiguousarray(grid, dtype=np.float32)
    z = np.empty_like(x32)
    getattr(_lib(), fn_name)(x32.ctypes.data, g32.ctypes.data, int(g32.size), int(x32.size), z.ctypes.data)
    return z


def scan_quant(x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """R_K: quant_cuda.quant's element rule (quant/quant_kernel.cu:25-37)."""
    return _round_with("oracle_scan_quant_f32", x, grid)


def argmin_quant(x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """R_A: quantize_to_nearest_grid (qu.py:209-230)."""
    return _round_with("oracle_argmin_quant_f32", x, grid)


def scan_quant_py(x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """Vectorised numpy restatement of the same scan (cross-check of the C loop)."""
    x32 = np.asarray(x, dtype=np.float32)
    best = np.full(x32.shape, np.float32(102400.0), dtype=np.float32)
    z = np.zeros(x32.shape, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        for g in np.asarray(grid, dtype=np.float32):
            d = np.abs(x32 - g)
            take = d <= best
            best = np.where(take, d, best)
            z = np.where(take, g, z)
    return z


_ROUND = {"kernel": scan_quant, "argmin": argmin_quant}


# --------------------------------------------------------------------------------------
# dtype-faithful arithmetic helpers
# --------------------------------------------------------------------------------------
def _div(a: np.ndarray, b: np.ndarray, dtype) -> np.ndarray:
    """torch ``a / b`` for tensors of ``dtype``: fp32 divide, then round to ``dtype``."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore", under="ignore"):
        return (a.astype(np.float32) / b.astype(np.float32)).astype(dtype)


def _absmax_lastdim(g: np.ndarray) -> np.ndarray:
    """``g.abs().max(dim=-1, keepdim=True)[0]`` -- NaN propagates, dtype preserved."""
    return np.max(np.abs(g), axis=-1, keepdims=True)


def _as_groups(x: np.ndarray, group_size):
    if group_size is None:                       # per_token / per_channel: the row is the last dim
        return x.reshape(-1, x.shape[-1])
    assert x.size % group_size == 0
    return x.reshape(-1, group_size)


# --------------------------------------------------------------------------------------
# Symmetric formats
# --------------------------------------------------------------------------------------
def fake_quant(x: np.ndarray, fmt: str, group_size=128, tie: str = "kernel",
               clamp3: bool = False, out_dtype=None) -> np.ndarray:
    """absmax scale -> divide -> grid rounding -> multiply back.

    tie="kernel": the ``*_cuda`` functions, e.g. fp_quant_e2_per_group_cuda qu.py:313-330,
      fp6_quant_e2m3_per_token_cuda qu.py:503-517 (``group_size=None``; those four FP6
      functions force ``out_dtype=float16``, qu.py:516,533,553,573).  Output dtype defaults
      to the input dtype (qu.py:329).
    tie="argmin": the torch-only functions, e.g. fp_quant_e2_per_group qu.py:298-310; the
      e1/e3 group variants and all per_token variants clamp to +-3 first (``clamp3``,
      qu.py:241,254,289,337,350).  ``quantized_x * scale`` promotes to fp32 (the grid is
      fp32), so the output is fp32 whatever the input dtype.
    """
    x = np.asarray(x)
    assert x.dtype in (np.float16, np.float32)
    dt = x.dtype
    grid = GRIDS[fmt]
    if clamp3:
        x = np.clip(x, dt.type(-3), dt.type(3))            # torch.clamp keeps NaN, as np.clip does
    g = _as_groups(x, group_size)
    a = _absmax_lastdim(g)                                  # dtype dt
    s = _div(a, np.full_like(a, grid_absmax(fmt), dtype=np.float32), dt)    # qu.py:320
    v = _div(g, s, dt)                                      # qu.py:321
    q = _ROUND[tie](v.astype(np.float32), grid).reshape(g.shape)           # qu.py:323-326 / :307
    with np.errstate(invalid="ignore", over="ignore", under="ignore"):
        out = q * s.astype(np.float32)                      # qu.py:328 (fp32 * dt -> fp32)
    if out_dtype is None:
        out_dtype = dt if tie == "kernel" else np.float32
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        return out.reshape(x.shape).astype(out_dtype)


# --------------------------------------------------------------------------------------
# Sign-split formats (fc2 inputs)
# --------------------------------------------------------------------------------------
def _global_clip(x: np.ndarray, clipping_strength: float) -> np.ndarray:
    """qu.py:421-422 -- ``clip = strength * x.abs().max(); clamp(x, -clip, clip)``.

    Identity for finite data at strength 1.0.  If the tensor holds a NaN the clip value is NaN
    and torch.clamp with NaN bounds turns EVERY element into NaN."""
    dt = x.dtype
    with np.errstate(invalid="ignore", over="ignore"):
        clip = (np.float32(clipping_strength) * np.max(np.abs(x)).astype(np.float32)).astype(dt)
    if np.isnan(clip):
        return np.full_like(x, np.nan)
    return np.minimum(np.maximum(x, -clip), clip)


def fake_quant_signsplit(x: np.ndarray, fmt: str, group_size=128, tie: str = "kernel",
                         clipping_strength=1.0) -> np.ndarray:
    """fp_quant_e1m2_neg_e2m1_pos_per_group_cuda qu.py:415-452 (tie="kernel"),
    fp_quant_e1m2_neg_e2m1_pos_per_group qu.py:381-412 (tie="argmin"),
    fp6_quant_int_neg_e2m3_pos_per_{group,token}_cuda qu.py:577-646 (no global clip:
    ``clipping_strength=None``), fp4_afpq_per_group_cuda qu0.py:498-535."""
    x = np.asarray(x)
    assert x.dtype in (np.float16, np.float32)
    dt = x.dtype
    gneg, gpos = (GRIDS[n] for n in SPLIT[fmt])
    if clipping_strength is not None:
        x = _global_clip(x, clipping_strength)
    g = _as_groups(x, group_size)
    zero = np.zeros_like(g)
    with np.errstate(invalid="ignore"):
        xn = np.where(g <= 0, g, zero)                      # qu.py:428 (NaN -> 0)
        xp = np.where(g > 0, g, zero)                       # qu.py:429
    sn = _div(_absmax_lastdim(xn), np.full((1, 1), np.max(np.abs(gneg)), np.float32), dt)   # :432
    sp = _div(_absmax_lastdim(xp), np.full((1, 1), np.max(np.abs(gpos)), np.float32), dt)   # :433
    vn = _div(xn, sn, dt)                                   # :436
    vp = _div(xp, sp, dt)                                   # :437
    qn = _ROUND[tie](vn.astype(np.float32), gneg).reshape(g.shape)         # :443
    qp = _ROUND[tie](vp.astype(np.float32), gpos).reshape(g.shape)         # :444
    with np.errstate(invalid="ignore", over="ignore", under="ignore"):
        if tie == "kernel":
            out = qn * sn.astype(np.float32) + qp * sp.astype(np.float32)  # :450
            out = out.astype(dt)                                           # :451
        else:
            with np.errstate(invalid="ignore"):
                sel = np.where(g <= 0, sn.astype(np.float32), sp.astype(np.float32))
            out = (qn + qp) * sel                                          # :409-410 (fp32)
    return out.reshape(x.shape)


def fake_quant_neg_reverse(x: np.ndarray, group_size=128) -> np.ndarray:
    """fp_neg_reverse_quant_per_group_cuda qu0.py:454-495: negatives are shifted by |min| and
    quantized on the full e2m1 grid, positives on e2m1; the shift is subtracted from EVERY
    element afterwards (qu0.py:491-493)."""
    x = np.asarray(x)
    dt = x.dtype
    grid = GRIDS["e2m1"]
    g = _as_groups(x, group_size)
    zero = np.zeros_like(g)
    m = np.abs(np.min(g, axis=-1, keepdims=True))           # :464-465
    with np.errstate(invalid="ignore"):
        xn = np.where(g <= 0, g, zero)
        xp = np.where(g > 0, g, zero)
    with np.errstate(over="ignore", invalid="ignore"):
        xr = (xn + m).astype(dt)                            # :471
    six = np.full((1, 1), 6.0, np.float32)
    sr = _div(_absmax_lastdim(xr), six, dt)
    sp = _div(_absmax_lastdim(xp), six, dt)
    qr = scan_quant(_div(xr, sr, dt).astype(np.float32), grid).reshape(g.shape)
    qp = scan_quant(_div(xp, sp, dt).astype(np.float32), grid).reshape(g.shape)
    with np.errstate(invalid="ignore", over="ignore", under="ignore"):
        hat = qr * sr.astype(np.float32) - m.astype(np.float32)            # :491
        out = hat + qp * sp.astype(np.float32)                             # :493
        return out.astype(dt).reshape(x.shape)


# --------------------------------------------------------------------------------------
# Rotation (block random Hadamard) and GALT transform
# --------------------------------------------------------------------------------------
# torch.manual_seed(42); torch.randint(0, 2, (128,)) -- rotate_utils/hadamard_utils.py:95-96.
# 1 = +1, 0 = -1, index 0 first.  Recorded in SURVEY.md section 8 (a8); tests re-derive it
# from torch's CPU generator.
SIGN_BITS_SEED42_128 = (
    "0100010001000010111010111111110011101000001111101101010110000000"
    "0110111101011101010100101111111111100111111110101101011010100110"
)


def sign_vector(bits: str = SIGN_BITS_SEED42_128) -> np.ndarray:
    return np.array([1.0 if c == "1" else -1.0 for c in bits], dtype=np.float64)


def hadamard_butterfly(x: np.ndarray) -> np.ndarray:
    """matmul_hadU for a power-of-two last dim (rotate_utils/hadamard_utils.py:63-85):
    repeated (a+b, a-b) pair stages, then division by fp32 sqrt(n) (``torch.tensor(n).sqrt()``
    is an fp32 0-dim tensor, so the fp64 data is divided by the fp32-rounded root)."""
    n = x.shape[-1]
    assert n & (n - 1) == 0
    inp = x.astype(np.float64).reshape(-1, n, 1).copy()
    while inp.shape[1] > 1:
        inp = inp.reshape(inp.shape[0], inp.shape[1] // 2, 2, inp.shape[2])
        out = np.empty_like(inp)
        out[:, :, 0, :] = inp[:, :, 0, :] + inp[:, :, 1, :]
        out[:, :, 1, :] = inp[:, :, 0, :] - inp[:, :, 1, :]
        inp = out.reshape(inp.shape[0], inp.shape[1], -1)
    return inp.reshape(x.shape) / np.float64(np.sqrt(np.float32(n)))


def random_hadamard_matrix(signs: np.ndarray) -> np.ndarray:
    """hadamard_utils.py:92-99: matmul_hadU(diag(signs)) in fp64."""
    return hadamard_butterfly(np.diag(signs.astype(np.float64)))


def block_random_hadamard_matrix(total_size: int, block_size: int = 128,
                                 signs: np.ndarray | None = None) -> np.ndarray:
    """rotation_utils.py:69-104: every diagonal block is the SAME seed-42 matrix (the inner
    call reseeds with the same seed for each block, hadamard_utils.py:95)."""
    assert total_size % block_size == 0
    if signs is None:
        assert block_size == 128
        signs = sign_vector()
    blk = random_hadamard_matrix(signs)
    q = np.zeros((total_size, total_size), dtype=np.float64)
    for i in range(total_size // block_size):
        q[i * block_size:(i + 1) * block_size, i * block_size:(i + 1) * block_size] = blk
    return q


def rotate_weight(w: np.ndarray, q: np.ndarray) -> np.ndarray:
    """rotate_mat_qkv / rotate_fc1 (rotation_utils.py:129-154): fp64 matmul, cast back."""
    return (w.astype(np.float64) @ q).astype(w.dtype)


def transform_weight(w: np.ndarray, s: np.ndarray) -> np.ndarray:
    """transform_mat_qkv / transform_fc1 (learnable_transformation/transform_model_utils.py:8-21)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return (w / s.astype(w.dtype)).astype(w.dtype)


def transform_rotate_activation_f64(x: np.ndarray, s: np.ndarray, q: np.ndarray) -> np.ndarray:
    """Exact-arithmetic statement of basic_var.py:263,266: ``matmul(x.mul(s), Q)``.  The
    reference runs this under fp16 autocast (error ~3e-3); the product path is checked against
    this fp64 value with a stated fp32 tolerance ("Rotation is
    tolerance-checked")."""
    xs = (x.astype(np.float32) * s.astype(np.float32)).astype(np.float64)   # fp32 mul, as the reference
    return xs @ q


def adaln_modulate(ln_out: np.ndarray, scale: np.ndarray, shift: np.ndarray) -> np.ndarray:
    """basic_var.py:263,266 ``self.ln_wo_grad(x).mul(scale1.add(1)).add_(shift1)``: three fp32 elementwise ops,
    each rounded to fp32 ([B, L, C] with [B, 1, C] operands)."""
    a = (scale.astype(np.float32) + np.float32(1)).astype(np.float32)
    return ((ln_out.astype(np.float32) * a).astype(np.float32) + shift.astype(np.float32)).astype(np.float32)


# --------------------------------------------------------------------------------------
# Format scoring (search scripts)
# --------------------------------------------------------------------------------------
def tensor_mse(x: np.ndarray, xq: np.ndarray) -> float:
    """compute_quant_error (search/search_fp4_format.py:472-476), accumulated in fp64."""
    d = x.astype(np.float64) - xq.astype(np.float64)
    return float(np.mean(d * d))
