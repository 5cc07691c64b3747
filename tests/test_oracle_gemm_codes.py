"""CPU checks of the oracle for the packed low-bit operands and the low-bit GEMM (oracle/lowbit.py, oracle/gemm_codes.c),
anchored on the oracle's statement of the reference quantizer (oracle.fake_quant = fp_quant_*_per_group_cuda,
models_fp_quant_transform_rotate/quant_utils.py:265-378, 537-574), which the golden vectors pin."""
import ctypes
import os

import numpy as np
import pytest

from oracle import lowbit as LB
from oracle import oracle as O

FMTS = ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2"]


def bits(a):
    return a.view({2: np.uint16, 4: np.uint32}[a.dtype.itemsize])


def adversarial(rng, rows, k, dtype):
    x = rng.standard_normal((rows, k)).astype(dtype)
    x[0, :128] = 0                                   # all-zero group
    x[1, 5] = np.inf
    x[2, 130] = np.nan
    x[3, :128] = (x[3, :128].astype(np.float32) * 2.0 ** -20).astype(dtype)
    x[4, :8] = np.asarray([0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 6.0], dtype=dtype)      # e2m1 midpoints when absmax = 6
    return x


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_codes_times_scale_is_the_fake_quantized_tensor(fmt, dtype):
    rng = np.random.default_rng(1)
    x = adversarial(rng, 37, 384, dtype)
    q, s = LB.quantize_codes(x, fmt)
    out_dt = np.float16 if fmt in ("e2m3", "e3m2") else dtype       # qu.py:553,573 force fp16 for the FP6 functions
    want = O.fake_quant(x, fmt, 128, "kernel", out_dtype=out_dt)
    got = LB.dequantize(q, s, out_dt)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("fmt", FMTS)
def test_every_grid_value_is_an_e4m3_number(fmt):
    g = O.GRIDS[fmt]
    b = LB.e4m3_encode(g)
    assert np.array_equal(LB.e4m3_decode(b), g.astype(np.float32))


def test_e4m3_decode_matches_the_format_definition():
    # spot values: 0x38 = 1.0, 0x30 = 0.5, 0x4c = 6.0, 0x7e = 448 (max), 0x01 = 2^-9, 0xb8 = -1.0
    b = np.asarray([0x38, 0x30, 0x4C, 0x7E, 0x01, 0xB8, 0x00], dtype=np.uint8)
    assert np.array_equal(LB.e4m3_decode(b), np.asarray([1, .5, 6, 448, 2.0 ** -9, -1, 0], dtype=np.float32))


def test_blocked_layout_round_trip_and_tile_contiguity():
    rng = np.random.default_rng(2)
    rows, k = 200, 384
    c = rng.integers(0, 256, (rows, k), dtype=np.uint8)
    flat = LB.to_blocked(c)
    assert flat.size == LB.rows_padded(rows) * k
    assert np.array_equal(LB.from_blocked(flat, rows, k), c)
    # the (128 rows x 128 K) tile (tile row 1, slab 2) is 16 KB contiguous, core matrices of 8 rows x 16 bytes
    rp = LB.rows_padded(rows)
    base = (2 * (rp // 8) + 16) * 1024
    tile = flat[base:base + 16384].reshape(16, 8, 8, 16)               # [row block, chunk, row, byte]
    for rb, ch, r in [(0, 0, 0), (3, 5, 7), (8, 7, 1)]:
        row = 128 + rb * 8 + r
        want = c[row, 256 + ch * 16:256 + ch * 16 + 16] if row < rows else np.zeros(16, np.uint8)
        assert np.array_equal(tile[rb, ch, r], want)


@pytest.mark.parametrize("fmt", ["e2m1", "e1m2", "e3m0"])
def test_nibble_storage_is_lossless(fmt):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((16, 256)).astype(np.float16)
    q, _ = LB.quantize_codes(x, fmt)
    codes = LB.e4m3_encode(q).reshape(-1)
    nib = LB.codes_to_nibbles(codes, fmt)
    assert nib.size * 2 == codes.size
    assert np.array_equal(LB.nibbles_to_codes(nib, fmt), codes)


@pytest.mark.parametrize("fmt_a,fmt_w", [("e2m1", "e2m1"), ("e1m2", "e3m0"), ("e2m3", "e2m3")])
def test_fixed_order_gemm_against_float64(fmt_a, fmt_w):
    rng = np.random.default_rng(4)
    m, n, k = 33, 24, 640
    x = rng.standard_normal((m, k)).astype(np.float16)
    w = (rng.standard_normal((n, k)) * 0.05).astype(np.float32)
    qa, sa = LB.quantize_codes(x, fmt_a)
    qw, sw = LB.quantize_codes(w, fmt_w)
    bias = rng.standard_normal(n).astype(np.float32)
    c = LB.gemm_codes(qa, sa, qw, sw, bias)
    ref = LB.linear_f64(qa, sa, qw, sw, bias)
    # k/128 fma steps, each with one extra rounding of P * sa: error <= slabs * 2^-23 * sum_t |term_t| (+ the bias add)
    mag = (np.abs(qa.astype(np.float64)).reshape(m, -1, 128) * sa[:, :, None]).reshape(m, k) @ \
          (np.abs(qw.astype(np.float64)).reshape(n, -1, 128) * sw[:, :, None]).reshape(n, k).T + np.abs(bias)
    assert np.all(np.abs(c - ref) <= (k // 128 + 2) * 2.0 ** -23 * mag)
    # and it is the reference's expression: F.linear on the fake-quantized fp16/fp32 tensors, up to their own roundings
    xq = O.fake_quant(x, fmt_a, 128, "kernel").astype(np.float64)
    wq = O.fake_quant(w, fmt_w, 128, "kernel").astype(np.float64)
    lin = xq @ wq.T + bias
    assert np.all(np.abs(c - lin) <= 2.0 ** -10 * mag)               # fp16 rounding of q * s on the activation side


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_row_scaled_codes_are_the_per_token_functions(fmt, dtype):
    """scale_group=None: one scale per row, fp6_quant_*_per_token_cuda / the per_channel weight functions (qu.py:503-534)."""
    rng = np.random.default_rng(6)
    x = adversarial(rng, 9, 384, dtype)
    q, s = LB.quantize_codes(x, fmt, None)
    assert s.shape == (9, 1)
    out_dt = np.float16 if fmt in ("e2m3", "e3m2") else dtype
    want = O.fake_quant(x, fmt, None, "kernel", out_dtype=out_dt)
    assert np.array_equal(bits(LB.dequantize(q, s, out_dt)), bits(want))
    codes, scales = LB.pack_codes(x, fmt, None)
    assert codes.size == 128 * 384 and scales.shape == (1, 128)
    assert np.array_equal(bits(scales[0, :9]), bits(s[:, 0])) and not scales[0, 9:].any()


def test_row_scaled_gemm_oracle_against_float64():
    rng = np.random.default_rng(8)
    m, n, k = 17, 24, 640
    x = rng.standard_normal((m, k)).astype(np.float16)
    w = (rng.standard_normal((n, k)) * 0.05).astype(np.float32)
    (qa, sa), (qw, sw) = LB.quantize_codes(x, "e2m1", None), LB.quantize_codes(w, "e2m1", None)
    c = LB.gemm_codes(qa, sa, qw, sw)
    ref = LB.linear_f64(qa, sa, qw, sw)
    # one exact sum, two roundings (P * sa, then * sw): 2^-22 relative
    assert np.all(np.abs(c - ref) <= 2.0 ** -22 * np.abs(ref) + 1e-30)


def test_gemm_oracle_slab_order_matters_only_in_rounding():
    rng = np.random.default_rng(5)
    qa = rng.choice(O.GRIDS["e2m1"], (8, 256)).astype(np.float32)
    qw = rng.choice(O.GRIDS["e2m1"], (8, 256)).astype(np.float32)
    sa = np.ones((8, 2), np.float32)
    sw = np.ones((8, 2), np.float32)
    c = LB.gemm_codes(qa, sa, qw, sw)
    assert np.array_equal(c, (qa.astype(np.float64) @ qw.astype(np.float64).T).astype(np.float32))     # unit scales: exact


def test_c_abi_exports_the_low_bit_entry_points():
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fpqvar_b200", "libfpq_b200.so")
    if not os.path.exists(so):
        pytest.skip("libfpq_b200.so is not built")
    lib = ctypes.CDLL(so)
    for name in ("fpq_codes_rows_padded", "fpq_pack_codes", "fpq_unpack_codes", "fpq_codes_to_nibbles", "fpq_nibbles_to_codes",
                 "fpq_gemm_codes"):
        assert hasattr(lib, name)
    lib.fpq_codes_rows_padded.restype = ctypes.c_size_t
    lib.fpq_codes_rows_padded.argtypes = [ctypes.c_size_t]
    assert lib.fpq_codes_rows_padded(68000) == 68096 and lib.fpq_codes_rows_padded(128) == 128


def test_lowbit_host_layer_has_no_cpu_path():
    """The product raises on CPU tensors and on malformed PackedCodes before anything reaches the library (no fallback)."""
    import torch
    from fpqvar_b200 import _lib as L, lowbit
    with pytest.raises(L.FpqError):
        lowbit.pack_codes(torch.randn(8, 128), "e2m1")
    p = lowbit.PackedCodes(torch.zeros(128 * 128, dtype=torch.uint8), torch.zeros(1, 128), 8, 128, "e2m1")
    with pytest.raises(L.FpqError):
        lowbit.linear_codes(p, p)
    with pytest.raises(L.FpqError):
        p.dequantize()
    assert lowbit.rows_padded(1) == 128 and lowbit.rows_padded(129) == 256
