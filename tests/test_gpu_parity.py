"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the
reference-generated golden vectors.  Bit-exact for every quantized tensor."""
import numpy as np
import pytest
import torch

from conftest import bits_equal, mismatch_report
from oracle import oracle as O

pytestmark = pytest.mark.gpu

NP = {"f32": np.float32, "f16": np.float16}
TD = {"f32": torch.float32, "f16": torch.float16}
SYM = {"e1m2": 1, "e2m1": 2, "e3m0": 3}


@pytest.fixture(scope="module")
def ops():
    from fpqvar_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def assert_bits(got: np.ndarray, want: np.ndarray, tag: str):
    assert got.dtype == want.dtype, f"{tag}: dtype {got.dtype} vs {want.dtype}"
    assert bits_equal(got, want), f"{tag}\n" + mismatch_report(got, want)


# ----------------------------------------------------------------------------------------
# 1. closed-form rounding == literal scan for all 2^32 inputs
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("tie", ["kernel", "argmin"])
@pytest.mark.parametrize("code", [0, 1, 2, 3, 4, 16, 17, 18, 19, 20])
def test_rounding_exhaustive(ops, code, tie):
    bad, first = ops.selftest_rounding(code, tie)
    assert bad == 0, f"format code {code} tie {tie}: {bad} mismatches, first at bits 0x{first:08x}"


@pytest.mark.parametrize("code", [0, 1, 2, 3, 4, 16, 17, 18, 32 + 0, 32 + 1, 32 + 3, 32 + 4])
def test_f16_flow_exhaustive(ops, code):
    """Packed fp16 fast path (division-free, magic-number rounding) == literal reference sequence
    for every (x, scale) pair of fp16 values that can occur in a regular group.  Codes 32 + format: the element function on
    the FP4 / FP6 conversion hardware (e2m1, e1m2, e2m3, e3m2) that the group, rotate and scoring kernels use; codes 0..4: the
    magic-number element function (e3m0 everywhere, and the per-token row kernels); 16 + split format: the sign-split element
    function (conversion hardware on the positive side, packed fp16 magic constant on a uniform negative side)."""
    bad, first = ops.selftest_f16_flow(code)
    assert bad == 0, (f"format code {code}: {bad} mismatches, first at scale bits 0x{0x0400 + (first >> 16):04x}, "
                      f"x bits 0x{first & 0xffff:04x}")


# ----------------------------------------------------------------------------------------
# 2. golden vectors produced by the reference's own Python functions
# ----------------------------------------------------------------------------------------
def _golden_check(golden, tag, got: torch.Tensor):
    want = golden[f"out/{tag}"]
    ref_dtype = str(golden[f"dtype/{tag}"])
    assert str(got.dtype) == ref_dtype, f"{tag}: reference returns {ref_dtype}, kernel path returns {got.dtype}"
    g = host(got.to(torch.float32))
    assert_bits(g, want, tag)


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["adv", "rnd"])
def test_golden_group(ops, golden, dn, iname):
    with np.errstate(over="ignore"):
        x = dev(golden[f"in/{iname}"].astype(NP[dn]))
    for fmt, e in SYM.items():
        _golden_check(golden, f"fp_quant_e{e}_per_group_cuda/{iname}/{dn}", ops.fake_quant(x, fmt, 128, "kernel"))
        _golden_check(golden, f"fp_quant_e{e}_per_group/{iname}/{dn}", ops.fake_quant(x, fmt, 128, "argmin", clamp3=(e != 2)))
    for fmt in ("e2m3", "e3m2"):
        _golden_check(golden, f"fp6_quant_{fmt}_per_group_cuda/{iname}/{dn}",
                      ops.fake_quant(x, fmt, 128, "kernel", out_dtype=torch.float16))
    _golden_check(golden, f"fp6_quant_int_neg_e2m3_pos_per_group_cuda/{iname}/{dn}",
                  ops.fake_quant_signsplit(x, "int_neg_e2m3_pos", 128, "kernel"))


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["rows_tok", "kv", "adv"])
def test_golden_per_token(ops, golden, dn, iname):
    with np.errstate(over="ignore"):
        x = dev(golden[f"in/{iname}"].astype(NP[dn]))
    for fmt, e in SYM.items():
        _golden_check(golden, f"fp_quant_e{e}_per_token/{iname}/{dn}", ops.fake_quant(x, fmt, None, "argmin", clamp3=True))
    for fmt in ("e2m3", "e3m2"):
        _golden_check(golden, f"fp6_quant_{fmt}_per_token_cuda/{iname}/{dn}",
                      ops.fake_quant(x, fmt, None, "kernel", out_dtype=torch.float16))
    _golden_check(golden, f"fp6_quant_int_neg_e2m3_pos_per_token_cuda/{iname}/{dn}",
                  ops.fake_quant_signsplit(x, "int_neg_e2m3_pos", None, "kernel"))


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["finite", "rnd", "adv"])
def test_golden_signsplit(ops, golden, dn, iname):
    with np.errstate(over="ignore"):
        x = dev(golden[f"in/{iname}"].astype(NP[dn]))
    _golden_check(golden, f"fp_quant_e1m2_neg_e2m1_pos_per_group_cuda/{iname}/{dn}",
                  ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "kernel", global_clip=True))
    _golden_check(golden, f"fp_quant_e1m2_neg_e2m1_pos_per_group/{iname}/{dn}",
                  ops.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "argmin", global_clip=True))
    _golden_check(golden, f"fp4_afpq_per_group_cuda/{iname}/{dn}",
                  ops.fake_quant_signsplit(x, "afpq_e2m1", 128, "kernel", global_clip=True))


@pytest.mark.parametrize("gname", ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2", "int_neg", "e2m3_pos", "e1m2_neg", "e2m1_pos"])
def test_golden_quantize_to_nearest_grid(ops, golden, gname):
    probe = golden["in/probe"]
    got = host(ops.quant_grid(dev(probe), dev(golden[f"grid/{gname}"]), "argmin"))
    assert_bits(got, golden[f"out/quantize_to_nearest_grid/{gname}"], gname)


# ----------------------------------------------------------------------------------------
# 3. oracle on seeded data, larger sizes, every format / dtype / tie rule
# ----------------------------------------------------------------------------------------
def _mixed_input(seed: int, n_groups: int, dn: str) -> np.ndarray:
    rng = np.random.default_rng(seed)
    scales = np.exp(rng.uniform(np.log(1e-3), np.log(3e2), size=(n_groups, 1))).astype(np.float32)
    x = rng.standard_normal((n_groups, 128)).astype(np.float32) * scales
    x[::97] = 0.0
    x[5::211, ::7] = 0.0
    with np.errstate(over="ignore"):
        return x.astype(NP[dn])


@pytest.mark.parametrize("tie", ["kernel", "argmin"])
@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("fmt", ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2"])
def test_oracle_group(ops, fmt, dn, tie):
    x = _mixed_input(hash((fmt, dn, tie)) % 2 ** 31, 4096 + 3, dn)
    want = O.fake_quant(x, fmt, 128, tie)
    got = host(ops.fake_quant(dev(x), fmt, 128, tie))
    assert_bits(got, want, f"{fmt}/{dn}/{tie}")


@pytest.mark.parametrize("tie", ["kernel", "argmin"])
@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("split", ["e1m2_neg_e2m1_pos", "int_neg_e2m3_pos", "afpq_e2m1"])
def test_oracle_signsplit(ops, split, dn, tie):
    x = _mixed_input(hash((split, dn, tie)) % 2 ** 31, 2048 + 5, dn)
    x[7] = np.abs(x[7])                 # a group without negatives
    x[9] = -np.abs(x[9])                # a group without positives
    gelu = np.where(x[100:600] > 0, x[100:600], x[100:600] * 0.03).astype(x.dtype)
    x[100:600] = gelu
    want = O.fake_quant_signsplit(x, split, 128, tie, clipping_strength=None)
    got = host(ops.fake_quant_signsplit(dev(x), split, 128, tie))
    assert_bits(got, want, f"{split}/{dn}/{tie}")


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("row_len", [64, 192, 1920, 2304, 7680, 1000, 1])
def test_oracle_rows(ops, dn, row_len):
    rng = np.random.default_rng(row_len)
    x = (rng.standard_normal((37, row_len)) * 2.0).astype(NP[dn])
    for fmt in ("e2m3", "e3m2", "e2m1"):
        want = O.fake_quant(x, fmt, None, "kernel", out_dtype=np.float16)
        got = host(ops.fake_quant(dev(x), fmt, None, "kernel", out_dtype=torch.float16))
        assert_bits(got, want, f"rows {fmt}/{dn}/{row_len}")
    want = O.fake_quant_signsplit(x, "int_neg_e2m3_pos", None, "kernel", clipping_strength=None)
    got = host(ops.fake_quant_signsplit(dev(x), "int_neg_e2m3_pos", None, "kernel"))
    assert_bits(got, want, f"rows split/{dn}/{row_len}")


def test_config1_shape_fp32_e2m1(ops):
    """BASELINE config 1: randn(4096, 4096) seed 0, g=128, fp_e2 -- both tie rules, full size."""
    torch.manual_seed(0)
    x = torch.randn(4096, 4096)
    xn = x.numpy()
    for tie in ("kernel", "argmin"):
        want = O.fake_quant(xn, "e2m1", 128, tie)
        got = host(ops.fake_quant(x.cuda(), "e2m1", 128, tie))
        assert_bits(got, want, f"config1/{tie}")
    # fp16 flow of the same tensor (the activation dtype of the rotated models)
    xh = xn.astype(np.float16)
    assert_bits(host(ops.fake_quant(dev(xh), "e2m1", 128, "kernel")), O.fake_quant(xh, "e2m1", 128, "kernel"), "config1/f16")


def test_fp16_all_inputs_for_scale_sweep(ops):
    """Every finite fp16 x against a sweep of group maxima: checks the half(x*RN(1/s)) ==
    half(x/s) claim of fpq_common.cuh on the real hardware, all formats."""
    allh = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16)            # every non-negative finite fp16
    rng = np.random.default_rng(7)
    maxima = np.unique(np.concatenate([allh[rng.integers(0x0400, 0x7C00, 96)], allh[[0x0400, 0x0401, 0x3C00, 0x4600, 0x7BFF]]]))
    rows = []
    for m in maxima:
        cand = allh[allh <= m]
        pick = cand[rng.integers(0, cand.size, 126)]
        sign = np.where(rng.random(126) < 0.5, -1, 1).astype(np.float16)
        rows.append(np.concatenate([pick * sign, [m, -m]]).astype(np.float16))
    x = np.stack(rows)
    for fmt in ("e2m1", "e1m2", "e3m0", "e2m3", "e3m2"):
        assert_bits(host(ops.fake_quant(dev(x), fmt, 128, "kernel")), O.fake_quant(x, fmt, 128, "kernel"), f"f16 sweep {fmt}")


def test_empty_and_errors(ops):
    from fpqvar_b200._lib import FpqError
    e = torch.empty(0, 128, device="cuda")
    assert ops.fake_quant(e, "e2m1").shape == (0, 128)
    assert ops.quant_grid(torch.empty(0, device="cuda"), torch.tensor([0.0, 1.0], device="cuda")).numel() == 0
    with pytest.raises(FpqError):
        ops.fake_quant(torch.zeros(4, 100, device="cuda"), "e2m1", 128)      # numel not a multiple of the group
    with pytest.raises(FpqError):
        ops.fake_quant(torch.zeros(4, 128), "e2m1")                           # CPU tensor: no fallback
    with pytest.raises(FpqError):
        ops.fake_quant(torch.zeros(4, 128, device="cuda", dtype=torch.bfloat16), "e2m1")


def test_unknown_grid_and_unaligned(ops):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(100003) * 3).astype(np.float32)
    grid = np.array([0.3, -2.0, 1.0, 1.0, 7.5, -0.1], dtype=np.float32)       # unsorted, duplicated
    for tie, fn in (("kernel", O.scan_quant), ("argmin", O.argmin_quant)):
        got = host(ops.quant_grid(dev(x), dev(grid), tie))
        assert_bits(got, fn(x, grid), f"unknown grid {tie}")
    base = dev(np.concatenate([[0.0], x]).astype(np.float32))
    got = host(ops.quant_grid(base[1:], dev(O.GRIDS["e2m1"]), "kernel"))      # 4-byte-aligned view
    assert_bits(got, O.scan_quant(x, O.GRIDS["e2m1"]), "unaligned")
