"""CPU: pin the oracle against fixtures produced by the reference's own Python functions
(tests/golden/make_golden.py).  Bit-exact, including signed zeros and NaN positions."""
import numpy as np
import pytest

from oracle import oracle as O
from conftest import bits_equal, mismatch_report

DT = {"f32": np.float32, "f16": np.float16}


def _check(golden, tag, got):
    want = golden[f"out/{tag}"]
    ref_dtype = str(golden[f"dtype/{tag}"])
    assert ref_dtype == {np.dtype(np.float16): "torch.float16", np.dtype(np.float32): "torch.float32"}[got.dtype], \
        f"{tag}: reference returns {ref_dtype}, oracle returns {got.dtype}"
    got32 = got.astype(np.float32)
    assert bits_equal(got32, want), f"{tag}\n" + mismatch_report(got32, want)


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["adv", "rnd"])
@pytest.mark.parametrize("e,fmt", [(1, "e1m2"), (2, "e2m1"), (3, "e3m0")])
def test_fp4_group(golden, dn, iname, e, fmt):
    x = golden[f"in/{iname}"].astype(DT[dn])
    _check(golden, f"fp_quant_e{e}_per_group_cuda/{iname}/{dn}", O.fake_quant(x, fmt, 128, "kernel"))
    _check(golden, f"fp_quant_e{e}_per_group/{iname}/{dn}",
           O.fake_quant(x, fmt, 128, "argmin", clamp3=(e != 2)))


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["adv", "rnd"])
@pytest.mark.parametrize("fmt", ["e2m3", "e3m2"])
def test_fp6_group(golden, dn, iname, fmt):
    x = golden[f"in/{iname}"].astype(DT[dn])
    _check(golden, f"fp6_quant_{fmt}_per_group_cuda/{iname}/{dn}",
           O.fake_quant(x, fmt, 128, "kernel", out_dtype=np.float16))


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["rows_tok", "kv", "adv"])
def test_per_token(golden, dn, iname):
    x = golden[f"in/{iname}"].astype(DT[dn])
    for e, fmt in ((1, "e1m2"), (2, "e2m1"), (3, "e3m0")):
        _check(golden, f"fp_quant_e{e}_per_token/{iname}/{dn}", O.fake_quant(x, fmt, None, "argmin", clamp3=True))
    for fmt in ("e2m3", "e3m2"):
        _check(golden, f"fp6_quant_{fmt}_per_token_cuda/{iname}/{dn}",
               O.fake_quant(x, fmt, None, "kernel", out_dtype=np.float16))
    _check(golden, f"fp6_quant_int_neg_e2m3_pos_per_token_cuda/{iname}/{dn}",
           O.fake_quant_signsplit(x, "int_neg_e2m3_pos", None, "kernel", clipping_strength=None))


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["finite", "rnd", "adv"])
def test_signsplit_fp4(golden, dn, iname):
    x = golden[f"in/{iname}"].astype(DT[dn])
    _check(golden, f"fp_quant_e1m2_neg_e2m1_pos_per_group_cuda/{iname}/{dn}",
           O.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "kernel"))
    _check(golden, f"fp_quant_e1m2_neg_e2m1_pos_per_group/{iname}/{dn}",
           O.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "argmin"))
    _check(golden, f"fp4_afpq_per_group_cuda/{iname}/{dn}", O.fake_quant_signsplit(x, "afpq_e2m1", 128, "kernel"))


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["adv", "rnd"])
def test_signsplit_fp6_and_neg_reverse(golden, dn, iname):
    x = golden[f"in/{iname}"].astype(DT[dn])
    _check(golden, f"fp6_quant_int_neg_e2m3_pos_per_group_cuda/{iname}/{dn}",
           O.fake_quant_signsplit(x, "int_neg_e2m3_pos", 128, "kernel", clipping_strength=None))
    _check(golden, f"fp_neg_reverse_quant_per_group_cuda/{iname}/{dn}", O.fake_quant_neg_reverse(x, 128))


@pytest.mark.parametrize("gname", ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2", "int_neg", "e2m3_pos", "e1m2_neg", "e2m1_pos"])
def test_grids_and_argmin_rule(golden, gname):
    assert bits_equal(O.GRIDS[gname], golden[f"grid/{gname}"])
    probe = golden["in/probe"]
    got = O.argmin_quant(probe, O.GRIDS[gname])
    want = golden[f"out/quantize_to_nearest_grid/{gname}"]
    assert bits_equal(got, want), mismatch_report(got, want)


def test_scan_c_equals_numpy_restatement(golden):
    probe = golden["in/probe"]
    for gname, grid in O.GRIDS.items():
        assert bits_equal(O.scan_quant(probe, grid), O.scan_quant_py(probe, grid)), gname


def test_rotation_matrix(golden):
    assert np.array_equal(O.sign_vector(), golden["rot/signs128"])
    q = O.block_random_hadamard_matrix(256, 128)
    assert bits_equal(q, golden["rot/q256"]), "block Hadamard matrix differs from the reference's"
    blk = q[:128, :128]
    assert np.abs(blk @ blk.T - np.eye(128)).max() < 1e-6
    # entries are +-1/fl32(sqrt(128)) and every block is identical
    assert np.allclose(np.abs(blk), 1.0 / float(np.sqrt(np.float32(128))), rtol=0, atol=0)
    assert np.array_equal(q[128:, 128:], blk)


def _requantize_row_family(fmt, ai):
    """Every fp16 x with |x| <= a (a = the ai-th positive fp16) quantized with the scale of absmax a, then quantized again
    with the scale re-derived from the quantized absmax element.  Returns (first pass, second pass)."""
    allpos = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16)
    grid, vmax = O.GRIDS[fmt], np.float32(O.grid_absmax(fmt))

    def q(x, s):
        with np.errstate(all="ignore"):
            v = (x.astype(np.float32) / np.float32(s)).astype(np.float16)
            return (O.scan_quant(v.astype(np.float32), grid) * np.float32(s)).astype(np.float16)

    a = allpos[ai]
    x = allpos[:ai + 1]
    with np.errstate(all="ignore"):
        y = q(x, (np.float32(a) / vmax).astype(np.float16))
        a2 = np.max(np.abs(y))
        y2 = q(y, (np.float32(a2) / vmax).astype(np.float16))
    return y, y2


@pytest.mark.parametrize("fmt", ["e2m3", "e2m1"])
def test_requantization_is_idempotent(fmt):
    """Why the KV cache may be quantized incrementally (fpqvar_b200/kv_cache.py): the reference re-quantizes the whole
    cache at every scale (basic_var.py:192-200), and a second pass over quantized rows is the identity.  The exhaustive
    run over every (absmax, x) pair is tests/idempotence_exhaustive.py; this is every 61st absmax plus the known exceptions."""
    for ai in list(range(128, 0x7BFF, 61)) + [0x7BFE]:
        y, y2 = _requantize_row_family(fmt, ai)
        assert bits_equal(y, y2), f"{fmt}: absmax index {ai}"
    # the exceptions: deep-subnormal scales and the overflowing absmax 65504
    for ai in {"e2m3": (57, 64, 65, 72, 79, 94, 0x7BFF), "e2m1": (9, 0x7BFF)}[fmt]:
        y, y2 = _requantize_row_family(fmt, ai)
        assert not bits_equal(y, y2), f"{fmt}: absmax index {ai} was expected to drift"
    assert float(np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16)[94]) < 2.0 ** -17      # the guard of kv_cache.py covers them
