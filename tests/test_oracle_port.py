"""CPU: the multi-threaded C port that bench.py times as the CPU baseline (oracle/fakequant_port.c) must
agree bit for bit with the numpy oracle, which is pinned to the reference-generated fixtures."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import port as P
from conftest import bits_equal, mismatch_report

DT = {"f32": np.float32, "f16": np.float16}


@pytest.mark.parametrize("tie", ["kernel", "argmin"])
@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("fmt", ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2"])
def test_port_sym(golden, fmt, dn, tie):
    with np.errstate(over="ignore"):
        x = np.concatenate([golden["in/adv"], golden["in/rnd"]]).astype(DT[dn])
    for clamp3 in (False, True):
        want = O.fake_quant(x, fmt, 128, tie, clamp3=clamp3)
        got = P.fake_quant(x, fmt, 128, tie, clamp3=clamp3)
        assert bits_equal(got, want), mismatch_report(got, want)
    rows = golden["in/rows_tok"].astype(DT[dn])
    want = O.fake_quant(rows, fmt, None, tie, out_dtype=np.float16)
    got = P.fake_quant(rows, fmt, None, tie, out_dtype=np.float16)
    assert bits_equal(got, want), mismatch_report(got, want)


@pytest.mark.parametrize("tie", ["kernel", "argmin"])
@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("split", ["e1m2_neg_e2m1_pos", "int_neg_e2m3_pos", "afpq_e2m1"])
def test_port_signsplit(golden, split, dn, tie):
    with np.errstate(over="ignore"):
        x = np.concatenate([golden["in/adv"], golden["in/rnd"]]).astype(DT[dn])
    want = O.fake_quant_signsplit(x, split, 128, tie, clipping_strength=None)
    got = P.fake_quant_signsplit(x, split, 128, tie)
    assert bits_equal(got, want), mismatch_report(got, want)


def test_port_transform_rotate():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((33, 384)).astype(np.float32)
    s = np.exp(rng.uniform(-1, 1, 384)).astype(np.float32)
    q = O.block_random_hadamard_matrix(384, 128)
    want = O.transform_rotate_activation_f64(x, s, q)
    out, rot = P.transform_rotate_quant(x, s, "e2m1", return_rotated=True)
    # fp16 rounding of an fp32-accumulated product.  Stated tolerance: half an fp16 ulp of the exact
    # value (the final rounding) + 2e-6 * max|x*s| (fp32 accumulation over 128 terms)
    ulp = np.spacing(np.abs(want).astype(np.float16)).astype(np.float64)
    tol = 0.5 * ulp + 2e-6 * np.abs(x * s).max()
    assert np.all(np.abs(rot.astype(np.float64) - want) <= tol)
    # the quantizer half is the plain fp16 group quantizer applied to the port's own rotated values
    assert bits_equal(out, O.fake_quant(rot, "e2m1", 128, "kernel"))
    assert P.num_threads() >= 1
