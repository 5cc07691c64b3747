"""SURVEY.md section 8 row a10, pinned to the reference's own search scripts (tests/golden/make_golden_search.py ->
tests/golden/reference_search.npz: FPQuant.forward, FPQuant_e1m2_neg_e2m1_pos.forward, fp4_quant, fp6_quant,
compute_quant_error and one pass of the per-layer search loop, run by importing search/search_fp{4,6}_format.py).

CPU: the oracle against the fixture.  GPU (-m gpu): `fpqvar_b200.search` against the fixture."""
import os

import numpy as np
import pytest

from conftest import ROOT, bits_equal, mismatch_report
from oracle import oracle as O

DT = {"f32": np.float32, "f16": np.float16}
FP4 = ("e1m2", "e2m1", "e3m0")


@pytest.fixture(scope="module")
def S():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_search.npz"))


def _want(S, tag, dtype_name):
    assert str(S[f"dtype/{tag}"]) == dtype_name, f"{tag}: reference returns {S[f'dtype/{tag}']}"
    return S[f"out/{tag}"]


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["finite", "rnd", "gelu"])
def test_oracle_fpquant_and_fp4_quant(S, iname, dn):
    x = S[f"in/{iname}"].astype(DT[dn])
    for fmt in FP4:
        # FPQuant.forward (search_fp4_format.py:340-363): whole-tensor clip (identity at 1.0), argmin rounding, NO +-3 clamp, fp32 out
        got = O.fake_quant(x, fmt, 128, "argmin", clamp3=False)
        want = _want(S, f"FPQuant/{fmt}/{iname}/{dn}", "torch.float32")
        assert bits_equal(got.astype(np.float32), want), f"FPQuant {fmt} {iname} {dn}\n" + mismatch_report(got.astype(np.float32), want)
        # fp4_quant (:544-553) -> the script's own *_cuda copies: kernel rounding, input dtype out
        got = O.fake_quant(x, fmt, 128, "kernel")
        want = _want(S, f"fp4_quant/{fmt}/{iname}/{dn}", {"f32": "torch.float32", "f16": "torch.float16"}[dn])
        assert bits_equal(got.astype(np.float32), want), f"fp4_quant {fmt} {iname} {dn}"
    got = O.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "argmin")
    want = _want(S, f"FPQuant_e1m2_neg_e2m1_pos/{iname}/{dn}", "torch.float32")
    assert bits_equal(got.astype(np.float32), want), f"split {iname} {dn}\n" + mismatch_report(got.astype(np.float32), want)


@pytest.mark.parametrize("dn", ["f32", "f16"])
def test_oracle_fp6_quant_per_token(S, dn):
    x = S["in/tok"].astype(DT[dn])
    for fmt in ("e2m3", "e3m2"):
        got = O.fake_quant(x, fmt, None, "kernel", out_dtype=np.float16)
        want = _want(S, f"fp6_quant/{fmt}/tok/{dn}", "torch.float16")                # search_fp6_format.py:513-554: always fp16
        assert bits_equal(got.astype(np.float32), want), f"fp6_quant {fmt} {dn}"


def test_oracle_clipping_strength(S):
    x = S["in/rnd"]
    clip = np.float32(0.8) * np.abs(x).max()
    xc = np.clip(x, -clip, clip)
    assert bits_equal(O.fake_quant(xc, "e2m1", 128, "argmin"), S["out/FPQuant/e2m1/rnd/f32/clip0.8"])
    assert bits_equal(O.fake_quant_signsplit(x, "e1m2_neg_e2m1_pos", 128, "argmin", clipping_strength=0.8),
                      S["out/FPQuant_e1m2_neg_e2m1_pos/rnd/f32/clip0.8"])


def test_oracle_quant_error_and_search_loop(S):
    x = S["in/rnd"]
    q = O.fake_quant(x, "e2m1", 128, "kernel")
    assert np.isclose(O.tensor_mse(x, q), float(S["out/compute_quant_error/f32"][0]), rtol=1e-5)
    assert str(S["dtype/compute_quant_error/f16"]) == "torch.float16"                # fp16 in -> fp16 loss, as torch.mean
    w = S["loop/w"]
    acts = [S[f"loop/x{i}"] for i in range(3)]
    table = np.zeros((3, 3))
    for wi, wf in enumerate(FP4):
        wq = O.fake_quant(w, wf, 128, "kernel").astype(np.float64)
        for ai, af in enumerate(FP4):
            for t in acts:
                tq = O.fake_quant(t, af, 128, "kernel").astype(np.float64)
                d = t.astype(np.float64) @ w.astype(np.float64).T - tq @ wq.T
                table[wi, ai] += np.mean(d * d) / len(acts)
    assert np.allclose(table, S["loop/loss"], rtol=1e-4)
    assert np.unravel_index(np.argmin(table), table.shape) == np.unravel_index(np.argmin(S["loop/loss"]), (3, 3))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["finite", "rnd", "gelu"])
def test_gpu_search_mirrors(S, iname, dn):
    import torch
    from fpqvar_b200 import search
    x = torch.from_numpy(S[f"in/{iname}"]).to({"f32": torch.float32, "f16": torch.float16}[dn]).cuda()
    for fmt in FP4:
        got = search.FPQuant.apply(x.clone(), 4, 128, fmt)
        assert str(got.dtype) == str(S[f"dtype/FPQuant/{fmt}/{iname}/{dn}"])
        assert bits_equal(got.float().cpu().numpy(), S[f"out/FPQuant/{fmt}/{iname}/{dn}"]), f"FPQuant {fmt}"
        got = search.fp4_quant(x.clone(), fmt)
        assert str(got.dtype) == str(S[f"dtype/fp4_quant/{fmt}/{iname}/{dn}"])
        assert bits_equal(got.float().cpu().numpy(), S[f"out/fp4_quant/{fmt}/{iname}/{dn}"]), f"fp4_quant {fmt}"
    got = search.FPQuant_e1m2_neg_e2m1_pos.apply(x.clone(), 4, 128)
    assert str(got.dtype) == str(S[f"dtype/FPQuant_e1m2_neg_e2m1_pos/{iname}/{dn}"])
    assert bits_equal(got.float().cpu().numpy(), S[f"out/FPQuant_e1m2_neg_e2m1_pos/{iname}/{dn}"])


@pytest.mark.gpu
def test_gpu_search_fp6_clip_and_loop(S):
    import torch
    from fpqvar_b200 import search
    for dn, dt in (("f32", torch.float32), ("f16", torch.float16)):
        x = torch.from_numpy(S["in/tok"]).to(dt).cuda()
        for fmt in ("e2m3", "e3m2"):
            got = search.fp6_quant(x, fmt)
            assert got.dtype == torch.float16
            assert bits_equal(got.float().cpu().numpy(), S[f"out/fp6_quant/{fmt}/tok/{dn}"])
    x = torch.from_numpy(S["in/rnd"]).cuda()
    assert bits_equal(search.FPQuant.apply(x.clone(), 4, 128, "e2m1", 0.8).cpu().numpy(), S["out/FPQuant/e2m1/rnd/f32/clip0.8"])
    assert bits_equal(search.FPQuant_e1m2_neg_e2m1_pos.apply(x.clone(), 4, 128, 0.8).cpu().numpy(), S["out/FPQuant_e1m2_neg_e2m1_pos/rnd/f32/clip0.8"])
    w = torch.from_numpy(S["loop/w"]).cuda()
    acts = [torch.from_numpy(S[f"loop/x{i}"]).cuda() for i in range(3)]
    for fn in (search.search_layer, search.search_layer_batched):
        loss = fn(w, acts).cpu().numpy()
        assert np.allclose(loss, S["loop/loss"], rtol=2e-4), fn.__name__
