"""GPU: incremental KV-cache fake quantization (fpqvar_b200/kv_cache.py) against the reference's schedule
(basic_var.py:188-203: re-quantize the whole cache at every scale, then append) -- bit-exact at every scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PATCH = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)


def reference_schedule(ks, vs, kv_bit):
    """the reference, literally: cached = quant(cached); cached = cat(cached, new) -- with this repo's drop-in quantizers"""
    from fpqvar_b200 import quant_utils as Q
    quant = (lambda t: Q.fp6_quant_e2m3_per_token_cuda(t, 6)) if kv_bit == 6 else (lambda t: Q.fp_quant_e2_per_group_cuda(t, 4))
    ck = cv = None
    out = []
    for k, v in zip(ks, vs):
        if ck is None:
            ck, cv = k, v
        else:
            ck, cv = quant(ck), quant(cv)
            ck, cv = torch.cat((ck, k), dim=1), torch.cat((cv, v), dim=1)
        out.append((ck, cv))
    return out


def same(a, b):
    return torch.equal(a.contiguous().view(torch.int16), b.contiguous().view(torch.int16))


@pytest.mark.parametrize("kv_bit", [6, 4])
def test_incremental_kv_quant_equals_the_reference_schedule(kv_bit):
    from fpqvar_b200.kv_cache import IncrementalKVQuant
    from fpqvar_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(kv_bit)
    B, H, hd = 6, 4, 64
    ks = [torch.nn.functional.normalize(torch.randn(B, p * p, H, hd, device="cuda", generator=g), dim=-1).half() for p in PATCH]
    vs = [(torch.randn(B, p * p, H, hd, device="cuda", generator=g) * torch.exp(torch.randn(B, p * p, H, 1, device="cuda", generator=g) * 2)).half()
          for p in PATCH]
    vs[2][0, 3] = 0                                                  # an all-zero token
    ref = reference_schedule(ks, vs, kv_bit)
    inc, full = IncrementalKVQuant(kv_bit, sum(p * p for p in PATCH)), IncrementalKVQuant(kv_bit, sum(p * p for p in PATCH), incremental=False)
    n0 = ops.launch_count()
    for (k, v), (rk, rv) in zip(zip(ks, vs), ref):
        ik, iv = inc.append(k, v)
        assert same(ik, rk) and same(iv, rv)
    n_inc = ops.launch_count() - n0
    for (k, v), (rk, rv) in zip(zip(ks, vs), ref):
        fk, fv = full.append(k, v)
        assert same(fk, rk) and same(fv, rv)
    assert inc.exact and full.exact
    assert n_inc == 2 * (len(PATCH) - 1)                             # one quantizer launch per tensor and scale, on the new rows only
    # the history really is quantized: a further pass changes nothing (idempotence, tests/idempotence_exhaustive.py)
    from fpqvar_b200 import quant_utils as Q
    hist = inc.k[:, :inc.done].contiguous()
    again = Q.fp6_quant_e2m3_per_token_cuda(hist, 6) if kv_bit == 6 else Q.fp_quant_e2_per_group_cuda(hist, 4)
    assert same(again, hist)


def test_rows_outside_the_idempotent_range_are_reported():
    from fpqvar_b200.kv_cache import IncrementalKVQuant
    B, H, hd = 2, 2, 64
    k = [torch.randn(B, n, H, hd, device="cuda").half() for n in (1, 4, 9)]
    v = [torch.randn(B, n, H, hd, device="cuda").half() for n in (1, 4, 9)]
    v[0][0, 0, 0] = torch.linspace(-1, 1, hd, device="cuda").half() * 4.29e-6      # absmax 4.29e-6: the scale is deep-subnormal fp16
    inc, full = IncrementalKVQuant(6, 14), IncrementalKVQuant(6, 14, incremental=False)
    ref = reference_schedule(k, v, 6)
    for (kk, vv), (rk, rv) in zip(zip(k, v), ref):
        inc.append(kk, vv)
        fk, fv = full.append(kk, vv)
        assert same(fk, rk) and same(fv, rv)                                       # the reference's schedule is always available
    assert not inc.exact and full.exact
    inc.reset()
    assert inc.exact


def test_kv_cache_argument_checks():
    from fpqvar_b200.kv_cache import IncrementalKVQuant
    c = IncrementalKVQuant(4, 8)
    x = torch.zeros(1, 2, 3, 64, device="cuda", dtype=torch.float16)                 # 3 * 64 is not a multiple of 128
    with pytest.raises(ValueError):
        c.append(x, x)
    with pytest.raises(NotImplementedError):
        IncrementalKVQuant(8, 8)
    c6 = IncrementalKVQuant(6, 3)
    y = torch.zeros(1, 2, 2, 64, device="cuda", dtype=torch.float16)
    c6.append(y, y)
    with pytest.raises(ValueError):
        c6.append(y, y)


@pytest.mark.parametrize("kv_bit", [6, 4])
def test_incremental_kv_quant_equals_the_oracle_run_of_the_reference_schedule(kv_bit):
    """The whole schedule on the ORACLE side (numpy restatement of fp6_quant_e2m3_per_token_cuda / fp_quant_e2_per_group_cuda,
    re-quantizing the entire cache before every append exactly as basic_var.py:188-203 does), against the incremental
    in-place CUDA path at every scale."""
    import numpy as np
    from oracle import oracle as O
    from fpqvar_b200.kv_cache import IncrementalKVQuant
    rng = np.random.default_rng(40 + kv_bit)
    B, H, hd = 3, 2, 64
    patch = PATCH[:7]
    ks = [rng.standard_normal((B, p * p, H, hd)).astype(np.float16) for p in patch]
    vs = [(rng.standard_normal((B, p * p, H, hd)) * np.exp(2 * rng.standard_normal((B, p * p, H, 1)))).astype(np.float16) for p in patch]
    vs[1][0, 2] = 0

    def oq(t):
        if kv_bit == 6:
            return O.fake_quant(t.reshape(-1, hd), "e2m3", hd, "kernel").reshape(t.shape)
        return O.fake_quant(t.reshape(-1, 128), "e2m1", 128, "kernel").reshape(t.shape)

    inc = IncrementalKVQuant(kv_bit, sum(p * p for p in patch))
    ck = cv = None
    for k, v in zip(ks, vs):
        if ck is None:
            ck, cv = k, v
        else:
            ck, cv = np.concatenate((oq(ck), k), axis=1), np.concatenate((oq(cv), v), axis=1)
        ik, iv = inc.append(torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda())
        assert np.array_equal(ik.cpu().numpy().view(np.uint16), ck.view(np.uint16))
        assert np.array_equal(iv.cpu().numpy().view(np.uint16), cv.view(np.uint16))


@pytest.mark.parametrize("row_len,fmt", [(64, "e2m3"), (128, "e2m1"), (128, "e3m2"), (64, "e1m2")])
def test_fake_quant_segments_in_place_and_pitched(row_len, fmt):
    """fpq_fake_quant_segments: a slice [B, lo:hi] of a larger tensor, in place == the plain kernel on a gathered copy; the
    rest of the tensor is untouched; out-of-place with a different pitch works too."""
    from fpqvar_b200 import ops, _lib as L
    g = torch.Generator(device="cuda").manual_seed(row_len)
    B, Lmax, inner = 5, 37, 384
    buf = (torch.randn(B, Lmax, inner, device="cuda", generator=g) * 3).half()
    buf[1, 9, :128] = 0
    buf[2, 10, 5] = float("inf")
    buf[3, 11, 7] = float("nan")
    orig = buf.clone()
    lo, hi = 8, 29
    want = ops.fake_quant(orig[:, lo:hi].contiguous(), fmt, row_len, "kernel")
    ops.fake_quant_segments_(buf, lo, hi, fmt, row_len)
    a, b = buf[:, lo:hi].contiguous().view(torch.int16), want.view(torch.int16)
    nan = torch.isnan(want)
    assert torch.equal(torch.isnan(buf[:, lo:hi]), nan) and torch.equal(a[~nan], b[~nan])
    assert torch.equal(buf[:, :lo].view(torch.int16), orig[:, :lo].view(torch.int16))
    assert torch.equal(buf[:, hi:].view(torch.int16), orig[:, hi:].view(torch.int16))
    # out of place into a dense tensor (pitch_out = segment size)
    dense = torch.empty(B, hi - lo, inner, device="cuda", dtype=torch.float16)
    seg = (hi - lo) * inner
    rc = L.lib().fpq_fake_quant_segments(orig.data_ptr() + lo * inner * 2, dense.data_ptr(), B, seg // row_len, row_len, Lmax * inner, seg,
                                         L.FMT[fmt], torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    d = dense.view(torch.int16)
    assert torch.equal(d[~nan], b[~nan])
    # partial overlap is refused; so is a row length the packed kernels do not take
    assert L.lib().fpq_fake_quant_segments(orig.data_ptr(), orig.data_ptr() + 256, B, seg // row_len, row_len, Lmax * inner, Lmax * inner,
                                           L.FMT[fmt], None) == L.FPQ_ERR_ARG
    assert L.lib().fpq_fake_quant_segments(orig.data_ptr(), dense.data_ptr(), 1, 4, 96, 4096, 4096, L.FMT[fmt], None) == L.FPQ_ERR_UNSUPPORTED
