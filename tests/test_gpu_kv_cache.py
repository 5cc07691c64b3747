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
