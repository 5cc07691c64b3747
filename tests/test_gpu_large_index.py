"""GPU: tensors with more than 2^31 (and 2^32) elements -- the "maximum sizes" edge of the path.  Every kernel indexes
with 64-bit offsets; a 32-bit product anywhere would wrap exactly here.  The oracle cannot run at this size, so the check
is group / row locality: slices of the big result (start, around the 2^31 and 2^32 element boundaries, end) must equal,
bit for bit, the same slices quantized on their own (which the parity tests pin to the oracle)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N31 = 1 << 31


@pytest.fixture(scope="module")
def ops():
    from fpqvar_b200 import ops as _ops
    return _ops


def bits(t):
    return t.view(torch.int16 if t.dtype == torch.float16 else torch.int32)


def fill(n, dtype, row_len):
    """n elements of cheap deterministic data with group-dependent magnitudes (a wrapped index would pick another group)."""
    x = torch.empty(n, dtype=dtype, device="cuda")
    chunk = 1 << 27
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        idx = torch.arange(i, i + m, device="cuda", dtype=torch.int64)
        g = (idx // row_len) % 8191
        v = ((idx * 2654435761) % 2001 - 1000).to(torch.float32) * (g.to(torch.float32) + 1.0) * 1e-3
        x[i:i + m] = v.to(dtype)
    return x


def windows(n, row_len, span_rows=64):
    """row-aligned windows: start, both sides of the 2^31 / 2^32 element boundaries that exist, end"""
    out = [0]
    for b in (N31, 2 * N31):
        if b < n:
            out.append((b // row_len - span_rows // 2) * row_len)
    out.append((n // row_len - span_rows) * row_len)
    return [(o, o + span_rows * row_len) for o in out]


@pytest.mark.parametrize("case", ["sym_f16_group", "split_f16_group", "sym_f32_group", "sym_f16_rows7680", "split_f16_rows7680",
                                  "sym_f16_rows64"])
def test_more_than_2pow31_elements(ops, case):
    free, _ = torch.cuda.mem_get_info()
    dtype = torch.float32 if "f32" in case else torch.float16
    row_len = 7680 if "7680" in case else (64 if "rows64" in case else 128)
    per_row = "rows" in case
    n = ((2 * N31 if dtype == torch.float16 else N31) // row_len + 1000) * row_len        # fp16: past 2^32 elements as well
    if free < n * dtype.itemsize * 2 + (4 << 30):
        pytest.skip("not enough free HBM")
    x = fill(n, dtype, row_len).view(-1, row_len)

    def q(t):
        if case.startswith("split"):
            return ops.fake_quant_signsplit(t, "int_neg_e2m3_pos" if per_row else "e1m2_neg_e2m1_pos", None if per_row else 128, "kernel")
        return ops.fake_quant(t, "e2m3" if per_row else "e2m1", None if per_row else 128, "kernel")

    y = q(x).view(-1)
    assert y.numel() == n
    xf = x.view(-1)
    for lo, hi in windows(n, row_len):
        part = q(xf[lo:hi].view(-1, row_len).clone()).view(-1)
        assert torch.equal(bits(y[lo:hi]), bits(part)), f"{case}: window [{lo}, {hi})"
        assert bool((part != 0).any())
    del x, y
    torch.cuda.empty_cache()


def test_rotate_more_than_2pow31_elements(ops):
    from fpqvar_b200.hotpath import seed42_sign_bits
    free, _ = torch.cuda.mem_get_info()
    C = 1920
    rows = N31 // C + 500                                                  # > 2^31 fp32 elements: 8.6 GB in, 4.3 GB out
    if free < rows * C * 6 + (6 << 30):
        pytest.skip("not enough free HBM")
    x = fill(rows * C, torch.float32, C).view(rows, C)
    s = torch.exp(torch.linspace(-1, 1, C, device="cuda"))
    bits_ = seed42_sign_bits()
    y = ops.transform_rotate_quant(x, s, bits_, "e2m1")
    for lo, hi in windows(rows * C, C, 32):
        part = ops.transform_rotate_quant(x.view(-1)[lo:hi].view(-1, C).clone(), s, bits_, "e2m1")
        assert torch.equal(bits(y.view(-1)[lo:hi]), bits(part.view(-1))), f"rotate window [{lo}, {hi})"
    # adaLN variant: batches of 4096 rows
    B = rows // 4096
    xm = x[:B * 4096].view(B, 4096, C)
    sc = torch.randn(B, 1, C, device="cuda") * 0.3
    sh = torch.randn(B, 1, C, device="cuda") * 0.5
    ym = ops.modulate_transform_rotate_quant(xm, sc, sh, s, bits_, "e2m1")
    for b in (0, B // 2, B - 1):
        part = ops.modulate_transform_rotate_quant(xm[b:b + 1].clone(), sc[b:b + 1].clone(), sh[b:b + 1].clone(), s, bits_, "e2m1")
        assert torch.equal(bits(ym[b:b + 1]), bits(part)), f"modulate batch {b}"
