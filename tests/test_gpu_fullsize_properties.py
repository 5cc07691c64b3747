"""GPU: size-independent properties at the FULL sizes of BASELINE.json's configs (the oracle is too slow there):
power-of-two scale equivariance, permutation equivariance inside a group, group locality, outputs on grid x scale,
and orthogonality / linearity of the rotation.  Every property is exact (bit-level) unless stated."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from fpqvar_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def sign_bits():
    from fpqvar_b200.hotpath import seed42_sign_bits
    return seed42_sign_bits()


def same_bits(a, b):
    return torch.equal(a.view(torch.int16 if a.dtype == torch.float16 else torch.int32), b.view(torch.int16 if b.dtype == torch.float16 else torch.int32))


FULL = {          # (rows, cols, dtype, op, fmt): the largest call of each site in configs[1..3]
    "d16_qkv_fp32": (32768, 1024, torch.float32, "sym", "e2m1"),
    "d30_proj_fp16": (25600, 1920, torch.float16, "sym", "e2m1"),
    "d30_fc2_fp16": (25600, 7680, torch.float16, "split", "e1m2_neg_e2m1_pos"),
    "d36_fc2_fp16": (20480, 9216, torch.float16, "split", "int_neg_e2m3_pos"),
    "d36_proj_fp6": (20480, 2304, torch.float16, "sym", "e2m3"),
    "config0_fp32": (4096, 4096, torch.float32, "sym", "e2m1"),
}


def run(ops, x, op, fmt):
    if op == "sym":
        return ops.fake_quant(x, fmt, 128, "kernel")
    return ops.fake_quant_signsplit(x, fmt, 128, "kernel")


@pytest.mark.parametrize("name", list(FULL))
def test_fullsize_quantizer_properties(ops, name):
    rows, cols, dt, op, fmt = FULL[name]
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 2 ** 31)
    x = torch.randn(rows, cols, device="cuda", generator=g)
    if op == "split":
        x = torch.nn.functional.gelu(x, approximate="tanh")
    x = x.to(dt)
    q = run(ops, x, op, fmt)
    assert q.shape == x.shape and q.dtype == x.dtype and torch.isfinite(q).all()
    # 1. power-of-two scale equivariance (no rounding is involved in scaling by 2^k inside the normal range)
    for k in (3, -4):
        assert same_bits(run(ops, x * (2.0 ** k), op, fmt), q * (2.0 ** k)), f"scale 2^{k}"
    # 2. permutation equivariance inside groups: the same permutation of the 128 positions in every group
    perm = torch.randperm(128, device="cuda", generator=g)
    xp = x.view(-1, 128)[:, perm].contiguous().view_as(x)
    assert same_bits(run(ops, xp, op, fmt), q.view(-1, 128)[:, perm].contiguous().view_as(x))
    # 3. group locality: rewriting one group changes that group's outputs only
    x2 = x.clone()
    x2.view(-1, 128)[12345] = torch.linspace(-3, 5, 128, device="cuda").to(dt)
    q2 = run(ops, x2, op, fmt)
    diff = (q2.view(-1, 128).view(torch.int16 if dt == torch.float16 else torch.int32) != q.view(-1, 128).view(torch.int16 if dt == torch.float16 else torch.int32)).any(1)
    assert diff.nonzero().flatten().tolist() in ([12345], [])
    # 4. every group uses at most as many distinct values as its format has levels, and reaches its absmax exactly when the
    #    format's largest level times the scale is representable (symmetric formats: |q|max == |x|max up to the scale rounding)
    levels = {"e2m1": 15, "e1m2": 15, "e3m0": 15, "e2m3": 63, "e3m2": 63, "e1m2_neg_e2m1_pos": 15, "int_neg_e2m3_pos": 64}[fmt]
    sample = q.view(-1, 128)[:: max(1, q.numel() // 128 // 4096)].float()
    n_distinct = torch.tensor([row.unique().numel() for row in sample[:512]])
    assert int(n_distinct.max()) <= levels
    # 5. quantization error is bounded by half the largest step of the format times the scale (+ the fp16 roundings of the
    #    scale and of the normalised value, which can turn a near-midpoint into an exact tie: 2^-9 relative slack)
    amax = x.view(-1, 128).float().abs().amax(1, keepdim=True)
    rel_step = {"e2m1": 2 / 6, "e3m0": 8 / 16, "e2m3": 0.5 / 7.5, "e1m2_neg_e2m1_pos": 2 / 6, "int_neg_e2m3_pos": 0.5 / 7.5}[fmt]
    err = (q.view(-1, 128).float() - x.view(-1, 128).float()).abs()
    assert bool((err <= 0.5 * rel_step * amax * 1.01 + amax * 2.0 ** -9 + 1e-30).all())


def test_fullsize_rotation_properties(ops, sign_bits):
    """configs[2] mat_qkv / fc1 input at the last stage: [25600, 1920] fp32."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(25600, 1920, device="cuda", generator=g)
    s = torch.exp(torch.rand(1920, device="cuda", generator=g) * 2 - 1)
    q, rot = ops.transform_rotate_quant(x, s, sign_bits, "e2m1", return_rotated=True)
    # orthogonality: the rotation preserves the norm of every 128-chunk of x*s (fp16 output: relative 2^-11 per element)
    n_in = (x * s).view(-1, 128).double().norm(dim=1)
    n_out = rot.view(-1, 128).double().norm(dim=1)
    assert torch.allclose(n_in, n_out, rtol=2e-3)
    # power-of-two equivariance of the whole fused op: bit-exact for the quantized output, and for every rotated value
    # that is a normal fp16 number (a subnormal fp16 result has fewer bits than its 8x counterpart)
    q8, rot8 = ops.transform_rotate_quant(x * 8.0, s, sign_bits, "e2m1", return_rotated=True)
    normal = rot.abs() >= 2.0 ** -14
    assert torch.equal(rot8[normal], (rot * 8.0)[normal]) and float((rot8[~normal].float() - 8 * rot[~normal].float()).abs().max()) <= 8 * 2.0 ** -25
    assert same_bits(q8, q * 8.0)
    # the fused quantizer equals the standalone fp16 quantizer on the rotated values (bit-exact)
    assert same_bits(q, ops.fake_quant(rot, "e2m1", 128, "kernel"))
    # linearity before the fp16 rounding: rotate(x) + rotate(y) ~= rotate(x + y) within 3 fp16 roundings
    y = torch.randn(25600, 1920, device="cuda", generator=g)
    ry = ops.transform_rotate_quant(y, s, sign_bits, None)
    rxy = ops.transform_rotate_quant(x + y, s, sign_bits, None)
    tol = 3 * 2.0 ** -11 * (rot.float().abs() + ry.float().abs() + rxy.float().abs()) + 1e-6
    assert bool(((rot.float() + ry.float() - rxy.float()).abs() <= tol).all())


def test_fullsize_weight_rotation_round_trip(ops, sign_bits):
    """configs[2] fc1 weight [7680, 1920]: rotating twice with the involution H diag(sigma) ... is not the identity, but
    W Q Q^T = W: apply Q, then Q^T = (FWHT then signs) emulated with the dense block on the GPU in fp64."""
    from fpqvar_b200 import rotation_utils as R
    g = torch.Generator(device="cuda").manual_seed(6)
    w = torch.randn(7680, 1920, device="cuda", generator=g) * 0.02
    wr = ops.transform_rotate_weight(w, None, sign_bits)
    q = R.block_random_hadamard_matrix(1920, 128, "cuda", 42)
    back = (wr.double() @ q.T).float()
    assert torch.allclose(back, w, rtol=0, atol=4e-9 + 2e-7 * float(w.abs().max()))
