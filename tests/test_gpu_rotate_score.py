"""GPU parity: fused transform+rotate(+quant), weight-side transform+rotate, batched format scoring,
and the host-buffer pipeline -- all through the C ABI, checked against the oracle.

Tolerances (stated here, as north_star asks): rotated/transformed VALUES are floating point and are
compared with the exact fp64 statement of the reference op:
  * activation path, fp16 output:  |y - y64| <= 0.5 ulp_fp16(y64) + 2e-6 * max|x*s|
    (final fp16 rounding + fp32 butterfly accumulation over 128 terms; the reference's own
    autocast-fp16 GEMM is ~1000x looser)
  * weight path, fp32 output from fp64 butterflies: |w - w64| <= 1 ulp_fp32(w64) + 128*2.3e-16*max|W/s|
    (the reference's fp64 GEMM sums in a different order; both are within 0.5 ulp + fp64 noise of exact)
QUANTIZED tensors are bit-exact: checked against the oracle quantizer applied to the kernel's own
rotated values."""
import numpy as np
import pytest
import torch

from conftest import bits_equal, mismatch_report
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from fpqvar_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


@pytest.fixture(scope="module")
def sign_bits():
    from fpqvar_b200.hotpath import seed42_sign_bits
    return seed42_sign_bits()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def test_sign_vector_constant_matches_torch_seed42():
    from fpqvar_b200.hotpath import SIGN_BITS_SEED42_128
    torch.manual_seed(42)
    bits = "".join(str(int(v)) for v in torch.randint(low=0, high=2, size=(128,)).tolist())
    assert bits == SIGN_BITS_SEED42_128 == O.SIGN_BITS_SEED42_128


@pytest.mark.parametrize("cols", [128, 1024, 1920, 2304])
@pytest.mark.parametrize("fmt", ["e2m1", "e2m3", None])
@pytest.mark.parametrize("with_smooth", [True, False])
def test_transform_rotate_quant(ops, sign_bits, cols, fmt, with_smooth):
    rng = np.random.default_rng(cols + (7 if with_smooth else 0))
    rows = 131
    x = (rng.standard_normal((rows, cols)) * np.exp(rng.uniform(-2, 2, (rows, 1)))).astype(np.float32)
    x[5] = 0.0
    s = np.exp(rng.uniform(-3, 1, cols)).astype(np.float32) if with_smooth else None
    if with_smooth:
        s[3] = -0.0883                                         # best_lambda_var36 holds non-positive entries
    out, rot = ops.transform_rotate_quant(dev(x), dev(s) if with_smooth else None, sign_bits, fmt, return_rotated=True)
    out, rot = host(out), host(rot)
    q = O.block_random_hadamard_matrix(cols, 128)
    want = O.transform_rotate_activation_f64(x, s if with_smooth else np.ones(cols, np.float32), q)
    xs = x * (s if with_smooth else np.float32(1))
    tol = 0.5 * np.spacing(np.abs(want).astype(np.float16)).astype(np.float64) + 2e-6 * np.abs(xs).max(axis=1, keepdims=True)
    err = np.abs(rot.astype(np.float64) - want)
    assert np.all(err <= tol), f"max excess {np.max(err - tol)}"
    if fmt is None:
        assert bits_equal(out, rot)
    else:
        want_q = O.fake_quant(rot, fmt, 128, "kernel")
        assert bits_equal(out, want_q), mismatch_report(out, want_q)


def test_transform_rotate_quant_matches_unfused_path(ops, sign_bits):
    """Fused kernel == (rotate only) followed by the standalone fp16 group quantizer, bit for bit."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((4097, 1920)).astype(np.float32)
    s = np.exp(rng.uniform(-1, 1, 1920)).astype(np.float32)
    fused = ops.transform_rotate_quant(dev(x), dev(s), sign_bits, "e2m1")
    rot = ops.transform_rotate_quant(dev(x), dev(s), sign_bits, None)
    two_step = ops.fake_quant(rot, "e2m1", 128, "kernel")
    assert torch.equal(fused.view(torch.int16), two_step.view(torch.int16))


@pytest.mark.parametrize("fmt", ["e2m1", "e2m3", None])
@pytest.mark.parametrize("B,Lr,C", [(3, 7, 256), (100, 16, 1920), (2, 1, 2304), (5, 33, 128)])
def test_modulate_transform_rotate_quant(ops, sign_bits, B, Lr, C, fmt):
    """adaLN modulate fused in: bit-identical to the reference's three ATen ops
    followed by the un-modulated fused kernel, and within the stated tolerance of the fp64 statement."""
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + C)
    x = torch.randn(B, Lr, C, device="cuda", generator=g)
    scale = torch.randn(B, 1, C, device="cuda", generator=g) * 0.3
    shift = torch.randn(B, 1, C, device="cuda", generator=g) * 0.5
    s = torch.exp(torch.rand(C, device="cuda", generator=g) * 2 - 1)
    fused, frot = ops.modulate_transform_rotate_quant(x, scale, shift, s, sign_bits, fmt, return_rotated=True)
    mod = x.mul(scale.add(1)).add_(shift)                                 # basic_var.py:263, on the GPU with ATen
    two, trot = ops.transform_rotate_quant(mod, s, sign_bits, fmt, return_rotated=True)
    assert torch.equal(frot.view(torch.int16), trot.view(torch.int16))
    assert torch.equal(fused.view(torch.int16), two.view(torch.int16))
    # oracle: the modulate restated in numpy, then the fp64 rotation
    mo = O.adaln_modulate(host(x), host(scale), host(shift))
    assert bits_equal(mo, host(mod))
    want = O.transform_rotate_activation_f64(mo.reshape(-1, C), host(s), O.block_random_hadamard_matrix(C, 128))
    xs = mo.reshape(-1, C) * host(s)
    tol = 0.5 * np.spacing(np.abs(want).astype(np.float16)).astype(np.float64) + 2e-6 * np.abs(xs).max(axis=1, keepdims=True)
    assert np.all(np.abs(host(frot).reshape(-1, C).astype(np.float64) - want) <= tol)
    if fmt is not None:
        assert bits_equal(host(fused), O.fake_quant(host(frot), fmt, 128, "kernel"))


@pytest.mark.parametrize("B,Lr,C", [(3, 7, 256), (100, 16, 1920), (4, 33, 2304)])
def test_modulate_with_fp16_adaln_tensors_matches_the_autocast_call_site(ops, sign_bits, B, Lr, C):
    """Under the reference's fp16 autocast (evaluate_fp_quant_transform_rotate.py:195) scale1 / shift1 come out of an
    autocast Linear as fp16 and `scale1.add(1)` is an fp16 add (basic_var.py:263): the fused call must reproduce
    that rounding (FPQ_MOD_GAIN), bit for bit, including scale == -1, tiny scales (1 + s rounds to 1) and -0."""
    g = torch.Generator(device="cuda").manual_seed(B + C)
    x = torch.randn(B, Lr, C, device="cuda", generator=g)
    scale = (torch.randn(B, 1, C, device="cuda", generator=g) * 0.3).half()
    shift = (torch.randn(B, 1, C, device="cuda", generator=g) * 0.5).half()
    scale[0, 0, :8] = torch.tensor([-1.0, 2.0 ** -12, -(2.0 ** -12), 6e-8, -6e-8, 0.0, -0.0, -1.0005], device="cuda").half()
    shift[0, 0, :8] = torch.tensor([0.0, -0.0, 0.0, -0.0, 1.0, 0.0, -0.0, 0.0], device="cuda").half()
    s = torch.exp(torch.rand(C, device="cuda", generator=g) * 2 - 1)
    fused, frot = ops.modulate_transform_rotate_quant(x, scale, shift, s, sign_bits, "e2m1", return_rotated=True)
    mod = x.mul(scale.add(1)).add_(shift)                                 # fp32 * fp16 -> fp32, as ATen promotes it
    assert mod.dtype == torch.float32
    two, trot = ops.transform_rotate_quant(mod, s, sign_bits, "e2m1", return_rotated=True)
    assert torch.equal(frot.view(torch.int16), trot.view(torch.int16))
    assert torch.equal(fused.view(torch.int16), two.view(torch.int16))
    # and it is NOT what an fp32 `+ 1` gives wherever the fp16 add rounds: the flag matters
    wide = ops.modulate_transform_rotate_quant(x, scale.float(), shift.float(), s, sign_bits, None)
    assert not torch.equal(wide.view(torch.int16), frot.view(torch.int16))
    # oracle: numpy restatement with the fp16 add
    gain = (host(scale).astype(np.float16) + np.float16(1)).astype(np.float16).astype(np.float32)
    mo = (host(x) * gain).astype(np.float32) + host(shift).astype(np.float32)
    assert bits_equal(mo.astype(np.float32), host(mod))


def test_modulate_argument_checks(ops, sign_bits):
    from fpqvar_b200._lib import FpqError
    x = torch.randn(4, 3, 256, device="cuda")
    ok = torch.zeros(4, 1, 256, device="cuda")
    with pytest.raises(FpqError):
        ops.modulate_transform_rotate_quant(x, torch.zeros(3, 1, 256, device="cuda"), ok, None, sign_bits, "e2m1")
    with pytest.raises(FpqError):
        ops.modulate_transform_rotate_quant(x, ok, ok.cpu(), None, sign_bits, "e2m1")
    out = ops.modulate_transform_rotate_quant(x, ok, ok, None, sign_bits, "e2m1")       # scale 0, shift 0 == no modulate
    assert torch.equal(out.view(torch.int16), ops.transform_rotate_quant(x, None, sign_bits, "e2m1").view(torch.int16))


@pytest.mark.parametrize("shape", [(384, 256), (1920 * 3, 1920), (100, 128)])
@pytest.mark.parametrize("with_smooth", [True, False])
def test_transform_rotate_weight(ops, sign_bits, shape, with_smooth):
    rng = np.random.default_rng(shape[0])
    w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    s = np.exp(rng.uniform(-3, 1, shape[1])).astype(np.float32) if with_smooth else None
    got = host(ops.transform_rotate_weight(dev(w), dev(s) if with_smooth else None, sign_bits))
    wt = O.transform_weight(w, s) if with_smooth else w
    want64 = wt.astype(np.float64) @ O.block_random_hadamard_matrix(shape[1], 128)
    want = want64.astype(np.float32)
    # 1 fp32 ulp of the exact value + the fp64 accumulation bound of a 128-term sum (matters only
    # where the sum cancels to ~0 and an fp32 ulp of the result is smaller than fp64 noise of the terms)
    tol = np.spacing(np.abs(want)).astype(np.float64) + 128 * 2.3e-16 * np.abs(wt).max()
    assert np.all(np.abs(got.astype(np.float64) - want64) <= tol)
    # and in practice nearly always the identical float
    assert np.mean(got.view(np.uint32) == want.view(np.uint32)) > 0.999
    # in-place variant
    wd = dev(w)
    r = ops.transform_rotate_weight(wd, dev(s) if with_smooth else None, sign_bits, inplace=True)
    assert r.data_ptr() == wd.data_ptr() and bits_equal(host(wd), got)


def test_rotation_is_orthogonal_round_trip(ops, sign_bits):
    """Size-independent property: Q Q^T = I, so rotating the rotated weight with the transposed
    block (= FWHT then signs) returns the input; here: ||W Q|| == ||W|| row-wise."""
    rng = np.random.default_rng(2)
    w = rng.standard_normal((2048, 2304)).astype(np.float32)
    r = host(ops.transform_rotate_weight(dev(w), None, sign_bits)).astype(np.float64)
    n0 = np.linalg.norm(w.astype(np.float64).reshape(-1, 128), axis=1)
    n1 = np.linalg.norm(r.reshape(-1, 128), axis=1)
    assert np.allclose(n0, n1, rtol=1e-6)


ALL_FORMATS = ["e2m1", "e1m2", "e3m0", "e2m3", "e3m2", "e1m2_neg_e2m1_pos", "int_neg_e2m3_pos", "afpq_e2m1"]


def _oracle_sse(x, fmt, tie):
    if fmt in O.SPLIT:
        xq = O.fake_quant_signsplit(x, fmt, 128, tie, clipping_strength=None)
    else:
        xq = O.fake_quant(x, fmt, 128, tie)
    d = x.astype(np.float64) - xq.astype(np.float64)
    return float((d * d).sum())


@pytest.mark.parametrize("tie", ["kernel", "argmin"])
@pytest.mark.parametrize("dn", ["f32", "f16"])
def test_score_formats(ops, dn, tie):
    rng = np.random.default_rng(17)
    x = rng.standard_normal((3001, 128)).astype(np.float32)
    x[100:900] = 0.5 * x[100:900] * (1 + np.tanh(0.79788456 * (x[100:900] + 0.044715 * x[100:900] ** 3)))   # GELU-skewed groups
    x[7] = 0
    x = x.astype({"f32": np.float32, "f16": np.float16}[dn])
    sse = host(ops.score_formats(dev(x), ALL_FORMATS, tie))
    for v, fmt in zip(sse, ALL_FORMATS):
        want = _oracle_sse(x, fmt, tie)
        # per-element errors are exact; only the summation order differs (fp32 partial sums of 16, then fp64)
        assert abs(v - want) <= 2e-6 * want, (fmt, v, want)
    # ranking = the search's argmin (search/search_fp4_format.py:818-821)
    fp4 = host(ops.score_formats(dev(x), ["e1m2", "e2m1", "e3m0"], tie))
    want_rank = int(np.argmin([_oracle_sse(x, f, tie) for f in ("e1m2", "e2m1", "e3m0")]))
    assert int(np.argmin(fp4)) == want_rank


def test_score_formats_accumulates_and_is_linear(ops):
    """Size-independent property at full size: sse(x1 ++ x2) == sse(x1) + sse(x2), and accumulation
    into a caller-provided table."""
    torch.manual_seed(3)
    x = torch.randn(4096 * 4096 // 128, 128, device="cuda")
    fm = ["e2m1", "e1m2", "e3m0"]
    whole = ops.score_formats(x, fm)
    acc = ops.score_formats(x[: x.shape[0] // 3], fm)
    ops.score_formats(x[x.shape[0] // 3:], fm, sse=acc)
    assert torch.allclose(whole, acc, rtol=1e-9)
    # agrees with quantize-then-MSE on the GPU path itself
    for i, f in enumerate(fm):
        d = (x - ops.fake_quant(x, f, 128, "kernel")).double()
        assert abs(float(whole[i]) - float((d * d).sum())) <= 1e-6 * float(whole[i])


def test_host_pipeline_matches_device_path(ops):
    from fpqvar_b200.hotpath import HostPipeline, run_call
    from fpqvar_b200.var_workload import VarHotPath
    hot = VarHotPath("tiny", 2, 1, (1, 2, 3, 4), True)
    calls = hot.calls()
    devc = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    smooth = {s: torch.exp(torch.rand(hot.width, generator=g) * 2 - 1).cuda() for s in ("mat_qkv", "fc1")}
    pipe = HostPipeline(devc, max(c.in_bytes for c in calls), max(c.out_bytes for c in calls), smooth)
    h_in, h_out, xs = [], [], []
    for c in calls:
        x = torch.randn(c.rows, c.cols, generator=g)
        if c.site == "fc2":
            x = torch.nn.functional.gelu(x, approximate="tanh")
        x = x.to(torch.float16 if c.in_dtype == "f16" else torch.float32)
        xs.append(x)
        h_in.append(x.view(-1).view(torch.uint8).pin_memory())
        h_out.append(torch.empty(c.out_bytes, dtype=torch.uint8).pin_memory())
    pipe.run(calls, h_in, h_out)
    pipe.synchronize()
    for c, x, ho in zip(calls, xs, h_out):
        want = run_call(c, x.cuda(), smooth.get(c.site)).cpu().view(-1).view(torch.uint8)
        assert torch.equal(ho, want), (c.site, c.stage)


def test_device_replay_in_a_cuda_graph_matches_oracle():
    """The bench's device path at small scale: every call of a tiny VAR pass captured in ONE CUDA graph (as bench.py
    does), replayed twice, every output checked against the oracle."""
    from fpqvar_b200.hotpath import DeviceReplay
    from fpqvar_b200.var_workload import VarHotPath
    hot = VarHotPath("tiny", 2, 1, (1, 2, 3), True)
    calls = hot.calls()
    dev_ = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    smooth = {s: torch.exp(torch.rand(hot.width, generator=g) * 2 - 1).cuda() for s in ("mat_qkv", "fc1")}
    rep = DeviceReplay(dev_, smooth)
    xs, outs = [], []
    for c in calls:
        x = torch.randn(c.rows, c.cols, generator=g)
        if c.site == "fc2":
            x = torch.nn.functional.gelu(x, approximate="tanh")
        xs.append(x.to(torch.float16 if c.in_dtype == "f16" else torch.float32).cuda())
        outs.append(torch.zeros(c.rows, c.cols, dtype=torch.float16 if c.out_dtype == "f16" else torch.float32, device="cuda"))
    side = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for c, x, o in zip(calls[:2], xs, outs):
            rep.launch(c, x.data_ptr(), o.data_ptr(), side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(gr, stream=side):
            for c, x, o in zip(calls, xs, outs):
                rep.launch(c, x.data_ptr(), o.data_ptr(), side.cuda_stream)
    for o in outs:
        o.zero_()
    gr.replay()
    gr.replay()
    torch.cuda.synchronize()
    q = O.block_random_hadamard_matrix(hot.width, 128)
    for c, x, o in zip(calls, xs, outs):
        xn, got = host(x), host(o)
        if c.op == "group":
            want = O.fake_quant(xn, c.fmt, 128, "kernel")
        elif c.op == "signsplit":
            want = O.fake_quant_signsplit(xn, c.fmt, 128, "kernel")
        else:
            # rotated values are tolerance-checked elsewhere; here: the graph path equals the eager op bit for bit
            want = host(ops_module().transform_rotate_quant(x, smooth[c.site], rep.sign_bits, c.fmt))
            assert q.shape[0] == hot.width
        assert bits_equal(got, want), (c.site, c.stage, c.block)


def ops_module():
    from fpqvar_b200 import ops as _ops
    return _ops


def test_rotation_bits_do_not_depend_on_the_launch_size(ops, sign_bits):
    """The launcher picks a different kernel layout for small tensors; a row must rotate and quantize to the same
    bits whether it is processed alone, in a small batch or in a large one."""
    torch.manual_seed(9)
    x = torch.randn(40000, 256, device="cuda")                      # 80000 chunks: the streaming kernel
    s = torch.exp(torch.rand(256, device="cuda") * 2 - 1)
    big, big_rot = ops.transform_rotate_quant(x, s, sign_bits, "e2m1", return_rotated=True)
    for n in (1, 7, 500, 12288):                                     # <= 40000 chunks: the small-launch kernel
        small, small_rot = ops.transform_rotate_quant(x[:n].contiguous(), s, sign_bits, "e2m1", return_rotated=True)
        assert torch.equal(small_rot.view(torch.int16), big_rot[:n].view(torch.int16)), n
        assert torch.equal(small.view(torch.int16), big[:n].view(torch.int16)), n


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_sse_rows_matches_the_unfused_loss(dtype):
    """fpq_sse_rows == sum_r w[r] * sum_c (a - b)^2 in float64 (search_fp4_format.py:472-476 compute_quant_error, row-weighted
    for the row-stacked calibration set); stated tolerance 2e-6 relative (fp32 products inside a lane, fp64 above)."""
    from fpqvar_b200 import ops, _lib as L
    g = torch.Generator(device="cuda").manual_seed(3)
    for rows, cols in ((1, 128), (37, 1920), (1000, 5760), (5, 8)):
        a = (torch.randn(rows, cols, device="cuda", generator=g) * 3).to(dtype)
        b = (a.float() + 0.01 * torch.randn(rows, cols, device="cuda", generator=g)).to(dtype)
        w = torch.rand(rows, device="cuda", generator=g, dtype=torch.float64)
        d2 = (a.double() - b.double()).square().sum(dim=1)
        got = ops.sse_rows(a, b)
        assert abs(float(got) - float(d2.sum())) <= 2e-6 * float(d2.sum()) + 1e-30
        acc = torch.full((), 5.0, dtype=torch.float64, device="cuda")
        ops.sse_rows(a, b, w, out=acc)
        want = 5.0 + float((d2 * w).sum())
        assert abs(float(acc) - want) <= 2e-6 * want
    x = torch.zeros(4, 12, device="cuda", dtype=dtype)               # 12 columns do not fill 16-byte vectors in fp16; 12 fp32 do
    if dtype == torch.float16:
        with pytest.raises(L.FpqError):
            ops.sse_rows(x, x)
    with pytest.raises(L.FpqError):
        ops.sse_rows(x, x.to(torch.float32 if dtype == torch.float16 else torch.float16))


@pytest.mark.gpu
def test_smooth_of_any_dtype_in_a_back_to_back_launch_sequence(ops, sign_bits):
    """GALT factors that are not fp32-contiguous (fp16, fp64, a strided slice) are converted by a cast kernel launched
    right in front of the rotate kernel, which is launched with programmatic stream serialization: the kernels may read
    `smooth` only after their dependency wait.  Forty back-to-back launches of every kind must equal the fp32 result."""
    g = torch.Generator(device="cuda").manual_seed(11)
    cols = 1920
    for rows in (64, 4096):                                       # small-launch kernel and streaming kernel
        x = torch.randn(rows, cols, device="cuda", generator=g)
        s32 = torch.exp(torch.rand(cols, device="cuda", generator=g) * 2 - 1)
        wide = torch.zeros(cols, 2, device="cuda")
        variants = {"f16": s32.half(), "f64": s32.half().double(), "strided": None}
        base = s32.half().float()                                  # all variants hold the same values
        wide[:, 0] = base
        variants["strided"] = wide[:, 0]
        want = ops.transform_rotate_quant(x, base, sign_bits, "e2m1")
        sc = torch.randn(rows // 64, 1, cols, device="cuda", generator=g) * 0.3
        sh = torch.randn(rows // 64, 1, cols, device="cuda", generator=g) * 0.5
        want_mod = ops.modulate_transform_rotate_quant(x.view(rows // 64, 64, cols), sc, sh, base, sign_bits, "e2m1")
        for name, s in variants.items():
            for _ in range(40):
                got = ops.transform_rotate_quant(x, s, sign_bits, "e2m1")
                got_mod = ops.modulate_transform_rotate_quant(x.view(rows // 64, 64, cols), sc, sh, s, sign_bits, "e2m1")
            assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (rows, name)
            assert torch.equal(got_mod.view(torch.int16), want_mod.view(torch.int16)), (rows, name)


@pytest.mark.gpu
def test_rotation_against_the_reference_autocast_gemm_is_bounded(ops, sign_bits):
    """Under the reference's fp16 autocast `torch.matmul(x.mul(s), Q)` (basic_var.py:263) rounds BOTH GEMM inputs to fp16
    first (x*s, and Q, whose 1/sqrt(128) becomes 0.088379: a gain error of 1.07e-4), accumulates in fp32 and rounds the
    result to fp16.  This library keeps x*s in fp32 and uses the exact Q (the fp32 statement the north star asks for), so
    rotated activations are NOT bit-identical to that path.  The difference is pinned here: at most 2^-8 * max|x*s| per
    element, and below 1e-3 relative RMS over a tensor (expected 4e-4: the gain error plus two 2^-11 roundings)."""
    from fpqvar_b200 import rotation_utils
    g = torch.Generator(device="cuda").manual_seed(12)
    rows, cols = 2048, 1920
    x = torch.randn(rows, cols, device="cuda", generator=g) * 2
    s = torch.exp(torch.rand(cols, device="cuda", generator=g) * 2 - 1)
    ours = ops.transform_rotate_quant(x, s, sign_bits, None).float()
    Q = rotation_utils.block_random_hadamard_matrix(total_size=cols, block_size=128, device="cuda", seed=42).to(torch.float32)
    with torch.autocast("cuda", enabled=True, dtype=torch.float16):
        ref = torch.matmul(x.mul(s), Q)
    assert ref.dtype == torch.float16
    ref = ref.float()
    xs_max = float((x * s).abs().max())
    assert float((ours - ref).abs().max()) <= 4 * 2.0 ** -10 * xs_max
    assert float((ours - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()) < 1e-3


@pytest.mark.gpu
def test_gelu_device_function_reproduces_atens_fp16_tanh_gelu_for_every_input():
    """All 65 536 fp16 bit patterns: the library's GELU(tanh) == torch.nn.functional.gelu(x, approximate="tanh") on a CUDA Half
    tensor, bit for bit (NaN in, NaN out) -- the precondition for fusing `self.act` (basic_var.py:108) into the fc2 quantizer."""
    from fpqvar_b200 import ops
    x = torch.arange(65536, device="cuda", dtype=torch.int32).to(torch.int16).view(torch.float16)
    want = torch.nn.functional.gelu(x, approximate="tanh")
    got = ops.gelu_table()
    nan = torch.isnan(want)
    assert torch.equal(torch.isnan(got), nan)
    bad = (got.view(torch.int16) != want.view(torch.int16)) & ~nan
    assert int(bad.sum()) == 0, f"{int(bad.sum())} of 65536 inputs differ, first at bit pattern {int(bad.nonzero()[0])}"


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["e1m2_neg_e2m1_pos", "int_neg_e2m3_pos", "afpq_e2m1"])
def test_gelu_fused_signsplit_equals_the_two_step_sequence(fmt):
    """fpq_gelu_fake_quant_signsplit == fake_quant_signsplit(F.gelu(x, approximate="tanh")) bit for bit, special values
    included (zero groups, tiny scales, inf, a NaN group with and without the whole-tensor clip)."""
    from fpqvar_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(21)
    x = (torch.randn(4096, 7680, device="cuda", generator=g) * 1.3).half()
    x[0, :128] = 0
    x[1, :128] = x[1, :128] * 1e-4
    x[2, 5] = float("inf")
    x[3, :256] = -x[3, :256].abs()                       # all-negative groups: GELU output in (-0.17, 0]
    x[4, :128] = x[4, :128].abs() * 8
    for clip in (False, True):
        want = ops.fake_quant_signsplit(torch.nn.functional.gelu(x, approximate="tanh"), fmt, 128, "kernel", global_clip=clip)
        got = ops.gelu_fake_quant_signsplit(x, fmt, global_clip=clip)
        nan = torch.isnan(want)
        assert torch.equal(torch.isnan(got), nan)
        assert torch.equal(got.view(torch.int16)[~nan], want.view(torch.int16)[~nan])
    y = x.clone()
    y[7, 300] = float("nan")
    for clip in (False, True):
        want = ops.fake_quant_signsplit(torch.nn.functional.gelu(y, approximate="tanh"), fmt, 128, "kernel", global_clip=clip)
        got = ops.gelu_fake_quant_signsplit(y, fmt, global_clip=clip)
        nan = torch.isnan(want)
        assert torch.equal(torch.isnan(got), nan)
        assert torch.equal(got.view(torch.int16)[~nan], want.view(torch.int16)[~nan])
