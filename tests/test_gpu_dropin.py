"""GPU: the drop-in layer (reference names and signatures) against the reference-generated golden
vectors -- these read like the tests the reference would have: call the reference's function name on
the fixture input, compare with what the reference's own Python returned (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
from torch import nn

from conftest import bits_equal, mismatch_report
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f16": torch.float16}


@pytest.fixture(scope="module")
def Q():
    from fpqvar_b200 import quant_utils
    return quant_utils


def _run(golden, Q, name, iname, dn, *args):
    with np.errstate(over="ignore"):
        x = torch.from_numpy(golden[f"in/{iname}"].copy()).to(DT[dn]).cuda()
    y = getattr(Q, name)(x, *args)
    tag = f"{name}/{iname}/{dn}"
    assert str(y.dtype) == str(golden[f"dtype/{tag}"]), f"{tag}: dtype {y.dtype} vs reference {golden[f'dtype/{tag}']}"
    got = y.detach().to(torch.float32).cpu().numpy()
    assert bits_equal(got, golden[f"out/{tag}"]), f"{tag}\n" + mismatch_report(got, golden[f"out/{tag}"])


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["adv", "rnd"])
def test_group_functions(golden, Q, iname, dn):
    for e in (1, 2, 3):
        _run(golden, Q, f"fp_quant_e{e}_per_group_cuda", iname, dn, 4, 128)
        _run(golden, Q, f"fp_quant_e{e}_per_group", iname, dn, 4, 128)
    for f in ("e2m3", "e3m2"):
        _run(golden, Q, f"fp6_quant_{f}_per_group_cuda", iname, dn, 6, 128)
    _run(golden, Q, "fp6_quant_int_neg_e2m3_pos_per_group_cuda", iname, dn, 6, 128)
    _run(golden, Q, "fp_neg_reverse_quant_per_group_cuda", iname, dn, 4, 128)


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["finite", "rnd", "adv"])
def test_signsplit_functions(golden, Q, iname, dn):
    _run(golden, Q, "fp_quant_e1m2_neg_e2m1_pos_per_group_cuda", iname, dn, 4, 128)
    _run(golden, Q, "fp_quant_e1m2_neg_e2m1_pos_per_group", iname, dn, 4, 128)
    _run(golden, Q, "fp4_afpq_per_group_cuda", iname, dn, 4, 128)


@pytest.mark.parametrize("dn", ["f32", "f16"])
@pytest.mark.parametrize("iname", ["rows_tok", "kv", "adv"])
def test_per_token_functions(golden, Q, iname, dn):
    for e in (1, 2, 3):
        _run(golden, Q, f"fp_quant_e{e}_per_token", iname, dn, 4)
    for f in ("e2m3", "e3m2"):
        _run(golden, Q, f"fp6_quant_{f}_per_token_cuda", iname, dn, 6)
    _run(golden, Q, "fp6_quant_int_neg_e2m3_pos_per_token_cuda", iname, dn, 6)


def test_quantize_to_nearest_grid_and_quant_cuda(golden, Q):
    from fpqvar_b200.dropin import quant_cuda
    probe = torch.from_numpy(golden["in/probe"]).cuda()
    for gname, grid in (("e2m1", Q.fp4_e2m1_grid), ("e3m2", Q.fp6_e3m2_grid), ("int_neg", Q.int_neg_grid)):
        got = Q.quantize_to_nearest_grid(probe, grid.cuda()).cpu().numpy()
        assert bits_equal(got, golden[f"out/quantize_to_nearest_grid/{gname}"]), gname
        z, idx = quant_cuda.quant(probe, grid.cuda())
        assert bits_equal(z.cpu().numpy(), O.scan_quant(golden["in/probe"], grid.numpy())), gname
        assert idx.shape == probe.shape and idx.dtype == probe.dtype and float(idx.abs().max()) == 0.0


def test_e2_per_group_mutates_its_argument_like_the_reference(Q):
    torch.manual_seed(1)
    x = torch.randn(16, 256, device="cuda")
    keep = x.clone()
    out = Q.fp_quant_e2_per_group(x, 4, 128)
    g = keep.view(-1, 128)
    want_x = (g / (g.abs().max(dim=-1, keepdim=True)[0] / 6.0)).view_as(keep)
    assert torch.equal(x, want_x)                                    # qu.py:306 x.div_(scale)
    assert bits_equal(out.cpu().numpy(), O.fake_quant(keep.cpu().numpy(), "e2m1", 128, "argmin"))


def test_clipping_strength_other_than_one(Q):
    torch.manual_seed(2)
    x = torch.nn.functional.gelu(torch.randn(64, 512, device="cuda") * 2).half()
    got = Q.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda(x, 4, 128, clipping_strength=0.5).cpu().numpy()
    want = O.fake_quant_signsplit(x.cpu().numpy(), "e1m2_neg_e2m1_pos", 128, "kernel", clipping_strength=0.5)
    assert bits_equal(got, want), mismatch_report(got, want)


def test_quantized_linear_against_reference_fixture(golden, Q):
    lin = nn.Linear(256, 384).cuda()
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(golden["ql/w"]))
        lin.bias.copy_(torch.from_numpy(golden["ql/b"]))
    kw = dict(weight_quant="per_group", act_quant="per_group", w_bit=4, a_bit=4, activation_fp_quant=True, weight_fp_quant=True,
              weight_fp_type="fp_e2")
    ql = Q.QuantizedLinear.from_float(lin, act_quant_sym=True, act_fp_type="fp_e2", **kw)
    assert repr(ql) == str(golden["ql/repr"])
    assert bits_equal(ql.weight.cpu().numpy(), golden["ql/wq"])
    assert ql.bias is lin.bias
    x = torch.from_numpy(golden["ql/x"]).cuda()
    y = ql(x).cpu().numpy()
    # the quantized operands are bit-identical; the GEMM itself is a library call (fp32 accumulation order differs CPU vs GPU)
    assert np.allclose(y, golden["ql/y"], rtol=2e-5, atol=2e-5)
    ql2 = Q.QuantizedLinear_fc2.from_float(lin, act_quant_sym=False, act_fp_type="fp_e1m2_neg_e2m1_pos", **kw)
    assert repr(ql2) == str(golden["ql2/repr"])
    y2 = ql2(torch.from_numpy(golden["ql2/x"]).cuda()).cpu().numpy()
    assert np.allclose(y2, golden["ql2/y"], rtol=2e-5, atol=2e-5)
    ql3 = Q.QuantizedLinear.from_float(lin, weight_quant="per_channel", act_quant="per_token", w_bit=6, a_bit=6, act_quant_sym=True,
                                       activation_fp_quant=True, weight_fp_quant=True, act_fp_type="fp6_e2m3", weight_fp_type="fp6_e2m3")
    assert str(ql3.weight.dtype) == str(golden["ql3/wq_dtype"])
    assert bits_equal(ql3.weight.float().cpu().numpy(), golden["ql3/wq"])
    with pytest.raises(ValueError):
        Q.QuantizedLinear.from_float(lin, weight_quant="bogus", act_quant="per_group", activation_fp_quant=True, weight_fp_quant=True,
                                     act_fp_type="fp_e2", weight_fp_type="fp_e2")
    # .to() keeps the module usable in half precision (evaluate_fp_quant_transform_rotate.py:131 `var.half()`)
    qh = ql.to(torch.float16)
    assert qh.weight.dtype == torch.float16
    assert qh(x.half()).dtype == torch.float16


class FFN(nn.Module):                      # class NAMES are what quantize_VAR keys on (reference: basic_var.py FFN / SelfAttention)
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c, 4 * c)
        self.fc2 = nn.Linear(4 * c, c)

    def forward(self, x):
        return self.fc2(torch.nn.functional.gelu(self.fc1(x), approximate="tanh"))


class SelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.mat_qkv = nn.Linear(c, 3 * c, bias=False)
        self.proj = nn.Linear(c, c)


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attn = SelfAttention(c)
        self.ffn = FFN(c)


class TinyVAR(nn.Module):
    def __init__(self, c=256, depth=2):
        super().__init__()
        self.C = c
        self.blocks = nn.ModuleList([Block(c) for _ in range(depth)])
        self.head = nn.Linear(c, 16)


def test_quantize_var_rotate_transform_end_to_end(Q):
    from fpqvar_b200 import rotation_utils as R, transform_model_utils as T
    torch.manual_seed(3)
    model = TinyVAR().cuda()
    ref_w = {n: p.detach().clone() for n, p in model.named_parameters()}
    g = torch.Generator().manual_seed(4)
    s_qkv = [torch.exp(torch.rand(256, generator=g) * 2 - 1).cuda() for _ in model.blocks]
    s_fc1 = [torch.exp(torch.rand(256, generator=g) * 2 - 1).cuda() for _ in model.blocks]
    T.transform_model(model, s_qkv, s_fc1)
    R.rotate_model(model, "cuda", True)
    q64 = O.block_random_hadamard_matrix(256, 128)
    for i, blk in enumerate(model.blocks):
        for name, lin, s in ((f"blocks.{i}.attn.mat_qkv.weight", blk.attn.mat_qkv, s_qkv[i]), (f"blocks.{i}.ffn.fc1.weight", blk.ffn.fc1, s_fc1[i])):
            wt = O.transform_weight(ref_w[name].cpu().numpy(), s.cpu().numpy())
            want64 = wt.astype(np.float64) @ q64
            got = lin.weight.detach().cpu().numpy()
            tol = np.spacing(np.abs(want64.astype(np.float32))).astype(np.float64) + 128 * 2.3e-16 * np.abs(wt).max()
            assert np.all(np.abs(got.astype(np.float64) - want64) <= tol), name
    # fused transform+rotate == the two-step path, bit for bit
    model2 = TinyVAR().cuda()
    model2.load_state_dict({k: v for k, v in ref_w.items()})
    T.transform_rotate_model(model2, s_qkv, s_fc1)
    for a, b in zip(model.parameters(), model2.parameters()):
        assert torch.equal(a, b)
    Q.quantize_VAR(model, weight_quant="per_group", act_quant="per_group", w_bit=4, a_bit=4, act_quant_sym=True,
                   activation_fp_quant=True, weight_fp_quant=True, act_fp_type="fp_e2", weight_fp_type="fp_e2",
                   fc2_fp_type="fp_e1m2_neg_e2m1_pos")
    for blk in model.blocks:
        assert isinstance(blk.ffn.fc1, Q.QuantizedLinear) and isinstance(blk.ffn.fc2, Q.QuantizedLinear_fc2)
        assert isinstance(blk.attn.mat_qkv, Q.QuantizedLinear) and isinstance(blk.attn.proj, Q.QuantizedLinear)
        assert blk.ffn.fc2.act_quant.func is Q.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda
        w = blk.ffn.fc1.weight
        assert bits_equal(w.cpu().numpy(), O.fake_quant(model2.blocks[0].ffn.fc1.weight.detach().cpu().numpy() if blk is model.blocks[0]
                                                        else model2.blocks[1].ffn.fc1.weight.detach().cpu().numpy(), "e2m1", 128, "kernel"))
    assert isinstance(model.head, nn.Linear)                                  # head / word_embed / ada_lin stay FP (qu.py:1142-1164)
    # online site: fused transform+rotate+quant feeds the quantized mat_qkv weight
    x = torch.randn(5, 7, 256, device="cuda")
    xq = R.transform_rotate_quant_activation(x, s_qkv[0], "fp_e2")
    assert xq.dtype == torch.float16 and xq.shape == x.shape
    with pytest.raises(ValueError, match="Unsupported fp_type"):
        R.transform_rotate_quant_activation(x, None, "fp_bogus")
    with pytest.raises(NotImplementedError):
        R.rotate_model(model2, "cuda", False)


def test_search_layer_and_tensor_scores():
    from fpqvar_b200 import search
    torch.manual_seed(5)
    w = (torch.randn(384, 256, device="cuda") * 0.05)
    acts = [torch.randn(2, 9, 256, device="cuda") for _ in range(3)]
    loss = search.search_layer(w, acts)
    assert loss.shape == (3, 3)
    for wi, wf in enumerate(search.FP4_FORMATS):
        wq = search.fp4_quant(w, wf)
        for ai, af in enumerate(search.FP4_FORMATS):
            ref = sum(search.compute_quant_error(x @ w.T, search.fp4_quant(x, af) @ wq.T).double() for x in acts) / len(acts)
            assert abs(float(loss[wi, ai]) - float(ref)) <= 1e-9 + 1e-6 * float(ref)
    best = search.best_formats(loss, search.FP4_FORMATS, search.FP4_FORMATS)
    assert best["loss"] == float(loss.min()) and best["weight_format"] in search.FP4_FORMATS
    x = acts[0].reshape(-1, 128)
    mse = search.score_tensor_formats(x, search.FP4_FORMATS)
    for v, f in zip(mse.tolist(), search.FP4_FORMATS):
        ref = float(search.compute_quant_error(x.double(), search.fp4_quant(x, f).double()))
        assert abs(v - ref) <= 1e-6 * ref
    res = search.search_layers([{"name": "l0", "weight": w, "activations": acts}])
    assert res[0]["weight_format"] == best["weight_format"] and res[0]["activation_format"] == best["activation_format"]
    with pytest.raises(NotImplementedError):
        search.fp4_quant(w, "e5m9")


@pytest.mark.parametrize("dtype,fmts", [(torch.float32, "FP4"), (torch.float16, "FP4"), (torch.float16, "FP6")])
def test_search_layer_batched_matches_the_per_tensor_loop(dtype, fmts):
    """Row-stacking the calibration tensors leaves every row's quantized values unchanged (scales live on the last dim), so the
    batched table equals the reference-order loop up to summation order; the ragged [2, pn^2, C] shapes of the reference's
    calibration set exercise the segment means."""
    from fpqvar_b200 import search
    g = torch.Generator(device="cuda").manual_seed(11)
    formats = search.FP4_FORMATS if fmts == "FP4" else search.FP6_FORMATS
    per = "group" if fmts == "FP4" else "token"
    w = (torch.randn(384, 256, device="cuda", generator=g) * 0.05).to(dtype)
    acts = [torch.randn(2, pn * pn, 256, device="cuda", generator=g).to(dtype) for pn in (1, 2, 3, 4, 5, 6, 8, 10, 13, 16, 1, 3)]
    ref = search.search_layer(w, acts, formats, formats, per)
    for max_rows in (32768, 100):                                          # one slab / slabs that cut through tensors
        got = search.search_layer_batched(w, acts, formats, formats, per, max_rows=max_rows)
        assert torch.allclose(got, ref, rtol=2e-3 if dtype == torch.float16 else 1e-4, atol=0)
        assert search.best_formats(got, formats, formats)["weight_format"] == search.best_formats(ref, formats, formats)["weight_format"]
    # quantized rows are bit-identical whether a tensor is quantized alone or inside the stack
    X = torch.cat([a.reshape(-1, 256) for a in acts])
    q_stack = search.quantize(X, formats[0], per)
    q_each = torch.cat([search.quantize(a, formats[0], per).reshape(-1, 256) for a in acts])
    assert torch.equal(q_stack.view(torch.int16 if q_stack.dtype == torch.float16 else torch.int32),
                       q_each.view(torch.int16 if q_each.dtype == torch.float16 else torch.int32))


def test_fpquant_autograd_functions():
    """search_fp4_format.py:340-422 FPQuant / FPQuant_e1m2_neg_e2m1_pos: argmin rounding forward, straight-through backward."""
    from fpqvar_b200 import search
    torch.manual_seed(8)
    x = (torch.randn(32, 256, device="cuda") * 2).requires_grad_()
    for fmt in search.FP4_FORMATS:
        y = search.FPQuant.apply(x, 4, 128, fmt, 1.0)
        want = O.fake_quant(x.detach().cpu().numpy(), fmt, 128, "argmin")
        assert bits_equal(y.detach().cpu().numpy(), want), fmt
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    h = torch.nn.functional.gelu(torch.randn(32, 256, device="cuda")).requires_grad_()
    z = search.FPQuant_e1m2_neg_e2m1_pos.apply(h, 4, 128, 1.0)
    want = O.fake_quant_signsplit(h.detach().cpu().numpy(), "e1m2_neg_e2m1_pos", 128, "argmin")
    assert bits_equal(z.detach().cpu().numpy(), want)
    (z * 2).sum().backward()
    assert torch.equal(h.grad, torch.full_like(h, 2.0))
    with pytest.raises(ValueError):
        search.FPQuant.apply(x, 4, 128, "e5m2", 1.0)


def test_var_generation_harness_modes():
    """tools/var_generate.py: the three quantized modes run the hot path (launch counts: 4 quantizer calls per block
    per scale) and the fused mode agrees with the module-level mode up to the fp16-GEMM vs fp32-FWHT rotation."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("var_generate", os.path.join(root, "tools", "var_generate.py"))
    vg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vg)
    from fpqvar_b200 import ops
    dev = torch.device("cuda")
    outs = {}
    for mode in ("fp16", "modules", "fused"):
        with torch.device(dev):
            model = vg.Var(2, (1, 2, 3, 4), False).eval()
        model.init_weights(0)
        vg.prepare(model, mode, 4)
        n0 = ops.launch_count()
        f_hat = model.generate(3, torch.tensor([1, 2, 3], device=dev), torch.Generator(device=dev).manual_seed(0))
        assert f_hat.shape == (3, 32, 4, 4) and bool(torch.isfinite(f_hat).all())
        assert ops.launch_count() - n0 == (0 if mode == "fp16" else 4 * 2 * 4)
        outs[mode] = model
    # same quantized weights in both quantized modes (transform + rotate + quantize happen offline, identically)
    for a, b in zip(outs["modules"].blocks, outs["fused"].blocks):
        assert torch.equal(a.attn.mat_qkv.weight, b.attn.mat_qkv.weight) and torch.equal(a.ffn.fc2.weight, b.ffn.fc2.weight)
    # one block, same input: fused == modules up to the rotation's precision (fp16 GEMM vs fp32 butterflies)
    x = torch.randn(6, 9, 128, device=dev)
    cond = torch.randn(6, 128, device=dev).half()
    ys = []
    for mode in ("modules", "fused"):
        m = outs[mode]
        for blk in m.blocks:
            blk.attn.reset_cache(6, 30, dev)
        with torch.no_grad():
            ys.append(m._block(0, m.blocks[0], x, cond, None))
    assert torch.allclose(ys[0], ys[1], atol=2e-2, rtol=0)
