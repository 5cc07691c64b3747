"""Generate tests/golden/*.npz by importing the REFERENCE's own Python code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's quantizer functions run unmodified on CPU tensors.  Two modules the reference
imports but does not ship are shimmed (``dist``, top-level ``quant_utils``; SURVEY.md appendix
B), and ``quant_cuda`` -- a CUDA-only extension -- is replaced by a CPU stand-in that calls
the oracle's plain-C restatement of its scan (oracle/scan_quant.c).  That stand-in is itself
checked against the real compiled reference extension on the GPU box
(tests/test_gpu_ref_ext.py), so the chain reference -> golden is closed.

Every fixture stores the inputs next to the reference outputs so tests never need the
reference at run time.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FPQ_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def _install_shims():
    d = types.ModuleType("dist")
    d.get_device = lambda: "cpu"
    d.initialized = lambda: False
    sys.modules["dist"] = d
    sys.modules["quant_utils"] = types.ModuleType("quant_utils")

    qc = types.ModuleType("quant_cuda")

    def quant(x, y):
        assert x.dtype == torch.float32 and x.dim() == 1
        z = O.scan_quant(x.detach().numpy(), y.detach().to(torch.float32).numpy())
        return torch.from_numpy(z), torch.zeros_like(x)

    qc.quant = quant
    sys.modules["quant_cuda"] = qc
    sys.path.insert(0, REF)


def adversarial_groups(rng: np.random.Generator) -> np.ndarray:
    """[G,128] fp32 rows that hit every special case of the path."""
    rows = []
    for scale in (1.0, 0.03, 37.0, 2.0 ** -20, 1e-30, 3e4):
        rows.append(rng.standard_normal(128).astype(np.float32) * np.float32(scale))
    # exact midpoints and grid points of every format relative to a group max of 6 / 1.75 / 16 / 7.5 / 28
    for gmax, name in ((6.0, "e2m1"), (1.75, "e1m2"), (16.0, "e3m0"), (7.5, "e2m3"), (28.0, "e3m2")):
        g = np.unique(np.abs(O.GRIDS[name]))
        mids = (g[1:] + g[:-1]) / 2
        vals = np.concatenate([g, mids, -g, -mids]).astype(np.float32)
        for s in (1.0, 0.37, 2.0 ** -9):
            row = np.zeros(128, np.float32)
            v = (vals * np.float32(s))[:127]
            row[:v.size] = v
            row[127] = np.float32(gmax * s)          # pins absmax so the scale is exactly s
            rows.append(row)
            rows.append(-row)
    rows.append(np.zeros(128, np.float32))                         # all-zero group
    z = np.zeros(128, np.float32); z[5] = -0.0; z[9] = 1e-45       # signed zero + denormal
    rows.append(z)
    r = rng.standard_normal(128).astype(np.float32); r[3] = np.inf
    rows.append(r)
    r = rng.standard_normal(128).astype(np.float32); r[77] = -np.inf
    rows.append(r)
    r = rng.standard_normal(128).astype(np.float32); r[0] = np.nan
    rows.append(r)
    rows.append(np.abs(rng.standard_normal(128)).astype(np.float32))          # all positive
    rows.append(-np.abs(rng.standard_normal(128)).astype(np.float32))         # all negative
    gelu_like = rng.standard_normal(128).astype(np.float32)
    gelu_like = np.where(gelu_like > 0, gelu_like * 3, gelu_like * 0.05).astype(np.float32)
    rows.append(gelu_like)
    tiny16 = rng.standard_normal(128).astype(np.float32) * np.float32(2e-7)   # fp16 scale underflow
    rows.append(tiny16)
    rows.append(rng.standard_normal(128).astype(np.float32) * np.float32(5e-5))  # fp16 subnormal scale
    return np.stack(rows)


def main():
    _install_shims()
    import importlib
    qu = importlib.import_module("models_fp_quant_transform_rotate.quant_utils")
    qu0 = importlib.import_module("models_fp_quant.quant_utils")
    from rotate_utils import rotation_utils

    rng = np.random.default_rng(20261018)
    adv = adversarial_groups(rng)
    finite = adv[np.isfinite(adv).all(axis=1)]
    rnd = rng.standard_normal((48, 128)).astype(np.float32)
    rows_tok = rng.standard_normal((6, 5, 192)).astype(np.float32) * 2.5        # per_token rows of 192
    kv = rng.standard_normal((2, 7, 3, 64)).astype(np.float32)                  # KV cache rows of 64
    inputs = {"adv": adv, "finite": finite, "rnd": rnd, "rows_tok": rows_tok, "kv": kv}

    out = {}
    for k, v in inputs.items():
        out[f"in/{k}"] = v

    def run(tag, fn, x_np, dtype, *args, **kw):
        x = torch.from_numpy(x_np.copy()).to(dtype)
        with np.errstate(all="ignore"):
            y = fn(x.clone(), *args, **kw)
        out[f"out/{tag}"] = y.detach().to(torch.float32).numpy() if y.dtype != torch.float64 else y.numpy()
        out[f"dtype/{tag}"] = np.array(str(y.dtype))

    dts = {"f32": torch.float32, "f16": torch.float16}
    group_inputs = ("adv", "rnd")
    for dn, dt in dts.items():
        for iname in group_inputs:
            x = inputs[iname]
            for e in (1, 2, 3):
                run(f"fp_quant_e{e}_per_group_cuda/{iname}/{dn}", getattr(qu, f"fp_quant_e{e}_per_group_cuda"), x, dt, 4, 128)
                run(f"fp_quant_e{e}_per_group/{iname}/{dn}", getattr(qu, f"fp_quant_e{e}_per_group"), x, dt, 4, 128)
            for f in ("e2m3", "e3m2"):
                run(f"fp6_quant_{f}_per_group_cuda/{iname}/{dn}", getattr(qu, f"fp6_quant_{f}_per_group_cuda"), x, dt, 6, 128)
            run(f"fp6_quant_int_neg_e2m3_pos_per_group_cuda/{iname}/{dn}", qu.fp6_quant_int_neg_e2m3_pos_per_group_cuda, x, dt, 6, 128)
            run(f"fp_neg_reverse_quant_per_group_cuda/{iname}/{dn}", qu0.fp_neg_reverse_quant_per_group_cuda, x, dt, 4, 128)
        # the two functions with the whole-tensor clip: finite data and the NaN-poisoned tensor separately
        for iname in ("finite", "rnd", "adv"):
            x = inputs[iname]
            run(f"fp_quant_e1m2_neg_e2m1_pos_per_group_cuda/{iname}/{dn}", qu.fp_quant_e1m2_neg_e2m1_pos_per_group_cuda, x, dt, 4, 128)
            run(f"fp_quant_e1m2_neg_e2m1_pos_per_group/{iname}/{dn}", qu.fp_quant_e1m2_neg_e2m1_pos_per_group, x, dt, 4, 128)
            run(f"fp4_afpq_per_group_cuda/{iname}/{dn}", qu0.fp4_afpq_per_group_cuda, x, dt, 4, 128)
        for iname in ("rows_tok", "kv", "adv"):
            x = inputs[iname]
            for e in (1, 2, 3):
                run(f"fp_quant_e{e}_per_token/{iname}/{dn}", getattr(qu, f"fp_quant_e{e}_per_token"), x, dt, 4)
            for f in ("e2m3", "e3m2"):
                run(f"fp6_quant_{f}_per_token_cuda/{iname}/{dn}", getattr(qu, f"fp6_quant_{f}_per_token_cuda"), x, dt, 6)
            run(f"fp6_quant_int_neg_e2m3_pos_per_token_cuda/{iname}/{dn}", qu.fp6_quant_int_neg_e2m3_pos_per_token_cuda, x, dt, 6)

    # raw element rules
    probe = np.concatenate([adv.reshape(-1)[:4096], np.linspace(-30, 30, 4001, dtype=np.float32),
                            np.array([np.nan, np.inf, -np.inf, 102405.0, 102406.5, -102430.0, 1e9], np.float32)])
    out["in/probe"] = probe
    for gname in ("e2m1", "e1m2", "e3m0", "e2m3", "e3m2", "int_neg", "e2m3_pos", "e1m2_neg", "e2m1_pos"):
        ref_grid = {
            "e2m1": qu.fp4_e2m1_grid, "e1m2": qu.fp4_e1m2_grid, "e3m0": qu.fp4_e3m0_grid,
            "e2m3": qu.fp6_e2m3_grid, "e3m2": qu.fp6_e3m2_grid, "int_neg": qu.int_neg_grid,
            "e2m3_pos": qu.e2m3_pos_grid,
            "e1m2_neg": torch.tensor([-1.75, -1.5, -1.25, -1.0, -0.75, -0.5, -0.25, 0.0]),
            "e2m1_pos": torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]),
        }[gname]
        out[f"grid/{gname}"] = ref_grid.numpy().astype(np.float32)
        with np.errstate(all="ignore"):
            out[f"out/quantize_to_nearest_grid/{gname}"] = qu.quantize_to_nearest_grid(torch.from_numpy(probe), ref_grid).numpy()

    # rotation: the seed-42 128-block and a 256 block-diagonal matrix
    q256 = rotation_utils.block_random_hadamard_matrix(256, 128, "cpu", 42)
    out["rot/q256"] = q256.numpy()
    torch.manual_seed(42)
    out["rot/signs128"] = (torch.randint(low=0, high=2, size=(128,)).to(torch.float64) * 2 - 1).numpy()

    # QuantizedLinear / QuantizedLinear_fc2 end to end on a small nn.Linear (CPU, fp32)
    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 384)
    x = torch.randn(3, 5, 256)
    ql = qu.QuantizedLinear.from_float(lin, weight_quant="per_group", act_quant="per_group", w_bit=4, a_bit=4,
                                       act_quant_sym=True, activation_fp_quant=True, weight_fp_quant=True,
                                       act_fp_type="fp_e2", weight_fp_type="fp_e2")
    out["ql/w"] = lin.weight.detach().numpy(); out["ql/b"] = lin.bias.detach().numpy(); out["ql/x"] = x.numpy()
    out["ql/wq"] = ql.weight.detach().numpy()
    out["ql/y"] = ql(x).detach().numpy()
    out["ql/repr"] = np.array(repr(ql))
    ql2 = qu.QuantizedLinear_fc2.from_float(lin, weight_quant="per_group", act_quant="per_group", w_bit=4, a_bit=4,
                                            act_quant_sym=False, activation_fp_quant=True, weight_fp_quant=True,
                                            act_fp_type="fp_e1m2_neg_e2m1_pos", weight_fp_type="fp_e2")
    xg = torch.nn.functional.gelu(x, approximate="tanh")
    out["ql2/x"] = xg.numpy()
    out["ql2/y"] = ql2(xg).detach().numpy()
    out["ql2/repr"] = np.array(repr(ql2))
    ql3 = qu.QuantizedLinear.from_float(lin, weight_quant="per_channel", act_quant="per_token", w_bit=6, a_bit=6,
                                        act_quant_sym=True, activation_fp_quant=True, weight_fp_quant=True,
                                        act_fp_type="fp6_e2m3", weight_fp_type="fp6_e2m3")
    out["ql3/wq"] = ql3.weight.detach().to(torch.float32).numpy()
    out["ql3/wq_dtype"] = np.array(str(ql3.weight.dtype))

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
