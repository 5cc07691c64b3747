"""Generate tests/golden/reference_search.npz by importing the REFERENCE's search scripts.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_search.py

`search/search_fp4_format.py` and `search/search_fp6_format.py` are importable (their sweeps sit under `__main__`).  Their
scorer functions run unmodified on CPU tensors with two shims: `quant_cuda` (the CPU stand-in of make_golden.py) and
`Tensor.cuda()` made a no-op (FPQuant.forward builds its grids with `.cuda()`, search_fp4_format.py:346-350).
Captured: FPQuant.forward (:340-363, three formats), FPQuant_e1m2_neg_e2m1_pos.forward (:378-411), fp4_quant (:544-553),
compute_quant_error (:472-476), fp6_quant (search_fp6_format.py:547-554), on random, adversarial and GELU-skewed inputs.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    MG._install_shims()
    torch.Tensor.cuda = lambda self, *a, **k: self                        # FPQuant.forward: grids are built with .cuda()
    s4 = _load(os.path.join(MG.REF, "search", "search_fp4_format.py"), "ref_search_fp4")
    s6 = _load(os.path.join(MG.REF, "search", "search_fp6_format.py"), "ref_search_fp6")

    rng = np.random.default_rng(20261019)
    adv = MG.adversarial_groups(rng)
    finite = adv[np.isfinite(adv).all(axis=1)]
    rnd = rng.standard_normal((40, 128)).astype(np.float32) * np.exp(rng.uniform(-3, 3, (40, 1))).astype(np.float32)
    g = torch.nn.functional.gelu(torch.from_numpy(rng.standard_normal((40, 128)).astype(np.float32) * 2), approximate="tanh").numpy()
    tok = rng.standard_normal((5, 7, 192)).astype(np.float32) * 2.0       # per-token rows of 192 (fp6_quant)
    inputs = {"finite": finite, "rnd": rnd, "gelu": g, "tok": tok}
    out = {f"in/{k}": v for k, v in inputs.items()}

    def put(tag, y):
        out[f"out/{tag}"] = y.detach().to(torch.float32).numpy()
        out[f"dtype/{tag}"] = np.array(str(y.dtype))

    for dn, dt in (("f32", torch.float32), ("f16", torch.float16)):
        for iname in ("finite", "rnd", "gelu"):
            x = torch.from_numpy(inputs[iname]).to(dt)
            with np.errstate(all="ignore"):
                for fmt in ("e1m2", "e2m1", "e3m0"):
                    put(f"FPQuant/{fmt}/{iname}/{dn}", s4.FPQuant.apply(x.clone(), 4, 128, fmt))
                    put(f"fp4_quant/{fmt}/{iname}/{dn}", s4.fp4_quant(x.clone(), fmt))
                put(f"FPQuant_e1m2_neg_e2m1_pos/{iname}/{dn}", s4.FPQuant_e1m2_neg_e2m1_pos.apply(x.clone(), 4, 128))
        xt = torch.from_numpy(tok).to(dt)
        for fmt in ("e2m3", "e3m2"):
            put(f"fp6_quant/{fmt}/tok/{dn}", s6.fp6_quant(xt.clone(), fmt))
    # clipping_strength != 1 (the GALT trainers pass it)
    x = torch.from_numpy(rnd)
    put("FPQuant/e2m1/rnd/f32/clip0.8", s4.FPQuant.apply(x.clone(), 4, 128, "e2m1", 0.8))
    put("FPQuant_e1m2_neg_e2m1_pos/rnd/f32/clip0.8", s4.FPQuant_e1m2_neg_e2m1_pos.apply(x.clone(), 4, 128, 0.8))
    # compute_quant_error: a tensor (not .item()), fp16 stays fp16
    a, b = torch.from_numpy(rnd), s4.fp4_quant(torch.from_numpy(rnd), "e2m1")
    put("compute_quant_error/f32", s4.compute_quant_error(a, b).reshape(1))
    put("compute_quant_error/f16", s4.compute_quant_error(a.half(), b.half()).reshape(1))
    # the search loop's per-layer loss (search_fp4_format.py:798-816) on a small layer
    w = torch.from_numpy(rng.standard_normal((256, 128)).astype(np.float32) * 0.05)
    acts = [torch.from_numpy(rng.standard_normal((2, n, 128)).astype(np.float32)) for n in (1, 4, 9)]
    out["loop/w"] = w.numpy()
    for i, t in enumerate(acts):
        out[f"loop/x{i}"] = t.numpy()
    table = np.zeros((3, 3))
    for wi, wf in enumerate(("e1m2", "e2m1", "e3m0")):
        wq = s4.fp4_quant(w, wf)
        for ai, af in enumerate(("e1m2", "e2m1", "e3m0")):
            loss = 0.0
            for t in acts:
                loss += s4.compute_quant_error(torch.matmul(t, w.T), torch.matmul(s4.fp4_quant(t, af), wq.T))
            table[wi, ai] = float(loss / len(acts))
    out["loop/loss"] = table

    path = os.path.join(HERE, "reference_search.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
