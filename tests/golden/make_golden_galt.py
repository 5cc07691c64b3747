"""Generate tests/golden/reference_galt.npz by importing the REFERENCE's GALT trainer (second half).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_galt.py

`learnable_transformation/learnable_transformation_mat_qkv_fp4.py` is importable (its training loop sits under `__main__`).  Captured on
CPU with its own functions, unmodified: FPQuant.apply (:76-100, straight-through), compute_quant_error_v1 (:122-138) = loss of one
activation tensor, its gradient with respect to the smoothing vector, and three steps of the trainer's loop (:271-291: AdamW, lr 0.01,
one step per activation tensor).  The rotation matrix is an input of the fixture (the script draws it with the global torch RNG).
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
    sys.path.insert(0, MG.REF)                       # `from rotate_utils import rotation_utils`
    if "matplotlib" not in sys.modules:              # imported by the script, never used by the functions captured here; not in this image
        import types
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    g = _load(os.path.join(MG.REF, "learnable_transformation", "learnable_transformation_mat_qkv_fp4.py"), "ref_galt_qkv_fp4")

    rng = np.random.default_rng(20261020)
    C, O = 256, 96
    acts = [rng.standard_normal((2, pn * pn, C)).astype(np.float32) * np.exp(rng.uniform(-1, 1, C)).astype(np.float32) for pn in (2, 3, 5)]
    w = (rng.standard_normal((O, C)) * 0.05).astype(np.float32)
    s0 = np.exp(rng.uniform(-0.3, 0.3, C)).astype(np.float32)
    # a block-diagonal randomized Hadamard rotation like the evaluation path's (any orthogonal matrix does for the trainer)
    from rotate_utils import rotation_utils as RU
    Q = RU.block_random_hadamard_matrix(C, 128, "cpu", 42).to(torch.float32)
    out = {"w": w, "s0": s0, "Q": Q.numpy()}
    for i, a in enumerate(acts):
        out[f"act{i}"] = a

    # FPQuant forward / backward
    x = torch.from_numpy(acts[1]).clone().requires_grad_(True)
    y = g.FPQuant.apply(x)
    y.backward(torch.ones_like(y))
    out["fpquant_out"] = y.detach().numpy()
    out["fpquant_grad_is_identity"] = np.array(bool(torch.equal(x.grad, torch.ones_like(x))))

    # one loss + gradient
    s = torch.nn.Parameter(torch.from_numpy(s0).clone())
    loss = g.compute_quant_error_v1(torch.from_numpy(acts[0]), torch.from_numpy(w), s, Q)
    loss.backward()
    out["loss0"] = np.array(loss.item(), dtype=np.float64)
    out["grad0"] = s.grad.numpy().copy()

    # the trainer's loop, 3 epochs over the 3 activation tensors (learnable_transformation_mat_qkv_fp4.py:271-291)
    s = torch.nn.Parameter(torch.ones(C))
    opt = torch.optim.AdamW([s], lr=0.01)
    hist = []
    for _ in range(3):
        ep = 0.0
        for a in acts:
            loss = g.compute_quant_error_v1(torch.from_numpy(a), torch.from_numpy(w), s, Q)
            loss.backward()
            opt.step()
            opt.zero_grad()
            ep += loss.item()
        hist.append(ep / len(acts))
    out["loop_epoch_loss"] = np.asarray(hist, dtype=np.float64)
    out["loop_s_final"] = s.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "reference_galt.npz"), **out)
    print("wrote reference_galt.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})
    print("epoch losses", hist)


if __name__ == "__main__":
    main()
