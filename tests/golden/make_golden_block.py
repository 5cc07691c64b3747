"""Generate tests/golden/reference_block.npz: ONE transformer block of the REFERENCE, run by the reference's own code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_block.py

What runs, all of it the reference's own Python (CPU, fp32; `quant_cuda` is the CPU stand-in of make_golden.py):
  * `models_fp_quant_transform_rotate.basic_var.AdaLNSelfAttn` (basic_var.py:225-270) built with its constructor,
  * the offline pipeline of evaluate_fp_quant_transform_rotate.py:87-130 on it: `transform_model` ->
    `rotate_model(block_rotate=True)` -> `quantize_VAR(per_group, W4A4, fp_e2, fc2 fp_e1m2_neg_e2m1_pos)`,
  * one forward with `rotation_matrix = block_random_hadamard_matrix(C, 128, seed 42)` and the two smoothing vectors.
Captured: the original and the transformed+rotated+quantized weights, the LayerNorm outputs and adaLN tensors of both
call sites (basic_var.py:263,266), the tensors entering each quantized linear (x_1 / x_2 = after modulate, smoothing and
rotation), what their `act_quant` made of them, and the block output.  This pins the CALL SITE the fused kernel
replaces — operand order, which tensors are smoothed / rotated, which quantizer sits where — to the reference's module
code rather than to a reading of it.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG._install_shims()
    import importlib
    bv = importlib.import_module("models_fp_quant_transform_rotate.basic_var")
    qu = importlib.import_module("models_fp_quant_transform_rotate.quant_utils")
    from rotate_utils import rotation_utils
    from learnable_transformation import transform_model_utils

    C, H, B, L = 256, 4, 2, 5
    torch.manual_seed(7)
    block = bv.AdaLNSelfAttn(block_idx=0, last_drop_p=0, embed_dim=C, cond_dim=C, shared_aln=False,
                             norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=H, mlp_ratio=4.0,
                             attn_l2_norm=True, flash_if_available=False, fused_if_available=False).eval()
    with torch.no_grad():                                       # make the adaLN branch and the biases non-trivial
        block.ada_lin[1].weight.normal_(0, 0.05)
        block.ada_lin[1].bias.normal_(0, 0.3)
        block.attn.q_bias.normal_(0, 0.1)
        block.attn.v_bias.normal_(0, 0.1)

    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = torch.nn.ModuleList([block])
            self.C = C
            self.mlp_ratio = 4.0

    model = Wrapper()
    out = {}
    for name in ("attn.mat_qkv", "attn.proj", "ffn.fc1", "ffn.fc2"):
        lin = block.get_submodule(name)
        out[f"w0/{name}"] = lin.weight.detach().numpy().copy()
        if lin.bias is not None:
            out[f"b/{name}"] = lin.bias.detach().numpy().copy()
    out["ada/w"] = block.ada_lin[1].weight.detach().numpy().copy()
    out["ada/b"] = block.ada_lin[1].bias.detach().numpy().copy()
    out["attn/q_bias"] = block.attn.q_bias.detach().numpy().copy()
    out["attn/v_bias"] = block.attn.v_bias.detach().numpy().copy()
    out["attn/scale_mul_1H11"] = block.attn.scale_mul_1H11.detach().numpy().copy()

    s_qkv = torch.exp(torch.randn(C) * 0.4)
    s_fc1 = torch.exp(torch.randn(C) * 0.4)
    out["s/mat_qkv"], out["s/fc1"] = s_qkv.numpy(), s_fc1.numpy()

    # ---- offline: evaluate_fp_quant_transform_rotate.py:87-130 --------------------------------------
    transform_model_utils.transform_model(model, [s_qkv], [s_fc1])
    rotation_utils.rotate_model(model, "cpu", True)
    for name in ("attn.mat_qkv", "ffn.fc1"):
        out[f"w_rot/{name}"] = block.get_submodule(name).weight.detach().numpy().copy()
    qu.quantize_VAR(model, weight_quant="per_group", act_quant="per_group", quantize_bmm_input=False, w_bit=4, a_bit=4,
                    act_quant_sym=True, fc2_act_log2_quant=False, quant_kv=False, kv_bit=8, activation_fp_quant=True,
                    weight_fp_quant=True, act_fp_type="fp_e2", weight_fp_type="fp_e2", fc2_fp_type="fp_e1m2_neg_e2m1_pos")
    for name in ("attn.mat_qkv", "attn.proj", "ffn.fc1", "ffn.fc2"):
        q = block.get_submodule(name)
        out[f"wq/{name}"] = q.weight.detach().to(torch.float32).numpy().copy()
        out[f"class/{name}"] = np.array(type(q).__name__)

    # ---- online: one forward, everything at the call sites captured ----------------------------------
    x = torch.randn(B, L, C) * 1.5
    cond = torch.randn(B, C)
    Q = rotation_utils.block_random_hadamard_matrix(total_size=C, block_size=128, device="cpu", seed=42).to(torch.float32)
    out["x"], out["cond"], out["Q"] = x.numpy(), cond.numpy(), Q.numpy()

    ln_outs = []
    block.ln_wo_grad.register_forward_hook(lambda m, i, o: ln_outs.append(o.detach().clone()))
    ada = []
    block.ada_lin.register_forward_hook(lambda m, i, o: ada.append(o.detach().clone()))
    for name in ("attn.mat_qkv", "attn.proj", "ffn.fc1", "ffn.fc2"):
        q = block.get_submodule(name)
        orig = q.act_quant

        def spy(t, _orig=orig, _name=name):
            out[f"act_in/{_name}"] = t.detach().to(torch.float32).numpy().copy()
            r = _orig(t.clone())
            out[f"act_q/{_name}"] = r.detach().to(torch.float32).numpy().copy()
            return r

        q.act_quant = spy
        q.register_forward_hook(lambda m, i, o, _name=name: out.__setitem__(f"lin_out/{_name}", o.detach().numpy().copy()))
    block.attn.kv_caching(False)
    with torch.no_grad():
        y = block(x=x, cond_BD=cond, attn_bias=None, step_idx=0, quant_KV=False, kv_bit=None, rotation_matrix=Q,
                  mat_qkv_best_s=s_qkv, fc1_best_s=s_fc1)
    out["y"] = y.numpy()
    assert len(ln_outs) == 2 and len(ada) == 1
    out["ln/1"], out["ln/2"] = ln_outs[0].numpy(), ln_outs[1].numpy()
    g1, g2, sc1, sc2, sh1, sh2 = ada[0].view(-1, 1, 6, C).unbind(2)
    for k, v in (("gamma1", g1), ("gamma2", g2), ("scale1", sc1), ("scale2", sc2), ("shift1", sh1), ("shift2", sh2)):
        out[f"ada/{k}"] = v.numpy().copy()

    path = os.path.join(HERE, "reference_block.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
