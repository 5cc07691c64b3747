"""Drop-in mirror of learnable_transformation/transform_model_utils.py: apply the GALT smoothing
factors to the weights (W[:, i] /= s[i], fp32, before the rotation).

Standalone it is a single IEEE division per element (the same ATen op the reference runs, bit for
bit); `transform_rotate_model` does transform + block rotation in one pass of the fused weight kernel,
which is what evaluate_fp_quant_transform_rotate.py:87-106 does in two."""
from __future__ import annotations

from . import rotation_utils


def transform_mat_qkv(layer, mat_qkv_best_s):
    w = layer.attn.mat_qkv.weight.data                                   # transform_model_utils.py:8-13
    layer.attn.mat_qkv.weight.data = (w / mat_qkv_best_s).to(w.dtype)


def transform_fc1(layer, fc1_best_s):
    w = layer.ffn.fc1.weight.data                                        # transform_model_utils.py:16-21
    layer.ffn.fc1.weight.data = (w / fc1_best_s).to(w.dtype)


def transform_model(model, mat_qkv_best_s, fc1_best_s):
    for idx, layer in enumerate(model.blocks):                           # transform_model_utils.py:24-28
        transform_mat_qkv(layer, mat_qkv_best_s[idx])
        transform_fc1(layer, fc1_best_s[idx])


def transform_rotate_model(model, mat_qkv_best_s, fc1_best_s, seed: int = 42):
    """transform_model followed by rotate_model(block_rotate=True), fused: one kernel per weight."""
    bits = rotation_utils.block_sign_bits(128, seed)
    for idx, layer in enumerate(model.blocks):
        for lin, s in ((layer.attn.mat_qkv, mat_qkv_best_s[idx]), (layer.ffn.fc1, fc1_best_s[idx])):
            w = lin.weight.data
            lin.weight.data = rotation_utils.rotate_weight(w.contiguous(), s.detach().to(w.device), bits)
