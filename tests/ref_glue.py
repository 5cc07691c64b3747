"""TEST INFRASTRUCTURE: torch restatement of the reference's Python glue around quant_cuda.quant,
operation for operation, so that the UNMODIFIED reference extension (oracle/_ref/ref_quant_cuda*.so)
can be driven on the GPU box, where /root/reference does not exist.  Follows
models_fp_quant_transform_rotate/quant_utils.py:313-330 (symmetric), :415-452 / :577-611
(sign-split).  `quant` is the extension's entry point (the reference's, or this repo's drop-in)."""
import glob
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ref_ext():
    """The unmodified reference extension built by oracle/build_ref.sh, or None."""
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_quant_cuda*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("ref_quant_cuda", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sym_group_cuda(quant, x, grid, group_size=128, out_dtype=None):
    quant_grid = grid.to(x.device)
    x_shape = x.shape
    x = x.reshape(-1, group_size)
    x_shape_1 = x.shape
    scale = x.abs().max(dim=-1, keepdim=True)[0] / quant_grid.abs().max()
    x = x / scale
    quant_array = x.view(-1).to(torch.float32)
    quant_grid = quant_grid.type_as(quant_array)
    quant_array, _ = quant(quant_array, quant_grid)
    quant_array = quant_array.view(x_shape_1)
    output = quant_array * scale
    return output.view(x_shape).to(x.dtype if out_dtype is None else out_dtype)


def signsplit_group_cuda(quant, x, grid_neg, grid_pos, group_size=128, clipping_strength=1.0):
    grid_neg = grid_neg.to(x.device)
    grid_pos = grid_pos.to(x.device)
    if clipping_strength is not None:
        clip_value = clipping_strength * x.abs().max()
        x = torch.clamp(x, -clip_value, clip_value)
    x_shape = x.shape
    x = x.reshape(-1, group_size)
    x_shape_1 = x.shape
    x_neg = torch.where(x <= 0, x, torch.zeros_like(x))
    x_pos = torch.where(x > 0, x, torch.zeros_like(x))
    scale_neg = x_neg.abs().max(dim=-1, keepdim=True)[0] / grid_neg.abs().max()
    scale_pos = x_pos.abs().max(dim=-1, keepdim=True)[0] / grid_pos.abs().max()
    x_neg_normalized = (x_neg / scale_neg).view(-1).to(torch.float32)
    x_pos_normalized = (x_pos / scale_pos).view(-1).to(torch.float32)
    quantized_neg, _ = quant(x_neg_normalized, grid_neg)
    quantized_pos, _ = quant(x_pos_normalized, grid_pos)
    quantized_neg = quantized_neg.view(x_shape_1)
    quantized_pos = quantized_pos.view(x_shape_1)
    output = quantized_neg * scale_neg + quantized_pos * scale_pos
    return output.view(x_shape).to(x.dtype)
