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


def _absmax_scale(t, grid):
    return t.abs().max(dim=-1, keepdim=True)[0] / grid.abs().max()


def _kernel_round(quant, t, grid):
    """`.view(-1).to(float32)` -> quant_cuda.quant -> back to the group shape (quant_utils.py:323-327)."""
    flat, _unused = quant(t.view(-1).to(torch.float32), grid.type_as(t.view(-1).to(torch.float32)))
    return flat.view(t.shape)


def sym_group_cuda(quant, x, grid, group_size=128, out_dtype=None):
    """fp_quant_e2_per_group_cuda & friends (quant_utils.py:313-330): scale = absmax/max|grid|; x/scale; kernel; *scale."""
    grid = grid.to(x.device)
    groups = x.reshape(-1, group_size)
    scale = _absmax_scale(groups, grid)
    q = _kernel_round(quant, groups / scale, grid)
    return (q * scale).view(x.shape).to(x.dtype if out_dtype is None else out_dtype)


def signsplit_group_cuda(quant, x, grid_neg, grid_pos, group_size=128, clipping_strength=1.0):
    """fp_quant_e1m2_neg_e2m1_pos_per_group_cuda (quant_utils.py:415-452) / fp6_quant_int_neg_e2m3_pos_per_group_cuda
    (:577-611, clipping_strength=None): whole-tensor clip, where()-split, one scale and one kernel call per side, sum."""
    grid_neg, grid_pos = grid_neg.to(x.device), grid_pos.to(x.device)
    if clipping_strength is not None:
        bound = clipping_strength * x.abs().max()
        x = torch.clamp(x, -bound, bound)
    groups = x.reshape(-1, group_size)
    zeros = torch.zeros_like(groups)
    neg, pos = torch.where(groups <= 0, groups, zeros), torch.where(groups > 0, groups, zeros)
    s_neg, s_pos = _absmax_scale(neg, grid_neg), _absmax_scale(pos, grid_pos)
    q_neg = _kernel_round(quant, neg / s_neg, grid_neg)
    q_pos = _kernel_round(quant, pos / s_pos, grid_pos)
    return (q_neg * s_neg + q_pos * s_pos).view(x.shape).to(x.dtype)
