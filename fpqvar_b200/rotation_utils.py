"""Drop-in mirror of rotate_utils/rotation_utils.py (+ the parts of hadamard_utils.py it uses) for the
128-block random-Hadamard rotation, on the fused kernels.

The reference materialises Q = I (x) diag(sigma) H_128 / fl32(sqrt(128)) as a dense fp64 C x C matrix
and multiplies by it (weights offline in fp64, activations online through a dense GEMM,
basic_var.py:263,266).  Here the weights go through `fpq_transform_rotate_weight` (fp64 butterflies,
one pass) and the activations through `fpq_transform_rotate_quant` (fp32 butterflies fused with the
GALT multiply and the following activation quantizer).  The dense matrix is still available for
callers that want it (`block_random_hadamard_matrix`), built on the host side once.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops
from ._lib import FpqError


def sign_vector(size: int, seed: int) -> torch.Tensor:
    """The +-1 diagonal of random_hadamard_matrix (hadamard_utils.py:92-97): torch.manual_seed(seed);
    torch.randint(0, 2, (size,)) * 2 - 1.  Like the reference this RESETS THE GLOBAL CPU RNG."""
    torch.manual_seed(seed)
    return torch.randint(low=0, high=2, size=(size,)).to(torch.float64) * 2 - 1


def _sylvester(n: int) -> torch.Tensor:
    if n & (n - 1):
        raise FpqError(f"Hadamard size {n} is not a power of two (the K != 1 tables of hadamard_utils.py are not needed "
                       "by the block rotation and are not implemented)")
    h = torch.ones(1, 1, dtype=torch.float64)
    while h.shape[0] < n:
        h = torch.cat([torch.cat([h, h], 1), torch.cat([h, -h], 1)], 0)
    return h


def random_hadamard_matrix(size, device, seed):
    """hadamard_utils.py:92-99: diag(sigma) @ H_size / fl32(sqrt(size)), fp64."""
    s = sign_vector(size, seed)
    q = (s[:, None] * _sylvester(size)) / float(torch.tensor(size).sqrt())       # torch.tensor(n).sqrt() is an fp32 0-dim tensor
    return q.to(device)


def block_random_hadamard_matrix(total_size=1920, block_size=128, device="cuda", seed=42, force_identity=False):
    """rotation_utils.py:69-104.  Every diagonal block is the SAME matrix: the inner call reseeds with the
    same seed for each block (hadamard_utils.py:95)."""
    assert total_size % block_size == 0, "size mismatch"
    n_blocks = total_size // block_size
    blk = torch.eye(block_size, dtype=torch.float64) if force_identity else random_hadamard_matrix(block_size, "cpu", seed)
    return torch.block_diag(*([blk] * n_blocks)).to(device)


def random_orthogonal_matrix(size, device):
    """rotation_utils.py:38-55: QR of a standard normal fp64 matrix with the signs of diag(R) folded in (draws from the
    global CPU RNG, like the reference)."""
    random_matrix = torch.randn(size, size, dtype=torch.float64).to(device)
    q, r = torch.linalg.qr(random_matrix)
    q *= torch.sign(torch.diag(r)).unsqueeze(0)
    return q


def get_orthogonal_matrix(size, mode="hadamard", device=None, seed=42):
    """rotation_utils.py:58-64.  mode="hadamard" is the full-width randomized Hadamard of `--rotate` without
    `--block_rotate`: available for power-of-two sizes; the reference's K = 60 / 36 factor tables for C = 1920 / 2304
    (hadamard_utils.py:28-39) are outside the hot path, so those sizes raise FpqError (use `--block_rotate`)."""
    if mode == "random":
        return random_orthogonal_matrix(size, device)
    if mode == "hadamard":
        return random_hadamard_matrix(size, device, seed)
    raise ValueError(f"Unknown mode {mode}")


def cleanup_memory() -> None:
    """rotation_utils.py:11-35 (called by evaluate_fp_quant_transform_rotate.py:105 after rotate_model): run the garbage
    collector and hand cached GPU memory back.  The fused weight kernels build no dense C x C temporaries, so there is
    little to free; the call is kept so the reference's scripts run unchanged."""
    import gc
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def block_sign_bits(block_size: int = 128, seed: int = 42):
    """The packed sign mask the kernels take (bit set = +1) for the reference's seed."""
    if block_size != 128:
        raise FpqError("the fused rotation kernels implement the reference's block size 128 only")
    return ops.pack_sign_bits(sign_vector(block_size, seed))


def rotate_weight(w: torch.Tensor, smooth: Optional[torch.Tensor] = None, sign_bits=None, inplace: bool = False) -> torch.Tensor:
    """(W / smooth).double() @ Q_block -> W.dtype (transform_model_utils.py:8-21 + rotation_utils.py:129-154)."""
    if sign_bits is None:
        sign_bits = block_sign_bits()
    if w.dtype != torch.float32:
        out = ops.transform_rotate_weight(w.to(torch.float32), smooth, sign_bits).to(w.dtype)
        if inplace:
            w.copy_(out)
            return w
        return out
    return ops.transform_rotate_weight(w, smooth, sign_bits, inplace=inplace)


def rotate_mat_qkv(layer, Q=None, sign_bits=None):
    """rotation_utils.py:129-144.  `Q` is accepted for signature compatibility and ignored: the kernel
    applies the same block matrix without materialising it."""
    w = layer.attn.mat_qkv.weight.data
    layer.attn.mat_qkv.weight.data = rotate_weight(w.contiguous(), None, sign_bits)


def rotate_fc1(layer, Q=None, sign_bits=None):
    """rotation_utils.py:147-154."""
    w = layer.ffn.fc1.weight.data
    layer.ffn.fc1.weight.data = rotate_weight(w.contiguous(), None, sign_bits)


def rotate_fc2(layer, Q):
    """rotation_utils.py:155-163 (defined by the reference, its call in rotate_model is commented out): W_fc2 @ Q in fp64.
    Q is a dense [4C, 4C] matrix here -- this path is offline, rarely used host code, not a hot-path kernel."""
    w = layer.ffn.fc2.weight.data
    layer.ffn.fc2.weight.data = torch.matmul(w.to(torch.float64), Q.to(w.device)).to(w.dtype)


def rotate_ada_lin(layer, Q):
    """rotation_utils.py:166-208 (also unused by the reference's rotate_model): the shift rows of the adaLN projection are
    rotated (Q^T W for the weight rows, b Q for the bias); gamma and scale rows stay as they are -- the reference computes
    rotated scale rows and then discards them (:191-192, :206-207), which is reproduced by leaving them untouched."""
    lin = layer.ada_lin[1]
    w, b = lin.weight.data, lin.bias.data
    C = w.shape[1]
    Q = Q.to(device=w.device, dtype=torch.float64)
    w64, b64 = w.to(torch.float64), b.to(torch.float64)
    w64 = torch.cat([w64[:4 * C], Q.T @ w64[4 * C:5 * C], Q.T @ w64[5 * C:6 * C]], dim=0)
    b64 = torch.cat([b64[:4 * C], b64[4 * C:5 * C] @ Q, b64[5 * C:6 * C] @ Q], dim=0)
    lin.weight.data, lin.bias.data = w64.to(w.dtype), b64.to(b.dtype)


def block_diag(blocks):
    """rotation_utils.py (helper of block_random_hadamard_matrix): equally sized square blocks on the diagonal."""
    return torch.block_diag(*[b.to(device=blocks[0].device, dtype=blocks[0].dtype) for b in blocks])


def rotate_model(model, device, block_rotate):
    """rotation_utils.py:211-240.  Only the block rotation of the README commands is implemented; the
    full-width randomized Hadamard (`block_rotate=False`, K = 60 / 36 Kronecker tables) is not on the
    hot path."""
    if not block_rotate:
        raise NotImplementedError("rotate_model(block_rotate=False): the full C x C randomized Hadamard of the reference "
                                  "(hadamard_utils.py K-tables) is outside the hot path; use --block_rotate")
    if model.C % 128:
        raise FpqError(f"model width {model.C} is not a multiple of the block size 128")
    bits = block_sign_bits(128, 42)
    for layer in model.blocks:
        rotate_mat_qkv(layer, sign_bits=bits)
        rotate_fc1(layer, sign_bits=bits)


def transform_rotate_quant_activation(x: torch.Tensor, smooth: Optional[torch.Tensor], act_fp_type: Optional[str] = "fp_e2",
                                      sign_bits=None) -> torch.Tensor:
    """The online site basic_var.py:263,266 + the act_quant of the QuantizedLinear that follows
    (qu.py:764-769) as ONE kernel:  fp16( (x * smooth) @ Q_block ) -> per-group-128 fake quant.
    x: fp32 [..., C] (adaLN-modulated LayerNorm output); returns fp16.  `act_fp_type=None` returns the
    rotated values unquantized."""
    fmt = {None: None, "fp_e1": "e1m2", "fp_e2": "e2m1", "fp_e3": "e3m0", "fp6_e2m3": "e2m3", "fp6_e3m2": "e3m2"}.get(act_fp_type, "?")
    if fmt == "?":
        raise ValueError("Unsupported fp_type.")
    if sign_bits is None:
        sign_bits = block_sign_bits()
    return ops.transform_rotate_quant(x, smooth, sign_bits, fmt)


def adaln_transform_rotate_quant_activation(ln_out: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, smooth: Optional[torch.Tensor],
                                            act_fp_type: Optional[str] = "fp_e2", sign_bits=None) -> torch.Tensor:
    """basic_var.py:263,266 in ONE kernel, adaLN modulate included:
        act_quant( matmul( ln_out.mul(scale.add(1)).add_(shift).mul(smooth), Q_block ) )
    ln_out: fp32 [B, L, C] (`self.ln_wo_grad(x)`); scale, shift: [B, 1, C] (scale1/shift1 or scale2/shift2), fp32, or fp16 as
    under the reference's fp16 autocast, where `scale.add(1)` is an fp16 add (reproduced bit for bit, see ops)."""
    fmt = {None: None, "fp_e1": "e1m2", "fp_e2": "e2m1", "fp_e3": "e3m0", "fp6_e2m3": "e2m3", "fp6_e3m2": "e3m2"}.get(act_fp_type, "?")
    if fmt == "?":
        raise ValueError("Unsupported fp_type.")
    if sign_bits is None:
        sign_bits = block_sign_bits()
    return ops.modulate_transform_rotate_quant(ln_out, scale, shift, smooth, sign_bits, fmt)


assert math  # keep the import the reference has (callers sometimes reach through the module)
