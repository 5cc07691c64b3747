"""GALT smoothing-vector training with the fused quantizer (second half).

Mirror of the reference's trainers, learnable_transformation/learnable_transformation_{mat_qkv,fc1}_fp4.py: the loss of one calibration
tensor (`compute_quant_error_v1`, :122-138) is

    mean( (x W^T  -  Q_a((x * s) R)  Q_w((W / s) R)^T)^2 ),      R = the rotation matrix, Q = straight-through FP4 fake quant

and the loop (:271-291) takes one AdamW step (lr 0.01) per calibration tensor, `epochs` times over the set.  The only thing replaced is
the quantizer inside FPQuant: the reference's distance tensor + torch.argmin + gather (:76-89) is one fused launch here
(ops.fake_quant, argmin tie rule, bit-identical); the GEMMs and the optimiser are torch's, as in the reference.  No CPU path.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import ops


class FPQuant(torch.autograd.Function):
    """learnable_transformation_mat_qkv_fp4.py:76-100: e2m1 in groups of 128, torch.argmin rounding, identity gradient."""

    @staticmethod
    def forward(ctx, x, n_bits=4, group_size=128):
        assert n_bits == 4
        return ops.fake_quant(x.contiguous(), "e2m1", group_size, "argmin").to(x.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None


def compute_quant_error(x: torch.Tensor, w: torch.Tensor, learnable_s: torch.Tensor, Q: torch.Tensor) -> torch.Tensor:
    """learnable_transformation_mat_qkv_fp4.py:104-138 (`compute_quant_error` == `compute_quant_error_v1`)."""
    target = x @ w.T                                                   # the full-precision layer output
    act_q = FPQuant.apply((x * learnable_s) @ Q)                       # smoothed, rotated, fake-quantized activation
    wgt_q = FPQuant.apply((w / learnable_s) @ Q)                       # inverse-smoothed, rotated, fake-quantized weight
    return ((target - act_q @ wgt_q.T) ** 2).mean()


compute_quant_error_v1 = compute_quant_error


def train_smoothing(activations: Sequence[torch.Tensor], weight: torch.Tensor, Q: torch.Tensor, epochs: int = 50, lr: float = 0.01,
                    on_epoch: Optional[Callable[[int, float], None]] = None) -> Tuple[torch.Tensor, float, List[float]]:
    """The per-block loop of the reference trainer (:271-301): s starts at ones, AdamW(lr), one step per calibration tensor.

    Returns (s, best_loss, epoch_losses).  The reference keeps `best_s = learnable_s`, i.e. the PARAMETER itself, so what it saves
    is the vector after the last epoch whatever epoch was best; `s` here is that same final vector."""
    if not activations:
        raise ValueError("train_smoothing: no calibration tensors")
    s = torch.nn.Parameter(torch.ones(weight.shape[1], device=weight.device, dtype=weight.dtype))
    opt = torch.optim.AdamW([s], lr=lr)
    best, hist = float("inf"), []
    for epoch in range(epochs):
        total = 0.0
        for x in activations:
            loss = compute_quant_error(x, weight, s, Q)
            loss.backward()
            opt.step()
            opt.zero_grad()
            total += loss.item()
        avg = total / len(activations)
        best = min(best, avg)
        hist.append(avg)
        if on_epoch is not None:
            on_epoch(epoch, avg)
    return s.detach(), best, hist
