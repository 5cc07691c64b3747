"""Drop-in mirror of the parts of rotate_utils/hadamard_utils.py the block rotation uses (power-of-two sizes).

`matmul_hadU` / `random_hadamard_matrix` are host-side torch code used to BUILD matrices offline, not hot-path
kernels; the K != 1 Kronecker tables (had12 ... had172, lines 164-4204) belong to
the full-width rotation, which is out of scope, and raise NotImplementedError here."""
from __future__ import annotations

import torch

from .rotation_utils import random_hadamard_matrix  # noqa: F401  (hadamard_utils.py:92-99)


def is_pow2(n):
    return (n & (n - 1) == 0) and (n > 0)


def get_hadK(n, transpose=False):
    """hadamard_utils.py:7-60 for the only case the block rotation needs: n a power of two -> (None, 1)."""
    if not is_pow2(n):
        raise NotImplementedError(f"get_hadK({n}): the Kronecker-factor tables of the full-width rotation are out of scope")
    return None, 1


def matmul_hadU(X, transpose=False):
    """X @ H_n / fl32(sqrt(n)) along the last dim, H_n the Sylvester matrix (what hadamard_utils.py:63-85 computes with
    in-place butterflies; H_n is symmetric, so `transpose` changes nothing).  Written as one fp64 product: for the
    +-1 diagonal inputs the rotation code feeds it, every partial sum is a small integer and the result is exact
    either way (the block matrix equals the reference's bit for bit, tests/test_host_logic.py)."""
    from .rotation_utils import _sylvester
    n = X.shape[-1]
    get_hadK(n, transpose)
    h = _sylvester(n).to(device=X.device, dtype=X.dtype)
    return (X.reshape(-1, n) @ h).view(X.shape) / torch.tensor(n).sqrt()


def matmul_hadUt(X):
    return matmul_hadU(X, transpose=True)


def apply_exact_had_to_linear(module, had_dim=-1, output=False):
    """hadamard_utils.py:119-154 (imported by rotation_utils.py:7, used only by the full-width rotation of fc2 / proj
    outputs, which the reference's own rotate_model leaves commented out): W <- W @ H / sqrt(n) on the input side, or
    H @ W on the output side, for power-of-two sizes, in float64 on the weight's device.  `had_dim != -1` needs the
    external fast_hadamard_transform package in the reference and is not provided."""
    assert isinstance(module, torch.nn.Linear)
    if had_dim != -1:
        raise NotImplementedError("apply_exact_had_to_linear(had_dim != -1): the chunked fast_hadamard_transform path is out of scope")
    w = module.weight.data
    w64 = w.double()
    w64 = matmul_hadU(w64.t(), False).t() if output else matmul_hadU(w64, False)
    module.weight.data = w64.to(device=w.device, dtype=w.dtype)


def matmul_hadU_cuda(X, hadK=None, K=1, transpose=False):
    """hadamard_utils.py:88-113 calls the external fast_hadamard_transform CUDA package for the power-of-two part and a
    dense product with the K-table otherwise.  Only K == 1 (power-of-two sizes) exists here, computed by `matmul_hadU`
    on X's device -- same value up to the summation order of the fp32 butterflies the external package uses."""
    if K != 1 or hadK is not None:
        raise NotImplementedError("matmul_hadU_cuda: the Kronecker-factor tables of the full-width rotation are out of scope")
    return matmul_hadU(X, transpose)


def matmul_hadUt_cuda(X, hadK=None, K=1):
    return matmul_hadU_cuda(X, hadK, K, transpose=True)
