"""Drop-in mirror of the parts of rotate_utils/hadamard_utils.py the block rotation uses (power-of-two sizes).

`matmul_hadU` is the reference's butterfly (hadamard_utils.py:63-85) -- host-side torch code used to BUILD
matrices offline, not a hot-path kernel; the K != 1 Kronecker tables (had12 ... had172, lines 164-4204) belong to
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
    """X @ H_n / fl32(sqrt(n)) along the last dim by repeated (a+b, a-b) stages (hadamard_utils.py:63-85)."""
    n = X.shape[-1]
    get_hadK(n, transpose)
    inp = X.clone().reshape(-1, n, 1)
    out = inp.clone()
    while inp.shape[1] > 1:
        inp = inp.view(inp.shape[0], inp.shape[1] // 2, 2, inp.shape[2])
        out = out.view(inp.shape)
        out[:, :, 0, :] = inp[:, :, 0, :] + inp[:, :, 1, :]
        out[:, :, 1, :] = inp[:, :, 0, :] - inp[:, :, 1, :]
        out = out.view(inp.shape[0], inp.shape[1], -1)
        inp, out = out, inp
    return inp.view(X.shape) / torch.tensor(n).sqrt()


def matmul_hadUt(X):
    return matmul_hadU(X, transpose=True)
