"""rotate_utils/block_rotation_utils.py is a copy of rotation_utils.py with a `__main__` orthogonality check
(lines 75-110); the same names are re-exported here, and the check lives in
tests/test_gpu_rotate_score.py::test_rotation_is_orthogonal_round_trip."""
from .rotation_utils import *  # noqa: F401,F403
from .rotation_utils import block_random_hadamard_matrix, random_hadamard_matrix, rotate_fc1, rotate_mat_qkv, rotate_model  # noqa: F401


def block_diag(blocks):
    """block_rotation_utils.py:54-72: equally sized square blocks on the diagonal, zeros elsewhere (device / dtype of the first)."""
    import torch
    return torch.block_diag(*[b.to(device=blocks[0].device, dtype=blocks[0].dtype) for b in blocks])
