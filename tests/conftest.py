import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    return np.load(path)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-exact comparison: distinguishes +0/-0, treats every NaN as equal to every NaN."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    ui = {2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
    an, bn = np.isnan(a), np.isnan(b)
    if not np.array_equal(an, bn):
        return False
    return bool(np.array_equal(a.view(ui)[~an], b.view(ui)[~bn]))


def mismatch_report(got: np.ndarray, want: np.ndarray, k: int = 5) -> str:
    got = np.asarray(got).reshape(-1)
    want = np.asarray(want).reshape(-1)
    ui = {2: np.uint16, 4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
    bad = np.nonzero((got.view(ui) != want.view(ui)) & ~(np.isnan(got) & np.isnan(want)))[0]
    lines = [f"{bad.size} / {got.size} mismatches"]
    for i in bad[:k]:
        lines.append(f"  [{i}] got {got[i]!r} want {want[i]!r}")
    return "\n".join(lines)
