import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the in-tree CUDA library: skip (not fail) them elsewhere, so a plain
    `pytest tests` is green on a CPU host.  On a GPU box a missing library is a hard failure, not a skip."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200) — run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def unit_ball(rng, b, n):
    """Uniform points in the unit ball (SURVEY.md §8d synthetic inputs), float32 (b,n,3)."""
    out = np.empty((b, n, 3), np.float32)
    for i in range(b):
        got = 0
        while got < n:
            c = rng.uniform(-1.0, 1.0, (2 * n + 16, 3)).astype(np.float32)
            c = c[(c.astype(np.float64) ** 2).sum(1) < 1.0]
            k = min(len(c), n - got)
            out[i, got:got + k] = c[:k]
            got += k
    return out


def lattice(rng, b, n, q=8):
    """Points on a coarse integer lattice: many exactly equal distances (tie-break stress)."""
    return (rng.integers(0, q, (b, n, 3)).astype(np.float32) / np.float32(q)).astype(np.float32)


@pytest.fixture
def rng():
    return np.random.default_rng(1234)
