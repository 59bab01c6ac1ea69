import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_adv():
    return dict(np.load(os.path.join(GOLDEN, "advantages.npz")))


@pytest.fixture(scope="session")
def golden_loss():
    return dict(np.load(os.path.join(GOLDEN, "losses.npz")))


@pytest.fixture(scope="session")
def golden_head():
    return dict(np.load(os.path.join(GOLDEN, "head.npz")))


def assert_close(actual, expected, rel=1e-5, floor=1.0, what=""):
    """|a - e| <= rel * max(|e|, floor), the tolerance semantics of SURVEY §8d."""
    import torch

    a = torch.as_tensor(np.asarray(actual)).double().cpu() if not hasattr(actual, "double") else actual.detach().double().cpu()
    e = torch.as_tensor(np.asarray(expected)).double().cpu() if not hasattr(expected, "double") else expected.detach().double().cpu()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    tol = rel * e.abs().clamp(min=floor)
    bad = (a - e).abs() > tol
    if bad.any():
        idx = bad.nonzero()[0].tolist()
        raise AssertionError(
            f"{what}: {int(bad.sum())} of {a.numel()} elements exceed rel={rel}; first at {idx}: "
            f"got {a[tuple(idx)].item()!r}, want {e[tuple(idx)].item()!r}; max abs err {float((a - e).abs().max())}"
        )
