import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (entries of LL of zero-mean noise sit near 0)."""
    import torch
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    denom = float(torch.linalg.norm(b))
    return float(torch.linalg.norm(a - b)) / (denom if denom > 0 else 1.0)
