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


def rel_err(a, b, floor=0.0):
    """Norm-wise relative error ||a-b|| / max(||b||, floor*sqrt(numel)).

    Norm-wise because entries of e.g. LL of zero-mean noise sit near 0.  `floor` (an RMS magnitude) guards
    quantities that are zero in exact arithmetic -- e.g. the time-embedding gradient through a GroupNorm with
    one channel per group -- where the reference itself holds only round-off noise."""
    import torch
    a = a.detach().double().flatten().cpu()
    b = b.detach().double().flatten().cpu()
    denom = max(float(torch.linalg.norm(b)), floor * float(b.numel()) ** 0.5)
    return float(torch.linalg.norm(a - b)) / (denom if denom > 0 else 1.0)


@pytest.fixture()
def emulated_ops(monkeypatch):
    """Route `unet_design_b200` through the CPU emulation of its op contract (tests/_emulated_ops.py) so the
    Python side of the drop-in modules can be checked without a GPU.  Test-only; never used with -m gpu."""
    from _emulated_ops import EmulatedOps
    import unet_design_b200._lib as lib
    monkeypatch.setattr(lib, "_ops", EmulatedOps())
    return lib._ops
