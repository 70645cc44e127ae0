import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor) -- the tolerance metric of BASELINE.json north_star (1e-4,
    fp32).  ``floor`` is an absolute scale for quantities that are mathematically zero (e.g. the
    gradient of a conv bias that feeds a batch-norm): they are compared against the scale of
    the other gradients instead of against their own rounding noise."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if not b.numel():
        return 0.0
    denom = max(b.abs().max().item(), floor, 1e-30)
    return (a - b).abs().max().item() / denom


def grad_floor(g, prefix="grad/", frac=1e-3):
    """frac x the largest reference gradient entry: the floor used for near-zero gradients.  The
    GPU tests use frac=0.1 (absolute tolerance 1e-5 of the largest gradient): a mathematically zero
    gradient is a cancelling sum whose rounding noise scales with its terms, not with its value."""
    return frac * max(float(np.abs(v).max()) for k, v in g.items() if k.startswith(prefix) and v.size)
