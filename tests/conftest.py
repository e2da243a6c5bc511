import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# Tolerances stated by BASELINE.json:north_star, norm-wise per tensor:
#   |x - ref| <= rtol * max|ref|   with rtol 1e-5 on forwards, 1e-4 on losses and gradients.
FWD_RTOL = 1e-5
GRAD_RTOL = 1e-4


def normwise_err(x, ref):
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30))


def assert_close(x, ref, rtol, what=""):
    e = normwise_err(x, ref)
    assert e <= rtol, f"{what}: normwise error {e:.3e} > {rtol:.1e}"
