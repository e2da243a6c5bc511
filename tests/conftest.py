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


_REPORT = []


def assert_close(x, ref, rtol, what=""):
    e = normwise_err(x, ref)
    _REPORT.append((os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], what, e, rtol))
    assert e <= rtol, f"{what}: normwise error {e:.3e} > {rtol:.1e}"


def pytest_sessionfinish(session, exitstatus):
    """Every norm-wise comparison of the session with its measured error and its tolerance -> gpurun_out/parity_report.json
    (only when that scratch directory exists and GAC_PARITY_REPORT is set): the margins behind the green ticks."""
    out = os.path.join(ROOT, "gpurun_out")
    if not _REPORT or not os.environ.get("GAC_PARITY_REPORT") or not os.path.isdir(out):
        return
    import json
    rows = [{"test": t, "what": w, "err": e, "rtol": r, "margin": (r / e if e > 0 else None)} for t, w, e, r in _REPORT]
    worst = sorted((r for r in rows if r["err"] > 0), key=lambda r: r["rtol"] / r["err"])[:40]
    with open(os.path.join(out, os.environ["GAC_PARITY_REPORT"]), "w") as f:
        json.dump({"comparisons": len(rows), "tightest": worst, "all": rows}, f, indent=1)
