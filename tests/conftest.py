import glob
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_op_files():
    return sorted(glob.glob(os.path.join(GOLD, "ops_*.pt")))


def load_golden(path):
    return torch.load(path, map_location="cpu", weights_only=False)


def rel_max(a, b):
    """max|a-b| / max|b|  -- the parity metric of SURVEY.md section 8c."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / den)


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def assert_close(a, b, tol, what=""):
    assert tuple(a.shape) == tuple(b.shape), "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    rm, rl = rel_max(a, b), rel_l2(a, b)
    assert rm <= tol and rl <= tol, "%s: rel_max %.3e rel_l2 %.3e > %.1e" % (what, rm, rl, tol)


def assert_scalar_close(a, b, tol, what=""):
    a, b = float(a), float(b)
    assert abs(a - b) <= tol * max(abs(b), 1e-30) + 1e-12, "%s: %.9g vs %.9g" % (what, a, b)
