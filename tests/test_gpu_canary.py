"""Out-of-bounds guard for every wrapper-allocated buffer (compute-sanitizer is not available on the GPU pool, so the check is
our own): while a group of tests/sanitize_ops.py runs -- every hot-path entry point forward and backward at shapes with partial
tiles, ragged rows, far-out-of-frame flows -- each torch.empty / empty_like / zeros_like the wrappers of ocflow_b200.{ops,data,
metrics} make (outputs, gradients, masks, statistics and reduction workspaces) is carved out of a larger allocation whose 4 KB
margins on both sides carry a canary pattern.  A kernel that stores or reduces past either end of a buffer it was handed
changes a canary."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

MARGIN_BYTES = 4096     # keeps the 16-byte (TMA: 128-byte) alignment of the carved-out view
CANARY = {torch.float32: 1234.5, torch.float64: 1234.5, torch.uint8: 0xA5}


class _GuardedTorch:
    """Stands in for the `torch` name inside a wrapper module: allocation calls return guarded views, the rest is torch."""

    def __init__(self, registry):
        self._registry = registry

    def __getattr__(self, name):
        return getattr(torch, name)

    def _carve(self, shape, dtype, device, zero=False):
        dtype = dtype or torch.float32
        if dtype not in CANARY:
            t = torch.zeros(shape, dtype=dtype, device=device) if zero else torch.empty(shape, dtype=dtype, device=device)
            return t
        n = int(math.prod(shape))
        pad = MARGIN_BYTES // torch.empty((), dtype=dtype).element_size()
        big = torch.full((n + 2 * pad,), CANARY[dtype], dtype=dtype, device=device)
        if zero:
            big[pad:pad + n] = 0
        self._registry.append((big, pad, n, tuple(shape), dtype))
        return big[pad:pad + n].view(tuple(shape))

    def empty(self, *size, dtype=None, device=None, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        return self._carve(tuple(int(s) for s in shape), dtype, device)

    def empty_like(self, t, **kw):
        return self._carve(tuple(t.shape), t.dtype, t.device)

    def zeros_like(self, t, **kw):
        return self._carve(tuple(t.shape), t.dtype, t.device, zero=True)


@pytest.mark.parametrize("group", ["corr", "normalize", "warp", "scatter", "loss", "ssim", "census", "resize", "level", "data", "big"])
def test_kernels_stay_inside_the_buffers_they_are_handed(group, monkeypatch):
    import sanitize_ops as S
    from ocflow_b200 import data, metrics, ops

    assert group in S.GROUPS
    registry = []
    guarded = _GuardedTorch(registry)
    for mod in (ops, data, metrics):
        monkeypatch.setattr(mod, "torch", guarded)
    S.GROUPS[group]()
    torch.cuda.synchronize()
    assert len(registry) > 0, "no wrapper allocation went through the guard"
    bad = []
    for i, (big, pad, n, shape, dtype) in enumerate(registry):
        c = CANARY[dtype]
        lo = int((big[:pad] != c).sum())
        hi = int((big[pad + n:] != c).sum())
        if lo or hi:
            bad.append("allocation %d %s %s: %d elements changed below, %d above" % (i, shape, dtype, lo, hi))
    assert not bad, "\n".join(bad)
