"""Developer probe: A/B of the tiled correlation backward's staging knobs inside ONE process (OCF_KNOBS_DYNAMIC=1 makes the
library re-read them on every call).  For each geometry: median time of ocf_corr_bwd (sign bitmask, both gradients, L2
flushed before every launch) under every knob combination, and whether the gradients are bit-identical to the default's.
usage: python tools/probe_bwd_knobs.py [--reps 15]"""
import argparse
import ctypes
import itertools
import os
import statistics
import sys

os.environ["OCF_KNOBS_DYNAMIC"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ocflow_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=15)
ap.add_argument("--geoms", default="8x32x96x128,8x64x48x64,8x128x96x128,8x32x188x620")
ap.add_argument("--combos", default="", help="GDIRECT:PREFETCH pairs, e.g. 0:0,0:1,3:1 (default: the full grid)")
a = ap.parse_args()
dev = "cuda"
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 * 1024 * 1024, device=dev)


def P(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


KNOBS = [("OCF_BWD_GDIRECT", (0, 2, 1, 3)), ("OCF_BWD_PREFETCH", (0, 1, 148))]
for geom in a.geoms.split(","):
    B, C, h, w = (int(v) for v in geom.split("x"))
    g = torch.Generator(device=dev).manual_seed(7)
    f1 = torch.randn(B, C, h, w, device=dev, generator=g)
    f2 = torch.randn(B, C, h, w, device=dev, generator=g)
    gout = torch.randn(B, 81, h, w, device=dev, generator=g)
    out = torch.empty(B, 81, h, w, device=dev)
    msk = torch.zeros(B, 81, h, (w + 7) // 8, device=dev, dtype=torch.uint8)
    _lib.call("ocf_corr_fwd", P(f1), P(f2), P(out), B, C, h, w, 4, 0, 0.1, None, P(msk), st)   # a real sign mask
    ref = None
    print("geometry %s" % geom)
    combos = [tuple(int(v) for v in c.split(":")) for c in a.combos.split(",")] if a.combos else list(itertools.product(*[v for _, v in KNOBS]))
    for combo in combos:
        for (name, _), val in zip(KNOBS, combo):
            os.environ[name] = str(val)
        d1, d2 = torch.full_like(f1, float("nan")), torch.full_like(f2, float("nan"))
        fn = lambda: _lib.call("ocf_corr_bwd", P(gout), None, P(f1), P(f2), P(d1), P(d2), B, C, h, w, 4, 0, 0, 0.1, P(msk), st)  # noqa: E731
        med, best = timeit(fn)
        torch.cuda.synchronize()
        if ref is None:
            ref = (d1.clone(), d2.clone())
            same = "reference"
        else:
            same = "bit-identical" if torch.equal(d1, ref[0]) and torch.equal(d2, ref[1]) else \
                "DIFFERENT (max abs %.3e)" % max((d1 - ref[0]).abs().max().item(), (d2 - ref[1]).abs().max().item())
        print("  " + " ".join("%s=%-3d" % (n[4:], v) for (n, _), v in zip(KNOBS, combo)) + "  median %7.2f us  min %7.2f us  %s" % (med, best, same))
for name, _ in KNOBS:
    os.environ.pop(name, None)
