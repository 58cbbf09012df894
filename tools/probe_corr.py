"""Developer probe: cost model of the correlation kernels.  Times ocf_corr_fwd / ocf_corr_bwd at the L2 geometry for
several channel counts, with / without the fused LeakyReLU mask and for one / both gradients, so that the per-tile
prologue (coefficient staging) can be separated from the per-channel main loop:  t(C) = a + b * C.
usage: python tools/probe_corr.py [--hw 96x128] [--batch 8]"""
import argparse
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ocflow_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hw", default="96x128")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--channels", default="8,16,32,64,128")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
h, w = (int(v) for v in a.hw.split("x"))
B = a.batch
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
    return statistics.median(ts)


print("geometry B=%d %dx%d" % (B, h, w))
print("%5s %10s %10s %10s %10s %10s %10s %10s" % ("C", "fwd+mask", "fwd", "bwd+act", "bwd+mask", "bwd", "bwd_df1", "bwd_df2"))
for C in (int(v) for v in a.channels.split(",")):
    f1 = torch.randn(B, C, h, w, device=dev)
    f2 = torch.randn(B, C, h, w, device=dev)
    out = torch.empty(B, 81, h, w, device=dev)
    g = torch.randn(B, 81, h, w, device=dev)
    msk = torch.zeros(B, 81, h, (w + 7) // 8, device=dev, dtype=torch.uint8)
    d1, d2 = torch.empty_like(f1), torch.empty_like(f2)
    t_f = timeit(lambda: _lib.call("ocf_corr_fwd", P(f1), P(f2), P(out), B, C, h, w, 4, 0, 0.1, None, P(msk), st))
    t_fn = timeit(lambda: _lib.call("ocf_corr_fwd", P(f1), P(f2), P(out), B, C, h, w, 4, 0, 0.1, None, None, st))
    t_ba = timeit(lambda: _lib.call("ocf_corr_bwd", P(g), P(out), P(f1), P(f2), P(d1), P(d2), B, C, h, w, 4, 0, 0, 0.1, None, st))
    t_bm = timeit(lambda: _lib.call("ocf_corr_bwd", P(g), None, P(f1), P(f2), P(d1), P(d2), B, C, h, w, 4, 0, 0, 0.1, P(msk), st))
    t_b = timeit(lambda: _lib.call("ocf_corr_bwd", P(g), None, P(f1), P(f2), P(d1), P(d2), B, C, h, w, 4, 0, 0, 1.0, None, st))
    t_b1 = timeit(lambda: _lib.call("ocf_corr_bwd", P(g), None, P(f1), P(f2), P(d1), None, B, C, h, w, 4, 0, 0, 1.0, None, st))
    t_b2 = timeit(lambda: _lib.call("ocf_corr_bwd", P(g), None, P(f1), P(f2), None, P(d2), B, C, h, w, 4, 0, 0, 1.0, None, st))
    print("%5d %10.2f %10.2f %10.2f %10.2f %10.2f %10.2f %10.2f" % (C, t_f, t_fn, t_ba, t_bm, t_b, t_b1, t_b2))
