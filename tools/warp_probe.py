"""Developer aid: warp forward / backward and range map through the C ABI for different KINDS of flow field (zero, gently
varying, the bench's up-sampled noise, per-pixel white noise), at a shape larger than the L2 so that the fixed launch cost
does not hide the streaming rate.  usage: python tools/warp_probe.py [--shape B,C,H,W] [--kinds zero,gentle,bench,noise]"""
import argparse
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from ocflow_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="8,32,188,620")
ap.add_argument("--kinds", default="zero,gentle,bench,noise")
ap.add_argument("--ops", default="warp_fwd,warp_bwd,warp_bwd_flow_only,range_map")
a = ap.parse_args()
SHAPES = [tuple(int(v) for v in sh.split(",")) for sh in a.shape.split(";")]   # several shapes: "8,32,96,128;8,64,48,64"
B, C, H, W = SHAPES[0]
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6548.8
g = torch.Generator(device="cuda").manual_seed(3)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(1 << 28, device="cuda")


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def flow_of(kind):
    if kind == "zero":
        return torch.zeros(B, 2, H, W, device="cuda")
    if kind == "noise":
        return torch.randn(B, 2, H, W, device="cuda", generator=g) * 2
    coarse = torch.randn(B, 2, max(H // 4, 2), max(W // 4, 2), device="cuda", generator=g) * 2.0
    if kind == "gentle":   # ~0.05 px / px: a field like the decoders' on real image pairs (piecewise smooth motion)
        coarse = torch.randn(B, 2, max(H // 32, 2), max(W // 32, 2), device="cuda", generator=g) * 3.0
    return F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True).contiguous()


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return statistics.median(ts)


for B, C, H, W in SHAPES:
    n = B * H * W
    img = torch.randn(B, C, H, W, device="cuda", generator=g)
    gout = torch.randn(B, C, H, W, device="cuda", generator=g)
    out = torch.empty_like(img)
    dimg = torch.empty_like(img)
    rm = torch.empty(B, 1, H, W, device="cuda")
    print("shape", (B, C, H, W), "peak", PEAK)
    for kind in a.kinds.split(","):
        fl = flow_of(kind)
        dfl = torch.empty_like(fl)
        cases = {
            "warp_fwd": (lambda: _lib.call("ocf_warp_fwd", P(img), P(fl), None, P(out), B, C, H, W, 0, 1.0, st), 4 * n * (2 * C + 2)),
            "warp_bwd": (lambda: _lib.call("ocf_warp_bwd", P(gout), P(img), P(fl), None, P(dimg), P(dfl), None, B, C, H, W, 0, 1.0, st), 4 * n * (3 * C + 4)),
            "warp_bwd_flow_only": (lambda: _lib.call("ocf_warp_bwd", P(gout), P(img), P(fl), None, None, P(dfl), None, B, C, H, W, 0, 1.0, st), 4 * n * (2 * C + 4)),
            "range_map": (lambda: _lib.call("ocf_range_map", P(fl), P(rm), None, B, H, W, st), 4 * n * 4),
        }
        for name in a.ops.split(","):
            fn, nbytes = cases[name]
            t = timed(fn)
            print("%-8s %-20s %9.1f us  %8.1f GB/s  %.3f of peak" % (kind, name, t * 1e6, nbytes / t / 1e9, nbytes / t / 1e9 / PEAK))
