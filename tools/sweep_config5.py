"""BASELINE.json config 5 (and the op-level part of config 4): correlation / warp microbenchmark at the native KITTI
375x1242 pyramid (ceil-halving: 188x621 ... 6x20), channel widths 16..196, batch 4..32, through the PUBLIC Python API
(ocflow_b200.compute_cost_volume / network_warp, forward and backward), CUDA-event timed with a 1 GiB L2 flush before
every launch.  Prints one line per (level, C, B): time, algorithmic GB/s (SURVEY.md section 8d byte counts) and the
fraction of the measured HBM peak.

    python tools/sweep_config5.py [--quick] > profiles/r1_config5_sweep.txt
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ocflow_b200 as ocf  # noqa: E402
from ocflow_b200 import ops  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, flush, reps=5):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    pk = peak()
    flush = torch.empty(256 * 1024 * 1024, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [("KITTI L1", 188, 621), ("KITTI L2", 94, 311), ("KITTI L3", 47, 156), ("KITTI L4", 24, 78), ("KITTI L5", 12, 39),
              ("KITTI L6", 6, 20), ("Sintel L2", 109, 256), ("Sintel L3", 55, 128)]
    chans = [16, 32, 64, 96, 128, 196]
    batches = [4, 8, 16, 32]
    if a.quick:
        chans, batches = [16, 64, 196], [8, 32]
    print("# measured HBM peak %.1f GB/s; times are medians of 5 cold-L2 launches (python wrapper + autograd included)" % pk)
    print("%-10s %4s %3s | %-26s | %-26s | %-26s | %-26s" % ("level", "C", "B", "corr fwd", "corr fwd+bwd", "warp fwd", "warp fwd+bwd"))
    for name, h, w in shapes:
        for C in chans:
            for B in batches:
                n = B * h * w
                if n * C * 4 > 3e9:
                    continue
                f1 = torch.randn(B, C, h, w, device="cuda", generator=g).requires_grad_(True)
                f2 = torch.randn(B, C, h, w, device="cuda", generator=g).requires_grad_(True)
                fl = (torch.randn(B, 2, h, w, device="cuda", generator=g) * 2).requires_grad_(True)
                cot = torch.randn(B, 81, h, w, device="cuda", generator=g)
                cotw = torch.randn(B, C, h, w, device="cuda", generator=g)

                def corr_f():
                    with torch.no_grad():
                        return ops.cost_volume(f1, f2, 4, 0.1)

                def corr_fb():
                    out = ops.cost_volume(f1, f2, 4, 0.1)
                    torch.autograd.grad(out, (f1, f2), cot)

                def warp_f():
                    with torch.no_grad():
                        return ocf.network_warp(f2, fl)

                def warp_fb():
                    out = ocf.network_warp(f2, fl)
                    torch.autograd.grad(out, (f2, fl), cotw)

                cells = []
                for fn, nbytes in ((corr_f, 4 * n * (2 * C + 81)), (corr_fb, 4 * n * (2 * C + 81) + 4 * n * (81 + 4 * C)),
                                   (warp_f, 4 * n * (2 * C + 2)), (warp_fb, 4 * n * (2 * C + 2) + 4 * n * (3 * C + 4))):
                    t = timed(fn, flush)
                    cells.append("%8.1f us %6.0f GB/s %.2f" % (t * 1e6, nbytes / t / 1e9, nbytes / t / 1e9 / pk))
                print("%-10s %4d %3d | %s" % (name, C, B, " | ".join(cells)), flush=True)
                del f1, f2, fl, cot, cotw


if __name__ == "__main__":
    main()
