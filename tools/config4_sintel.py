"""BASELINE.json config 4: Sintel-shape (436x1024) forward-backward flow with occlusion mask and Charbonnier / census /
SSIM photometric terms, op level at the native shape (the network itself needs multiples of 64).

Per image pair: fw = randn*8 px, bw = -fw + randn*0.5; range map of bw -> occlusion mask ->
(a) the fused occlusion-weighted Charbonnier pass (warp + mask + loss + d/dflow in one kernel),
(b) warp(img2, fw) -> census(occ) and SSIM terms -> backward through the warp to the flow.
Every stage is timed with CUDA events (1 GiB L2 flush before each); the line per stage gives time, algorithmic GB/s and
the fraction of the measured HBM peak.  Under torchrun every rank runs an independent replica (no collective on the
data path); rank 0 prints the max-over-ranks time of the whole pipeline and the aggregate pairs/s.

    python tools/config4_sintel.py [--batch 8] > profiles/r1_config4_sintel.txt
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ocflow_b200 as ocf  # noqa: E402
from ocflow_b200 import ops  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def pipeline(img1, img2, fw, bw, census_w=1.0, ssim_w=1.0):
    """Returns (total, parts).  Gradient flows to fw only (images are data, the mask is computed under no_grad)."""
    with torch.no_grad():
        rmap, occ = ops.range_map(bw, with_occlusion=True)
    photo, photo_occ, _, _ = ops.occ_photo_fused(img1, img2, fw, rmap)
    warped = ops.warp(img2, fw, align_corners=True)
    census = ocf.census_loss(warped, img1, occ, 3)
    ssim = ocf.ssim_photometric_loss(warped, img1, 11)
    return photo + census_w * census + ssim_w * ssim, (photo, photo_occ, census, ssim)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=436)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, H, W = a.batch, a.height, a.width
    n = B * H * W
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    img1 = torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1
    img2 = torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1
    fw = (torch.randn(B, 2, H, W, device="cuda", generator=g) * 8).requires_grad_(True)
    bw = -fw.detach() + torch.randn(B, 2, H, W, device="cuda", generator=g) * 0.5
    flush = torch.empty(256 * 1024 * 1024, device="cuda")
    pk = peak()

    def timed(fn, reps=a.reps):
        for _ in range(3):
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

    with torch.no_grad():
        rmap, occ = ops.range_map(bw, with_occlusion=True)
        warped = ops.warp(img2, fw, align_corners=True)
    wreq = warped.clone().requires_grad_(True)
    cot = torch.randn_like(warped)
    Ho, Wo = H, W   # odd SSIM window: map has the image size
    stages = [
        ("range_map + mask", lambda: ops.range_map(bw, with_occlusion=True), 4 * n * (2 + 2 + 1)),
        ("fused occ Charbonnier fwd+dflow", lambda: ops.occ_photo_fused(img1, img2, fw.detach().requires_grad_(True), rmap), 4 * n * (3 + 3 + 2 + 1 + 2)),
        ("warp fwd (3 ch)", lambda: ops.warp(img2, fw.detach(), align_corners=True), 4 * n * (2 * 3 + 2)),
        ("warp bwd (d flow)", lambda: torch.autograd.grad(ops.warp(img2, fw, align_corners=True), fw, cot), 4 * n * (2 * 3 + 2) + 4 * n * (2 * 3 + 4)),
        ("census fwd (7x7)", lambda: ocf.census_loss(warped, img1, occ, 3), 4 * n * (3 + 3 + 1)),
        ("census fwd+bwd", lambda: torch.autograd.grad(ocf.census_loss(wreq, img1, occ, 3), wreq), 4 * n * (3 + 3 + 1) * 2 + 4 * n * 3),
        ("ssim fwd (11x11)", lambda: ocf.ssim(warped, img1, 11), 4 * n * 6),
        ("ssim fwd+bwd", lambda: torch.autograd.grad(ocf.ssim(wreq, img1, 11), wreq), 4 * n * 6 + 4 * n * 3 * 4 * 2 + 4 * n * (6 + 3)),
    ]
    lines = []
    for name, fn, nbytes in stages:
        t = timed(fn)
        lines.append("%-34s %9.1f us  %7.0f GB/s  %.3f of measured HBM peak" % (name, t * 1e6, nbytes / t / 1e9, nbytes / t / 1e9 / pk))

    def whole():
        total, _ = pipeline(img1, img2, fw, bw)
        (gfw,) = torch.autograd.grad(total, fw)
        return gfw

    t = timed(whole)
    tt = torch.tensor([t], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        total, parts = pipeline(img1, img2, fw, bw)
        print("# config 4: Sintel %dx%d, batch %d per GPU, %d GPU(s) as independent replicas; measured HBM peak %.1f GB/s" % (H, W, B, world, pk))
        print("# losses: photo %.6f photo_occ %.6f census %.6f ssim %.6f" % tuple(float(p) for p in parts))
        for ln in lines:
            print(ln)
        print("whole pipeline fwd+bwd (cold L2, python + autograd included): %.1f us per batch -> %.0f pairs/s on %d GPU(s)" % (
            float(tt) * 1e6, world * B / float(tt), world))
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
