"""Every hot-path entry point once, forward and backward, at small shapes that still reach each code path (persistent /
cluster-split / tensor-core / d = 10 / generic correlation, both warp conventions, scatter kernels, fused loss, SSIM, census,
resize, fused level, input packing, metrics), with a few results checked against the oracle.

Written for compute-sanitizer (`compute-sanitizer --tool memcheck python tests/sanitize_ops.py`); that tool is closed on this
GPU pool, so the groups are driven by tests/test_gpu_canary.py instead, which carves every wrapper-allocated buffer out of a
canary-filled allocation and checks that nothing was written past either end.  Stand-alone it also runs with
PYTORCH_NO_CUDA_MEMORY_CACHING=1 (every tensor its own cudaMalloc).  Not collected by pytest (no test_ prefix); test
infrastructure like tests/insitu.py."""
import argparse
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ocflow_b200 as ocf  # noqa: E402
from ocflow_b200 import _lib, data, metrics, ops  # noqa: E402
from oracle import ocflow_oracle as O  # noqa: E402

G = torch.Generator().manual_seed(7)


def rnd(*shape, scale=1.0):
    return torch.randn(*shape, generator=G) * scale


def rel(mine, want):
    want = want.detach()
    return float((mine.detach().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-12))


def run_bwd(out, *leaves):
    outs = out if isinstance(out, (list, tuple)) else [out]
    total = sum((o * torch.randn(o.shape, generator=G).cuda()).sum() for o in outs)
    return torch.autograd.grad(total, [l for l in leaves if l.requires_grad], allow_unused=True)


def leaf(t):
    return t.clone().cuda().requires_grad_(True)


def leaky(x, s=0.1):
    return torch.where(x > 0, x, x * s)


def group_corr():
    # (B, C, H, W, d, slope): persistent forward with TMA stores (>= 148 tiles), cluster-split (2 tiles, many channels), tensor-core
    # kernel (ragged narrow rows), re-pitched ragged rows, d = 10 tiled, generic d = 2
    for B, C, H, W, d, slope in ((5, 12, 64, 128, 4, 0.1), (1, 96, 16, 32, 4, 0.1), (2, 8, 7, 9, 4, 0.1), (2, 8, 12, 30, 4, 1.0),
                                 (1, 8, 12, 32, 10, 1.0), (2, 5, 9, 13, 2, 1.0)):
        f1, f2 = rnd(B, C, H, W), rnd(B, C, H, W)
        a, b = leaf(f1), leaf(f2)
        out = ops.cost_volume(a, b, d, slope)
        run_bwd(out, a, b)
        want = O.cost_volume(f1, f2, d)
        e = rel(out, leaky(want, slope) if slope != 1.0 else want)
        assert e < 1e-4, ("corr", B, C, H, W, d, e)
    layer = ocf.CostVolumeLayer(4)
    a, b = leaf(rnd(1, 16, 8, 32)), leaf(rnd(1, 16, 8, 32))
    run_bwd(layer(a, b), a, b)


def group_normalize():
    f1, f2 = rnd(2, 8, 16, 24) * 2 + 1, rnd(2, 8, 16, 24)
    for n in (True, False):
        for c in (True, False):
            for ch in (True, False):
                for im in (True, False):
                    a, b = leaf(f1), leaf(f2)
                    ys = ocf.normalize_features([a, b], n, c, ch, im)
                    run_bwd(list(ys), a, b)
                    want = O.normalize_features([f1, f2], n, c, ch, im)
                    assert rel(ys[0], want[0]) < 1e-4 and rel(ys[1], want[1]) < 1e-4
    x = leaf(rnd(1, 5, 3, 7))   # odd length: scalar kernels
    run_bwd(list(ocf.normalize_features([x])), x)


def group_warp():
    for B, C, H, W in ((2, 3, 32, 48), (2, 32, 24, 32), (1, 5, 7, 9)):
        img, flow = rnd(B, C, H, W), rnd(B, 2, H, W, scale=3.0)
        flow[:, :, 0, 0] = 100.0   # far outside the frame
        for ac in (True, False):
            a, f = leaf(img), leaf(flow)
            out = ops.warp(a, f, align_corners=ac)
            run_bwd(out, a, f)
            assert rel(out, O.warp(img, flow, ac)) < 1e-4
        a, f = leaf(img), leaf(flow)
        run_bwd(ops.warp(a, f, is_mask=True), a, f)
        occ = (torch.rand(B, 1, H, W, generator=G) < 0.5).float().cuda()
        a, f = leaf(img), leaf(flow)
        run_bwd(ops.warp(a, f, align_corners=False, occ=occ, flow_scale=1.25), a, f)
        ocf.backwarp(img.cuda(), flow.cuda())
        ocf.network_warp(img.cuda(), flow.cuda())


def group_scatter():
    for B, H, W in ((2, 32, 48), (1, 7, 9), (2, 64, 64)):
        flow = rnd(B, 2, H, W, scale=4.0)
        flow[:, :, -1, -1] = -1000.0
        assert rel(ocf.compute_range_map(flow.cuda()), O.range_map(flow)) < 1e-4
        ocf.occlusion_mask(flow.cuda())
        ocf.flow_to_warp(flow.permute(0, 2, 3, 1).contiguous().cuda())


def group_loss():
    for B, H, W in ((2, 32, 48), (1, 9, 13)):
        i1, i2 = rnd(B, 3, H, W), rnd(B, 3, H, W)
        flow = rnd(B, 2, H, W, scale=2.0)
        occ = torch.rand(B, 1, H, W, generator=G)
        a, b, o = leaf(i1), leaf(i2), leaf(occ)
        run_bwd(ocf.photometric_error(a, b, o), a, b, o)
        a, b = leaf(i1), leaf(i2)
        out = ocf.photometric_error(a, b)
        run_bwd(out, a, b)
        assert abs(float(out) - float(O.photometric_error(i1, i2))) < 1e-3 * abs(float(O.photometric_error(i1, i2)))
        x = leaf(i1)
        run_bwd(ocf.robust_l1(x), x)
        x = leaf(i1)
        run_bwd(ocf.charbonnier_loss(x), x)
        x = leaf(i1)
        run_bwd(list(ocf.gradient(x)), x)
        for fn, ofn in ((ocf.first_order_smoothness_loss, O.first_order_smoothness_loss),
                        (ocf.second_order_smoothness_loss, O.second_order_smoothness_loss)):
            smooth_img = i1 * 0.01      # gentle image gradients: exp(-100 |dI|) does not vanish
            a, f = leaf(smooth_img), leaf(flow)
            out = fn(a, f)
            run_bwd(out, a, f)
            want = float(ofn(smooth_img, flow))
            assert abs(float(out) - want) <= 1e-3 * abs(want) + 1e-12, (float(out), want)
        p, t = torch.rand(B, 1, H, W, generator=G) * 0.98 + 0.01, (torch.rand(B, 1, H, W, generator=G) < 0.3).float()
        for fn, x, y in ((ocf.flow_mse_loss, flow, rnd(B, 2, H, W)), (ocf.flow_l1_loss, flow, rnd(B, 2, H, W)),
                         (ocf.occlusion_bce_loss, p, t), (ocf.occlusion_focal_loss, p, t)):
            a = leaf(x)
            run_bwd(fn(a, y.cuda()), a)
        # the fused occlusion-aware loss chain: range map of the backward flow -> everything else in one kernel
        f = leaf(flow)
        rmap = ops.range_map(rnd(B, 2, H, W, scale=2.0).cuda())
        photo, photo_occ, mse, bce = ops.occ_photo_fused(i1.cuda(), i2.cuda(), f, rmap, flow_gt=rnd(B, 2, H, W).cuda(), occ_gt=t.cuda())
        run_bwd(photo, f)
        ops.occ_photo_fused(i1.cuda(), i2.cuda(), flow.cuda())      # no range map, no ground truth


def group_ssim():
    for B, C, H, W, ws in ((2, 3, 32, 48, 11), (1, 3, 20, 23, 4), (1, 1, 9, 9, 3)):
        i1, i2 = torch.rand(B, C, H, W, generator=G), torch.rand(B, C, H, W, generator=G)
        a, b = leaf(i1), leaf(i2)
        out = ocf.ssim(a, b, ws)
        run_bwd(out, a, b)
        assert abs(float(out) - float(O.ssim(i1, i2, ws))) < 1e-4
    a, b = leaf(torch.rand(2, 3, 16, 16, generator=G)), leaf(torch.rand(2, 3, 16, 16, generator=G))
    run_bwd(ocf.ssim(a, b, 11, size_average=False), a, b)
    a, b = leaf(torch.rand(1, 3, 24, 24, generator=G)), leaf(torch.rand(1, 3, 24, 24, generator=G))
    run_bwd(ocf.ssim_photometric_loss(a, b), a, b)


def group_census():
    for B, H, W, md in ((2, 32, 48, 3), (1, 9, 13, 2), (1, 17, 16, 1)):
        i1, i2 = torch.rand(B, 3, H, W, generator=G) * 2 - 1, torch.rand(B, 3, H, W, generator=G) * 2 - 1
        occ = torch.rand(B, 1, H, W, generator=G)
        a, b = leaf(i1), leaf(i2)
        out = ocf.census_loss(a, b, occ.cuda(), md)
        run_bwd(out, a, b)
        want = float(O.census_loss(i1, i2, occ, md))
        assert abs(float(out) - want) < 1e-3 * abs(want), (float(out), want)
        a, b = leaf(i1), leaf(i2)
        run_bwd(ocf.census_loss(a, b, None, md), a, b)


def group_resize():
    for shape, kw, mul in (((2, 2, 12, 16), dict(scale_factor=4), 20.0), ((1, 3, 32, 48), dict(scale_factor=0.25), 1.0),
                           ((1, 2, 5, 7), dict(size=(9, 11)), 1.0)):
        x0 = rnd(*shape)
        x = leaf(x0)
        out = ops.resize_bilinear(x, mul=mul, **kw)
        run_bwd(out, x)
        assert rel(out, O.resize_bilinear(x0, **kw) * mul) < 1e-4


def group_level():
    for B, C, H, W, with_up in ((2, 32, 24, 32, True), (2, 196, 6, 8, False), (1, 16, 12, 30, True)):
        c1, c2 = leaf(rnd(B, C, H, W)), leaf(rnd(B, C, H, W))
        if with_up:
            uf, ft = leaf(rnd(B, 2, H, W, scale=2.0)), leaf(rnd(B, 2, H, W))
            run_bwd(ops.level_fused(c1, c2, uf, ft, flow_scale=1.25), c1, c2, uf, ft)
        else:
            run_bwd(ops.level_fused(c1, c2), c1, c2)
    x, bias = leaf(rnd(2, 16, 8, 8)), leaf(rnd(16))
    y = ops.bias_leaky_relu_(x * 1.0, bias)
    run_bwd(y, x, bias)


def group_data():
    B, H, W = 2, 70, 130
    i1 = torch.randint(0, 256, (B, H, W, 3), generator=G, dtype=torch.uint8)
    i2 = torch.randint(0, 256, (B, H, W, 3), generator=G, dtype=torch.uint8)
    fl = rnd(B, H, W, 2)
    imgs, flow = data.pack_pairs(i1.cuda(), i2.cuda(), fl.cuda())
    want_i, want_f = O.pack_pairs(i1, i2, fl)
    assert rel(imgs, want_i) < 1e-6 and rel(flow, want_f) < 1e-6
    data.pack_occ(torch.randint(0, 256, (B, H, W), generator=G, dtype=torch.uint8).cuda())
    gt, pr = rnd(B, 2, 24, 32, scale=3.0).cuda(), rnd(B, 2, 24, 32, scale=3.0).cuda()
    metrics.batch_epe(pr, gt)
    metrics.evaluate_flow(gt[0].permute(1, 2, 0).contiguous(), pr[0].permute(1, 2, 0).contiguous())


def group_big():
    # BASELINE shapes (config-2 L2 level, a ragged KITTI level, the loss level): several generations of CTAs per SM, the
    # correlation backward's next-generation prefetch, the persistent forward's tile loop.  No oracle here (the parity and
    # property tests cover these shapes); this group exists for the out-of-bounds guard.
    for B, C, H, W in ((8, 32, 96, 128), (4, 16, 188, 621), (2, 128, 12, 16)):
        a, b = leaf(rnd(B, C, H, W)), leaf(rnd(B, C, H, W))
        run_bwd(ops.cost_volume(a, b, 4, 0.1), a, b)
        img, flow = leaf(rnd(B, C, H, W)), leaf(rnd(B, 2, H, W, scale=6.0))
        run_bwd(ops.warp(img, flow, align_corners=False), img, flow)
        c1, c2, uf, ft = leaf(rnd(B, C, H, W)), leaf(rnd(B, C, H, W)), leaf(rnd(B, 2, H, W, scale=2.0)), leaf(rnd(B, 2, H, W))
        run_bwd(ops.level_fused(c1, c2, uf, ft, flow_scale=1.25), c1, c2, uf, ft)
    B, H, W = 8, 384, 512
    i1, i2, flow = rnd(B, 3, H, W), rnd(B, 3, H, W), rnd(B, 2, H, W, scale=8.0)
    f = leaf(flow)
    rmap = ops.range_map(rnd(B, 2, H, W, scale=8.0).cuda())
    photo = ops.occ_photo_fused(i1.cuda(), i2.cuda(), f, rmap, flow_gt=rnd(B, 2, H, W).cuda(),
                                occ_gt=(torch.rand(B, 1, H, W, generator=G) < 0.3).float().cuda())[0]
    run_bwd(photo, f)
    a, fl = leaf(i1 * 0.01), leaf(flow)
    run_bwd(ocf.first_order_smoothness_loss(a, fl), a, fl)
    a, b = leaf(torch.rand(2, 3, 436, 1024, generator=G)), leaf(torch.rand(2, 3, 436, 1024, generator=G))
    run_bwd(ocf.ssim(a, b, 11), a, b)
    a, b = leaf(torch.rand(2, 3, 436, 1024, generator=G)), leaf(torch.rand(2, 3, 436, 1024, generator=G))
    run_bwd(ocf.census_loss(a, b, torch.rand(2, 1, 436, 1024, generator=G).cuda(), 3), a, b)


GROUPS = {"corr": group_corr, "normalize": group_normalize, "warp": group_warp, "scatter": group_scatter, "loss": group_loss,
          "ssim": group_ssim, "census": group_census, "resize": group_resize, "level": group_level, "data": group_data, "big": group_big}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated groups (default: all): " + ",".join(GROUPS))
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a CUDA device"
    _lib.load()
    names = [n for n in args.only.split(",") if n] or list(GROUPS)
    before = _lib.launch_count
    failed = []
    for n in names:
        t0 = time.time()
        try:
            GROUPS[n]()
            torch.cuda.synchronize()
            print("group %-10s ok  (%.1f s, %d C-ABI calls so far)" % (n, time.time() - t0, _lib.launch_count - before), flush=True)
        except Exception:   # keep going: the sanitizer's report on the other groups is still wanted
            failed.append(n)
            print("group %-10s FAILED" % n, flush=True)
            traceback.print_exc()
    print("sanitize_ops: %d groups, %d failed %s, %d C-ABI calls" % (len(names), len(failed), failed, _lib.launch_count - before))
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
