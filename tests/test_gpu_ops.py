"""GPU parity tests: the CUDA path (through the C ABI) vs the reference fixtures and vs the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): 1e-4 relative (max|d|/max|ref| and L2) for correlation / warp / normalisation
outputs and gradients; 1e-3 relative for loss scalars.
"""
import os

import pytest
import torch

from conftest import assert_close, assert_scalar_close, golden_op_files, load_golden
from oracle import ocflow_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4
LOSS_TOL = 1e-3


def cuda(t):
    return t.cuda()


def _grads(fn, inputs, cots):
    leaves = [t.clone().cuda().requires_grad_(True) for t in inputs]
    out = fn(*leaves)
    outs = out if isinstance(out, (list, tuple)) else [out]
    total = sum((o * c.cuda()).sum() for o, c in zip(outs, cots))
    return [o.detach() for o in outs], torch.autograd.grad(total, leaves, allow_unused=True)


@pytest.mark.parametrize("path", golden_op_files(), ids=os.path.basename)
def test_cuda_ops_match_reference_fixtures(path):
    import ocflow_b200 as ocf
    from ocflow_b200 import ops

    c = load_golden(path)
    f1, f2, flow, i1, i2, occ_soft = c["f1"], c["f2"], c["flow"], c["i1"], c["i2"], c["occ_soft"]
    for d in (4, 10):
        if "ref_corr_d%d" % d not in c:
            continue
        outs, grads = _grads(lambda a, b: ocf.compute_cost_volume(a, b, d), [f1, f2], [c["cot_corr_d%d" % d]])
        assert_close(outs[0], c["ref_corr_d%d" % d], TOL, "corr d=%d" % d)
        for g, r in zip(grads, c["ref_corr_d%d_grads" % d]):
            assert_close(g, r, TOL, "corr grad d=%d" % d)
    for i, kw in enumerate(c["norm_flags"]):
        outs, grads = _grads(lambda a, b: ocf.normalize_features([a, b], **kw), [f1, f2], c["cot_norm_%d" % i])
        for o, r in zip(outs, c["ref_norm_%d" % i]):
            assert_close(o, r, TOL, "normalize %s" % kw)
        for g, r in zip(grads, c["ref_norm_%d_grads" % i]):
            assert_close(g, r, 3e-4, "normalize grad %s" % kw)   # fp32 reference grads carry ~1e-4 cancellation noise
    for key, fn in (("ref_warp_ac1", lambda a, f: ocf.warp(a, f)), ("ref_warp_ac0", lambda a, f: ocf.network_warp(a, f)),
                    ("ref_warp_mask", lambda a, f: ocf.warp(a, f, is_mask=True))):
        outs, grads = _grads(fn, [f2, flow], [c["cot_warp"]])
        assert_close(outs[0], c[key], TOL, key)
        for g, r in zip(grads, c[key + "_grads"]):
            assert_close(g, r, TOL, key + " grad")
    assert_close(ocf.backwarp(cuda(f2), cuda(flow)), c["ref_backwarp"], TOL, "backwarp")
    assert_close(ocf.compute_range_map(cuda(flow)), c["ref_range_map"], TOL, "range map")
    assert_close(ocf.flow_to_warp(cuda(flow.permute(0, 2, 3, 1).contiguous())), c["ref_flow_to_warp"], 0, "flow_to_warp")
    rmap, occ = ocf.occlusion_mask(cuda(flow))
    assert_close(occ, c["ref_occ"], TOL, "occ")

    outs, grads = _grads(lambda a, b: ocf.photometric_error(a, b), [i2, i1], [torch.ones(())])
    assert_scalar_close(outs[0], c["ref_photo"], LOSS_TOL)
    outs, grads = _grads(lambda a, b, o: ocf.photometric_error(a, b, o), [i2, i1, occ_soft], [torch.ones(())])
    assert_scalar_close(outs[0], c["ref_photo_occ"], LOSS_TOL)
    assert_scalar_close(ocf.photometric_error(cuda(i2), cuda(i1), cuda(c["ref_occ"])), c["ref_photo_hardocc"], LOSS_TOL)
    outs, grads = _grads(lambda a: ocf.robust_l1(a), [f1], [c["cot_robust_l1"]])
    assert_close(outs[0], c["ref_robust_l1"], 1e-5)
    assert_close(grads[0], c["ref_robust_l1_grads"][0], TOL)
    assert_scalar_close(ocf.charbonnier_loss(cuda(f1)), c["ref_charbonnier"], LOSS_TOL)
    assert_close(ocf.charbonnier_loss(cuda(f1), reduction=False), c["ref_charbonnier_map"], 1e-5)
    img_s = c["img_smooth"]
    assert_scalar_close(ocf.first_order_smoothness_loss(cuda(img_s), cuda(flow)), c["ref_smooth1"], LOSS_TOL)
    assert_scalar_close(ocf.second_order_smoothness_loss(cuda(img_s), cuda(flow)), c["ref_smooth2"], LOSS_TOL)
    assert_scalar_close(ocf.first_order_smoothness_loss(cuda(img_s), cuda(occ_soft)), c["ref_smooth1_1ch"], LOSS_TOL)
    gx, gy = ocf.gradient(cuda(i1), 2)
    assert_close(gx, c["ref_gradient_s2"][0], 0)
    assert_close(gy, c["ref_gradient_s2"][1], 0)
    p, t = c["occ_prob"], c["occ_tgt"]
    assert_scalar_close(ocf.occlusion_bce_loss(cuda(p), cuda(t)), c["ref_bce"], LOSS_TOL)
    assert_scalar_close(ocf.occlusion_focal_loss(cuda(p), cuda(t)), c["ref_focal"], LOSS_TOL)
    assert_scalar_close(ocf.flow_l1_loss(cuda(flow), cuda(flow * 0.5 + 0.1)), c["ref_l1"], LOSS_TOL)
    assert_scalar_close(ocf.flow_mse_loss(cuda(flow), cuda(flow * 0.5 + 0.1)), c["ref_mse"], LOSS_TOL)


@pytest.mark.parametrize("path", golden_op_files(), ids=os.path.basename)
def test_cuda_loss_grads_match_reference_fixtures(path):
    """Gradients of the scalar losses wrt every input, against the real reference's autograd."""
    import ocflow_b200 as ocf

    c = load_golden(path)
    flow, i1, i2, occ_soft, img_s = c["flow"], c["i1"], c["i2"], c["occ_soft"], c["img_smooth"]
    one = [torch.ones(())]
    # the fixture's cotangent for scalar outputs is a random scalar; recover it from the generator seed is not
    # possible here, so compare grads normalised by the cotangent through a second oracle evaluation
    for key, fn_cuda, fn_orc, inputs in (
        ("photo", lambda a, b: ocf.photometric_error(a, b), lambda a, b: O.photometric_error(a, b), [i2, i1]),
        ("photo_occ", lambda a, b, o: ocf.photometric_error(a, b, o), lambda a, b, o: O.photometric_error(a, b, o), [i2, i1, occ_soft]),
        ("smooth1", lambda a, f: ocf.first_order_smoothness_loss(a, f), lambda a, f: O.first_order_smoothness_loss(a, f), [img_s, flow]),
        ("smooth2", lambda a, f: ocf.second_order_smoothness_loss(a, f), lambda a, f: O.second_order_smoothness_loss(a, f), [img_s, flow]),
        ("smooth1_1ch", lambda a, f: ocf.first_order_smoothness_loss(a, f), lambda a, f: O.first_order_smoothness_loss(a, f), [img_s, occ_soft]),
    ):
        _, gc = _grads(fn_cuda, inputs, one)
        leaves = [t.clone().double().requires_grad_(True) for t in inputs]
        go = torch.autograd.grad(fn_orc(*leaves), leaves)
        ref = c["ref_%s_grads" % key]
        for a, b, r in zip(gc, go, ref):
            assert_close(a, b, 2e-4, key + " grad vs fp64 oracle")
            # direction check against the stored reference grads (scaled by an unknown scalar cotangent)
            cos = float((a.cpu().double() * r.double()).sum() / (a.cpu().double().norm() * r.double().norm()).clamp_min(1e-30))
            assert abs(abs(cos) - 1.0) < 1e-5, key


SHAPES = [
    # B, C, H, W, flow scale   (pyramid levels of config 2, KITTI/Sintel odd widths, tiny and ragged cases)
    (2, 32, 24, 32, 2.0), (1, 196, 6, 8, 1.0), (2, 16, 47, 39, 3.0), (1, 3, 33, 65, 6.0), (2, 64, 12, 20, 1.5),
    (1, 1, 1, 1, 0.5), (1, 5, 2, 3, 1.0), (1, 96, 9, 311, 2.0), (3, 128, 12, 16, 1.0), (1, 17, 8, 36, 40.0),
    # BASELINE.json shapes in full: config-2 level L2 (B=8, C=32, 96x128) and a KITTI level of config 5 (B=4, C=16, 188x621)
    (8, 32, 96, 128, 2.0), (4, 16, 188, 621, 2.0),
]


@pytest.mark.parametrize("B,C,H,W,fs", SHAPES)
def test_cuda_ops_match_oracle(B, C, H, W, fs):
    import ocflow_b200 as ocf
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(B * 1000003 + C * 1009 + H * 31 + W)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.randn(B, C, H, W, generator=g) + 0.3
    flow = torch.randn(B, 2, H, W, generator=g) * fs
    cot = torch.randn(B, 81, H, W, generator=g)
    cotf = torch.randn(B, C, H, W, generator=g)

    def ref_grads(fn, inputs, cots, dtype=torch.float32):
        leaves = [t.clone().to(dtype).requires_grad_(True) for t in inputs]
        out = fn(*leaves)
        outs = out if isinstance(out, (list, tuple)) else [out]
        total = sum((o * c.to(dtype)).sum() for o, c in zip(outs, cots))
        return [o.detach() for o in outs], torch.autograd.grad(total, leaves, allow_unused=True)

    # correlation (d = 4 tiled kernel; d = 2 generic kernel), with and without the fused LeakyReLU
    for d, slope in ((4, 1.0), (4, 0.1), (2, 1.0)):
        nd = (2 * d + 1) ** 2
        ct = cot[:, :nd]
        orc = (lambda a, b: torch.nn.functional.leaky_relu(O.cost_volume(a, b, d), slope)) if slope != 1.0 else (lambda a, b: O.cost_volume(a, b, d))
        ro, rg = ref_grads(orc, [f1, f2], [ct], torch.float64)
        co, cg = _grads(lambda a, b: ops.cost_volume(a, b, d, leaky_slope=slope), [f1, f2], [ct])
        assert_close(co[0], ro[0], TOL, "corr d=%d slope=%g" % (d, slope))
        for a, b in zip(cg, rg):
            assert_close(a, b, TOL, "corr grad d=%d slope=%g" % (d, slope))

    # warp, both conventions, mask, fused occ multiply + flow scale
    occ = torch.rand(B, 1, H, W, generator=g)
    for ac, mask in ((True, False), (False, False), (True, True)):
        ro, rg = ref_grads(lambda a, f: O.warp(a, f, ac, mask), [f2, flow], [cotf], torch.float32)
        co, cg = _grads(lambda a, f: ops.warp(a, f, align_corners=ac, is_mask=mask), [f2, flow], [cotf])
        assert_close(co[0], ro[0], TOL, "warp ac=%s mask=%s" % (ac, mask))
        assert_close(cg[0], rg[0], TOL, "warp d_img")
        assert_close(cg[1], rg[1], 2e-4, "warp d_flow")
    ro, rg = ref_grads(lambda a, f, o: O.warp(a, f * 1.25, False) * o, [f2, flow, occ], [cotf])
    co, cg = _grads(lambda a, f, o: ops.warp(a, f, align_corners=False, occ=o, flow_scale=1.25), [f2, flow, occ], [cotf])
    assert_close(co[0], ro[0], TOL, "woc warp")
    for a, b in zip(cg, rg):
        assert_close(a, b, 2e-4, "woc warp grads")

    # normalisation (default flags) + range map + occlusion
    ro, rg = ref_grads(lambda a, b: O.normalize_features([a, b]), [f1, f2], [cotf, cotf * 0.5], torch.float64)
    co, cg = _grads(lambda a, b: ocf.normalize_features([a, b]), [f1, f2], [cotf, cotf * 0.5])
    for a, b in zip(co, ro):
        assert_close(a, b, TOL, "normalize")
    for a, b in zip(cg, rg):
        assert_close(a, b, TOL, "normalize grad")
    assert_close(ocf.compute_range_map(flow.cuda()), O.range_map(flow), TOL, "range map")


def test_cuda_fused_occ_photo_matches_oracle_chain():
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 37, 52
    i1 = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    i2 = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    fw = torch.randn(B, 2, H, W, generator=g) * 4
    bw = -fw + torch.randn(B, 2, H, W, generator=g) * 0.5
    flow_gt = torch.randn(B, 2, H, W, generator=g) * 5
    occ_gt = (torch.rand(B, 1, H, W, generator=g) < 0.3).float()
    f = fw.clone().double().requires_grad_(True)
    rm = O.range_map(bw.double())
    occ = O.occlusion_from_range_map(rm)
    warped = O.warp(i2.double(), f, True)
    photo = O.photometric_error(warped, i1.double(), occ)
    photo_occ = O.photometric_error(warped, i1.double(), 1.0 - occ)
    mse = ((f - flow_gt.double()) ** 2).mean()
    bce = O.binary_cross_entropy(occ_gt.double(), occ).mean()
    (gref,) = torch.autograd.grad(photo, f)

    fc = fw.clone().cuda().requires_grad_(True)
    rmc = ops.range_map(bw.cuda())
    p, po, m, b = ops.occ_photo_fused(i1.cuda(), i2.cuda(), fc, rmc, flow_gt.cuda(), occ_gt.cuda())
    assert_scalar_close(p, photo, LOSS_TOL, "photo")
    assert_scalar_close(po, photo_occ, LOSS_TOL, "photo_occ")
    assert_scalar_close(m, mse, LOSS_TOL, "mse")
    assert_scalar_close(b, bce, LOSS_TOL, "bce")
    (gc,) = torch.autograd.grad(p, fc)
    assert_close(gc, gref, 2e-4, "d photo / d flow")


def test_cuda_rejects_cpu_and_bad_args():
    import ocflow_b200 as ocf
    from ocflow_b200 import _lib

    with pytest.raises(TypeError):
        ocf.compute_cost_volume(torch.zeros(1, 2, 3, 3), torch.zeros(1, 2, 3, 3))
    with pytest.raises(TypeError):
        ocf.warp(torch.zeros(1, 2, 3, 3, device="cuda", dtype=torch.float64), torch.zeros(1, 2, 3, 3, device="cuda"))
    with pytest.raises(RuntimeError):
        ocf.compute_cost_volume(torch.zeros(1, 2, 3, 3, device="cuda"), torch.zeros(1, 2, 3, 3, device="cuda"), 17)
    assert _lib.load().ocf_corr_fwd(None, None, None, 1, 1, 1, 1, 4, 0, 1.0, None, None, None) == -1


@pytest.mark.parametrize("shape", [(2, 32, 24, 32), (1, 16, 9, 13)])
def test_corr_backward_consumes_concat_gradient_slice_in_place(shape):
    """cost_volume_flow_net.py:173-180: LeakyReLU(corr) is concatenated with other maps, so its gradient arrives as a
    channel slice (dense per item, wider batch stride).  ocf_corr_bwd takes that stride; results must equal the oracle."""
    from ocflow_b200 import ops

    B, C, H, W = shape
    g = torch.Generator().manual_seed(77)
    f1, f2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    other = torch.randn(B, 7, H, W, generator=g)
    cot = torch.randn(B, 7 + 81 + 7, H, W, generator=g)

    def run(corr, dev):
        a, b, o = (t.clone().to(dev).requires_grad_(True) for t in (f1, f2, other))
        cat = torch.cat((o, corr(a, b), o * 2), 1)
        (cat * cot.to(dev)).sum().backward()
        return a.grad, b.grad

    mine = run(lambda a, b: ops.cost_volume(a, b, 4, 0.1), "cuda")
    want = run(lambda a, b: torch.nn.functional.leaky_relu(O.cost_volume(a, b, 4), 0.1), "cpu")
    for m, w in zip(mine, want):
        assert_close(m, w, TOL, "corr grad through a concat slice")


@pytest.mark.parametrize("shape,ws", [((2, 3, 37, 45), 11), ((1, 3, 16, 16), 11), ((2, 3, 21, 18), 4), ((1, 1, 9, 70), 7),
                                      ((2, 3, 64, 96), 11)])
def test_ssim_matches_oracle_value_and_gradients(shape, ws):
    """inpainting_metrics/ssim/ssim.py:17-37.  The oracle's ssim is pinned to the real reference by the golden fixtures
    (tests/test_oracle_golden.py: ref_ssim11, ref_ssim4_map_mean); here the CUDA kernel meets it, both reductions."""
    import ocflow_b200 as ocf

    g = torch.Generator().manual_seed(5)
    i1 = torch.rand(*shape, generator=g)
    i2 = (i1 + 0.2 * torch.randn(*shape, generator=g)).clamp(0, 1)
    for size_average in (True, False):
        a, b = i1.clone().requires_grad_(True), i2.clone().requires_grad_(True)
        want = O.ssim(a, b, ws, size_average)
        cot = torch.ones_like(want) if size_average else torch.linspace(0.5, 1.5, shape[0])
        (want * cot).sum().backward()
        ac, bc = i1.clone().cuda().requires_grad_(True), i2.clone().cuda().requires_grad_(True)
        mine = ocf.ssim(ac, bc, ws, size_average)
        (mine * cot.cuda()).sum().backward()
        assert_close(mine.reshape(-1), want.reshape(-1), LOSS_TOL, "ssim ws=%d" % ws)
        assert_close(ac.grad, a.grad, 2e-4, "d ssim / d img1")
        assert_close(bc.grad, b.grad, 2e-4, "d ssim / d img2")
    loss = ocf.ssim_photometric_loss(i1.cuda(), i2.cuda(), ws)
    assert_scalar_close(loss, (1 - O.ssim(i1, i2, ws)) / 2, LOSS_TOL)


@pytest.mark.parametrize("B,C,H,W", [(4, 20, 56, 96), (3, 8, 61, 100), (2, 41, 75, 128), (5, 32, 48, 72)])
@pytest.mark.parametrize("slope", [1.0, 0.1])
def test_corr_persistent_kernels_match_oracle(B, C, H, W, slope):
    """Shapes with at least one tile per SM: the forward takes the persistent kernel (several tiles per CTA, TMA ring running
    across tile boundaries) and the backward runs its one-tile CTAs in several rounds with the LeakyReLU mask fetched into
    register bits.  Channel tails (C % 8 != 0), row / column overhang and gradients that arrive as concat slices are
    covered; each gradient is also requested alone (one CTA per tile instead of two)."""
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(B * 7919 + C * 31 + H + W)
    f1, f2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g) + 0.2
    other = torch.randn(B, 3, H, W, generator=g)
    cot = torch.randn(B, 3 + 81, H, W, generator=g)
    orc = (lambda a, b: torch.nn.functional.leaky_relu(O.cost_volume(a, b, 4), slope)) if slope != 1.0 else (lambda a, b: O.cost_volume(a, b, 4))

    def run(corr, dev, need=(True, True), dtype=torch.float32):
        a = f1.clone().to(dev, dtype).requires_grad_(need[0])
        b = f2.clone().to(dev, dtype).requires_grad_(need[1])
        out = corr(a, b)
        cat = torch.cat((other.to(dev, dtype), out), 1)        # the gradient of `out` arrives as a channel slice
        (cat * cot.to(dev, dtype)).sum().backward()
        return out.detach(), a.grad, b.grad

    want = run(orc, "cpu", dtype=torch.float64)
    mine = run(lambda a, b: ops.cost_volume(a, b, 4, leaky_slope=slope), "cuda")
    for m, w, name in zip(mine, want, ("corr", "d f1", "d f2")):
        assert_close(m, w, TOL, name)
    only1 = run(lambda a, b: ops.cost_volume(a, b, 4, leaky_slope=slope), "cuda", need=(True, False))
    only2 = run(lambda a, b: ops.cost_volume(a, b, 4, leaky_slope=slope), "cuda", need=(False, True))
    assert only1[2] is None and only2[1] is None
    assert_close(only1[1], want[1], TOL, "d f1 alone")
    assert_close(only2[2], want[2], TOL, "d f2 alone")


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 24, 32), (1, 20, 17, 23), (1, 64, 48, 64), (2, 7, 5, 9), (1, 256, 12, 16), (3, 16, 30, 44)])
@pytest.mark.parametrize("slope", [1.0, 0.1])
def test_corr_d10_tiled_forward_matches_oracle(B, C, H, W, slope):
    """max_displacement = 10 (441 planes, the FlowNetC-family call sites flow_net_c.py:22-25 / flow_occ_net_c.py:26 /
    occlusion_net_c.py:24): the tiled forward (three dy-group CTAs per tile; TMA, 16-byte cp.async and ragged 4-byte staging)
    and the tiled backward (partial sums of the three dy groups accumulated with vector reds) against the fp64 oracle."""
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(B * 131 + C * 17 + H * 5 + W)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.randn(B, C, H, W, generator=g) + 0.2
    a, b = f1.double().requires_grad_(True), f2.double().requires_grad_(True)
    ref = O.cost_volume(a, b, 10)
    if slope != 1.0:
        ref = torch.nn.functional.leaky_relu(ref, slope)
    cot = torch.randn(ref.shape, generator=g)
    ga, gb = torch.autograd.grad((ref * cot.double()).sum(), (a, b))
    x1, x2 = f1.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
    out = ops.cost_volume(x1, x2, 10, leaky_slope=slope)
    assert out.shape == (B, 441, H, W)
    assert_close(out, ref, TOL, "corr d=10 slope=%g" % slope)
    g1, g2 = torch.autograd.grad((out * cot.cuda()).sum(), (x1, x2))
    assert_close(g1, ga, TOL, "corr d=10 grad f1")
    assert_close(g2, gb, TOL, "corr d=10 grad f2")


def test_empty_batches_return_empty_results_like_the_reference():
    """B = 0 (and zero-sized images): the reference's tensor ops return empty tensors of the right shape; so do the wrappers
    (no kernel is launched)."""
    from ocflow_b200 import ops

    f = torch.zeros(0, 8, 6, 7, device="cuda", requires_grad=True)
    fl = torch.zeros(0, 2, 6, 7, device="cuda")
    cv = ops.cost_volume(f, f, 4)
    assert cv.shape == (0, 81, 6, 7) and cv.requires_grad
    assert ops.cost_volume(f, f, 10).shape == (0, 441, 6, 7)
    assert ops.warp(f, fl, align_corners=False).shape == (0, 8, 6, 7)
    assert ops.range_map(fl).shape == (0, 1, 6, 7)
    ref = O.cost_volume(torch.zeros(0, 8, 6, 7), torch.zeros(0, 8, 6, 7), 4)
    assert ref.shape == cv.shape


@pytest.mark.parametrize("shape", [(2, 16, 24, 32), (3, 5, 7, 9), (1, 196, 6, 8), (8, 32, 96, 128)])
def test_bias_leaky_relu_epilogue_equals_the_aten_ops(shape):
    """ops.bias_leaky_relu_ (the FlowNetCV conv-block epilogue, cost_volume_flow_net.py:11-15) against conv-bias-add + LeakyReLU as
    ATen runs them: bit-identical forward, gradients to the input (exact) and to the bias (summation order differs)."""
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).cuda()
    b = torch.randn(shape[1], generator=g).cuda()
    cot = torch.randn(*shape, generator=g).cuda()
    x1, b1 = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.leaky_relu(x1 + b1[None, :, None, None], 0.1)
    gx_ref, gb_ref = torch.autograd.grad((ref * cot).sum(), (x1, b1))
    x2, b2 = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = ops.bias_leaky_relu_(x2 * 1.0, b2, 0.1)          # x2 * 1.0: a fresh tensor the op may overwrite, like a convolution output
    assert torch.equal(out, ref)
    gx, gb = torch.autograd.grad((out * cot).sum(), (x2, b2))
    assert torch.equal(gx, gx_ref)
    assert_close(gb, gb_ref, 1e-5, "bias gradient")


def test_corr_backward_staging_variants_are_bit_identical():
    """The developer variants of the tiled correlation backward (coefficients loaded directly, early first feature stage, L2
    prefetch on / off / other stride) must give the gradients of the default path bit for bit: they only change HOW the
    operands reach the registers.  Runs tools/probe_bwd_knobs.py in a subprocess (the library decides at its first call
    whether it re-reads its knobs on every call) on two small geometries, one of them with a partial tile row and C % 8 != 0, and
    on one with more CTAs than resident slots (768: the prefetch has a next generation to prefetch for)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "tools", "probe_bwd_knobs.py"), "--reps", "1", "--geoms", "2x32x24x64,3x20x20x36,8x16x96x128"],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    lines = [l for l in proc.stdout.splitlines() if "median" in l]
    assert len(lines) == 3 * 12, proc.stdout
    assert all(("bit-identical" in l) or ("reference" in l) for l in lines), proc.stdout
