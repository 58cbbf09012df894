"""CPU, build container only: the oracle restatement against the REAL reference imported in place (oracle/ref_loader.py),
on fresh seeded inputs that are NOT in the committed fixtures -- random shapes incl. odd sizes, flows leaving the frame, every
flag combination.  Skipped where the reference tree is absent (the GPU box); tests/test_oracle_golden.py covers that case
with the fixtures this same reference produced."""
import pytest
import torch

from conftest import assert_close, assert_scalar_close
from oracle import ocflow_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

SHAPES = [(1, 7, 9, 11, 2.0), (2, 16, 24, 32, 4.0), (2, 5, 17, 8, 15.0), (3, 32, 6, 8, 1.0)]


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


def _both(fn_ref, fn_orc, inputs, seed):
    """outputs and input gradients of both implementations under the same random cotangents"""
    res = []
    for fn in (fn_ref, fn_orc):
        leaves = [t.clone().requires_grad_(True) for t in inputs]
        out = fn(*leaves)
        outs = list(out) if isinstance(out, (list, tuple)) else [out]
        g = torch.Generator().manual_seed(seed)
        total = sum((o * torch.randn(o.shape, generator=g)).sum() for o in outs)
        grads = torch.autograd.grad(total, leaves, allow_unused=True)
        res.append(([o.detach() for o in outs], grads))
    return res


def _compare(res, tol, what):
    (ro, rg), (oo, og) = res
    for a, b in zip(oo, ro):
        assert_close(a, b, tol, what)
    for a, b in zip(og, rg):
        assert (a is None) == (b is None), what
        if a is not None:
            assert_close(a, b, tol, what + " grad")


@pytest.mark.parametrize("B,C,H,W,fs", SHAPES)
def test_ops_match_the_real_reference(R, B, C, H, W, fs):
    g = torch.Generator().manual_seed(B * 131 + C * 17 + H * 5 + W)
    f1, f2 = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g) + 0.1
    flow = torch.randn(B, 2, H, W, generator=g) * fs
    i1, i2 = torch.rand(B, 3, H, W, generator=g), torch.rand(B, 3, H, W, generator=g)
    occ = torch.rand(B, 1, H, W, generator=g)
    net = R.cost_volume_flow_net.FlowNetCV()
    stage = R.model.FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5})
    for d in (4, 10):
        _compare(_both(lambda a, b: R.correlation_layer.compute_cost_volume(a, b, d), lambda a, b: O.cost_volume(a, b, d), [f1, f2], 1), 2e-6,
                 "cost volume d=%d" % d)
    for kw in (dict(), dict(center=False), dict(normalize=False), dict(moments_across_channels=False),
               dict(moments_across_images=False), dict(moments_across_channels=False, moments_across_images=False)):
        _compare(_both(lambda a, b: R.correlation_layer.normalize_features([a, b], **kw), lambda a, b: O.normalize_features([a, b], **kw),
                       [f1, f2], 2), 1e-5, "normalize %s" % kw)
    _compare(_both(lambda a, f: net.warp(a, f), lambda a, f: O.warp(a, f, False), [f2, flow], 3), 1e-5, "network warp")
    _compare(_both(lambda a, f: stage.warp(a, f), lambda a, f: O.warp(a, f, True), [i2, flow], 4), 1e-5, "loss warp")
    _compare(_both(lambda a, f: R.utils.warp(a, f, True), lambda a, f: O.warp(a, f, True, True), [i2, flow], 5), 1e-5, "utils.warp is_mask")
    _compare(_both(lambda a, f: R.pwc_net.backwarp(a, f), lambda a, f: O.warp(a, f, True), [f2, flow], 6), 1e-5, "backwarp")
    with torch.no_grad():
        assert_close(O.range_map(flow), stage.compute_range_map(flow), 2e-6, "range map")
        assert_close(O.flow_to_warp(flow.permute(0, 2, 3, 1).contiguous()), stage.flow_to_warp(flow.permute(0, 2, 3, 1).contiguous()), 1e-7,
                     "flow_to_warp")
    _compare(_both(lambda p, q, o: R.model.photometric_error(p, q, o), lambda p, q, o: O.photometric_error(p, q, o), [i2, i1, occ], 7), 1e-5,
             "photometric(occ)")
    _compare(_both(lambda p, q: R.model.photometric_error(p, q), lambda p, q: O.photometric_error(p, q), [i2, i1], 8), 1e-5, "photometric")
    _compare(_both(lambda x: R.model.robust_l1(x), lambda x: O.robust_l1(x), [flow], 9), 1e-6, "robust_l1")
    _compare(_both(lambda x: R.utils.charbonnier_loss(x), lambda x: O.charbonnier_loss(x), [flow], 10), 1e-6, "charbonnier")
    if H > 2 and W > 2:
        _compare(_both(lambda im, f: R.model.first_order_smoothness_loss(im, f), lambda im, f: O.first_order_smoothness_loss(im, f),
                       [i1 * 0.05, flow], 11), 1e-5, "smooth1")
        _compare(_both(lambda im, f: R.model.second_order_smoothness_loss(im, f), lambda im, f: O.second_order_smoothness_loss(im, f),
                       [i1 * 0.05, flow], 12), 1e-5, "smooth2")
        for s in (1, 2):
            _compare(_both(lambda im: R.model.gradient(im, s), lambda im: O.gradient(im, s), [i1], 13), 1e-7, "gradient stride %d" % s)
    for ws in (11, 4):
        if min(H, W) >= 6:
            _compare(_both(lambda p, q: R.ssim.ssim(p, q, ws), lambda p, q: O.ssim(p, q, ws), [i1, i2], 14), 1e-5, "ssim %d" % ws)


def test_occlusion_aware_step_matches_the_real_reference(R):
    net = R.cost_volume_flow_net.FlowNetCV()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = O.deterministic_state_dict(shapes, seed=9, flow_gain=0.1)
    stage = R.model.FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5})
    stage.flow_pred.load_state_dict(sd)
    g = torch.Generator().manual_seed(99)
    imgs = torch.rand(1, 6, 64, 128, generator=g) * 2 - 1
    flow_gt = torch.randn(1, 2, 64, 128, generator=g) * 5
    occ_gt = (torch.rand(1, 1, 64, 128, generator=g) < 0.3).float()
    with torch.no_grad():
        ref = stage.general_step_occ_aware((imgs, flow_gt, occ_gt), 0, "train")
        mine = O.occ_aware_step(sd, (imgs, flow_gt, occ_gt))
        rf, rl2 = stage(imgs)
        of, ol2 = O.flownetcv_forward(sd, imgs)
    assert_close(of, rf, 1e-5, "flow1")
    assert_close(ol2, rl2, 1e-5, "flow_l2")
    for a, b in zip(mine, ref):
        assert_scalar_close(a, b, 1e-5)
