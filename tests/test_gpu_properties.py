"""GPU, BASELINE.json full sizes: size-independent properties (the CPU oracle is too slow / memory-hungry there)."""
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu

# (B, C, H, W): config-2 pyramid levels L2..L6 and config-5 KITTI level L1/L2 (odd widths)
LEVELS = [(8, 32, 96, 128), (8, 64, 48, 64), (8, 96, 24, 32), (8, 128, 12, 16), (8, 196, 6, 8), (4, 16, 188, 621), (4, 32, 94, 311)]


def _k(dy, dx, d=4):
    return (dy + d) * (2 * d + 1) + (dx + d)


@pytest.mark.parametrize("B,C,H,W", LEVELS)
def test_corr_symmetry_linearity_adjoint(B, C, H, W):
    from ocflow_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(7)
    f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
    f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    f3 = torch.randn(B, C, H, W, device="cuda", generator=g)
    c12 = ops.cost_volume(f1, f2)
    c21 = ops.cost_volume(f2, f1)
    # symmetry: corr(f1,f2)[k(dy,dx), y, x] == corr(f2,f1)[k(-dy,-dx), y+dy, x+dx]
    for dy, dx in ((0, 0), (-4, 4), (3, -2), (1, 0), (4, 4)):
        ya, yb = max(0, -dy), min(H, H - dy)
        xa, xb = max(0, -dx), min(W, W - dx)
        a = c12[:, _k(dy, dx), ya:yb, xa:xb]
        b = c21[:, _k(-dy, -dx), ya + dy:yb + dy, xa + dx:xb + dx]
        assert_close(a, b, 1e-5, "symmetry (%d,%d)" % (dy, dx))
    # centre plane is the plain channel mean of the product
    assert_close(c12[:, 40], (f1 * f2).mean(1), 1e-5, "centre plane")
    # out-of-image displacements are exactly zero
    assert float(c12[:, _k(-4, 0), :4].abs().max()) == 0.0 and float(c12[:, _k(0, 4), :, W - 4:].abs().max()) == 0.0
    # linearity in the second argument
    assert_close(ops.cost_volume(f1, f2 + 2.0 * f3), c12 + 2.0 * ops.cost_volume(f1, f3), 2e-5, "linearity")
    # adjoint identity <g, corr(f1,f2)> = <corr_bwd_f1(g), f1> = <corr_bwd_f2(g), f2>   (bilinear form)
    a1 = f1.clone().requires_grad_(True)
    a2 = f2.clone().requires_grad_(True)
    out = ops.cost_volume(a1, a2)
    cot = torch.randn(out.shape, device="cuda", generator=g)
    d1, d2 = torch.autograd.grad((out * cot).sum(), (a1, a2))
    s = float((out.double() * cot.double()).sum())
    assert abs(float((d1.double() * f1.double()).sum()) - s) <= 1e-5 * abs(s) + 1e-3
    assert abs(float((d2.double() * f2.double()).sum()) - s) <= 1e-5 * abs(s) + 1e-3


@pytest.mark.parametrize("B,C,H,W", [(8, 3, 384, 512), (8, 32, 96, 128), (2, 3, 436, 1024), (2, 16, 188, 621)])
def test_warp_and_range_map_properties(B, C, H, W):
    from ocflow_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(9)
    img = torch.randn(B, C, H, W, device="cuda", generator=g)
    zero = torch.zeros(B, 2, H, W, device="cuda")
    # zero flow, align_corners=True: identity (up to the fp32 normalise/un-normalise round trip)
    assert_close(ops.warp(img, zero, align_corners=True), img, 2e-4, "identity warp")
    # integer translation by (+3, -2): out[y, x] = img[y-2, x+3] inside, 0 outside
    shift = zero.clone()
    shift[:, 0] = 3.0
    shift[:, 1] = -2.0
    out = ops.warp(img, shift, align_corners=True)
    assert_close(out[:, :, 2:, :W - 3], img[:, :, :H - 2, 3:], 5e-4, "integer shift")
    assert float(out[:, :, :2].abs().max()) < 1e-3 and float(out[:, :, :, W - 3:].abs().max()) < 1e-3
    # warp is linear in the image; d_img is its adjoint:  <cot, warp(img)> == <d_img, img>
    flow = torch.randn(B, 2, H, W, device="cuda", generator=g) * 3
    for ac in (True, False):
        a = img.clone().requires_grad_(True)
        o = ops.warp(a, flow, align_corners=ac)
        cot = torch.randn(o.shape, device="cuda", generator=g)
        (d_img,) = torch.autograd.grad((o * cot).sum(), a)
        s = float((o.double() * cot.double()).sum())
        assert abs(float((d_img.double() * img.double()).sum()) - s) <= 1e-4 * abs(s) + 1e-2
    # range map: zero flow -> all ones; total mass of an in-frame flow == number of pixels; occ in [0,1]
    rm = ops.range_map(zero)
    assert float((rm - 1).abs().max()) == 0.0
    small = torch.rand(B, 2, H, W, device="cuda", generator=g) * 0.9
    small[:, 0, :, W - 1] = 0
    small[:, 1, H - 1, :] = 0
    rm, occ = ops.range_map(small, with_occlusion=True)
    assert abs(float(rm.double().sum()) - B * H * W) <= 1e-5 * B * H * W
    assert float(occ.min()) >= 0.0 and float(occ.max()) <= 1.0
    # integer flows splat whole pixels: range map of a +1 px shift is 1 everywhere except the first column
    one = zero.clone()
    one[:, 0] = 1.0
    rm = ops.range_map(one)
    assert float((rm[..., 1:] - 1).abs().max()) == 0.0 and float(rm[..., 0].abs().max()) == 0.0


def test_normalize_idempotent_and_moments():
    import ocflow_b200 as ocf

    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(8, 32, 96, 128, device="cuda", generator=g) * 3 + 1.5
    b = torch.randn(8, 32, 96, 128, device="cuda", generator=g) * 0.5 - 0.2
    na, nb = ocf.normalize_features([a, b])
    both = torch.stack((na, nb)).double()
    assert abs(float(both.mean())) < 1e-4
    mean_var = torch.stack([t.double().var(dim=(1, 2, 3), unbiased=False) for t in (na, nb)]).mean()
    assert abs(float(mean_var) - 1.0) < 1e-4
    na2, nb2 = ocf.normalize_features([na, nb])
    assert_close(na2, na, 1e-4, "idempotence")
    assert_close(nb2, nb, 1e-4, "idempotence")


def test_fused_loss_consistent_with_unfused_ops_full_size():
    import ocflow_b200 as ocf
    from ocflow_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W = 8, 384, 512
    i1 = torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1
    i2 = torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1
    fw = (torch.randn(B, 2, H, W, device="cuda", generator=g) * 8).requires_grad_(True)
    bw = -fw.detach() + torch.randn(B, 2, H, W, device="cuda", generator=g) * 0.5
    rm, occ = ops.range_map(bw, with_occlusion=True)
    p, po, _, _ = ops.occ_photo_fused(i1, i2, fw, rm)
    (gf,) = torch.autograd.grad(p, fw)
    fw2 = fw.detach().clone().requires_grad_(True)
    warped = ocf.warp(i2, fw2)
    p2 = ocf.photometric_error(warped, i1, occ)
    po2 = ocf.photometric_error(warped, i1, 1.0 - occ)
    (gf2,) = torch.autograd.grad(p2, fw2)
    assert abs(float(p) - float(p2)) <= 1e-5 * abs(float(p2))
    assert abs(float(po) - float(po2)) <= 1e-5 * abs(float(po2))
    assert_close(gf, gf2, 1e-4, "fused vs unfused d photo/d flow")
