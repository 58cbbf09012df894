"""GPU: ocf_resize_bilinear (align_corners=True bilinear resize, gather backward) against the fp64 oracle restatement of the
ATen op the reference calls (F.interpolate at cost_volume_flow_net.py:245 and models/model.py:396), values and gradients."""
import pytest
import torch

from conftest import assert_close
from oracle import ocflow_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,kw,mul", [((2, 2, 12, 16), dict(scale_factor=4), 20.0), ((2, 3, 64, 96), dict(scale_factor=0.25), 1.0),
                                          ((8, 2, 96, 128), dict(scale_factor=4), 20.0), ((2, 3, 384, 512), dict(scale_factor=0.25), 1.0),
                                          ((2, 1, 7, 5), dict(size=(13, 9)), 1.0), ((1, 2, 9, 11), dict(size=(4, 1)), 0.5),
                                          ((1, 1, 1, 6), dict(size=(3, 6)), 1.0), ((1, 2, 109, 256), dict(size=(436, 1024)), 20.0)])
def test_resize_bilinear_matches_oracle(shape, kw, mul):
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    ref_in = x.double().requires_grad_(True)
    ref = O.resize_bilinear(ref_in, **kw) * mul
    cot = torch.randn(ref.shape, generator=g)
    (ref_g,) = torch.autograd.grad((ref * cot.double()).sum(), ref_in)
    xc = x.cuda().requires_grad_(True)
    out = ops.resize_bilinear(xc, mul=mul, **kw)
    assert_close(out, ref, 1e-4, "resize output")   # fp32 source coordinates (r * dst, as ATen computes them) vs the fp64 oracle
    (gc,) = torch.autograd.grad((out * cot.cuda()).sum(), xc)
    assert_close(gc, ref_g, 1e-4, "resize gradient")


def test_resize_bilinear_equals_torch_interpolate_on_device():
    """Same values as the ATen CUDA kernel the unpatched reference runs (fp32 vs fp32)."""
    from ocflow_b200 import ops

    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 2, 96, 128, generator=g).cuda()
    ref = torch.nn.functional.interpolate(x, scale_factor=4, mode="bilinear", align_corners=True) * 20
    assert_close(ops.resize_bilinear(x, scale_factor=4, mul=20.0), ref, 1e-6, "vs ATen")
    img = torch.rand(2, 3, 384, 512, generator=g).cuda()
    ref = torch.nn.functional.interpolate(img, scale_factor=0.25, mode="bilinear", align_corners=True)
    assert_close(ops.resize_bilinear(img, scale_factor=0.25), ref, 1e-6, "vs ATen, x0.25")
