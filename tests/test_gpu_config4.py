"""GPU, BASELINE.json config 4: the Sintel-shape (436x1024) forward-backward occlusion pipeline -- range map of the
backward flow -> occlusion mask -> warp -> occlusion-weighted Charbonnier + census + SSIM terms, forward and the gradient
to the forward flow -- against the CPU oracle on the same seeded inputs at the FULL native shape (one pair)."""
import os
import sys

import pytest
import torch

from conftest import ROOT, assert_close, assert_scalar_close
from oracle import ocflow_oracle as O

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("H,W", [(436, 1024), (109, 256)])
def test_sintel_pipeline_matches_oracle(H, W):
    import config4_sintel as c4

    g = torch.Generator().manual_seed(4)
    img1 = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    img2 = torch.rand(1, 3, H, W, generator=g) * 2 - 1
    # smooth images as well: white-noise images make every bilinear cell a different function (ill-conditioned d/dflow)
    img1 = torch.nn.functional.avg_pool2d(img1, 5, 1, 2)
    img2 = torch.nn.functional.avg_pool2d(img2, 5, 1, 2)
    fw = torch.randn(1, 2, H, W, generator=g) * 8
    bw = -fw + torch.randn(1, 2, H, W, generator=g) * 0.5

    f_ref = fw.clone().requires_grad_(True)
    rmap = O.range_map(bw)
    occ = O.occlusion_from_range_map(rmap)
    warped = O.warp(img2, f_ref, True)
    ref_parts = (O.photometric_error(warped, img1, occ), O.photometric_error(warped, img1, 1 - occ),
                 O.census_loss(warped, img1, occ, 3), (1 - O.ssim(warped, img1, 11)) * 0.5)
    (ref_parts[0] + ref_parts[2] + ref_parts[3]).backward()

    f = fw.cuda().requires_grad_(True)
    total, parts = c4.pipeline(img1.cuda(), img2.cuda(), f, bw.cuda())
    for mine, ref, name in zip(parts, ref_parts, ("photo", "photo_occ", "census", "ssim")):
        assert_scalar_close(mine, ref, 1e-3, name)
    total.backward()
    assert_close(f.grad, f_ref.grad, 1e-4, "d total / d flow")
