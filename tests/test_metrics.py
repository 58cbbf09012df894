"""End-point-error metrics: oracle vs the fixture generated from the REAL reference's numpy
functions (tests/golden/metrics.pt, oracle/make_golden_metrics.py), and the CUDA reduction vs both."""
import os

import pytest
import torch

from conftest import GOLD, assert_scalar_close, load_golden
from oracle import ocflow_oracle as O

CASES = load_golden(os.path.join(GOLD, "metrics.pt"))


@pytest.mark.parametrize("c", CASES, ids=lambda c: "%dx%d" % tuple(c["gt"].shape[:2]))
def test_oracle_metrics_match_reference_fixture(c):
    gt, pred, gt3 = c["gt"], c["pred"], c["gt3"]
    assert_scalar_close(O.flow_error(gt[..., 0], gt[..., 1], pred[..., 0], pred[..., 1]), c["ref_epe"], 1e-6)
    ones = torch.ones(gt.shape[:2])
    for got, ref in ((O.flow_kitti_error(gt3[..., 0], gt3[..., 1], pred[..., 0], pred[..., 1], ones), c["ref_kitti2"]),
                     (O.flow_kitti_error(gt3[..., 0], gt3[..., 1], pred[..., 0], pred[..., 1], gt3[..., 2]), c["ref_kitti3"])):
        assert_scalar_close(got[0], ref[0], 1e-6)
        assert_scalar_close(got[1], ref[1], 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=lambda c: "%dx%d" % tuple(c["gt"].shape[:2]))
def test_cuda_metrics_match_reference_fixture(c):
    from ocflow_b200 import metrics as M

    gt, pred, gt3 = c["gt"].cuda(), c["pred"].cuda(), c["gt3"].cuda()
    assert_scalar_close(M.evaluate_flow(gt, pred), c["ref_epe"], 1e-5)
    e2, a2 = M.evaluate_kitti_flow(gt3[..., :2], pred)
    e3, a3 = M.evaluate_kitti_flow(gt3, pred)
    assert_scalar_close(e2, c["ref_kitti2"][0], 1e-5)
    assert_scalar_close(a2, c["ref_kitti2"][1], 1e-6)
    assert_scalar_close(e3, c["ref_kitti3"][0], 1e-5)
    assert_scalar_close(a3, c["ref_kitti3"][1], 1e-6)


@pytest.mark.gpu
def test_cuda_batch_epe_full_size():
    from ocflow_b200 import metrics as M

    g = torch.Generator().manual_seed(5)
    gt = torch.randn(8, 2, 384, 512, generator=g) * 5
    pred = gt + torch.randn(8, 2, 384, 512, generator=g)
    want = torch.sqrt(((gt - pred) ** 2).sum(1)).double().mean()
    assert_scalar_close(M.batch_epe(pred.cuda(), gt.cuda()), want, 1e-5)
    assert float(M.batch_epe(gt.cuda(), gt.cuda())) == 0.0
    with pytest.raises(TypeError):
        M.batch_epe(pred, gt)          # CPU tensors: there is no CPU path
