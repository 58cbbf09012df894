"""Loss helpers of models/model.py:27-114, utils.py:8-18 and the supervised steps (a-13)."""
import torch.nn as nn

from . import ops


def robust_l1(x, alpha=0.001):
    """(x^2 + alpha^2)^0.5 elementwise -- models/model.py:27-35."""
    return ops.robust_l1(x, alpha)


def charbonnier_loss(loss, alpha=0.001, reduction=True):
    """utils.py:8-18."""
    y = ops.robust_l1(loss, alpha)
    return y.mean() if reduction else y


def photometric_error(img_pred, img, occ=None):
    """models/model.py:37-46.  occ: 1 occluded, 0 non-occluded, [B,1,H,W]."""
    return ops.photometric_error(img_pred, img, occ)


class PhotometricLoss(nn.Module):
    """models/model.py:47-51."""

    def forward(self, img_pred, img):
        return photometric_error(img_pred, img)


def gradient(img, stride=1):
    """Forward differences (dx, dy) -- models/model.py:53-66."""
    return ops.gradient(img, stride)


def first_order_smoothness_loss(img, flow, alpha=100.0):
    """models/model.py:93-101."""
    return ops.smoothness_loss(img, flow, 1, alpha)


def second_order_smoothness_loss(img, flow, alpha=100.0):
    """models/model.py:103-114."""
    return ops.smoothness_loss(img, flow, 2, alpha)


def flow_mse_loss(flow_pred, flow):
    """F.mse_loss(flow_pred, flow) of FlowModel.general_step (models/flow_model.py:173-186)."""
    return ops.pair_loss(flow_pred, flow, ops.PAIR_MSE)


def flow_l1_loss(flow_pred, flow):
    """F.l1_loss(flow_pred, flow) of FlowOccModel.general_step (models/flow_occ_model.py:53)."""
    return ops.pair_loss(flow_pred, flow, ops.PAIR_L1)


def occlusion_bce_loss(occ_pred, occ):
    """F.binary_cross_entropy(occ_pred, occ) (models/flow_occ_model.py:54)."""
    return ops.pair_loss(occ_pred, occ, ops.PAIR_BCE)


def occlusion_focal_loss(occ_pred, occ):
    """Focal BCE (gamma=2) of OcclusionModel.general_step (models/occlusion_model.py:55-62)."""
    return ops.pair_loss(occ_pred, occ, ops.PAIR_FOCAL)


def ssim(img1, img2, window_size=11, size_average=True):
    """inpainting_metrics/ssim/ssim.py:64-75 (`ssim`): mean of the SSIM map (or per-item means when size_average=False)."""
    return ops.ssim(img1, img2, window_size, size_average)


class SSIM(nn.Module):
    """inpainting_metrics/ssim/ssim.py:39-62 (`SSIM` module; the cached window of the reference is internal to the kernel)."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average

    def forward(self, img1, img2):
        return ops.ssim(img1, img2, self.window_size, self.size_average)


def ssim_photometric_loss(img_pred, img, window_size=11):
    """SSIM-style photometric term named by north_star: (1 - SSIM(img_pred, img)) / 2, differentiable in both images.
    The reference defines SSIM only as a metric; this is the standard loss form built on its exact definition."""
    return (1.0 - ops.ssim(img_pred, img, window_size, True)) * 0.5


def census_loss(img_pred, img, occ=None, max_distance=3):
    """Census-style photometric term named by north_star.  The reference defines no census loss:
    this is the published UnFlow soft census with the occlusion weighting of photometric_error -- parity unpinned."""
    return ops.census_loss(img_pred, img, occ, max_distance)
