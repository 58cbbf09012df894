"""Flow evaluation metrics on the device -- mirror of models/data/utils/flow_utils.py:179-306.

The reference evaluates these in numpy on [H,W(,C)] arrays copied back to the host; here ground truth and prediction
stay on the GPU and the whole metric is one fused reduction (`ocf_flow_metrics`).  Same function names and argument
order; inputs are CUDA tensors in the reference's layouts ([H,W] maps, [H,W,2|3] flows) or batched [B,2,H,W].
Deviation: `flow_error(..., occ=...)` of the reference cannot run on 2-D maps as written (it indexes the [H,W] error map
with a flattened mask, flow_utils.py:227-230); `occ` is honoured here with the evident intent (mean over occ == 0).
"""
import torch

from . import _lib
from .ops import _p, _req, _stream


def _metrics(gt, pred, mask, kitti):
    gt = _req(gt, "gt_flow", 4)
    pred = _req(pred, "pred_flow", 4)
    if gt.shape != pred.shape or gt.shape[1] != 2:
        raise ValueError("gt and pred must both be [B,2,H,W] (got %s and %s)" % (tuple(gt.shape), tuple(pred.shape)))
    B, _, H, W = gt.shape
    if mask is not None:
        mask = _req(mask.to(torch.float32), "mask", 4)
        if mask.shape != (B, 1, H, W):
            raise ValueError("mask must be [B,1,H,W]")
    sums = torch.empty(3, device=gt.device, dtype=torch.float64)
    with torch.cuda.device_of(gt):
        _lib.call("ocf_flow_metrics", _p(gt), _p(pred), _p(mask), _p(sums), B, H, W, int(kitti), _stream())
    return sums


def _as_b2hw(*maps):
    """[H,W] component maps -> [1,2,H,W]."""
    return torch.stack([m.to(torch.float32) for m in maps], 0).unsqueeze(0)


def flow_error(tu, tv, u, v, occ=None):
    """Average end-point error, flow_utils.py:179-232 (ground truth above 1e7 marks unknown pixels: zeroed, still counted)."""
    gt, pred = _as_b2hw(tu, tv), _as_b2hw(u, v)
    if occ is None:
        s = _metrics(gt, pred, None, 0)
    else:
        unknown = (tu.abs() > 1e7) | (tv.abs() > 1e7)
        gt = gt.masked_fill(unknown, 0.0)
        pred = pred.masked_fill(unknown, 0.0)
        s = _metrics(gt, pred, (1 - occ.to(torch.float32)).reshape(1, 1, *tu.shape), 1)
    return (s[0] / s[1]).to(torch.float32)


def flow_kitti_error(tu, tv, u, v, mask):
    """(mean EPE over mask != 0, 1 - outlier ratio), flow_utils.py:234-271."""
    s = _metrics(_as_b2hw(tu, tv), _as_b2hw(u, v), mask.reshape(1, 1, *tu.shape), 1)
    return (s[0] / s[1]).to(torch.float32), (1.0 - s[2] / s[1]).to(torch.float32)


def evaluate_flow(gt_flow, pred_flow, occ=None):
    """flow_utils.py:289-296; gt_flow / pred_flow are [H,W,2]."""
    return flow_error(gt_flow[:, :, 0], gt_flow[:, :, 1], pred_flow[:, :, 0], pred_flow[:, :, 1], occ)


def evaluate_kitti_flow(gt_flow, pred_flow, rigid_flow=None):
    """flow_utils.py:299-310; a third ground-truth channel is the validity mask."""
    if gt_flow.shape[2] == 2:
        mask = torch.ones(gt_flow.shape[:2], device=gt_flow.device)
    elif gt_flow.shape[2] == 3:
        mask = gt_flow[:, :, 2]
    else:
        raise ValueError("gt_flow must be [H,W,2] or [H,W,3]")
    return flow_kitti_error(gt_flow[:, :, 0], gt_flow[:, :, 1], pred_flow[:, :, 0], pred_flow[:, :, 1], mask)


def batch_epe(flow_pred, flow_gt):
    """Mean end-point error of a [B,2,H,W] batch (the quantity the reference logs per validation image)."""
    s = _metrics(flow_gt, flow_pred, None, 0)
    return (s[0] / s[1]).to(torch.float32)
