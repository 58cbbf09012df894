"""Mirrors of the reference's 11 warp bodies."""
from . import ops


def warp(img, flow, is_mask=False):
    """utils.py:20-58 -- align_corners=True bilinear backward warp; is_mask zeroes pixels whose footprint leaves
    the frame.  Also the body of FlowStageModel.warp / FlowModel.warp / TwoStageModel.warp (models/model.py:191-221)."""
    return ops.warp(img, flow, align_corners=True, is_mask=is_mask)


def backwarp(img, flow):
    """models/networks/pwc_net.py:6-29 (align_corners=True)."""
    return ops.warp(img, flow, align_corners=True)


def network_warp(img, flow):
    """The align_corners=False warp of the network classes: cost_volume_flow_net.py:121-151, flow_net.py:57-87,
    cost_volume_flow_occ_net.py:137-167, flow_occ_net.py:96-126, inpainting_model.py:22-52.  Samples at
    (x+u)*W/(W-1) - 0.5 -- the reference's normalise/un-normalise mismatch is reproduced on purpose."""
    return ops.warp(img, flow, align_corners=False)


def network_warp_method(self, img, flow):
    """Unbound-method form used by patch_reference (signature `warp(self, img, flow)`)."""
    return ops.warp(img, flow, align_corners=False)


def loss_warp_method(self, img, flow):
    return ops.warp(img, flow, align_corners=True)
