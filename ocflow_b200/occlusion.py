"""Range-map occlusion estimation (models/model.py:223-305, 391 = models/flow_model.py:81-163)."""
from . import ops


def flow_to_warp(flow):
    """[B,H,W,2] flow -> [B,H,W,2] endpoints (x+u, y+v).  models/model.py:223-241."""
    return ops.flow_to_warp(flow)


def compute_range_map(flow):
    """[B,2,H,W] flow -> [B,1,H,W]: bilinear forward splat of ones along the flow (models/model.py:243-305)."""
    return ops.range_map(flow)


def occlusion_mask(flow):
    """1 - clamp(range_map(flow), 0, 1): 1 = occluded (models/model.py:388-391).  Returns (range_map, occ)."""
    return ops.range_map(flow, with_occlusion=True)


# unbound-method forms for patch_reference
def flow_to_warp_method(self, flow):
    return ops.flow_to_warp(flow)


def compute_range_map_method(self, flow):
    return ops.range_map(flow)
