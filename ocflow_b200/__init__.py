"""ocflow_b200 -- B200-native (sm_100a) implementation of OCFlow's per-pyramid-level hot path.

Python mirror of the reference's hot-path symbols over hand-written CUDA behind a C ABI
(include/ocflow_b200.h).  Layout follows the reference's modules:

    correlation_layer   compute_cost_volume, normalize_features   (models/networks/correlation_layer.py)
    cost_volume_net     CostVolumeLayer                            (models/networks/cost_volume_net.py -- missing upstream)
    warping             warp (utils.py), network_warp (FlowNetCV.warp ...), backwarp (pwc_net.py)
    occlusion           flow_to_warp, compute_range_map, occlusion_mask (models/model.py:223-305, 391)
    losses              robust_l1, photometric_error, charbonnier_loss, gradient, *_smoothness_loss, supervised losses,
                        ssim / SSIM (inpainting_metrics/ssim/ssim.py), census_loss (no upstream definition)
    metrics             flow_error, evaluate_flow, flow_kitti_error, evaluate_kitti_flow (models/data/utils/flow_utils.py)
    data                pack_pairs: uint8 frames (+ flow) -> cropped, normalised [B,6,H,W] on the device (datasets.py / datamodule)
    flow_io             read_flow, save_flow (.flo files, models/data/utils/flow_utils.py)
    flow_net_cv         FlowNetCV (state_dict compatible with the reference network)
    flow_model          FlowModel (models/flow_model.py, model='pwc': BASELINE config 1)
    flow_stage          FlowStageModel (general_step / general_step_occ / general_step_occ_aware / training_step)
    patch               patch_reference(): rebinds the reference's own symbols to this package
"""
from .correlation_layer import compute_cost_volume, normalize_features  # noqa: F401
from .cost_volume_net import CostVolumeLayer  # noqa: F401
from .warping import warp, network_warp, backwarp  # noqa: F401
from .occlusion import flow_to_warp, compute_range_map, occlusion_mask  # noqa: F401
from .losses import (robust_l1, photometric_error, PhotometricLoss, charbonnier_loss, gradient,  # noqa: F401
                     first_order_smoothness_loss, second_order_smoothness_loss, flow_mse_loss, flow_l1_loss,
                     occlusion_bce_loss, occlusion_focal_loss, ssim, SSIM, ssim_photometric_loss, census_loss)

from . import data, flow_io, metrics  # noqa: F401,E402

__version__ = "0.1.0"
