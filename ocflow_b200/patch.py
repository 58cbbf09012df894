"""patch_reference(): make an installed OCFlow reference tree run its hot path on ocflow_b200.

The reference binds its hot-path helpers at import time (`from ... import compute_cost_volume`) and copies `warp`
into 11 classes, so rebinding the defining module is not enough; the patch points are listed in SURVEY.md
section 8b.  This function only touches modules that are already importable (`models.*`, `utils` must be on
sys.path) and never imports anything from the oracle.

    import ocflow_b200.patch as p
    p.install_cost_volume_net()      # before importing the reference's FlowOcc* networks
    import models.model              # the reference
    p.patch_reference()
"""
import importlib
import sys
import types

from . import correlation_layer, cost_volume_net, losses, occlusion, warping


def install_cost_volume_net():
    """Provide the module the reference imports but does not ship (models/networks/cost_volume_net.py)."""
    name = "models.networks.cost_volume_net"
    if name not in sys.modules:
        mod = types.ModuleType(name)
        mod.CostVolumeLayer = cost_volume_net.CostVolumeLayer
        mod.__doc__ = "Injected by ocflow_b200.patch: CostVolumeLayer backed by the sm_100a correlation kernel."
        sys.modules[name] = mod
    return sys.modules[name]


def _try_import(name):
    try:
        return importlib.import_module(name)
    except Exception:  # missing third-party deps of unrelated reference modules must not break patching
        return None


_UNDO = []  # (object, attribute, original value) of every rebinding made by patch_reference(), newest last


def unpatch_reference():
    """Restore every symbol patch_reference() rebound (the injected cost_volume_net module stays: the reference has none).
    Returns the number of restored attributes."""
    n = len(_UNDO)
    while _UNDO:
        obj, attr, orig = _UNDO.pop()
        setattr(obj, attr, orig)
    return n


def patch_reference(verbose=False):
    """Rebind every hot-path symbol of the already-importable reference modules.  Returns the list of patched names;
    unpatch_reference() undoes it."""
    install_cost_volume_net()
    done = []

    def setp(obj, attr, value):
        if obj is not None and hasattr(obj, attr):
            orig = obj.__dict__[attr] if attr in getattr(obj, "__dict__", {}) else getattr(obj, attr)
            if orig is not value:
                _UNDO.append((obj, attr, orig))
            setattr(obj, attr, value)
            done.append("%s.%s" % (getattr(obj, "__name__", repr(obj)), attr))

    cl = _try_import("models.networks.correlation_layer")
    setp(cl, "compute_cost_volume", correlation_layer.compute_cost_volume)
    setp(cl, "normalize_features", correlation_layer.normalize_features)

    # modules that did `from models.networks.correlation_layer import ...`
    for modname in ("models.networks.cost_volume_flow_net", "models.networks.pwc_net", "models.networks.flow_net",
                    "models.networks.flow_net_c"):
        m = _try_import(modname)
        setp(m, "compute_cost_volume", correlation_layer.compute_cost_volume)
        setp(m, "normalize_features", correlation_layer.normalize_features)
    m = _try_import("models.networks.pwc_net")
    setp(m, "backwarp", warping.backwarp)

    # align_corners=False network warps (class attribute replaces all instances' bound method)
    for modname, clsnames in (("models.networks.cost_volume_flow_net", ("FlowNetCV",)),
                              ("models.networks.flow_net", ("FlowNet",)),
                              ("models.networks.cost_volume_flow_occ_net", ("FlowOccNetCV", "FlowOccNetCV2")),
                              ("models.networks.flow_occ_net", ("FlowOccNet",)),
                              ("models.inpainting_model", ("InpaintingModel",))):
        m = _try_import(modname)
        for cn in clsnames:
            cls = getattr(m, cn, None) if m is not None else None
            setp(cls, "warp", warping.network_warp_method)

    # FlowNetCV captures normalize_features per instance in __init__ (`self.normalize = normalize_features`, :49) and
    # FlowNet / FlowNetC capture compute_cost_volume (`self.correlation_layer`, `self.corr`): new instances pick up
    # the patched module globals; existing instances can be fixed with patch_instance().

    mm = _try_import("models.model")
    for name in ("robust_l1", "photometric_error", "gradient", "first_order_smoothness_loss", "second_order_smoothness_loss"):
        setp(mm, name, getattr(losses, name))
    for cn in ("FlowStageModel", "TwoStageModel", "TwoStageModelGC"):
        cls = getattr(mm, cn, None) if mm is not None else None
        setp(cls, "warp", warping.loss_warp_method)
        setp(cls, "flow_to_warp", occlusion.flow_to_warp_method)
        setp(cls, "compute_range_map", occlusion.compute_range_map_method)
    fm = _try_import("models.flow_model")
    cls = getattr(fm, "FlowModel", None) if fm is not None else None
    setp(cls, "warp", warping.loss_warp_method)
    setp(cls, "flow_to_warp", occlusion.flow_to_warp_method)
    setp(cls, "compute_range_map", occlusion.compute_range_map_method)

    ut = _try_import("utils")
    if ut is not None and hasattr(ut, "charbonnier_loss"):
        setp(ut, "warp", warping.warp)
        setp(ut, "charbonnier_loss", losses.charbonnier_loss)
    # SSIM (inpainting_metrics/ssim/ssim.py:39-75): the functional entry point and the module
    sm = _try_import("inpainting_metrics.ssim.ssim")
    setp(sm, "ssim", losses.ssim)
    setp(sm, "SSIM", losses.SSIM)
    if verbose:
        for d in done:
            print("patched", d)
    return done


def patch_instance(net):
    """Fix per-instance captures on an already-constructed reference network."""
    if hasattr(net, "normalize"):
        net.normalize = correlation_layer.normalize_features
    if hasattr(net, "correlation_layer") and callable(getattr(net, "correlation_layer")):
        net.correlation_layer = correlation_layer.compute_cost_volume
    if hasattr(net, "corr") and not hasattr(net.corr, "parameters"):
        net.corr = correlation_layer.compute_cost_volume
    return net
