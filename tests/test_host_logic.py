"""CPU: host-side mirror of the reference API -- argument checking, state_dict compatibility, patch points,
and that the product never imports the oracle or a CPU fallback."""
import os
import re

import pytest
import torch

from conftest import GOLD, ROOT, load_golden


def test_ops_refuse_cpu_and_wrong_dtype():
    import ocflow_b200 as ocf

    x = torch.zeros(1, 2, 4, 4)
    for fn in (lambda: ocf.compute_cost_volume(x, x), lambda: ocf.normalize_features([x, x]), lambda: ocf.warp(x, x),
               lambda: ocf.network_warp(x, x), lambda: ocf.compute_range_map(x), lambda: ocf.photometric_error(x, x),
               lambda: ocf.robust_l1(x), lambda: ocf.first_order_smoothness_loss(x, x), lambda: ocf.gradient(x),
               lambda: ocf.flow_mse_loss(x, x), lambda: ocf.CostVolumeLayer()(x, x),
               lambda: ocf.census_loss(x, x), lambda: ocf.ssim(x, x), lambda: ocf.metrics.batch_epe(x, x),
               lambda: ocf.metrics.evaluate_flow(torch.zeros(4, 4, 2), torch.zeros(4, 4, 2)),
               lambda: ocf.data.pack_pairs(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), torch.zeros(1, 4, 4, 3, dtype=torch.uint8))):
        with pytest.raises(TypeError, match="no CPU path"):
            fn()
    # the whole-network mirrors end up in the same ops: a CPU model cannot silently run anywhere else
    from ocflow_b200.flow_model import FlowModel
    with pytest.raises(TypeError, match="no CPU path"):
        FlowModel({"model": "pwc", "learning_rate": 1e-3})(torch.zeros(1, 6, 64, 64))


def test_flownetcv_state_dict_matches_reference_shapes():
    from ocflow_b200.flow_net_cv import FlowNetCV

    shapes = load_golden(os.path.join(GOLD, "net_2x64x64.pt"))["shapes"]
    net = FlowNetCV()
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert mine == shapes
    assert sum(p.numel() for p in net.parameters()) == 9374340
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 64, 64))


def test_flow_stage_model_hparams_and_api():
    from ocflow_b200.flow_stage import FlowStageModel

    m = FlowStageModel({"learning_rate": 1e-5, "model": "pwc"})
    assert (m.photo_weight, m.smooth1_weight, m.smooth2_weight, m.occ_aware, m.displacement) == (1.0, 0.0, 1.0, False, 4)
    for name in ("warp", "flow_to_warp", "compute_range_map", "general_step", "general_step_occ", "general_step_occ_aware",
                 "training_step", "validation_step", "test_step", "configure_optimizers", "save_state_dict"):
        assert callable(getattr(m, name))
    assert isinstance(m.configure_optimizers(), torch.optim.Adam)
    with pytest.raises(ValueError):
        FlowStageModel({"learning_rate": 1e-5, "model": "simple"})
    with pytest.raises(ValueError):
        m.general_step_occ_aware((1,), 0, "train")


def test_same_seed_gives_reference_initial_weights():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    R = ref_loader.load()
    from ocflow_b200.flow_net_cv import FlowNetCV

    torch.manual_seed(0)
    a = R.cost_volume_flow_net.FlowNetCV().state_dict()
    torch.manual_seed(0)
    b = FlowNetCV().state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_patch_reference_rebinds_every_patch_point():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    R = ref_loader.load()
    import importlib

    import ocflow_b200.patch as P
    from ocflow_b200 import correlation_layer, losses, warping

    import sys
    orig_ccv = R.correlation_layer.compute_cost_volume
    orig_ssim = R.ssim.ssim
    try:
        done = P.patch_reference()
        assert R.cost_volume_flow_net.compute_cost_volume is correlation_layer.compute_cost_volume
        assert R.cost_volume_flow_net.normalize_features is correlation_layer.normalize_features
        assert R.pwc_net.backwarp is warping.backwarp
        assert R.model.photometric_error is losses.photometric_error
        assert R.model.first_order_smoothness_loss is losses.first_order_smoothness_loss
        assert R.utils.warp is warping.warp
        assert R.cost_volume_flow_net.FlowNetCV.warp is warping.network_warp_method
        assert "models.networks.cost_volume_net" in sys.modules
        # the four importers of the missing module become importable
        occ_net = importlib.import_module("models.networks.cost_volume_flow_occ_net")
        assert hasattr(occ_net, "FlowOccNetCV")
        assert len(done) >= 25
        # a freshly constructed reference net now captures our normalize_features (cost_volume_flow_net.py:49)
        assert R.cost_volume_flow_net.FlowNetCV().normalize is correlation_layer.normalize_features
    finally:
        assert P.unpatch_reference() >= 25
    # everything is back: the reference's own functions again (other tests run the real reference after this one)
    assert R.correlation_layer.compute_cost_volume is orig_ccv and R.cost_volume_flow_net.compute_cost_volume is orig_ccv
    assert R.ssim.ssim is orig_ssim
    assert R.model.photometric_error is not losses.photometric_error
    assert R.cost_volume_flow_net.FlowNetCV.warp is not warping.network_warp_method
    assert P.unpatch_reference() == 0


def test_product_never_imports_oracle_or_falls_back():
    pkg = os.path.join(ROOT, "ocflow_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|ocflow_oracle|ref_loader|grid_sample|/root/reference", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, fn)).read()
                code = "\n".join(l for l in src.splitlines() if not l.strip().startswith(("#", "//", '"""')))
                m = bad.search(re.sub(r'""".*?"""', "", code, flags=re.S))
                assert m is None, "%s: %r" % (fn, m.group(0))


def test_conv_math_hparam_is_validated_and_scoped():
    """hparams['conv_math'] (not a reference hyper-parameter): 'fp32' is the parity mode, 'tf32' torch's default conv math;
    the scope sets cudnn.allow_tf32 for the convolutions dispatched inside it and restores the previous value."""
    import torch

    from ocflow_b200.flow_stage import FlowStageModel

    with pytest.raises(ValueError):
        FlowStageModel({"model": "pwc", "learning_rate": 1e-3, "conv_math": "bf16"})
    old = torch.backends.cudnn.allow_tf32
    try:
        for mode, want in (("fp32", False), ("tf32", True)):
            m = FlowStageModel({"model": "pwc", "learning_rate": 1e-3, "conv_math": mode})
            torch.backends.cudnn.allow_tf32 = not want
            with m.conv_math_scope():
                assert torch.backends.cudnn.allow_tf32 is want
            assert torch.backends.cudnn.allow_tf32 is (not want)
        assert FlowStageModel({"model": "pwc", "learning_rate": 1e-3}).conv_math == "fp32"
    finally:
        torch.backends.cudnn.allow_tf32 = old
