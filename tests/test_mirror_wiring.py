"""CPU, build container only: the WIRING of the host-side mirrors (FlowNetCV, FlowStageModel, FlowModel) against the REAL
reference.  The CUDA ops cannot run here, so `ops` is replaced by a stand-in built from the oracle's functions with the
same signatures; what is then compared with the real reference is everything the mirrors themselves decide: the pyramid
/ decoder data flow, flow scales per level, which flow feeds the range map, loss assembly of general_step,
general_step_occ and general_step_occ_aware (both encoder-sharing modes), gradients to the parameters.  (The ops
themselves are compared with the oracle / the reference's fixtures on the GPU.)"""
import types

import pytest
import torch

from conftest import assert_close, assert_scalar_close
from oracle import ocflow_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _oracle_ops():
    ns = types.SimpleNamespace(PAIR_L1=0, PAIR_MSE=1, PAIR_BCE=2, PAIR_FOCAL=3)

    def warp(img, flow, align_corners=True, is_mask=False, occ=None, flow_scale=1.0):
        out = O.warp(img, flow * flow_scale, align_corners, is_mask)
        return out if occ is None else out * occ

    def cost_volume(f1, f2, max_displacement=4, leaky_slope=1.0):
        out = O.cost_volume(f1, f2, max_displacement)
        return out if leaky_slope == 1.0 else torch.nn.functional.leaky_relu(out, leaky_slope)

    def range_map(flow, with_occlusion=False):
        r = O.range_map(flow.detach())
        return (r, O.occlusion_from_range_map(r)) if with_occlusion else r

    def occ_photo_fused(img1, img2, flow, rmap=None, flow_gt=None, occ_gt=None, alpha=0.001):
        warped = O.warp(img2, flow, True)
        occ = O.occlusion_from_range_map(rmap) if rmap is not None else torch.zeros_like(flow[:, :1])
        photo = O.photometric_error(warped, img1, occ)
        photo_occ = O.photometric_error(warped, img1, 1.0 - occ).detach()
        mse = ((flow - flow_gt) ** 2).mean().detach() if flow_gt is not None else torch.zeros(())
        bce = O.binary_cross_entropy(occ_gt, occ).mean().detach() if occ_gt is not None else torch.zeros(())
        return photo, photo_occ, mse, bce

    def smoothness_loss(img, flow, order, alpha=100.0, alpha_rho=0.001):
        return O.first_order_smoothness_loss(img, flow, alpha) if order == 1 else O.second_order_smoothness_loss(img, flow, alpha)

    def pair_loss(a, b, kind):
        assert kind == ns.PAIR_MSE
        return ((a - b) ** 2).mean()

    def level_fused(c1, c2, up_flow=None, up_feat=None, flow_scale=1.0, leaky_slope=0.1):
        # what ops.level_fused computes, spelled with the oracle's functions (cost_volume_flow_net.py:186-190)
        if up_flow is not None:
            c2 = O.warp(c2, up_flow * flow_scale, False)
        c1n, c2n = O.normalize_features([c1, c2])
        corr = torch.nn.functional.leaky_relu(O.cost_volume(c1n, c2n, 4), leaky_slope)
        return corr if up_flow is None else torch.cat((corr, c1n, up_flow, up_feat), 1)

    ns.level_fused = level_fused
    ns.resize_bilinear = lambda x, size=None, scale_factor=None, mul=1.0: O.resize_bilinear(x, size, scale_factor) * mul
    ns.warp, ns.cost_volume, ns.range_map, ns.occ_photo_fused = warp, cost_volume, range_map, occ_photo_fused
    ns.smoothness_loss, ns.pair_loss = smoothness_loss, pair_loss
    ns.normalize_features = lambda fl, **kw: O.normalize_features(fl, **kw)
    ns.photometric_error = lambda p, i, occ=None, alpha=0.001: O.photometric_error(p, i, occ)
    ns.flow_to_warp = O.flow_to_warp
    return ns


@pytest.fixture()
def mirrors(monkeypatch):
    from ocflow_b200 import flow_model, flow_net_cv, flow_stage

    fake = _oracle_ops()
    for mod in (flow_net_cv, flow_stage, flow_model):
        monkeypatch.setattr(mod, "ops", fake)
    return flow_stage.FlowStageModel, flow_model.FlowModel


def _batch(seed, B=1, H=64, W=128):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.rand(B, 6, H, W, generator=g) * 2 - 1
    flow_gt = torch.randn(B, 2, H, W, generator=g) * 5
    occ_gt = (torch.rand(B, 1, H, W, generator=g) < 0.3).float()
    return imgs, flow_gt, occ_gt


@pytest.mark.parametrize("hp,step", [(dict(), "general_step"), (dict(with_occ=True), "general_step_occ"),
                                     (dict(occ_aware=True), "general_step_occ_aware"),
                                     (dict(occ_aware=True, share_encoder=False), "general_step_occ_aware")])
def test_flowstage_mirror_wiring_matches_the_real_reference(mirrors, hp, step):
    FlowStageModel, _ = mirrors
    R = ref_loader.load()
    base = {"model": "pwc", "learning_rate": 1e-5, "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.25}
    ref_hp = dict(base, **{k: v for k, v in hp.items() if k != "share_encoder"})
    ref = R.model.FlowStageModel(ref_hp)
    mine = FlowStageModel(dict(base, **hp))
    shapes = {k: tuple(v.shape) for k, v in ref.flow_pred.state_dict().items()}
    sd = O.deterministic_state_dict(shapes, seed=4, flow_gain=0.1)
    ref.flow_pred.load_state_dict(sd)
    mine.flow_pred.load_state_dict(sd)
    batch = _batch(17)
    want = getattr(ref, step)(batch, 0, "train")
    got = getattr(mine, step)(batch, 0, "train")
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert_scalar_close(a, b, 2e-5)
    # the dispatch of training_step (models/model.py:411-424) and the gradients it sends to the parameters
    loss_ref = ref.photo_weight * want[0] + ref.smooth1_weight * want[1] + ref.smooth2_weight * want[2]
    loss = mine.training_step(batch, 0)
    assert_scalar_close(loss, loss_ref, 2e-5)
    loss_ref.backward()
    loss.backward()
    named = dict(mine.flow_pred.named_parameters())
    checked = 0
    for k, p in ref.flow_pred.named_parameters():
        if p.grad is None or float(p.grad.norm()) < 1e-7:
            continue
        g = named[k].grad
        cos = float((g.double() * p.grad.double()).sum() / (g.double().norm() * p.grad.double().norm()))
        assert cos > 0.999, (k, cos)
        checked += 1
    assert checked > 60


def test_flowmodel_mirror_wiring_matches_the_real_reference(mirrors):
    _, FlowModel = mirrors
    R = ref_loader.load()
    hp = {"model": "pwc", "learning_rate": 1e-3, "displacement": 4}
    ref, mine = R.flow_model.FlowModel(hp), FlowModel(hp)
    shapes = {k: tuple(v.shape) for k, v in ref.flow_pred.state_dict().items()}
    sd = O.deterministic_state_dict(shapes, seed=6, flow_gain=0.1)
    ref.flow_pred.load_state_dict(sd)
    mine.flow_pred.load_state_dict(sd)
    imgs, flow_gt, occ_gt = _batch(23, B=2, H=64, W=64)
    with torch.no_grad():
        assert_close(mine(imgs), ref(imgs), 1e-5, "FlowModel.forward")
        for batch in ((imgs, flow_gt), (imgs, flow_gt, occ_gt)):
            assert_scalar_close(mine.general_step(batch, 0, "val"), ref.general_step(batch, 0, "val"), 1e-5)
