"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the REAL OCFlow reference.

Run in the build container (needs /root/reference or $OCFLOW_REF):

    python oracle/make_golden.py

Every tensor stored under ``ref_*`` was produced by the unmodified reference functions (imported by
oracle/ref_loader.py); the inputs are stored beside them so the fixtures are self-contained on the GPU
box, where the reference tree does not exist.  The parity tests compare (1) the oracle restatement and
(2) the CUDA path against these.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import ocflow_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FLOW_GAIN = 0.1  # see ocflow_oracle.deterministic_state_dict: keeps the model-level fixture well conditioned


def _grads(fn, inputs, seed):
    """Run fn(*inputs) with grads; returns (outputs, grads wrt every float input that takes part)."""
    leaves = [t.clone().requires_grad_(True) for t in inputs]
    out = fn(*leaves)
    outs = out if isinstance(out, (list, tuple)) else [out]
    g = torch.Generator().manual_seed(seed)
    cot = [torch.randn(o.shape, generator=g) for o in outs]
    total = sum((o * c).sum() for o, c in zip(outs, cot))
    grads = torch.autograd.grad(total, leaves, allow_unused=True)
    return [o.detach() for o in outs], cot, [None if x is None else x.detach() for x in grads]


def op_case(R, name, B, C, H, W, seed, flow_scale, d_list=(4,), integer_flow=False):
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.randn(B, C, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g) * flow_scale
    if integer_flow:
        flow = torch.round(flow)
    i1 = torch.rand(B, 3, H, W, generator=g)
    i2 = torch.rand(B, 3, H, W, generator=g)
    occ_soft = torch.rand(B, 1, H, W, generator=g)
    case = dict(name=name, f1=f1, f2=f2, flow=flow, i1=i1, i2=i2, occ_soft=occ_soft)
    net = R.cost_volume_flow_net.FlowNetCV()
    stage = R.model.FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5})

    for d in d_list:
        o, cot, gr = _grads(lambda a, b: R.correlation_layer.compute_cost_volume(a, b, d), [f1, f2], seed + 1)
        case["ref_corr_d%d" % d] = o[0]
        case["cot_corr_d%d" % d] = cot[0]
        case["ref_corr_d%d_grads" % d] = gr

    flags = [dict(), dict(center=False), dict(normalize=False), dict(moments_across_channels=False),
             dict(moments_across_images=False), dict(moments_across_channels=False, moments_across_images=False)]
    case["norm_flags"] = flags
    for i, kw in enumerate(flags):
        o, cot, gr = _grads(lambda a, b: R.correlation_layer.normalize_features([a, b], **kw), [f1, f2], seed + 2 + i)
        case["ref_norm_%d" % i] = o
        case["cot_norm_%d" % i] = cot
        case["ref_norm_%d_grads" % i] = gr

    o, cot, gr = _grads(lambda a, f: R.utils.warp(a, f), [f2, flow], seed + 10)
    case["ref_warp_ac1"], case["cot_warp"], case["ref_warp_ac1_grads"] = o[0], cot[0], gr
    o, _, gr = _grads(lambda a, f: net.warp(a, f), [f2, flow], seed + 10)
    case["ref_warp_ac0"], case["ref_warp_ac0_grads"] = o[0], gr
    o, _, gr = _grads(lambda a, f: R.utils.warp(a, f, is_mask=True), [f2, flow], seed + 10)
    case["ref_warp_mask"], case["ref_warp_mask_grads"] = o[0], gr
    case["ref_backwarp"] = R.pwc_net.backwarp(f2, flow).detach()
    case["ref_stage_warp"] = stage.warp(i2, flow).detach()

    rm = stage.compute_range_map(flow)
    case["ref_range_map"] = rm
    case["ref_flow_to_warp"] = stage.flow_to_warp(flow.permute(0, 2, 3, 1))
    occ = 1.0 - torch.clamp(rm, 0.0, 1.0)
    case["ref_occ"] = occ

    o, cot, gr = _grads(lambda a, b: R.model.photometric_error(a, b), [i2, i1], seed + 20)
    case["ref_photo"], case["ref_photo_grads"] = o[0], gr
    o, cot, gr = _grads(lambda a, b, c: R.model.photometric_error(a, b, c), [i2, i1, occ_soft], seed + 20)
    case["ref_photo_occ"], case["ref_photo_occ_grads"] = o[0], gr
    case["ref_photo_hardocc"] = R.model.photometric_error(i2, i1, occ)
    o, cot, gr = _grads(lambda a: R.model.robust_l1(a), [f1], seed + 21)
    case["ref_robust_l1"], case["cot_robust_l1"], case["ref_robust_l1_grads"] = o[0], cot[0], gr
    case["ref_charbonnier"] = R.utils.charbonnier_loss(f1)
    case["ref_charbonnier_map"] = R.utils.charbonnier_loss(f1, reduction=False)

    img_s = i1 * 0.02  # alpha=100 on raw uniform noise underflows the edge weights to 0
    case["img_smooth"] = img_s
    o, cot, gr = _grads(lambda a, f: R.model.first_order_smoothness_loss(a, f), [img_s, flow], seed + 30)
    case["ref_smooth1"], case["ref_smooth1_grads"] = o[0], gr
    o, cot, gr = _grads(lambda a, f: R.model.second_order_smoothness_loss(a, f), [img_s, flow], seed + 31)
    case["ref_smooth2"], case["ref_smooth2_grads"] = o[0], gr
    o, cot, gr = _grads(lambda a, f: R.model.first_order_smoothness_loss(a, f), [img_s, occ_soft], seed + 32)
    case["ref_smooth1_1ch"], case["ref_smooth1_1ch_grads"] = o[0], gr
    gx, gy = R.model.gradient(i1, stride=2)
    case["ref_gradient_s2"] = [gx, gy]

    case["ref_ssim11"] = R.ssim.ssim(i1, i2, 11) if min(H, W) >= 6 else None
    win = R.ssim.create_window(4, 3)
    case["ref_ssim4_map_mean"] = R.ssim._ssim(i1, i2, win, 4, 3, True)

    # supervised losses (a-13): F.mse_loss / F.l1_loss / BCE / focal as the reference composes them
    import torch.nn.functional as F
    p = occ_soft.clamp(1e-4, 1 - 1e-4)
    tgt = (i1[:, :1] > 0.5).float()
    case["occ_prob"], case["occ_tgt"] = p, tgt
    case["ref_mse"] = F.mse_loss(flow, flow * 0.5 + 0.1)
    case["ref_l1"] = F.l1_loss(flow, flow * 0.5 + 0.1)
    case["ref_bce"] = F.binary_cross_entropy(p, tgt)
    bce = F.binary_cross_entropy(p, tgt, reduction="none")
    case["ref_focal"] = ((1 - torch.exp(-bce)) ** 2 * bce).mean()
    case["ref_bce_swapped"] = F.binary_cross_entropy(tgt, occ)  # models/model.py:407 argument order
    return case


def net_case(R, B, H, W, seed):
    net = R.cost_volume_flow_net.FlowNetCV()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = O.deterministic_state_dict(shapes, seed=seed, flow_gain=FLOW_GAIN)
    stage = R.model.FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5,
                                    "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.0})
    stage.flow_pred.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 7)
    imgs = torch.rand(B, 6, H, W, generator=g) * 2 - 1
    flow_gt = torch.randn(B, 2, H, W, generator=g) * 5
    occ_gt = (torch.rand(B, 1, H, W, generator=g) < 0.3).float()
    with torch.no_grad():
        flow1, flow_l2 = stage(imgs)
    losses = stage.general_step_occ_aware((imgs, flow_gt, occ_gt), 0, "train")
    loss = 4.0 * losses[0] + 0.5 * losses[1] + 0.0 * losses[2]
    loss.backward()
    keep = ["conv1a.0.weight", "conv3b.0.bias", "conv6_0.0.weight", "predict_flow6.weight", "upfeat5.weight",
            "conv2_4.0.bias", "predict_flow2.bias", "dc_conv7.weight", "deconv3.weight"]
    named = dict(stage.flow_pred.named_parameters())
    grads = {k: named[k].grad.detach().clone() for k in keep}
    gnorm = {k: float(p.grad.norm()) for k, p in named.items() if p.grad is not None}
    return dict(shapes=shapes, seed=seed, flow_gain=FLOW_GAIN, imgs=imgs, flow_gt=flow_gt, occ_gt=occ_gt, ref_flow1=flow1, ref_flow_l2=flow_l2,
                ref_losses=[x.detach() for x in losses], ref_total=loss.detach(), ref_grads=grads, ref_grad_norms=gnorm)


def main():
    R = ref_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    cases = [
        op_case(R, "survey_2x8x6x7", 2, 8, 6, 7, 1234, 1.5, d_list=(4, 10)),
        op_case(R, "odd_2x12x13x19", 2, 12, 13, 19, 77, 3.0, d_list=(4,)),
        op_case(R, "intflow_1x6x16x20", 1, 6, 16, 20, 99, 2.0, d_list=(4,), integer_flow=True),
        op_case(R, "bigflow_2x5x9x12", 2, 5, 9, 12, 5, 12.0, d_list=(4, 10)),
    ]
    for c in cases:
        torch.save(c, os.path.join(GOLD, "ops_%s.pt" % c["name"]))
    torch.save(net_case(R, 2, 64, 64, 3), os.path.join(GOLD, "net_2x64x64.pt"))
    # the survey's hand-checkable impulse known-answer
    f1 = torch.zeros(1, 2, 7, 7)
    f2 = torch.zeros(1, 2, 7, 7)
    f1[0, :, 3, 3] = 1
    f2[0, :, 1, 5] = 2
    torch.save(dict(f1=f1, f2=f2, ref=R.correlation_layer.compute_cost_volume(f1, f2, 4)), os.path.join(GOLD, "kat_impulse.pt"))
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
