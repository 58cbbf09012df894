"""GPU: the fused FlowNetCV level op (ops.level_fused: warp -> statistics -> [apply -> fp32 FMA correlation | tensor-core
correlation normalising on load], writing into the concat buffer) against the fp64 oracle chain of cost_volume_flow_net.py:186-190, outputs and gradients,
at pyramid-level shapes (regular and ragged), with and without the warp (coarsest level)."""
import pytest
import torch

from conftest import assert_close
from oracle import ocflow_oracle as O

pytestmark = pytest.mark.gpu


def _chain(c1, c2, up_flow, up_feat, scale):
    if up_flow is not None:
        c2 = O.warp(c2, up_flow * scale, False)
    c1n, c2n = O.normalize_features([c1, c2])
    corr = torch.nn.functional.leaky_relu(O.cost_volume(c1n, c2n, 4), 0.1)
    return corr if up_flow is None else torch.cat((corr, c1n, up_flow, up_feat), 1)


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("B,C,H,W,with_flow", [(2, 32, 24, 32, True), (1, 196, 6, 8, False), (2, 16, 47, 39, True), (3, 128, 12, 16, True),
                                               (2, 64, 48, 64, True), (1, 96, 9, 311, True), (2, 8, 16, 8, False)])
def test_level_fused_matches_oracle_chain(B, C, H, W, with_flow, tensor_cores, monkeypatch):
    from ocflow_b200 import ops

    # both forms of the level: statistics -> apply -> fp32 FMA correlation (default) and statistics -> tcgen05 correlation that
    # normalises on load (ragged rows always take the latter)
    monkeypatch.setattr(ops, "LEVEL_TENSOR_CORES", tensor_cores)

    g = torch.Generator().manual_seed(B * 7919 + C * 31 + H * 7 + W)
    c1 = torch.randn(B, C, H, W, generator=g) * 1.7 + 0.8       # un-normalised features: mean and std away from (0, 1)
    c2 = torch.randn(B, C, H, W, generator=g) * 1.3 + 0.5
    up_flow = torch.randn(B, 2, H, W, generator=g) * 1.5 if with_flow else None
    up_feat = torch.randn(B, 2, H, W, generator=g) if with_flow else None
    scale = 1.25
    ins = [t for t in (c1, c2, up_flow, up_feat) if t is not None]
    ref_in = [t.clone().double().requires_grad_(True) for t in ins]
    ref = _chain(*(ref_in + [None, None])[:4], scale) if with_flow else _chain(ref_in[0], ref_in[1], None, None, scale)
    cot = torch.randn(ref.shape, generator=g)
    ref_grads = torch.autograd.grad((ref * cot.double()).sum(), ref_in)

    cu_in = [t.clone().cuda().requires_grad_(True) for t in ins]
    out = ops.level_fused(cu_in[0], cu_in[1], cu_in[2] if with_flow else None, cu_in[3] if with_flow else None, flow_scale=scale)
    assert_close(out, ref, 1e-4, "level output")
    grads = torch.autograd.grad((out * cot.cuda()).sum(), cu_in)
    for name, a, b in zip(("c1", "c2", "up_flow", "up_feat"), grads, ref_grads):
        assert_close(a, b, 2e-4 if name == "up_flow" else 1e-4, "d level / d " + name)


def test_level_fused_equals_unfused_ops_inside_flownetcv():
    """FlowNetCV with fused_level on / off: same flows, same parameter gradients (same cuDNN convolutions in both runs)."""
    from ocflow_b200.flow_net_cv import FlowNetCV

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        net = FlowNetCV().cuda()
        g = torch.Generator().manual_seed(5)
        x = (torch.rand(2, 6, 128, 192, generator=g) * 2 - 1).cuda()
        res = {}
        for fused in (True, False):
            net.fused_level = fused
            net.zero_grad(set_to_none=True)
            f1, f2 = net(x)
            (f1.square().mean() + f2.square().mean()).backward()
            res[fused] = (f1.detach(), f2.detach(), torch.cat([p.grad.flatten() for p in net.parameters() if p.grad is not None]).double())   # deconv2 is unused (as upstream)
        assert_close(res[True][0], res[False][0], 1e-4, "flow1")
        assert_close(res[True][1], res[False][1], 1e-4, "flow_l2")
        a, b = res[True][2], res[False][2]
        assert float((a * b).sum() / (a.norm() * b.norm())) > 0.9999
    finally:
        torch.backends.cudnn.allow_tf32 = old
