"""GPU: the DROP-IN path.  The real, unmodified reference (baseline/_ref, installed by oracle/install_ref.py) runs on
the B200 twice on the same weights and inputs: once as shipped (torch's generic CUDA kernels: 81 x slice/mul/mean,
grid_sample, nonzero + scatter_add_ ...) and once after `ocflow_b200.patch.patch_reference()` has rebound its hot-path
symbols to our kernels.  Flows must agree to 1e-4, loss scalars to 1e-3 (BASELINE.json north_star), and the patched run
must actually launch our kernels.

Covers models/networks/cost_volume_flow_net.py (FlowNetCV), pwc_net.py (PWCNet, backwarp + cost volume),
flow_net.py (FlowNet, FPN), cost_volume_flow_occ_net.py (FlowOccNetCV, the "woc" chain with the injected
CostVolumeLayer), flow_occ_net.py (FlowOccNet, in-place `warped2 *= occ`) and models/model.py FlowStageModel
(general_step, general_step_occ, general_step_occ_aware)."""
import re

import pytest
import torch

from conftest import assert_close, assert_scalar_close
from oracle import ref_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False   # parity runs compare fp32 with fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False    # same heuristic conv algorithms in both runs
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old


@pytest.fixture(scope="module")
def R():
    if not ref_loader.available():
        pytest.skip("reference tree not installed (python -m oracle.install_ref)")
    return ref_loader.load()


def _batch(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.rand(B, 6, H, W, generator=g) * 2 - 1
    flow = torch.randn(B, 2, H, W, generator=g) * 5
    occ = (torch.rand(B, 1, H, W, generator=g) < 0.3).float()
    return imgs.cuda(), flow.cuda(), occ.cuda()


def _boost_flow(net, gain):
    """Scale the flow heads so that the predicted flows are a few pixels at every level (probed on CPU with the real
    reference: FlowNetCV / FlowNet / FlowOccNet already are with their default initialisation; PWCNet needs x40 and
    FlowOccNetCV x4) -- otherwise the warps would be tested at the identity only."""
    if gain == 1.0:
        return
    with torch.no_grad():
        for name, p in net.named_parameters():
            if re.search(r"predict_flow\d|net(Two|Thr|Fou|Fiv|Six)\.netSix\.", name):
                p.mul_(gain)


def _grad_cos(net_a, net_b):
    """Cosine between the parameter gradients of the two runs.  Returns the GLOBAL cosine (all parameters concatenated);
    asserts that every parameter carrying a significant share of the gradient (norm >= 1e-3 of the largest) agrees to
    0.99 on its own.  Per-parameter agreement cannot be held tighter: in the reference itself a 1e-6 relative
    perturbation of the weights moves individual parameter gradients by up to 5e-2 (conditioning note)."""
    pairs = []
    for (ka, pa), (kb, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        assert ka == kb
        if pa.grad is None or pb.grad is None:
            assert pa.grad is None and pb.grad is None, ka
            continue
        pairs.append((ka, pa.grad.double().flatten(), pb.grad.double().flatten()))
    biggest = max(float(b.norm()) for _, _, b in pairs)
    for k, a, b in pairs:
        nb = float(b.norm())
        if nb >= 1e-3 * biggest:
            c = float((a * b).sum() / (a.norm() * b.norm()))
            assert c > 0.99, "gradient of %s: cosine %.5f (|g| = %.3e, largest %.3e)" % (k, c, nb, biggest)
    A, Bv = torch.cat([a for _, a, _ in pairs]), torch.cat([b for _, _, b in pairs])
    return float((A * Bv).sum() / (A.norm() * Bv.norm()))


class _Patched:
    """patch_reference() for the duration of a with-block; counts the C-ABI kernel launches made inside."""

    def __enter__(self):
        import ocflow_b200.patch as P
        from ocflow_b200 import _lib

        self.P, self.lib = P, _lib
        P.patch_reference()
        self.before = _lib.launch_count
        return self

    def launches(self):
        return self.lib.launch_count - self.before

    def __exit__(self, *a):
        self.P.unpatch_reference()


def _pair(make_net, run, gain=1.0):
    """(result of the unpatched reference, result on our kernels, launches) for the same weights."""
    torch.manual_seed(0)
    ref_net = make_net().cuda()
    _boost_flow(ref_net, gain)
    out_ref = run(ref_net)
    with _Patched() as ctx:
        ours = make_net().cuda()          # constructed under the patch: per-instance captures point at our ops
        ours.load_state_dict(ref_net.state_dict())
        out_ours = run(ours)
        torch.cuda.synchronize()
        n = ctx.launches()
    return ref_net, ours, out_ref, out_ours, n


def _fwd_bwd(net, x):
    out = net(x)
    outs = out if isinstance(out, (tuple, list)) else (out,)
    sum(o.square().mean() for o in outs).backward()
    return [o.detach() for o in outs]


def test_flownetcv_patched_matches_unpatched(R):
    x = _batch(2, 128, 192)[0]
    ref_net, ours, a, b, n = _pair(R.cost_volume_flow_net.FlowNetCV, lambda net: _fwd_bwd(net, x))
    assert n >= 5 * 4 + 4 * 2, n   # 5 levels x (normalise 2 + corr) fwd+bwd, 4 warps fwd+bwd
    assert float(a[0].abs().max()) > 0.5, "flows too small to exercise the warps"
    assert_close(b[0], a[0], 1e-4, "flow1")
    assert_close(b[1], a[1], 1e-4, "flow_l2")
    assert _grad_cos(ours, ref_net) > 0.999


def test_pwcnet_patched_matches_unpatched(R):
    x = _batch(2, 128, 192, seed=7)[0]
    ref_net, ours, a, b, n = _pair(lambda: R.pwc_net.PWCNet(pre_train=False), lambda net: _fwd_bwd(net, x), gain=40.0)
    assert n >= 5 * 2 + 4 * 2, n
    assert float(a[0].abs().max()) > 0.5
    assert_close(b[0], a[0], 1e-4, "flow1")
    assert_close(b[1], a[1], 1e-4, "flow_l2")
    assert _grad_cos(ours, ref_net) > 0.999


def test_fpn_flownet_patched_matches_unpatched(R):
    x = _batch(2, 128, 192, seed=8)[0]
    ref_net, ours, a, b, n = _pair(R.flow_net.FlowNet, lambda net: _fwd_bwd(net, x))
    assert n >= 5 * 2 + 4 * 2, n
    assert_close(b[0], a[0], 1e-4, "predicted_flow")
    assert _grad_cos(ours, ref_net) > 0.999


def _ref_cost_volume_layer(R):
    """Stand-in for the module the reference imports but does not ship, backed by the REFERENCE's compute_cost_volume
    (correlation_layer.py:7-40): what the unpatched run of the FlowOcc* nets uses.  Parity of the layer itself is unpinned
    by construction."""
    ccv = R.correlation_layer.compute_cost_volume

    class CostVolumeLayer(torch.nn.Module):
        def __init__(self, max_displacement=4):
            super().__init__()
            self.max_displacement = max_displacement

        def forward(self, f1, f2):
            return ccv(f1, f2, self.max_displacement)

    return CostVolumeLayer


def _occ_net_pair(R, modname, clsname, x, gain):
    import importlib

    import ocflow_b200.patch as P
    from ocflow_b200.cost_volume_net import CostVolumeLayer as Ours

    P.install_cost_volume_net()               # makes the module importable at all
    mod = importlib.import_module(modname)
    cls = getattr(mod, clsname)
    mod.CostVolumeLayer = _ref_cost_volume_layer(R)
    try:
        torch.manual_seed(0)
        ref_net = cls().cuda()
        _boost_flow(ref_net, gain)
        a = _fwd_bwd(ref_net, x)
    finally:
        mod.CostVolumeLayer = Ours
    with _Patched() as ctx:
        ours = cls().cuda()
        ours.load_state_dict(ref_net.state_dict())
        b = _fwd_bwd(ours, x)
        torch.cuda.synchronize()
        n = ctx.launches()
    return ref_net, ours, a, b, n


def test_flowoccnetcv_woc_chain_patched_matches_unpatched(R):
    x = _batch(2, 128, 192, seed=9)[0]
    ref_net, ours, a, b, n = _occ_net_pair(R, "models.networks.cost_volume_flow_occ_net", "FlowOccNetCV", x, 4.0)
    assert n >= 5 * 2 + 4 * 2, n
    assert_close(b[0], a[0], 1e-4, "flow")
    assert_close(b[1], a[1], 1e-4, "occ")
    assert _grad_cos(ours, ref_net) > 0.999


def test_flowoccnet_fpn_patched_matches_unpatched(R):
    x = _batch(2, 128, 192, seed=10)[0]
    ref_net, ours, a, b, n = _occ_net_pair(R, "models.networks.flow_occ_net", "FlowOccNet", x, 1.0)
    assert n >= 5 * 2 + 4 * 2, n
    assert_close(b[0], a[0], 1e-4, "flow")
    # the occlusion head ends in sigmoid(10 * x) (flow_occ_net.py:61): 10x the sensitivity of the flow output
    assert_close(b[1], a[1], 1e-3, "occ")
    assert _grad_cos(ours, ref_net) > 0.999


HP = {"model": "pwc", "learning_rate": 1e-5, "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.0, "displacement": 4}


@pytest.mark.parametrize("mode", ["general_step", "general_step_occ", "general_step_occ_aware"])
def test_flow_stage_model_patched_matches_unpatched(R, mode):
    """models/model.py:315-409 of the REAL reference, unpatched vs patched, on CUDA."""
    batch = _batch(2, 128, 192, seed=11)
    hp = dict(HP, occ_aware=(mode == "general_step_occ_aware"), with_occ=(mode == "general_step_occ"))

    def run(m):
        b = batch if mode != "general_step" else batch[:2]
        losses = getattr(m, mode)(b, 0, "train")
        total = m.photo_weight * losses[0] + m.smooth1_weight * losses[1] + m.smooth2_weight * losses[2]
        total.backward()
        return [float(v) for v in losses] + [float(total)]

    torch.manual_seed(0)
    ref_m = R.model.FlowStageModel(hp).cuda()
    a = run(ref_m)
    with _Patched() as ctx:
        ours = R.model.FlowStageModel(hp).cuda()
        ours.load_state_dict(ref_m.state_dict())
        b = run(ours)
        torch.cuda.synchronize()
        n = ctx.launches()
    assert n >= 30, n
    for i, (x, y) in enumerate(zip(b, a)):
        assert_scalar_close(x, y, 1e-3, "%s[%d]" % (mode, i))
    assert _grad_cos(ours.flow_pred, ref_m.flow_pred) > 0.999


@pytest.mark.parametrize("mode", ["general_step", "general_step_occ"])
def test_mirror_steps_match_unpatched_reference(R, mode):
    """Our FlowStageModel mirror (ocflow_b200/flow_stage.py:78-95) against the real reference's step on CUDA: the two
    step variants that had no GPU test in round 1."""
    from ocflow_b200.flow_stage import FlowStageModel

    batch = _batch(2, 128, 192, seed=12)
    hp = dict(HP, occ_aware=False, with_occ=(mode == "general_step_occ"))
    torch.manual_seed(0)
    ref_m = R.model.FlowStageModel(hp).cuda()
    mine = FlowStageModel(hp).cuda()
    mine.flow_pred.load_state_dict(ref_m.flow_pred.state_dict())
    b = batch if mode != "general_step" else batch[:2]
    la = getattr(ref_m, mode)(b, 0, "train")
    lb = getattr(mine, mode)(b, 0, "train")
    for i, (x, y) in enumerate(zip(lb, la)):
        assert_scalar_close(x, y, 1e-3, "%s[%d]" % (mode, i))
    (4.0 * la[0] + 0.5 * la[1]).backward()
    (4.0 * lb[0] + 0.5 * lb[1]).backward()
    assert _grad_cos(mine.flow_pred, ref_m.flow_pred) > 0.999
