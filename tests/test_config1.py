"""BASELINE config 1: FlowModel({'model':'pwc'}) forward on one synthetic 2x3x256x256 image pair.  The fixture
(tests/golden/config1_flowmodel_256.pt, oracle/make_golden_config1.py) holds the REAL reference's CPU fp32 output, its
supervised MSE (general_step) and a few parameter gradients.  CPU: the oracle reproduces it; GPU: the FlowModel mirror on
the CUDA hot path reproduces it (flow 1e-4, loss 1e-3)."""
import os

import pytest
import torch

from conftest import GOLD, assert_close, assert_scalar_close, load_golden
from oracle import ocflow_oracle as O

C1 = load_golden(os.path.join(GOLD, "config1_flowmodel_256.pt"))


def _sd():
    return O.deterministic_state_dict(C1["shapes"], seed=C1["seed"], flow_gain=C1["flow_gain"])


def test_oracle_reproduces_config1_reference_output():
    x, flow_gt = C1["x"].float(), C1["flow_gt"].float()
    with torch.no_grad():
        flow, _ = O.flownetcv_forward(_sd(), x)
    assert_close(flow, C1["ref_flow"], 1e-5, "config-1 flow (oracle)")
    assert_scalar_close(((flow - flow_gt) ** 2).mean(), C1["ref_mse"], 1e-5, "config-1 mse (oracle)")


def test_flowmodel_mirror_api_on_cpu():
    from ocflow_b200.flow_model import FlowModel

    m = FlowModel({"model": "pwc", "learning_rate": 1e-3})
    assert sorted(m.flow_pred.state_dict().keys()) == sorted(C1["shapes"].keys())
    assert type(m.configure_optimizers()).__name__ == "Adam"
    with pytest.raises(ValueError):
        FlowModel({"model": "simple", "learning_rate": 1e-3})
    with pytest.raises(ValueError):
        m.general_step(torch.zeros(1), 0, "train")


@pytest.mark.gpu
def test_cuda_flowmodel_matches_config1_reference_output():
    from ocflow_b200.flow_model import FlowModel

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        m = FlowModel({"model": "pwc", "learning_rate": 1e-3, "displacement": 4})
        m.flow_pred.load_state_dict(_sd())
        m = m.cuda().eval()
        x, flow_gt = C1["x"].float().cuda(), C1["flow_gt"].float().cuda()
        with torch.no_grad():
            flow = m(x)
        assert_close(flow, C1["ref_flow"], 1e-4, "config-1 flow")
        loss = m.general_step((x, flow_gt), 0, "train")
        assert_scalar_close(loss, C1["ref_mse"], 1e-3, "config-1 mse")
        loss.backward()
        named = dict(m.flow_pred.named_parameters())
        for k, ref in C1["ref_grads"].items():
            g = named[k].grad.detach().cpu().double()
            cos = float((g * ref.double()).sum() / (g.norm() * ref.double().norm()))
            assert cos > 0.9995, "grad %s: cosine %.6f" % (k, cos)
        # batch of two pairs: per-sample results do not depend on the batch composition beyond normalize_features' statistics
        x2 = torch.cat((x, x.flip(0)), 0)
        with torch.no_grad():
            f2 = m(x2)
        assert_close(f2[:1], f2[1:], 5e-5, "identical pairs in one batch")   # cuDNN may tile the two items differently
    finally:
        torch.backends.cudnn.allow_tf32 = old
