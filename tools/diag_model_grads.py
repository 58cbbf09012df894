"""Tie-breaker for the model-level gradient parity: fp64 oracle vs (a) the fp32 reference fixture, (b) the CUDA path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ocflow_oracle as O
from ocflow_b200.flow_stage import FlowStageModel

c = torch.load(os.path.join(ROOT, "tests/golden/net_2x64x64.pt"), weights_only=False)
sd32 = O.deterministic_state_dict(c["shapes"], seed=c["seed"], flow_gain=c["flow_gain"])
sd = {k: v.double().requires_grad_(True) for k, v in sd32.items()}
batch = (c["imgs"].double(), c["flow_gt"].double(), c["occ_gt"].double())
O.total_loss(O.occ_aware_step(sd, batch)).backward()
torch.backends.cudnn.allow_tf32 = False
m = FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5, "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.0})
m.flow_pred.load_state_dict(sd32)
m = m.cuda()
m.training_step((c["imgs"].cuda(), c["flow_gt"].cuda(), c["occ_gt"].cuda()), 0).backward()
named = dict(m.flow_pred.named_parameters())
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())
for k, ref in c["ref_grads"].items():
    print("%-22s ref32-vs-fp64 %.2e/%.2e   cuda-vs-fp64 %.2e/%.2e   cuda-vs-ref32 %.2e/%.2e" % ((k,) + rel(ref, sd[k].grad) + rel(named[k].grad, sd[k].grad) + rel(named[k].grad, ref)))

# the ORACLE's own torch code on the GPU (cuDNN convs + ATen ops, none of our kernels): how much of the gap is GPU conv noise?
sdg = {k: v.cuda().requires_grad_(True) for k, v in sd32.items()}
O.total_loss(O.occ_aware_step(sdg, (c["imgs"].cuda(), c["flow_gt"].cuda(), c["occ_gt"].cuda()))).backward()
for k in c["ref_grads"]:
    print("%-22s oracle-on-GPU-vs-fp64 %.2e/%.2e   cuda-vs-oracle-on-GPU %.2e/%.2e" % ((k,) + rel(sdg[k].grad, sd[k].grad) + rel(named[k].grad, sdg[k].grad)))
