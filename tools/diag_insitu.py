"""Prints the in-situ parity table (tests/insitu.py) for one real training step on the GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import insitu
from oracle import ocflow_oracle as O
from ocflow_b200.flow_stage import FlowStageModel

torch.backends.cudnn.allow_tf32 = False
c = torch.load(os.path.join(ROOT, "tests/golden/net_2x64x64.pt"), weights_only=False)
gain = float(sys.argv[1]) if len(sys.argv) > 1 else c["flow_gain"]
seed = int(sys.argv[2]) if len(sys.argv) > 2 else c["seed"]
m = FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5, "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.0})
m.flow_pred.load_state_dict(O.deterministic_state_dict(c["shapes"], seed=seed, flow_gain=gain))
m = m.cuda()
with insitu.recording() as calls:
    loss = m.training_step((c["imgs"].cuda(), c["flow_gt"].cuda(), c["occ_gt"].cuda()), 0)
insitu.check(calls, loss, report=print)
