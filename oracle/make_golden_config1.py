"""TEST INFRASTRUCTURE ONLY -- tests/golden/config1_flowmodel_256.pt: BASELINE config 1 run through the REAL reference
(`FlowModel({'model':'pwc',...})` forward on a synthetic 2x3x256x256 pair, CPU fp32, plus its supervised general_step).
Run in the build container:  python oracle/make_golden_config1.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import ocflow_oracle as O  # noqa: E402


def main():
    R = ref_loader.load()
    m = R.flow_model.FlowModel({"model": "pwc", "learning_rate": 1e-3, "displacement": 4})
    shapes = {k: tuple(v.shape) for k, v in m.flow_pred.state_dict().items()}
    seed, gain = 21, 0.1
    m.flow_pred.load_state_dict(O.deterministic_state_dict(shapes, seed=seed, flow_gain=gain))
    g = torch.Generator().manual_seed(1234)
    x = (torch.rand(1, 6, 256, 256, generator=g) * 2 - 1).half().float()   # SURVEY.md section 8d config 1 (fp16-representable)
    flow_gt = (torch.randn(1, 2, 256, 256, generator=g) * 5).half().float()
    m.eval()
    with torch.no_grad():
        flow = m(x)
    loss = m.general_step((x, flow_gt), 0, "train")
    loss.backward()
    named = dict(m.flow_pred.named_parameters())
    keep = ["conv1a.0.weight", "conv6_0.0.weight", "predict_flow2.weight", "dc_conv7.weight"]
    out = os.path.join(ROOT, "tests", "golden", "config1_flowmodel_256.pt")
    torch.save(dict(seed=seed, flow_gain=gain, shapes=shapes, x=x.half(), flow_gt=flow_gt.half(), ref_flow=flow, ref_mse=loss.detach(),
                    ref_grads={k: named[k].grad.detach().clone() for k in keep},
                    note="x and flow_gt are stored in fp16 and were rounded to fp16 BEFORE the reference ran"), out)
    print(out, os.path.getsize(out), float(loss), float(flow.abs().max()))


if __name__ == "__main__":
    main()
