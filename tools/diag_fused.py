import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ocflow_oracle as O
from ocflow_b200 import ops
from ocflow_b200.flow_stage import FlowStageModel
torch.backends.cudnn.allow_tf32 = False
c = torch.load(os.path.join(ROOT, "tests/golden/net_2x64x64.pt"), weights_only=False)
sd32 = O.deterministic_state_dict(c["shapes"], seed=c["seed"], flow_gain=c["flow_gain"])
m = FlowStageModel({"model": "pwc", "occ_aware": True, "learning_rate": 1e-5, "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.0})
m.flow_pred.load_state_dict(sd32); m = m.cuda()
imgs = c["imgs"].cuda()
with torch.no_grad():
    flow, _ = m(imgs)
    back, _ = m(torch.cat((imgs[:, 3:], imgs[:, :3]), 1))
    rmap = ops.range_map(back)
i1, i2 = imgs[:, :3].contiguous(), imgs[:, 3:].contiguous()
f = flow.clone().requires_grad_(True)
p = ops.occ_photo_fused(i1, i2, f, rmap)[0]
(g,) = torch.autograd.grad(p, f)
res = {}
for dt in (torch.float32, torch.float64):
    fo = flow.cpu().to(dt).requires_grad_(True)
    occ = O.occlusion_from_range_map(rmap.cpu().to(dt))
    po = O.photometric_error(O.warp(i2.cpu().to(dt), fo, True), i1.cpu().to(dt), occ)
    (go,) = torch.autograd.grad(po, fo)
    res[dt] = go
    err = (g.cpu().double() - go.double()).abs()
    mx = go.abs().max()
    print(dt, "loss", float(p), float(po), "max err/max", float(err.max() / mx), "n>1e-5:", int((err / mx > 1e-5).sum()), "n>1e-6:", int((err / mx > 1e-6).sum()), "of", err.numel())
    idx = torch.nonzero(err / mx > 1e-5)
    for b, ch, y, x in idx[:8].tolist():
        u, v = float(flow[b, 0, y, x]), float(flow[b, 1, y, x])
        print("   b%d c%d (%d,%d) u=%.5f v=%.5f -> ix=%.6f iy=%.6f  mine %.6e oracle %.6e vis %.4f" % (b, ch, y, x, u, v, x + u, y + v, float(g[b, ch, y, x]), float(go[b, ch, y, x]), 1 - float((1 - rmap[b, 0, y, x].clamp(0, 1)))))
e32 = (res[torch.float32].double() - res[torch.float64]).abs().max() / res[torch.float64].abs().max()
print("oracle fp32 vs fp64:", float(e32))
