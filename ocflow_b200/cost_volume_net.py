"""The module the reference imports but does not ship: models/networks/cost_volume_net.py.

API inferred from the call sites (cost_volume_flow_occ_net.py:53,188,300; flow_occ_net.py:73; flow_occ_net_c.py:26,99;
occlusion_net_c.py:24): `CostVolumeLayer()` / `CostVolumeLayer(10)`, `forward(f1, f2) -> [B,(2d+1)^2,h,w]`, no
parameters.  Normalisation by the channel count follows compute_cost_volume -- parity for this symbol is unpinned
(SURVEY.md section 8a-3)."""
import torch.nn as nn

from . import ops


class CostVolumeLayer(nn.Module):
    def __init__(self, max_displacement=4):
        super().__init__()
        self.max_displacement = int(max_displacement)

    def forward(self, features1, features2):
        return ops.cost_volume(features1, features2, self.max_displacement)

    def extra_repr(self):
        return "max_displacement=%d" % self.max_displacement
