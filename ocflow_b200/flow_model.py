"""FlowModel on the B200 hot path -- mirror of models/flow_model.py:17-235 for model == 'pwc' (BASELINE config 1).

Same `hparams` keys (`learning_rate`, `model`, `displacement`), `forward(x[B,6,H,W]) -> flow [B,2,H,W]` (the first output
of FlowNetCV, :43-48), the helper methods of :49-163 (`warp`, `flow_to_warp`, `compute_range_map`), `general_step`
(supervised MSE against the ground-truth flow, :173-186) and the *_step / configure_optimizers entry points.  A plain
nn.Module: pytorch_lightning is neither needed nor imported.
"""
import torch
import torch.nn as nn
from torch.optim import Adam

from . import ops
from .flow_net_cv import FlowNetCV


class FlowModel(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = dict(hparams)
        self.lr = hparams["learning_rate"]
        model = hparams.get("model", "simple")
        self.model = model
        if model != "pwc":
            # the other reference networks are plain conv stacks without correlation/warping
            raise ValueError("Unsupported model: %s (ocflow_b200 implements the 'pwc' hot path)" % model)
        self.flow_pred = FlowNetCV(displacement=hparams.get("displacement", 4))

    def forward(self, x):
        out, _ = self.flow_pred(x)
        return out

    # ---- reference helper methods (models/flow_model.py:49-163) ----
    def warp(self, img, flow):
        return ops.warp(img, flow, align_corners=True)

    def flow_to_warp(self, flow):
        return ops.flow_to_warp(flow)

    def compute_range_map(self, flow):
        return ops.range_map(flow)

    @property
    def is_cuda(self):
        return next(self.parameters()).is_cuda

    def save_state_dict(self, path):
        torch.save(self.state_dict(), path)

    def general_step(self, batch, batch_idx, mode):
        if not isinstance(batch, (list, tuple)):
            raise ValueError("Not supported dataset")
        if len(batch) == 2:
            imgs, flow = batch
        elif len(batch) == 3:
            imgs, flow, _ = batch
        else:
            raise ValueError("Not supported dataset")
        return ops.pair_loss(self(imgs), flow, ops.PAIR_MSE)   # F.mse_loss(flow_pred, flow), :184

    def training_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "train")

    def validation_step(self, batch, batch_idx):
        with torch.no_grad():
            return self.general_step(batch, batch_idx, "val")

    test_step = validation_step

    def configure_optimizers(self):
        return Adam(self.parameters(), self.lr)
