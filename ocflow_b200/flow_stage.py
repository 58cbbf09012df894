"""FlowStageModel on the B200 hot path -- mirror of models/model.py:155-509 for model == 'pwc'.

Same `hparams` dict keys and defaults (models/model.py:161-170), same method names, argument meaning and return
tuples: `general_step` (:315-341), `general_step_occ` (:343-364), `general_step_occ_aware` (:366-409),
`training_step` (:411-436, minus the TensorBoard writes, which need a Lightning logger), `configure_optimizers`
(:508-509).  It is a plain nn.Module: pytorch_lightning is neither needed nor imported.

The occlusion-aware step collapses warp(img2, flow) -> range map -> occlusion mask -> photometric(occ) ->
photometric(1-occ) -> mse -> bce (≈120 ATen launches and 4 host syncs in the reference) into 2 kernels.
"""
import torch
import torch.nn as nn
from torch.optim import Adam

from . import ops
from .flow_net_cv import FlowNetCV


class _ConvMath:
    def __init__(self, allow_tf32):
        self.allow = bool(allow_tf32)

    def __enter__(self):
        self.old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = self.allow

    def __exit__(self, *exc):
        torch.backends.cudnn.allow_tf32 = self.old
        return False


class FlowStageModel(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = dict(hparams)
        self.lr = hparams["learning_rate"]
        self.photo_weight = hparams.get("photo_weight", 1.0)
        self.smooth1_weight = hparams.get("smooth1_weight", 0.0)
        self.smooth2_weight = hparams.get("smooth2_weight", 1.0)
        self.with_occ = hparams.get("with_occ", False)
        self.log_every_n_steps = hparams.get("log_every_n_steps", 20)
        self.occ_aware = hparams.get("occ_aware", False)
        self.displacement = hparams.get("displacement", 4)
        # not a reference hyper-parameter: False evaluates the network twice exactly as models/model.py:380-386 is written
        self.share_encoder = hparams.get("share_encoder", True)
        # not a reference hyper-parameter: math of the cuDNN convolution stacks (the hot-path kernels are fp32 either way).
        # 'fp32' = strict IEEE fp32 convolutions, the parity mode (matches the reference's CPU results to the tolerances of
        # tests/); 'tf32' = torch's own default on Ampere+ (cudnn.allow_tf32), i.e. what the unmodified reference gets on a GPU
        self.conv_math = hparams.get("conv_math", "fp32")
        if self.conv_math not in ("fp32", "tf32"):
            raise ValueError("conv_math must be 'fp32' or 'tf32' (got %r)" % (self.conv_math,))
        model = hparams.get("model", "simple")
        self.model = model
        if model != "pwc":
            # the other reference networks are plain conv stacks without correlation/warping
            raise ValueError("Unsupported model: %s (ocflow_b200 implements the 'pwc' hot path)" % model)
        self.flow_pred = FlowNetCV(displacement=self.displacement)
        self.last_scalars = {}

    def forward(self, x):
        return self.flow_pred(x)

    # ---- reference helper methods (models/model.py:191-305) ----
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

    # ---- steps ----
    @staticmethod
    def _unpack(batch):
        if len(batch) == 2:
            imgs, flow = batch
            return imgs, flow, None
        if len(batch) == 3:
            return batch
        raise ValueError("Not supported dataset")

    def _smoothness(self, img1, flow_l2):
        img1_l2 = ops.resize_bilinear(img1, scale_factor=0.25)            # F.interpolate(..., align_corners=True), :396
        smooth1 = ops.smoothness_loss(img1_l2, flow_l2, 1)
        smooth2 = ops.smoothness_loss(img1_l2, flow_l2, 2)
        return smooth1, smooth2

    def general_step(self, batch, batch_idx, mode):
        imgs, flow, _ = self._unpack(batch)
        img1, img2 = imgs[:, 0:3], imgs[:, 3:6]
        flow_pred, flow_l2 = self(imgs)
        photo, _, flow_error, _ = ops.occ_photo_fused(img1, img2, flow_pred, None, flow, None)
        # without an occlusion map the fused pass computes sum(rho)/(3*N): identical to torch.mean for 3 channels
        smooth1, smooth2 = self._smoothness(img1, flow_l2)
        return photo, smooth1, smooth2, flow_error

    def general_step_occ(self, batch, batch_idx, mode):
        imgs, flow, occ = batch
        img1, img2 = imgs[:, 0:3], imgs[:, 3:6]
        flow_pred, flow_l2 = self(imgs)
        img_warped = ops.warp(img2, flow_pred, align_corners=True)
        photo = ops.photometric_error(img_warped, img1, occ)
        smooth1, smooth2 = self._smoothness(img1, flow_l2)
        flow_error = ops.pair_loss(flow_pred, flow, ops.PAIR_MSE)
        return photo, smooth1, smooth2, flow_error

    def general_step_occ_aware(self, batch, batch_idx, mode):
        imgs, flow, occ = self._unpack(batch)
        img1, img2 = imgs[:, 0:3], imgs[:, 3:6]
        if self.share_encoder:
            # same values as the two evaluations below; the swapped pair reuses this evaluation's feature pyramids
            flow_pred, flow_l2, back_flow_pred = self.flow_pred.forward_bidirectional(imgs)
        else:
            flow_pred, flow_l2 = self(imgs)
            with torch.no_grad():
                back_flow_pred, _ = self(torch.cat((img2, img1), dim=1))
        with torch.no_grad():
            range_map = ops.range_map(back_flow_pred)
        photo, photo_occ, flow_error, occ_error = ops.occ_photo_fused(img1, img2, flow_pred, range_map, flow, occ)
        smooth1, smooth2 = self._smoothness(img1, flow_l2)
        if occ is not None:
            return photo, smooth1, smooth2, flow_error, photo_occ, occ_error
        return photo, smooth1, smooth2, flow_error, photo_occ

    def conv_math_scope(self):
        """Context manager applying hparams['conv_math'] to the cuDNN convolutions run inside it.  The flag is read when a
        convolution is DISPATCHED, so a backward pass must run inside the scope as well (train.TrainStep does)."""
        return _ConvMath(self.conv_math == "tf32")

    def _losses(self, batch, batch_idx, mode):
        with self.conv_math_scope():
            return self._losses_impl(batch, batch_idx, mode)

    def _losses_impl(self, batch, batch_idx, mode):
        if not self.occ_aware:
            if not self.with_occ:
                return self.general_step(batch, batch_idx, mode)
            return self.general_step_occ(batch, batch_idx, mode)
        return self.general_step_occ_aware(batch, batch_idx, mode)

    def training_step(self, batch, batch_idx):
        losses = self._losses(batch, batch_idx, "train")
        photo, smooth1, smooth2 = losses[0], losses[1], losses[2]
        loss = self.photo_weight * photo + self.smooth1_weight * smooth1 + self.smooth2_weight * smooth2
        names = ("photometric", "smooth1", "smooth2", "flow_error", "photometric_occ", "occ_error")
        self.last_scalars = {"train_" + n: v.detach() for n, v in zip(names, losses)}
        return loss

    def validation_step(self, batch, batch_idx):
        with torch.no_grad():
            losses = self._losses(batch, batch_idx, "val")
        return self.photo_weight * losses[0] + self.smooth1_weight * losses[1] + self.smooth2_weight * losses[2]

    test_step = validation_step

    def configure_optimizers(self):
        return Adam(self.parameters(), self.lr)
