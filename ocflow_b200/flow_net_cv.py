"""FlowNetCV on the B200 hot path -- host-side mirror of models/networks/cost_volume_flow_net.py:22-246.

Same constructor (`FlowNetCV(displacement=4)`), same parameter names/shapes and creation order (so a reference
checkpoint loads unchanged and `torch.manual_seed(s)` yields the reference's initial weights), same forward contract:
`forward(x[B,6,H,W]) -> (flow1 [B,2,H,W] pixels, flow_l2 [B,2,H/4,W/4] quarter-res pixels)`.

Per pyramid level the reference runs warp -> normalize_features -> compute_cost_volume -> LeakyReLU -> cat as ~1000 ATen
launches; here it is ONE fused op of 4 launches (`ops.level_fused`: warp with the `up_flow*scale` folded in, one
statistics pass, an apply pass that leaves the normalised c1 where the decoder's concat wants it, and the correlation,
which reads it in place and writes the cost volume into the same buffer).  The convolution stacks stay on cuDNN (out of
scope); the x4 flow up-sampling at the end (:245) is `ops.resize_bilinear`.
"""
import os

import torch
import torch.nn as nn

from . import ops

# (name, in, out, stride) of the feature pyramid, in the reference's creation order (:30-47)
_ENCODER = (
    ("conv1a", 3, 16, 2), ("conv1aa", 16, 16, 1), ("conv1b", 16, 16, 1),
    ("conv2a", 16, 32, 2), ("conv2aa", 32, 32, 1), ("conv2b", 32, 32, 1),
    ("conv3a", 32, 64, 2), ("conv3aa", 64, 64, 1), ("conv3b", 64, 64, 1),
    ("conv4a", 64, 96, 2), ("conv4aa", 96, 96, 1), ("conv4b", 96, 96, 1),
    ("conv5a", 96, 128, 2), ("conv5aa", 128, 128, 1), ("conv5b", 128, 128, 1),
    ("conv6aa", 128, 196, 2), ("conv6a", 196, 196, 1), ("conv6b", 196, 196, 1),
)
_PYRAMID = {1: ("conv1a", "conv1aa", "conv1b"), 2: ("conv2a", "conv2aa", "conv2b"), 3: ("conv3a", "conv3aa", "conv3b"),
            4: ("conv4a", "conv4aa", "conv4b"), 5: ("conv5a", "conv5aa", "conv5b"), 6: ("conv6aa", "conv6a", "conv6b")}
_LEVEL_FEAT = {6: 0, 5: 128, 4: 96, 3: 64, 2: 32}       # channels of c1 concatenated at the level
_DENSE_OUT = (128, 128, 96, 64, 32)
# flow is predicted in units of pixels/20 at full resolution: the warp at level l uses 20 / 2**l
_WARP_SCALE = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}
_CONTEXT = ((128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1))  # (out channels, dilation)


class _ConvBlock(nn.Sequential):
    """nn.Sequential(Conv2d(bias=True), LeakyReLU(0.1)) -- the reference's `conv` helper (cost_volume_flow_net.py:11-15), same
    parameter names ('<block>.0.weight', '<block>.0.bias').  On CUDA the convolution runs without its bias and the bias add +
    LeakyReLU are ONE in-place pass (ops.bias_leaky_relu_; backward: leaky_relu' and the bias-gradient sum in one pass) instead of
    ATen's bias-add kernel + LeakyReLU kernel (+ leaky_relu_backward + a separate reduction in the backward)."""
    fused_epilogue = os.environ.get("OCF_FUSED_CONV_EPILOGUE", "1") == "1"

    def forward(self, x):
        conv = self[0]
        if not (self.fused_epilogue and x.is_cuda and x.dtype == torch.float32):
            return super().forward(x)
        y = nn.functional.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        return ops.bias_leaky_relu_(y, conv.bias, self[1].negative_slope)


def _conv_block(cin, cout, stride=1, dilation=1):
    return _ConvBlock(nn.Conv2d(cin, cout, 3, stride=stride, padding=dilation, dilation=dilation, bias=True), nn.LeakyReLU(0.1))


class FlowNetCV(nn.Module):
    def __init__(self, displacement=4):
        super().__init__()
        for name, cin, cout, stride in _ENCODER:
            setattr(self, name, _conv_block(cin, cout, stride))
        self.leakyRELU = nn.LeakyReLU(0.1)
        self.displacement = int(displacement)
        nd = (2 * self.displacement + 1) ** 2
        for lvl in (6, 5, 4, 3, 2):
            width = nd + (_LEVEL_FEAT[lvl] + 4 if lvl < 6 else 0)
            for i, cout in enumerate(_DENSE_OUT):
                setattr(self, "conv%d_%d" % (lvl, i), _conv_block(width, cout))
                width += cout
            setattr(self, "predict_flow%d" % lvl, nn.Conv2d(width, 2, 3, 1, 1, bias=True))
            setattr(self, "deconv%d" % lvl, nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True))
            if lvl > 2:
                setattr(self, "upfeat%d" % lvl, nn.ConvTranspose2d(width, 2, 4, 2, 1, bias=True))
        cin = width
        for i, (cout, dil) in enumerate(_CONTEXT):
            setattr(self, "dc_conv%d" % (i + 1), _conv_block(cin, cout, 1, dil))
            cin = cout
        self.dc_conv7 = nn.Conv2d(cin, 2, 3, 1, 1, bias=True)

    # reference API: FlowNetCV.warp(img, flow) (cost_volume_flow_net.py:121-151, align_corners=False)
    def warp(self, img, flow):
        return ops.warp(img, flow, align_corners=False)

    @staticmethod
    def normalize(feature_list, **kw):
        return ops.normalize_features(feature_list, **kw)

    def pyramid(self, im):
        feats = {}
        t = im
        for lvl in range(1, 7):
            for name in _PYRAMID[lvl]:
                t = getattr(self, name)(t)
            feats[lvl] = t
        return feats

    # hparams-free switch (not a reference hyper-parameter): False runs the level as 5 separate ops + torch.cat
    fused_level = True
    # forward_bidirectional: run the swapped pair's no-grad decode on a side stream (False: one after the other)
    overlap_decodes = os.environ.get("OCF_OVERLAP_DECODES", "1") == "1"
    _side = None

    def _level(self, lvl, c1, c2, up_flow, up_feat):
        if self.fused_level and self.displacement == 4:
            x = ops.level_fused(c1, c2, up_flow, up_feat, flow_scale=_WARP_SCALE.get(lvl, 1.0), leaky_slope=0.1)
            for i in range(5):
                x = torch.cat((getattr(self, "conv%d_%d" % (lvl, i))(x), x), 1)
            return x, getattr(self, "predict_flow%d" % lvl)(x)
        if lvl < 6:
            c2 = ops.warp(c2, up_flow, align_corners=False, flow_scale=_WARP_SCALE[lvl])
        c1, c2 = ops.normalize_features([c1, c2])
        corr = ops.cost_volume(c1, c2, self.displacement, leaky_slope=0.1)
        x = corr if lvl == 6 else torch.cat((corr, c1, up_flow, up_feat), 1)
        for i in range(5):
            x = torch.cat((getattr(self, "conv%d_%d" % (lvl, i))(x), x), 1)
        flow = getattr(self, "predict_flow%d" % lvl)(x)
        return x, flow

    # set by train.TrainStep: called from the backward pass when the gradient of the encoder's output arrives, i.e. when every
    # decoder / context-network gradient is final and only the encoder's backward is left (autograd runs the nodes created first
    # -- the encoder -- last)
    decoder_grads_done_hook = None

    def encoder_parameters(self):
        for name, _, _, _ in _ENCODER:
            yield from getattr(self, name).parameters()

    def pyramids(self, x):
        """Feature pyramids of both images with ONE pass of the (weight-shared) encoder over a 2B batch: the reference runs
        the same 18 convolutions twice (cost_volume_flow_net.py:158-169); per-sample results are identical."""
        B, _, H, W = x.shape
        both = x.reshape(B, 2, 3, H, W).transpose(0, 1).reshape(2 * B, 3, H, W)
        feats = self.pyramid(both)
        if self.decoder_grads_done_hook is not None and feats[6].requires_grad:
            hook = self.decoder_grads_done_hook
            feats[6].register_hook(lambda g: hook())     # returns None: the gradient is not modified
        return {l: f[:B] for l, f in feats.items()}, {l: f[B:] for l, f in feats.items()}

    def decode(self, p1, p2):
        """Coarse-to-fine decoder of cost_volume_flow_net.py:171-246 on two given pyramids."""
        up_flow = up_feat = None
        for lvl in (6, 5, 4, 3, 2):
            feat, flow = self._level(lvl, p1[lvl], p2[lvl], up_flow, up_feat)
            if lvl > 2:
                up_flow = getattr(self, "deconv%d" % lvl)(flow)
                up_feat = getattr(self, "upfeat%d" % lvl)(feat)
        t = feat
        for i in range(1, 7):
            t = getattr(self, "dc_conv%d" % i)(t)
        flow2 = flow + self.dc_conv7(t)
        flow1 = ops.resize_bilinear(flow2, scale_factor=4, mul=20.0)      # F.interpolate(..., align_corners=True) * 20, :245
        return flow1, flow2 * 5.0

    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != 6:
            raise ValueError("FlowNetCV expects [B,6,H,W] (two RGB images), got %s" % (tuple(x.shape),))
        p1, p2 = self.pyramids(x)
        return self.decode(p1, p2)

    def forward_bidirectional(self, x):
        """(flow1, flow_l2) of forward(x) plus the no-grad full-resolution flow of the swapped pair, i.e. what
        models/model.py:380-386 obtains from `self(imgs)` and `self(cat(img2, img1))`.  The swapped evaluation needs the
        pyramids of the same two images (its c1 pyramid is this call's c2 pyramid and vice versa), so the encoder runs once
        instead of twice more; normalize_features statistics stay per evaluation."""
        if x.dim() != 4 or x.shape[1] != 6:
            raise ValueError("FlowNetCV expects [B,6,H,W] (two RGB images), got %s" % (tuple(x.shape),))
        p1, p2 = self.pyramids(x)
        if not (self.overlap_decodes and x.is_cuda):
            flow1, flow_l2 = self.decode(p1, p2)
            with torch.no_grad():
                back_flow1, _ = self.decode({l: f.detach() for l, f in p2.items()}, {l: f.detach() for l, f in p1.items()})
            return flow1, flow_l2, back_flow1
        # The two decodes are independent once the pyramids exist, and the coarse levels (6x8 ... 24x32 pixels per item) leave most
        # of the 148 SMs idle: the no-grad decode of the swapped pair runs on a side stream next to the main one (fork / join by
        # events, so it is captured as a parallel branch of the step's CUDA graph).
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            back_flow1, _ = self.decode({l: f.detach() for l, f in p2.items()}, {l: f.detach() for l, f in p1.items()})
        flow1, flow_l2 = self.decode(p1, p2)
        cur.wait_stream(side)
        back_flow1.record_stream(cur)     # allocated on the side stream, consumed on the current one
        return flow1, flow_l2, back_flow1
