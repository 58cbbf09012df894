"""Step driver: the occlusion-aware unsupervised training step (models/model.py:411-436 + Lightning's
backward / Adam.step, models/model.py:508-509), data-parallel over the GPUs of one box.

One process per GPU.  Work is sharded by batch of image pairs; every hot-path op is per-sample, so the only
exchange is the gradient all-reduce (mean) of the 9.37 M fp32 parameters: all parameter gradients live in ONE flat
buffer (each `p.grad` is a view), reduced in place by NCCL over NVLink/NVSwitch in two slices -- everything but the shared
encoder (36 of the 37.5 MB) on a side stream as soon as the decoders' backward is complete, overlapped with the encoder's
backward, then the encoder slice -- with the 1 / world factor folded into the loss; no copies.  (no collective follows a hot-path
kernel directly, so there is nothing to fuse a collective into.)

The whole step (2 network forwards, loss kernels, backward, all-reduce, Adam) can be captured in one CUDA graph
(`use_graph=True`): shapes are static, every C-ABI call enqueues on torch's current stream and never synchronises.
"""
import contextlib
import os

import torch
import torch.distributed as dist

from .flow_stage import FlowStageModel

DEFAULT_HPARAMS = {
    # shipped config/unsupervised_config.yml:11,23-26,31 of the reference
    "model": "pwc", "occ_aware": True, "photo_weight": 4.0, "smooth1_weight": 0.5, "smooth2_weight": 0.0,
    "displacement": 4, "learning_rate": 1e-5,
}


class FlatGrads:
    """All gradients of `params` as views of one contiguous buffer (zero-copy collectives).  `late` (optional): parameters whose
    gradients complete LAST in the backward pass (the shared encoder); they are placed first, so that flat[n_late:] -- everything
    else -- is one contiguous slice that can be reduced while the backward pass is still running."""

    def __init__(self, params, late=()):
        late_ids = {id(p) for p in late}
        params = [p for p in params if p.requires_grad]
        self.params = [p for p in params if id(p) in late_ids] + [p for p in params if id(p) not in late_ids]
        self.n_late = sum(p.numel() for p in self.params if id(p) in late_ids)
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, device=ref.device, dtype=ref.dtype)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero_(self):
        self.flat.zero_()

    def all_reduce_sum(self, lo=0, hi=None, group=None):
        """flat[lo:hi] <- sum over ranks, in place.  No-op without an initialised process group / with world size 1."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        part = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
        if part.numel():
            dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)

    def all_reduce_mean(self, group=None):
        """grad <- mean over ranks.  No-op without an initialised process group / with world size 1."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


class TrainStep:
    """model + Adam + (optional) whole-step CUDA graph.  `step(batch)` returns the detached loss (device tensor)."""

    def __init__(self, model, lr=None, use_graph=False, group=None):
        self.model = model
        self.group = group
        self.use_graph = bool(use_graph)
        lr = model.lr if lr is None else lr
        # Gradient exchange: the loss is scaled by 1 / world, so a SUM all-reduce yields the mean gradient (no extra pass over the
        # buffer).  With a FlowNetCV inside, everything but the shared encoder is reduced on a side stream as soon as the decoders'
        # backward is complete -- overlapped with the encoder's backward -- and the encoder slice follows at the end.
        net = getattr(model, "flow_pred", None)
        late = list(net.encoder_parameters()) if hasattr(net, "encoder_parameters") else []
        self.grads = FlatGrads(model.parameters(), late=late)
        self._overlap = bool(late) and hasattr(net, "decoder_grads_done_hook") and os.environ.get("OCF_EARLY_ALLREDUCE", "1") == "1"
        self._early_done = False
        self._side = None
        if self._overlap:
            net.decoder_grads_done_hook = self._reduce_early
        # one fused multi-tensor Adam kernel on CUDA (same update formula as the reference's torch.optim.Adam, models/model.py:508-509)
        on_cuda = next(model.parameters()).is_cuda
        fused = on_cuda and os.environ.get("OCF_FUSED_ADAM", "1") == "1"
        self.opt = torch.optim.Adam(model.parameters(), lr, capturable=self.use_graph, foreach=None if fused else True, fused=fused or None)
        self.graph = None
        self.static_batch = None
        self.static_loss = None

    def _world(self):
        if os.environ.get("OCF_DIAG_NO_ALLREDUCE") == "1":   # diagnosis only: ranks run as independent replicas (slowest-GPU step time)
            return 1
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def _reduce_early(self):
        """Called from the backward pass (autograd hook on the encoder's output) once every non-encoder gradient is final."""
        if self._world() == 1 or self._early_done:
            return
        cur = torch.cuda.current_stream() if self.grads.flat.is_cuda else None
        if cur is not None:
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.grads.all_reduce_sum(self.grads.n_late, None, self.group)
        else:
            self.grads.all_reduce_sum(self.grads.n_late, None, self.group)
        self._early_done = True

    def _eager(self, batch):
        self.grads.zero_()
        world = self._world()
        self._early_done = False
        scope = self.model.conv_math_scope() if hasattr(self.model, "conv_math_scope") else contextlib.nullcontext()
        with scope:   # forward AND backward convolutions under hparams['conv_math']
            loss = self.model.training_step(batch, 0)
            (loss if world == 1 else loss * (1.0 / world)).backward()
        if world > 1:
            if self._early_done:
                self.grads.all_reduce_sum(0, self.grads.n_late, self.group)          # the encoder slice
                if self._side is not None:
                    torch.cuda.current_stream().wait_stream(self._side)
            else:
                self.grads.all_reduce_sum(0, None, self.group)
        self.opt.step()
        return loss.detach()

    def close(self):
        """Release the captured graph (it holds the NCCL all-reduce of the step: the process group cannot be destroyed while
        a graph that references its communicator is alive) and wait for the device.  After this,
        torch.distributed.destroy_process_group() returns normally."""
        self.graph = None
        self.static_loss = None
        self.static_batch = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _capture(self, batch):
        self.static_batch = tuple(t.clone() for t in batch)
        # The warm-up runs real steps (cuDNN autotune, lazy module loads, Adam state allocation).  Weights and optimizer
        # state are snapshotted before and restored after it, so the captured trajectory is exactly the eager one: one
        # update per batch, Adam's step count starting at 0 (the reference / use_graph=False behaviour).
        model_state = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._eager(self.static_batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for k, v in self.model.state_dict().items():
                v.copy_(model_state[k])
            for st in self.opt.state.values():   # keep the (capturable, device-resident) state tensors, reset their contents
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager(self.static_batch)

    def step(self, batch):
        if not self.use_graph:
            return self._eager(batch)
        if self.graph is None:
            self._capture(batch)
        for dst, src in zip(self.static_batch, batch):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss


def synthetic_batch(batch_size, height, width, device, seed):
    """images uniform in [-1,1], flow_gt ~ 5*N(0,1), occ_gt ~ Bernoulli(0.3)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    imgs = torch.rand(batch_size, 6, height, width, generator=g) * 2 - 1
    flow = torch.randn(batch_size, 2, height, width, generator=g) * 5
    occ = (torch.rand(batch_size, 1, height, width, generator=g) < 0.3).float()
    return tuple(t.to(device) for t in (imgs, flow, occ))


def build_model(hparams=None, device="cuda", seed=0):
    torch.manual_seed(seed)
    model = FlowStageModel(dict(DEFAULT_HPARAMS, **(hparams or {})))
    return model.to(device)
