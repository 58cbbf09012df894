"""CPU, world_size 2 over gloo: the data-parallel step driver (ocflow_b200/train.py) -- batch sharded by rank, the only
exchange is ONE in-place all-reduce of the flat gradient buffer; afterwards every rank holds the mean gradient and
identical parameters.  (The hot-path kernels need a GPU; the driver logic does not, so a toy model stands in.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class ToyStage(nn.Module):
    """Anything with the FlowStageModel training contract: .lr and .training_step(batch, idx) -> scalar loss."""

    def __init__(self):
        super().__init__()
        self.lr = 1e-2
        self.net = nn.Sequential(nn.Conv2d(6, 8, 3, padding=1), nn.LeakyReLU(0.1), nn.Conv2d(8, 2, 3, padding=1))

    def training_step(self, batch, batch_idx):
        imgs, flow, occ = batch
        return ((self.net(imgs) - flow) ** 2 * (1 - occ)).mean()


class ToyNet(nn.Module):
    """Stand-in for FlowNetCV's contract with the step driver: an encoder whose gradients complete last, a hook that fires
    from the backward pass once everything else is final (flow_net_cv.FlowNetCV.pyramids registers it the same way)."""
    decoder_grads_done_hook = None

    def __init__(self):
        super().__init__()
        self.enc = nn.Sequential(nn.Conv2d(6, 8, 3, padding=1), nn.LeakyReLU(0.1))
        self.dec = nn.Conv2d(8, 2, 3, padding=1)

    def encoder_parameters(self):
        return self.enc.parameters()

    def forward(self, x):
        f = self.enc(x)
        if self.decoder_grads_done_hook is not None and f.requires_grad:
            hook = self.decoder_grads_done_hook
            f.register_hook(lambda g: hook())
        return self.dec(f)


class ToyStageOverlap(nn.Module):
    def __init__(self):
        super().__init__()
        self.lr = 1e-2
        self.flow_pred = ToyNet()

    def training_step(self, batch, batch_idx):
        imgs, flow, occ = batch
        return ((self.flow_pred(imgs) - flow) ** 2 * (1 - occ)).mean()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ocflow_b200.train import FlatGrads, TrainStep, synthetic_batch

        torch.manual_seed(0)                      # identical initial weights on every rank
        model = ToyStage()
        step = TrainStep(model, use_graph=False)
        assert isinstance(step.grads, FlatGrads) and step.grads.flat.numel() == sum(p.numel() for p in model.parameters())
        batch = synthetic_batch(2, 16, 16, "cpu", 1234 + rank)   # rank-dependent shard of the global batch
        # reference: local gradient of this rank, then mean over ranks via all_gather
        ref_model = ToyStage()
        ref_model.load_state_dict(model.state_dict())
        ref_model.training_step(batch, 0).backward()
        local = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean_grad = torch.stack(gathered).mean(0)
        assert not torch.allclose(gathered[0], gathered[1])       # shards really differ

        loss = step.step(batch)
        assert torch.isfinite(loss)
        assert torch.allclose(step.grads.flat, mean_grad, rtol=1e-6, atol=1e-8)
        # every p.grad is a view of the flat buffer (zero-copy single collective)
        for p in model.parameters():
            assert p.grad.data_ptr() >= step.grads.flat.data_ptr()
            assert p.grad.data_ptr() < step.grads.flat.data_ptr() + step.grads.flat.numel() * 4
        # parameters stay bit-identical across ranks after the optimiser step
        flat_p = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        allp = [torch.zeros_like(flat_p) for _ in range(world)]
        dist.all_gather(allp, flat_p)
        assert torch.equal(allp[0], allp[1])
        for _ in range(2):
            step.step(batch)
        # the two-slice exchange (everything but the encoder reduced from inside the backward pass, the encoder slice at the end)
        torch.manual_seed(1)
        m2 = ToyStageOverlap()
        s2 = TrainStep(m2, use_graph=False)
        assert s2._overlap and 0 < s2.grads.n_late < s2.grads.flat.numel()
        r2 = ToyStageOverlap()
        r2.load_state_dict(m2.state_dict())
        r2.training_step(batch, 0).backward()
        local2 = torch.cat([p.grad.reshape(-1) for p in list(r2.flow_pred.enc.parameters()) + list(r2.flow_pred.dec.parameters())])
        g2 = [torch.zeros_like(local2) for _ in range(world)]
        dist.all_gather(g2, local2)
        l2 = s2.step(batch)
        assert s2._early_done and torch.isfinite(l2)
        assert torch.allclose(s2.grads.flat, torch.stack(g2).mean(0), rtol=1e-6, atol=1e-8)
        out.put((rank, float(loss)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_step_averages_gradients_and_keeps_replicas_in_sync():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(out.get(timeout=5) for _ in range(world))
    assert [r for r, _ in got] == [0, 1]


def test_single_process_all_reduce_is_a_noop():
    from ocflow_b200.train import FlatGrads

    m = nn.Linear(3, 2)
    fg = FlatGrads(m.parameters())
    m(torch.ones(1, 3)).sum().backward()
    before = fg.flat.clone()
    fg.all_reduce_mean()
    assert torch.equal(before, fg.flat) and fg.flat.abs().sum() > 0
    fg.zero_()
    assert m.weight.grad.abs().sum() == 0
