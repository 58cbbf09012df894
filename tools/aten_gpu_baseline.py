"""Developer measurement ("like-for-like GPU baseline"): the SAME training step as bench.py, on the
same B200, but with the hot path expressed as plain ATen ops -- the oracle's restatement of the reference algorithm moved
to the GPU (81 slice-multiply-mean planes for the cost volume, index_put_ range map, elementwise losses ...), convolutions
on the same strict-fp32 cuDNN.  The reference itself cannot travel to the GPU box; this is its port run on CUDA, so the
difference to bench.py's `value` is what the hand-written kernels buy on identical hardware and conv math.

Also prints per-op times (ATen composition vs ocflow_b200 kernel) at the L2 level of config 2.
    python tools/aten_gpu_baseline.py [--steps 5] > profiles/r1_aten_gpu_baseline.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ocflow_oracle as O  # noqa: E402  (developer tool: the oracle is the ATen restatement being timed)
from ocflow_b200 import ops  # noqa: E402
from ocflow_b200.flow_net_cv import FlowNetCV  # noqa: E402
from ocflow_b200.train import DEFAULT_HPARAMS, TrainStep, build_model, synthetic_batch  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W = a.batch, 384, 512
    batch = synthetic_batch(B, H, W, "cuda", 1234)

    # ---- whole step, ATen hot path ----
    torch.manual_seed(0)
    net = FlowNetCV(DEFAULT_HPARAMS["displacement"])
    sd = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in net.state_dict().items()}
    opt = torch.optim.Adam(list(sd.values()), DEFAULT_HPARAMS["learning_rate"])

    def aten_step():
        opt.zero_grad(set_to_none=True)
        losses = O.occ_aware_step(sd, batch, DEFAULT_HPARAMS["displacement"])
        loss = O.total_loss(losses, DEFAULT_HPARAMS["photo_weight"], DEFAULT_HPARAMS["smooth1_weight"], DEFAULT_HPARAMS["smooth2_weight"])
        loss.backward()
        opt.step()
        return loss

    t_aten = timed(aten_step, reps=a.steps, warm=3)
    loss_aten = float(aten_step().detach())
    del sd, opt
    torch.cuda.empty_cache()

    # ---- whole step, ocflow_b200 kernels (eager and CUDA graph) ----
    ours = {}
    for use_graph in (False, True):
        step = TrainStep(build_model(seed=0), use_graph=use_graph)
        ours[use_graph] = timed(lambda: step.step(batch), reps=a.steps, warm=4)
        del step
        torch.cuda.empty_cache()
    print("# B200, strict fp32 convs, batch %d pairs of %dx%d, one occlusion-aware training step (fwd + bwd + Adam)" % (B, H, W))
    print("ATen hot path (oracle port on CUDA), eager : %9.2f ms/step  %7.1f pairs/s   (loss %.5f)" % (t_aten / 1e3, B / (t_aten * 1e-6), loss_aten))
    print("ocflow_b200 kernels, eager                 : %9.2f ms/step  %7.1f pairs/s" % (ours[False] / 1e3, B / (ours[False] * 1e-6)))
    print("ocflow_b200 kernels, whole-step CUDA graph : %9.2f ms/step  %7.1f pairs/s" % (ours[True] / 1e3, B / (ours[True] * 1e-6)))

    # ---- per-op, L2 level of config 2 (B=8, C=32, 96x128) and the loss level ----
    g = torch.Generator(device="cuda").manual_seed(1)
    C, h, w = 32, 96, 128
    f1 = torch.randn(B, C, h, w, device="cuda", generator=g).requires_grad_(True)
    f2 = torch.randn(B, C, h, w, device="cuda", generator=g).requires_grad_(True)
    fl = (torch.randn(B, 2, h, w, device="cuda", generator=g) * 2).requires_grad_(True)
    cot = torch.randn(B, 81, h, w, device="cuda", generator=g)
    cotw = torch.randn(B, C, h, w, device="cuda", generator=g)
    big = (torch.randn(B, 2, H, W, device="cuda", generator=g) * 5)
    lrelu = torch.nn.functional.leaky_relu
    rows = [
        ("cost volume + LeakyReLU fwd+bwd (L2)", lambda: torch.autograd.grad(lrelu(O.cost_volume(f1, f2, 4), 0.1), (f1, f2), cot),
         lambda: torch.autograd.grad(ops.cost_volume(f1, f2, 4, 0.1), (f1, f2), cot)),
        ("warp align_corners=False fwd+bwd (L2)", lambda: torch.autograd.grad(O.warp(f2, fl, False), (f2, fl), cotw),
         lambda: torch.autograd.grad(ops.warp(f2, fl, align_corners=False), (f2, fl), cotw)),
        ("normalize_features fwd+bwd (L2)", lambda: torch.autograd.grad(sum((y * cotw).sum() for y in O.normalize_features([f1, f2])), (f1, f2)),
         lambda: torch.autograd.grad(sum((y * cotw).sum() for y in ops.normalize_features([f1, f2])), (f1, f2))),
        ("range map (384x512)", lambda: O.range_map(big), lambda: ops.range_map(big)),
    ]
    print("%-42s %12s %12s %8s" % ("op", "ATen us", "ours us", "ratio"))
    for name, fa, fo in rows:
        ta, to = timed(fa), timed(fo)
        print("%-42s %12.1f %12.1f %8.1fx" % (name, ta, to, ta / to))


if __name__ == "__main__":
    main()
