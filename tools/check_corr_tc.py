"""Developer check (GPU box): the tensor-core correlation forward (csrc/corr_tc.cu) against the fp64 CPU oracle on a
list of shapes (regular, ragged, tiny, channel tails), the sign bitmask, the fused-normalisation variant, and timings of the
step's pyramid levels and the KITTI level (run with OCF_CORR_TC=0 to time the fp32 FMA kernels instead)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocflow_b200 import _lib  # noqa: E402
from oracle import ocflow_oracle as O  # noqa: E402


def P(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def run(f1, f2, slope=0.1, norm=None, want_mask=True):
    B, C, H, W = f1.shape
    out = torch.full((B, 81, H, W), float("nan"), device="cuda")
    mask = torch.zeros(B, 81, H, (W + 7) // 8, device="cuda", dtype=torch.uint8) if want_mask else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("ocf_level_corr_fwd", P(f1), P(f2), P(norm), P(out), 0, None, 0, None, P(mask), B, C, H, W, slope, st)
    torch.cuda.synchronize()
    return out, mask


def main():
    torch.manual_seed(0)
    print("tensor-core correlation forward (ocf_level_corr_fwd); OCF_TC_NO_TMA =", os.environ.get("OCF_TC_NO_TMA", "0"))
    worst = 0.0
    for (B, C, H, W) in [(1, 8, 16, 8), (2, 32, 24, 32), (1, 196, 6, 8), (2, 16, 47, 39), (1, 3, 33, 65), (2, 64, 12, 20), (1, 1, 1, 1),
                         (1, 5, 2, 3), (1, 96, 9, 311), (3, 128, 12, 16), (8, 32, 96, 128), (2, 16, 188, 621), (2, 13, 40, 52), (1, 70, 20, 12)]:
        f1 = torch.randn(B, C, H, W)
        f2 = torch.randn(B, C, H, W) + 0.3
        ref = torch.nn.functional.leaky_relu(O.cost_volume(f1.double(), f2.double(), 4), 0.1)
        out, mask = run(f1.cuda(), f2.cuda())
        err = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
        # sign bitmask: bit (x & 7) of byte x >> 3 == (pre-activation > 0)
        bits = torch.zeros(B, 81, H, ((W + 7) // 8) * 8, dtype=torch.bool)
        mc = mask.cpu()
        for j in range(8):
            bits[..., j::8] = ((mc >> j) & 1).bool()
        mask_bad = int((bits[..., :W] != (out.cpu() > 0)).sum())
        worst = max(worst, err)
        print("shape %-18s rel_max %.3e  nan %d  mask mismatches %d" % ((B, C, H, W), err, int(torch.isnan(out).sum()), mask_bad))
    # normalisation folded in: equals corr(normalize([f1, f2])) with the reference's zero padding AFTER normalisation
    f1 = torch.randn(2, 32, 24, 32) * 2 + 1.5
    f2 = torch.randn(2, 32, 24, 32) * 2 + 1.5
    n1, n2 = O.normalize_features([f1.double(), f2.double()])
    ref = torch.nn.functional.leaky_relu(O.cost_volume(n1, n2, 4), 0.1)
    mean = (f1.double().mean() + f2.double().mean()) / 2
    var = (f1.double().var(dim=(1, 2, 3), unbiased=False).mean() + f2.double().var(dim=(1, 2, 3), unbiased=False).mean()) / 2
    norm = torch.tensor([float(mean), float(1.0 / torch.sqrt(var + 1e-16))], device="cuda")
    out, _ = run(f1.cuda(), f2.cuda(), norm=norm)
    print("fused normalisation: rel_max %.3e" % float((out.cpu().double() - ref).abs().max() / ref.abs().max()))
    print("worst rel_max %.3e (bar 1e-4)" % worst)

    # timings (L2 flushed between launches)
    flush = torch.empty(256 * 1024 * 1024, device="cuda")
    for (B, C, H, W) in [(8, 196, 6, 8), (8, 128, 12, 16), (8, 96, 24, 32), (8, 64, 48, 64), (8, 32, 96, 128), (8, 128, 96, 128),
                         (8, 16, 188, 621), (8, 32, 188, 621), (32, 32, 188, 620), (8, 128, 188, 620)]:
        f1 = torch.randn(B, C, H, W, device="cuda")
        f2 = torch.randn(B, C, H, W, device="cuda")
        out = torch.empty(B, 81, H, W, device="cuda")
        mask = torch.zeros(B, 81, H, (W + 7) // 8, device="cuda", dtype=torch.uint8)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("ocf_level_corr_fwd", P(f1), P(f2), None, P(out), 0, None, 0, None, P(mask), B, C, H, W, 0.1, st)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        us = sum(ts) / len(ts)
        n = B * H * W
        nbytes = 4 * n * (2 * C + 81)
        print("time %-20s %8.2f us  %7.1f GB/s (%.3f of 6549)  %6.2f TFLOP/s useful" % ((B, C, H, W), us, nbytes / us / 1e3, nbytes / us / 1e3 / 6548.8,
                                                                                   2 * 81 * C * n / us / 1e6))


if __name__ == "__main__":
    main()
