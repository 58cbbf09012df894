"""Launch selected hot-path kernels of bench.kernel_table a few times (for ncu --set full captures).
usage: python tools/run_kernel.py corr_fwd_L2 corr_bwd_L2 [--reps 3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("names", nargs="+")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--height", type=int, default=384)
ap.add_argument("--width", type=int, default=512)
a = ap.parse_args()
table = bench.kernel_table(a, torch)
for n in a.names:
    fn = table[n][0]
    for _ in range(a.reps):
        fn()
torch.cuda.synchronize()
print("ran", a.names)
