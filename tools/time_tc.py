"""Developer aid: time the tensor-core correlation forward at a few shapes (L2 flushed).  usage: python tools/time_tc.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocflow_b200 import _lib
P = lambda t: ctypes.c_void_p(t.data_ptr())
flush = torch.empty(256 * 1024 * 1024, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (B, C, H, W) in [(8, 32, 96, 128), (8, 128, 96, 128), (8, 128, 188, 620)]:
    f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
    out = torch.empty(B, 81, H, W, device="cuda"); mask = torch.zeros(B, 81, H, (W + 7) // 8, device="cuda", dtype=torch.uint8)
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("ocf_level_corr_fwd", P(f1), P(f2), None, P(out), 0, None, 0, None, P(mask), B, C, H, W, 0.1, st)
        e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    print("%-22s %8.2f us" % ((B, C, H, W), sum(ts) / len(ts)))
