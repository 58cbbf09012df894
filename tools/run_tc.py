"""Developer aid: launch the tensor-core correlation forward (ocf_level_corr_fwd, no normalisation) a few times at one shape
(for ncu captures).  usage: python tools/run_tc.py B C H W [reps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocflow_b200 import _lib
B, C, H, W = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
out = torch.empty(B, 81, H, W, device="cuda")
msk = torch.zeros(B, 81, H, (W + 7) // 8, device="cuda", dtype=torch.uint8)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(reps):
    _lib.call("ocf_level_corr_fwd", P(f1), P(f2), None, P(out), 0, None, 0, None, P(msk), B, C, H, W, 0.1, st)
torch.cuda.synchronize()
print("ok")
