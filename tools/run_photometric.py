"""One launch of each photometric-term kernel at the Sintel shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ocflow_b200 as ocf
B, H, W = 8, 436, 1024
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand(B, 3, H, W, device="cuda", generator=g) * 2 - 1
b = (a + 0.05 * torch.randn(a.shape, device="cuda", generator=g)).requires_grad_(True)
occ = (torch.rand(B, 1, H, W, device="cuda", generator=g) < 0.3).float()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    torch.autograd.grad(ocf.census_loss(b, a, occ, 3), b)
    torch.autograd.grad(ocf.ssim(b, a, 11), b)
torch.cuda.synchronize()
print("ok")
