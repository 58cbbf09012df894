"""Developer aid (torchrun): time a plain NCCL all-reduce of the flat gradient buffer (9.37 M fp32 = 37.5 MB), eager, CUDA events."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
x = torch.randn(9374340, device="cuda")
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
if rank == 0: print("all-reduce 37.5 MB x%d ranks: %.1f us" % (dist.get_world_size(), e0.elapsed_time(e1) * 1e3 / 20))
dist.destroy_process_group()
